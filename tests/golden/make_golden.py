#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own code (its unmodified translation units behind
oracle/ref_harness.cpp, built from /root/reference by oracle/Makefile) on small seeded inputs.

Run in the authoring container (where /root/reference exists):   python tests/golden/make_golden.py
The fixtures hold the inputs (transcripts, reads) and the reference's outputs (per-read sketch sets,
per-read candidate lists, pi, NumReads, presence), so the C restatement and the CUDA path can be checked
against the reference wherever the fixtures travel.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle_py  # noqa: E402
from datasets import SKETCH, dataset  # noqa: E402


def run(name, ks, d, fraction=0.9, sketch=SKETCH, iters=20, tol=0.01):
    port = oracle_py.PortOracle()
    r = oracle_py.RefOracle(ks)
    r.set_transcripts(d["names"])
    # the index postings come from the reference's own sketch function
    postings = {}
    for k in ks:
        pairs = []
        for t, s in enumerate(d["tseqs"]):
            if len(s) < max(ks):
                continue
            for h in oracle_py.RefOracle.sketch_of(s, k, sketch):
                pairs.append((int(h), t))
        pairs.sort()
        keys = sorted(set(h for h, _ in pairs))
        off, tids, i = [0], [], 0
        for kk in keys:
            while i < len(pairs) and pairs[i][0] == kk:
                tids.append(pairs[i][1])
                i += 1
            off.append(len(tids))
        postings[k] = (np.array(keys, np.uint32), np.array(off, np.uint64), np.array(tids, np.uint32))
        r.set_postings(k, *postings[k])
    for i, s in enumerate(d["reads"]):
        r.add_read(b"r%d" % i, s, sketch)
    r.chain(fraction)
    r.em(iters, tol)
    r.assign()
    out = {"ks": np.array(ks, np.uint32), "sketch_fraction": np.float64(sketch), "chain_fraction": np.float64(fraction),
           "em_iters": np.int32(iters), "em_tol": np.float64(tol),
           "tseq_blob": np.frombuffer(b"".join(d["tseqs"]), np.uint8),
           "tseq_off": np.cumsum([0] + [len(s) for s in d["tseqs"]]).astype(np.uint64),
           "read_blob": np.frombuffer(b"".join(d["reads"]), np.uint8),
           "read_off": np.cumsum([0] + [len(s) for s in d["reads"]]).astype(np.uint64),
           "names": np.array(d["names"])}
    sk_off, sk = [0], []
    c_off, c_tid, c_score = [0], [], []
    admitted = []
    for i in range(len(d["reads"])):
        rid = b"r%d" % i
        cand = r.read_candidates(rid)
        admitted.append(cand is not None)
        for k in ks:
            s = r.read_sketch(rid, k)
            sk += [] if s is None else s.tolist()
            sk_off.append(len(sk))
        if cand is not None:
            c_tid += cand[0].tolist()
            c_score += cand[1].tolist()
        c_off.append(len(c_tid))
    rc, rp = r.counts()
    out.update(admitted=np.array(admitted), sketch_off=np.array(sk_off, np.uint64), sketch=np.array(sk, np.uint32),
               cand_off=np.array(c_off, np.uint64), cand_tid=np.array(c_tid, np.uint32),
               cand_score=np.array(c_score, np.int32), pi=r.pi(), numreads=rc, present=rp,
               R=np.uint64(r.num_reads()))
    for k in ks:
        out["post_keys_%d" % k], out["post_off_%d" % k], out["post_tid_%d" % k] = postings[k]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "T=%d reads=%d R=%d pairs=%d" % (len(d["names"]), len(d["reads"]), r.num_reads(), len(c_tid)))


if __name__ == "__main__":
    if not oracle_py.have_ref():
        raise SystemExit("oracle/_ref is not built: make -C oracle ref (needs /root/reference)")
    d = dataset(n_genes=12, n_reads=160, seed=3)
    # some reads the reference refuses: non-ACGT character, shorter than max k
    d = dict(d, reads=d["reads"] + [b"ACGTN" * 30, b"ACGT" * 5, d["reads"][0].lower(), d["reads"][1]])
    run("short_k31", [31], d)
    run("short_k21_25_31", [21, 25, 31], d)
    dl = dataset(n_genes=10, n_reads=40, long_reads=(800, 4000), err=0.05, exon_median=300, seed=4)
    run("long_k21_31", [21, 31], dl)
    run("short_k31_scale02", [31], d, sketch=float(np.float32(0.2)), fraction=0.75)
