"""Pins the C restatement against the reference's OWN code (its unmodified translation units behind
oracle/ref_harness.cpp, built by oracle/Makefile where /root/reference exists).  The harness .so travels
to the GPU box as a binary, so these run wherever it is present."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from datasets import SKETCH, csr_to_lists, dataset

pytestmark = pytest.mark.skipif(not oracle_py.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def _ref_run(d, ks, postings, fraction=0.9, iters=20, tol=0.01, sketch=SKETCH):
    r = oracle_py.RefOracle(ks)
    r.set_transcripts(d["names"])
    for k in ks:
        r.set_postings(k, *postings[k])
    for i, s in enumerate(d["reads"]):
        r.add_read(b"r%d" % i, s, sketch)
    r.chain(fraction)
    r.em(iters, tol)
    r.assign()
    return r


@pytest.mark.parametrize("ks", [[31], [21, 25, 31], [15]])
def test_sketch_matches_reference(port, ks):
    d = dataset()
    thr = port.threshold(SKETCH)
    for s in d["tseqs"][:20] + d["reads"][:60]:
        for k in ks:
            if len(s) < k:
                continue
            assert port.sketch(s, k, thr).tolist() == oracle_py.RefOracle.sketch_of(s, k, SKETCH).tolist()


@pytest.mark.parametrize("ks", [[31], [21, 25, 31]])
def test_quant_matches_reference(port, ks):
    d = dataset()
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    r = _ref_run(d, ks, postings)
    adm, off, tid, score, R = port.chain_batch(ks, thr, 0.9, postings, d["reads"])
    assert R == r.num_reads() == int(adm.sum())
    mine = csr_to_lists(off, tid, score)
    n_nonempty = 0
    for i in range(len(d["reads"])):
        rt, rs = r.read_candidates(b"r%d" % i)
        assert sorted(zip(rt.tolist(), rs.tolist())) == mine[i]
        n_nonempty += len(rt) > 0
    assert n_nonempty > len(d["reads"]) // 2
    T = len(d["names"])
    pi, iters = port.em(off, tid, score, R, T)
    nr, present = port.assign(off, tid, score, T, pi)
    np.testing.assert_allclose(pi, r.pi(), rtol=1e-9, atol=0)
    rc, rp = r.counts()
    assert present.tolist() == rp.tolist()
    np.testing.assert_allclose(nr, rc, rtol=1e-9, atol=1e-12)


def test_reference_binary_end_to_end(port, tmp_path, sqb):
    """index + quant with the reference PROGRAM (oracle/_ref/ref_test) and compare its CSV with the oracle"""
    if not os.path.exists(oracle_py.REF_BIN):
        pytest.skip("ref_test binary not built")
    d = dataset()
    fa, fq, idx, csv = (str(tmp_path / n) for n in ("t.fa", "r.fq", "t.idx", "o.csv"))
    with open(fa, "wb") as f:
        for nm, s in zip(d["names"], d["tseqs"]):
            f.write(b">" + nm.encode() + b" some description\n" + s + b"\n")
    with open(fq, "wb") as f:
        for i, s in enumerate(d["reads"]):
            f.write(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    subprocess.run([oracle_py.REF_BIN, "-k", "31", "-o", "index", fa, idx], check=True, capture_output=True)
    out = subprocess.run([oracle_py.REF_BIN, "-o", "quant", idx, fq, csv], check=True, capture_output=True, text=True)
    for marker in ("Loading index completed", "Loading read completed", "Sparse chaining completed",
                   "EM estimation completed", "Read assignment completed", "Output written to"):
        assert marker in out.stdout
    rows = {}
    lines = open(csv).read().splitlines()
    assert lines[0] == "Name,NumReads,EM_Abundance"
    for ln in lines[1:]:
        nm, a, b = ln.split(",")
        rows[nm] = (float(a), float(b))
    # the reference-written index parsed by our reader gives the same postings as the oracle
    ks, names, seqs, postings = sqb.index_io.read_index(idx)
    assert ks == [31] and sorted(names) == sorted(d["names"])
    thr = port.threshold(SKETCH)
    order = [d["names"].index(n) for n in names]
    mine = port.postings_from_sequences([d["tseqs"][i] for i in order], [31], thr)
    assert mine[31][0].tolist() == postings[31][0].tolist()
    assert mine[31][2].tolist() == postings[31][2].tolist()
    _, off, tid, score, R = port.chain_batch([31], thr, 0.9, postings, d["reads"])
    pi, _ = port.em(off, tid, score, R, len(names))
    nr, present = port.assign(off, tid, score, len(names), pi)
    assert set(rows) == {names[i] for i in range(len(names)) if present[i]}
    for i, nm in enumerate(names):
        if present[i]:
            assert rows[nm][0] == pytest.approx(nr[i], rel=1e-5)
            assert rows[nm][1] == pytest.approx(pi[i], rel=1e-5)
