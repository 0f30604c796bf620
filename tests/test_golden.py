"""Golden vectors produced by the reference's own code (tests/golden/make_golden.py): the C restatement must
reproduce them on the CPU, the CUDA path on the GPU."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    g = dict(np.load(path))
    blob, off = g["tseq_blob"].tobytes(), g["tseq_off"]
    g["tseqs"] = [blob[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    blob, off = g["read_blob"].tobytes(), g["read_off"]
    g["reads"] = [blob[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    g["ks_list"] = [int(k) for k in g["ks"]]
    g["postings"] = {k: (g["post_keys_%d" % k], g["post_off_%d" % k], g["post_tid_%d" % k]) for k in g["ks_list"]}
    return g


def _cand_lists(off, tid, score):
    return [sorted(zip(tid[int(off[r]):int(off[r + 1])].tolist(), score[int(off[r]):int(off[r + 1])].tolist()))
            for r in range(len(off) - 1)]


def test_fixtures_exist():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_reference_golden(port, path):
    g = _load(path)
    ks, thr = g["ks_list"], port.threshold(float(g["sketch_fraction"]))
    # index postings: reference sketch function == restatement
    mine = port.postings_from_sequences(g["tseqs"], ks, thr)
    for k in ks:
        for a, b in zip(mine[k], g["postings"][k]):
            assert a.tolist() == b.tolist()
    adm, off, tid, score, R = port.chain_batch(ks, thr, float(g["chain_fraction"]), g["postings"], g["reads"])
    assert adm.tolist() == g["admitted"].tolist() and R == int(g["R"])
    # per-read sketch sets
    i = 0
    for r, s in enumerate(g["reads"]):
        for k in ks:
            want = g["sketch"][int(g["sketch_off"][i]):int(g["sketch_off"][i + 1])]
            i += 1
            if adm[r]:
                assert port.sketch(s, k, thr).tolist() == want.tolist()
            else:
                assert len(want) == 0
    assert _cand_lists(off, tid, score) == _cand_lists(g["cand_off"], g["cand_tid"], g["cand_score"])
    T = len(g["names"])
    pi, _ = port.em(off, tid, score, R, T, int(g["em_iters"]), float(g["em_tol"]))
    nr, present = port.assign(off, tid, score, T, pi)
    np.testing.assert_allclose(pi, g["pi"], rtol=1e-9)
    np.testing.assert_allclose(nr, g["numreads"], rtol=1e-9, atol=1e-12)
    assert present.tolist() == g["present"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_cuda_reproduces_reference_golden(gpu_lib, sqb, path):
    g = _load(path)
    ks = g["ks_list"]
    reads = [s for s, a in zip(g["reads"], g["admitted"]) if a]
    assert [sqb.packing.admit(s, ks) for s in g["reads"]] == g["admitted"].tolist()
    T = len(g["names"])
    with sqb.Engine(ks, T, sketch_fraction=float(g["sketch_fraction"]), chain_fraction=float(g["chain_fraction"])) as e:
        # index built on the GPU from the transcript sequences
        idx = [i for i, s in enumerate(g["tseqs"]) if len(s) >= max(ks)]
        words, off, ln = sqb.packing.pack_reads([g["tseqs"][i] for i in idx])
        for ki, k in enumerate(ks):
            got = e.build_postings(ki, words, off, ln, np.asarray(idx, dtype=np.uint32))
            for a, b in zip(got, g["postings"][k]):
                assert a.tolist() == b.tolist()
            e.load_index(ki, *got)
        words, off, ln = sqb.packing.pack_reads(reads)
        counts, hashes = e.sketch(words, off, ln)
        e.push_reads(words, off, ln)
        coff, tid, score = e.candidates()
        pi, nr, present, it = e.finish(0, int(g["em_iters"]), float(g["em_tol"]))
    # sketch sets of admitted reads
    p, i, j = 0, 0, 0
    for r, a in enumerate(g["admitted"]):
        for ki in range(len(ks)):
            want = g["sketch"][int(g["sketch_off"][i]):int(g["sketch_off"][i + 1])]
            i += 1
            if a:
                c = int(counts[j, ki])
                assert sorted(set(hashes[p:p + c].tolist())) == want.tolist()
                p += c
        j += bool(a)
    keep = np.nonzero(g["admitted"])[0]
    want = _cand_lists(g["cand_off"], g["cand_tid"], g["cand_score"])
    assert _cand_lists(coff, tid, score) == [want[i] for i in keep]
    np.testing.assert_allclose(pi, g["pi"], rtol=1e-9)
    np.testing.assert_allclose(nr, g["numreads"], rtol=1e-9, atol=1e-12)
    assert present.tolist() == g["present"].tolist()
