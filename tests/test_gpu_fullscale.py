"""BASELINE config-2 shapes on the GPU (human-scale index: ~250 k transcripts, ~420 Mbp; millions of reads):
size-independent properties of the path plus spot checks of sampled reads against the CPU oracle.

  * every read's candidate list depends only on that read and the index -> a random sample of reads is
    re-derived by the oracle and must match bit for bit;
  * sum(NumReads) = number of reads with a candidate (each such read distributes exactly 1.0);
  * after any M-step, sum(pi) = (#reads with candidates) + T * ((double)(0.01f/(float)R) + (double)0.01f)
    (isoform_assignment.cpp:54-57), and every pi >= the pseudocount term;
  * pushing the reads in one batch or in many gives identical candidates (batching is invisible);
  * FracMinHash keeps ~5 % of the k-mers.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SKETCH = float(np.float32(0.05))


@pytest.fixture(scope="module")
def big(gpu_lib, sqb):
    syn = sqb.synth
    tx = syn.make_transcriptome(62500, seed=7, device="cuda:0")
    T = tx["t_off"].numel() - 1
    eng = sqb.Engine([31], T, sketch_fraction=SKETCH)
    tlen = tx["t_off"][1:] - tx["t_off"][:-1]
    keep = torch.nonzero(tlen >= 31).flatten()
    words, boff, ln = syn.pack_ragged(tx["codes"], tx["t_off"], align=4)
    post = eng.build_postings(0, syn.to_u32(words), syn.to_u32(boff[keep].contiguous()), syn.to_u32(ln[keep].contiguous()),
                              keep.cpu().numpy().astype(np.uint32))
    eng.load_index(0, *post)
    chunks = []
    for ch in syn.simulate_reads(tx, 3_000_000, 150, seed=99, err=0.005, chunk=1 << 20):
        w, b, l = syn.pack_ragged(ch["codes"], ch["r_off"], align=4)
        chunks.append((w, b, l, int(ch["r_off"][-1])))
    yield {"tx": tx, "T": T, "eng": eng, "post": post, "chunks": chunks}
    eng.close()


def _push_all(eng, chunks):
    eng.reset_reads()
    for w, b, l, nb in chunks:
        eng.push_reads_device(w.data_ptr(), w.numel(), b.data_ptr(), l.data_ptr(), l.numel(), nb + 4 * l.numel())


def test_fullscale_properties_and_sampled_parity(big, sqb, port):
    eng, T = big["eng"], big["T"]
    _push_all(eng, big["chunks"])
    off, tid, score = eng.candidates()
    R = len(off) - 1
    assert R == 3_000_000
    pi, nr, present, it = eng.finish(0, 20, 0.01)
    st = eng.stats()
    ncand = np.diff(off.astype(np.int64))
    with_c = int((ncand > 0).sum())
    assert with_c > 0.9 * R
    # each read with candidates hands out exactly one unit
    assert nr.sum() == pytest.approx(with_c, rel=1e-9)
    const = float(np.float32(0.01) / np.float32(R)) + float(np.float32(0.01))
    assert pi.sum() == pytest.approx(with_c + T * const, rel=1e-9)
    assert pi.min() >= const * (1 - 1e-12)
    assert it == 20
    assert set(np.nonzero(present)[0].tolist()) == set(np.unique(tid).tolist())
    # FracMinHash: ~5 % of the 120 k-mers of each read
    frac = st["sketch_hashes"] / (R * 120.0)
    assert 0.048 < frac < 0.052
    # candidate lists are ordered (score desc, transcript asc) and scores are positive
    same_read = np.repeat(np.arange(R), ncand)
    inner = same_read[1:] == same_read[:-1]
    assert np.all(score[:-1][inner] >= score[1:][inner])
    tie = inner & (score[:-1] == score[1:])
    assert np.all(tid[:-1][tie] < tid[1:][tie])
    assert score.min() >= 1
    # sampled reads, re-derived on the CPU
    rng = np.random.default_rng(1)
    w0, b0, l0, _ = big["chunks"][0]
    W = sqb.synth.to_u32(w0)
    b0, l0 = b0.cpu().numpy(), l0.cpu().numpy()
    sample = np.sort(rng.choice(l0.shape[0], 1500, replace=False))
    seqs = [sqb.packing.unpack_read(W, int(b0[i]), int(l0[i])) for i in sample]
    thr = port.threshold(SKETCH)
    _, ooff, otid, oscore, _ = port.chain_batch([31], thr, 0.9, {31: big["post"]}, seqs)
    for j, i in enumerate(sample):
        a, b = int(off[i]), int(off[i + 1])
        assert tid[a:b].tolist() == otid[int(ooff[j]):int(ooff[j + 1])].tolist(), i
        assert score[a:b].tolist() == oscore[int(ooff[j]):int(ooff[j + 1])].tolist(), i


def test_fingerprint_classes_equal_exact_classes(big, sqb):
    """EM classes found by 128-bit list fingerprints are the classes found by comparing the lists"""
    eng = big["eng"]
    _push_all(eng, big["chunks"])
    eng.set_option("exact_classes", 0)
    pi0, nr0, pr0, _ = eng.finish(0, 20, 0.01)
    c0 = eng.stats()["em_classes"]
    eng.set_option("exact_classes", 1)
    pi1, nr1, pr1, _ = eng.finish(0, 20, 0.01)
    c1 = eng.stats()["em_classes"]
    eng.set_option("exact_classes", 0)
    assert c0 == c1 and c0 < 0.5 * 3_000_000
    assert np.array_equal(pi0, pi1) and np.array_equal(nr0, nr1) and np.array_equal(pr0, pr1)


def test_batching_is_invisible(big, sqb):
    eng = big["eng"]
    _push_all(eng, big["chunks"][:1])
    off1, tid1, score1 = eng.candidates()
    pi1, nr1, *_ = eng.finish(0, 20, 0.01)
    # same reads again through the host path in small sub-batches
    w, b, l, nb = big["chunks"][0]
    W, B, L = (sqb.synth.to_u32(x) for x in (w, b, l))
    e2 = sqb.Engine([31], big["T"], sketch_fraction=SKETCH)
    e2.set_option("batch_bases", 1 << 24)
    e2.load_index(0, *big["post"])
    e2.push_reads(W, B, L)
    off2, tid2, score2 = e2.candidates()
    pi2, nr2, *_ = e2.finish(0, 20, 0.01)
    st2 = e2.stats()
    e2.close()
    assert st2["batches"] > 5
    assert np.array_equal(off1, off2) and np.array_equal(tid1, tid2) and np.array_equal(score1, score2)
    assert np.array_equal(pi1, pi2) and np.array_equal(nr1, nr2)  # deterministic reductions: bitwise equal
