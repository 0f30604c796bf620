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
    # The table finds the distinct lists.  The sort path starts a class wherever a read's list differs from its
    # predecessor's in (best candidate, 14-bit list hash) order, so two classes that share that key and interleave
    # are cut into more pieces (harmless: the pieces add up to the same terms), and it keeps one empty class for
    # the reads without candidates.  The two paths order the classes differently, so the sums agree up to
    # re-association.
    assert c0 <= c1 and c1 - c0 < 0.01 * c1 and c0 < 0.5 * 3_000_000
    np.testing.assert_allclose(pi0, pi1, rtol=1e-12)
    np.testing.assert_allclose(nr0, nr1, rtol=1e-12, atol=1e-12)
    assert np.array_equal(pr0, pr1)


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


def test_scrambled_ids_at_full_scale(big, sqb, port):
    """The same index loaded under a random relabelling of the transcripts (a reference-written index is in
    unordered_map order): sampled reads re-derived by the oracle match bit for bit, the quantification is the same
    up to the relabelling, and the share of reads the bit-sliced kernel hands on does not grow."""
    T = big["T"]
    perm = np.random.default_rng(11).permutation(T).astype(np.uint32)   # new id of transcript i
    keys, off, tids = big["post"]
    tid2 = perm[tids]
    # lists must stay ascending for the oracle's merge; the engine sorts them itself, the oracle gets sorted ones
    tid2s = tid2.copy()
    starts = off[:-1].astype(np.int64)
    order = np.lexsort((tid2s, np.repeat(np.arange(len(keys)), np.diff(off.astype(np.int64)))))
    tid2s = tid2s[order]
    eng = big["eng"]
    _push_all(eng, big["chunks"][:1])
    off1, tid1, score1 = eng.candidates()
    pi1, nr1, pr1, _ = eng.finish(0, 20, 0.01)
    st1 = eng.stats()
    e2 = sqb.Engine([31], T, sketch_fraction=SKETCH)
    e2.load_index(0, keys, off, tid2)          # unsorted lists, scrambled ids
    w, b, l, nb = big["chunks"][0]
    e2.push_reads_device(w.data_ptr(), w.numel(), b.data_ptr(), l.data_ptr(), l.numel(), nb + 4 * l.numel())
    off2, tid2c, score2 = e2.candidates()
    pi2, nr2, pr2, _ = e2.finish(0, 20, 0.01)
    st2 = e2.stats()
    e2.close()
    assert np.array_equal(off1, off2) and np.array_equal(score1, score2)
    np.testing.assert_allclose(pi2[perm], pi1, rtol=1e-9)
    np.testing.assert_allclose(nr2[perm], nr1, rtol=1e-9, atol=1e-12)
    assert np.array_equal(pr2[perm], pr1)
    assert st2["mid_reads"] <= 1.1 * st1["mid_reads"] + 1000 and st2["slow_reads"] <= 1.1 * st1["slow_reads"] + 1000
    # sampled reads against the oracle on the scrambled index
    rng = np.random.default_rng(2)
    W = sqb.synth.to_u32(w)
    bb, ll = b.cpu().numpy(), l.cpu().numpy()
    sample = np.sort(rng.choice(ll.shape[0], 800, replace=False))
    seqs = [sqb.packing.unpack_read(W, int(bb[i]), int(ll[i])) for i in sample]
    _, ooff, otid, oscore, _ = port.chain_batch([31], port.threshold(SKETCH), 0.9, {31: (keys, off, tid2s)}, seqs)
    for j, i in enumerate(sample):
        a, c = int(off2[i]), int(off2[i + 1])
        assert tid2c[a:c].tolist() == otid[int(ooff[j]):int(ooff[j + 1])].tolist(), i
        assert score2[a:c].tolist() == oscore[int(ooff[j]):int(ooff[j + 1])].tolist(), i


def test_long_reads_sampled_parity_at_scale(big, sqb, port):
    """config-4 shape: 100 k ONT-like reads (1-10 kb, 5 % substitutions) against the human-scale index; sampled
    reads re-derived by the oracle, sum(NumReads) = reads with a candidate"""
    syn = sqb.synth
    e = sqb.Engine([31], big["T"], sketch_fraction=SKETCH)
    e.load_index(0, *big["post"])
    first = None
    for ch in syn.simulate_reads(big["tx"], 100_000, seed=77, err=0.05, long_reads=(1000, 10000), chunk=1 << 15):
        w, b, l = syn.pack_ragged(ch["codes"], ch["r_off"], align=4)
        if first is None:
            first = (w, b, l)
        e.push_reads_device(w.data_ptr(), w.numel(), b.data_ptr(), l.data_ptr(), l.numel(), int(ch["r_off"][-1]) + 4 * l.numel())
        torch.cuda.synchronize()
    off, tid, score = e.candidates()
    pi, nr, present, it = e.finish(0, 20, 0.01)
    st = e.stats()
    e.close()
    assert len(off) - 1 == 100_000 and st["slow_reads"] < 0.05 * 100_000
    with_c = int((np.diff(off.astype(np.int64)) > 0).sum())
    assert nr.sum() == pytest.approx(with_c, rel=1e-9)
    w, b, l = first
    W = syn.to_u32(w)
    bb, ll = b.cpu().numpy(), l.cpu().numpy()
    sample = np.sort(np.random.default_rng(3).choice(ll.shape[0], 150, replace=False))
    seqs = [sqb.packing.unpack_read(W, int(bb[i]), int(ll[i])) for i in sample]
    _, ooff, otid, oscore, _ = port.chain_batch([31], port.threshold(SKETCH), 0.9, {31: big["post"]}, seqs)
    for j, i in enumerate(sample):
        a, c = int(off[i]), int(off[i + 1])
        assert tid[a:c].tolist() == otid[int(ooff[j]):int(ooff[j + 1])].tolist(), i
        assert score[a:c].tolist() == oscore[int(ooff[j]):int(ooff[j + 1])].tolist(), i


def test_exact_classes_at_20M_reads(gpu_lib, sqb):
    """config-2 size: the classes found through the fingerprint table are the classes found by comparing the
    candidate lists element by element, and give the same pi / NumReads bit for bit"""
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
    del words, boff, ln
    for ch in syn.simulate_reads(tx, 20_000_000, 150, seed=1000, err=0.005, chunk=1 << 21):
        w, b, l = syn.pack_ragged(ch["codes"], ch["r_off"], align=4)
        eng.push_reads_device(w.data_ptr(), w.numel(), b.data_ptr(), l.data_ptr(), l.numel(), int(ch["r_off"][-1]) + 4 * l.numel())
        eng.sync()
    pi0, nr0, pr0, _ = eng.finish(0, 20, 0.01)
    c0 = eng.stats()["em_classes"]
    eng.set_option("exact_classes", 1)
    pi1, nr1, pr1, _ = eng.finish(0, 20, 0.01)
    c1 = eng.stats()["em_classes"]
    eng.close()
    # the sort path may cut a class into several pieces (see test_fingerprint_classes_equal_exact_classes)
    assert c0 <= c1 and c1 - c0 < 0.01 * c1 and c0 < 0.2 * 20_000_000
    np.testing.assert_allclose(pi0, pi1, rtol=1e-12)
    np.testing.assert_allclose(nr0, nr1, rtol=1e-12, atol=1e-12)
    assert np.array_equal(pr0, pr1)


def test_vote_tiers_agree_at_full_scale(big, sqb):
    """every read through the bit-sliced kernel (+ its hand-ons), through the warp-per-read window kernel, or through
    the general kernel: the candidate lists are the same, bit for bit"""
    eng = big["eng"]
    got = {}
    for tier in (0, 1, 2):
        eng.set_option("vote_tier", tier)
        _push_all(eng, big["chunks"])
        got[tier] = eng.candidates()
    eng.set_option("vote_tier", 0)
    for tier in (0, 1):
        for a, b in zip(got[tier], got[2]):
            assert np.array_equal(a, b), tier
