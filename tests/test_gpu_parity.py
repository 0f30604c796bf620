"""Parity of the CUDA path (through the C ABI) against the CPU oracle: bit-exact hashes / sketches /
candidate sets / scores, and pi / NumReads within 1e-9 relative (the reference itself is only defined up to
floating-point re-association; north_star asks 1e-6)."""
import numpy as np
import pytest

import oracle_py
from datasets import SKETCH, csr_to_lists, dataset

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _sketch_multisets(sqb, seqs, ks, fraction=SKETCH, threshold=None):
    words, off, ln = sqb.packing.pack_reads(seqs)
    with sqb.Engine(ks, 1, sketch_fraction=fraction, threshold=threshold) as e:
        counts, hashes = e.sketch(words, off, ln)
    out, p = [], 0
    for r in range(len(seqs)):
        row = []
        for i in range(len(ks)):
            c = int(counts[r, i])
            row.append(np.sort(hashes[p:p + c]))
            p += c
        out.append(row)
    assert p == hashes.shape[0]
    return out


@pytest.mark.parametrize("ks", [[31], [21, 25, 31], [5], [33, 67], [16, 32, 48]])
def test_sketch_multiset_bit_exact(gpu_lib, sqb, port, ks):
    d = dataset()
    seqs = [s for s in d["reads"][:200] + d["tseqs"][:30] if len(s) >= max(ks)]
    thr = port.threshold(SKETCH)
    got = _sketch_multisets(sqb, seqs, ks)
    for r, s in enumerate(seqs):
        for i, k in enumerate(ks):
            want = np.sort(port.selected(s, k, thr))
            assert got[r][i].tolist() == want.tolist(), (r, k, len(s))


def test_all_hashes_bit_exact_threshold_max(gpu_lib, sqb, port):
    """threshold 0xFFFFFFFF keeps every k-mer: the whole hash stream must match, including ragged lengths"""
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(list(b"ACGT"), n).tolist()) for n in
            [31, 32, 33, 47, 48, 49, 63, 64, 65, 150, 255, 256, 257, 511, 512, 513, 1000, 5000, 12345]]
    for ks in ([31], [21, 25, 31], [1], [81]):
        ss = [s for s in seqs if len(s) >= max(ks)]
        got = _sketch_multisets(sqb, ss, ks, threshold=0xFFFFFFFF)
        for r, s in enumerate(ss):
            for i, k in enumerate(ks):
                want = np.sort(port.hash32_windows(s, k))
                assert got[r][i].shape[0] == len(s) - k + 1
                assert got[r][i].tolist() == want.tolist(), (len(s), k)


def test_unaligned_read_starts(gpu_lib, sqb, port):
    """reads may start at any base of the packed stream"""
    d = dataset()
    seqs = d["reads"][:64]
    thr = port.threshold(SKETCH)
    for align in (1, 4, 16):
        words, off, ln = sqb.packing.pack_reads(seqs, align=align)
        with sqb.Engine([21, 31], 1) as e:
            counts, hashes = e.sketch(words, off, ln)
        p = 0
        for r, s in enumerate(seqs):
            for i, k in enumerate((21, 31)):
                c = int(counts[r, i])
                assert np.sort(hashes[p:p + c]).tolist() == np.sort(port.selected(s, k, thr)).tolist()
                p += c


def _gpu_quant(sqb, d, ks, postings, fraction=0.9, iters=20, tol=0.01, options=None, sketch=SKETCH, chunks=1):
    e = sqb.Engine(ks, len(d["names"]), sketch_fraction=sketch, chain_fraction=fraction)
    for name, val in (options or {}).items():
        e.set_option(name, val)
    for i, k in enumerate(ks):
        if k in postings:
            e.load_index(i, *postings[k])
    reads = d["reads"]
    step = (len(reads) + chunks - 1) // chunks
    for c in range(0, len(reads), step):
        words, off, ln = sqb.packing.pack_reads(reads[c:c + step])
        e.push_reads(words, off, ln)
    off, tid, score = e.candidates()
    pi, nr, present, it = e.finish(0, iters, tol)
    st = e.stats()
    e.close()
    return off, tid, score, pi, nr, present, it, st


@pytest.mark.parametrize("ks,kw", [([31], {}), ([21, 25, 31], {}), ([31], dict(chunks=3)),
                                   ([25, 31], dict(options={"batch_bases": 4096})),
                                   ([31, 31], {}), ([15, 17, 19, 21, 23], {})])
def test_quant_parity_short_reads(gpu_lib, sqb, port, ks, kw):
    d = dataset()
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], sorted(set(ks)), thr)
    off, tid, score, pi, nr, present, it, st = _gpu_quant(sqb, d, ks, postings, **kw)
    _, ooff, otid, oscore, R = port.chain_batch(ks, thr, 0.9, postings, d["reads"])
    assert st["reads"] == R == len(d["reads"])
    assert st["bases"] == sum(len(s) for s in d["reads"])
    assert st["kmers"] == sum(max(len(s) - k + 1, 0) for s in d["reads"] for k in ks)
    assert csr_to_lists(off, tid, score) == csr_to_lists(ooff, otid, oscore)
    # ordering contract of the tap: score descending, then transcript id
    assert tid.tolist() == otid.tolist() and score.tolist() == oscore.tolist()
    T = len(d["names"])
    opi, oit = port.em(ooff, otid, oscore, R, T)
    onr, opres = port.assign(ooff, otid, oscore, T, opi)
    assert it == oit
    np.testing.assert_allclose(pi, opi, rtol=RTOL, atol=0)
    assert present.tolist() == opres.tolist()
    np.testing.assert_allclose(nr, onr, rtol=RTOL, atol=1e-12)


def test_quant_parity_long_reads(gpu_lib, sqb, port):
    """ONT-like reads (1-10 kb, 5% substitutions): multi-item reads, cross-chunk duplicate removal"""
    d = dataset(n_genes=60, n_reads=150, long_reads=(1000, 10000), err=0.05, exon_median=400, seed=9)
    ks = [21, 31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    off, tid, score, pi, nr, present, it, st = _gpu_quant(sqb, d, ks, postings)
    _, ooff, otid, oscore, R = port.chain_batch(ks, thr, 0.9, postings, d["reads"])
    assert csr_to_lists(off, tid, score) == csr_to_lists(ooff, otid, oscore)
    T = len(d["names"])
    opi, _ = port.em(ooff, otid, oscore, R, T)
    onr, opres = port.assign(ooff, otid, oscore, T, opi)
    np.testing.assert_allclose(pi, opi, rtol=RTOL)
    np.testing.assert_allclose(nr, onr, rtol=RTOL, atol=1e-12)
    assert max(len(s) for s in d["reads"]) > 2000


def test_overflow_path_many_transcripts(gpu_lib, sqb, port):
    """a k-mer shared by hundreds of transcripts overflows the shared-memory vote table: the large-table
    kernel must give the same answer"""
    rng = np.random.default_rng(11)
    core = bytes(rng.choice(list(b"ACGT"), 120).tolist())
    tseqs = [bytes(rng.choice(list(b"ACGT"), 60).tolist()) + core + bytes(rng.choice(list(b"ACGT"), 60).tolist())
             for _ in range(700)]
    names = ["T%d" % i for i in range(len(tseqs))]
    reads = [core, tseqs[3][:150], tseqs[5][40:200], core[:100]] * 5
    d = {"tseqs": tseqs, "names": names, "reads": reads}
    ks = [31]
    thr = port.threshold(float(np.float32(0.2)))
    sk = float(np.float32(0.2))
    postings = port.postings_from_sequences(tseqs, ks, thr)
    off, tid, score, pi, nr, present, it, st = _gpu_quant(sqb, d, ks, postings, sketch=sk)
    _, ooff, otid, oscore, R = port.chain_batch(ks, thr, 0.9, postings, reads)
    assert st["overflow_reads"] > 0
    assert csr_to_lists(off, tid, score) == csr_to_lists(ooff, otid, oscore)
    assert tid.tolist() == otid.tolist()
    opi, _ = port.em(ooff, otid, oscore, R, len(names))
    np.testing.assert_allclose(pi, opi, rtol=RTOL)


def test_offsets_derived_on_device(gpu_lib, sqb, port):
    """sq_push_reads with base_off = NULL: reads packed back to back on 4-base boundaries"""
    d = dataset()
    ks = [21, 31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    reads = d["reads"][:150] + [d["tseqs"][0][:31], d["tseqs"][1][:1234]]  # ragged lengths
    res = []
    for explicit in (True, False):
        with sqb.Engine(ks, len(d["names"])) as e:
            for i, k in enumerate(ks):
                e.load_index(i, *postings[k])
            words, off, ln = sqb.packing.pack_reads(reads, align=4)
            e.push_reads(words, off if explicit else None, ln)
            res.append(e.candidates())
    for a, b in zip(*res):
        assert a.tolist() == b.tolist()
    _, ooff, otid, oscore, _ = port.chain_batch(ks, thr, 0.9, postings, reads)
    assert res[1][1].tolist() == otid.tolist() and res[1][2].tolist() == oscore.tolist()


def test_fixed_length_push(gpu_lib, sqb, port):
    """sq_push_reads_fixed: equal-length reads, only the packed words are copied"""
    d = dataset()
    ks = [21, 31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    for L in (150, 33):  # 33: not a multiple of 4, the stride is padded
        reads = [r[:L] for r in d["reads"] if len(r) >= L][:180]
        assert len(reads) > 50
        with sqb.Engine(ks, len(d["names"])) as e:
            for i, k in enumerate(ks):
                e.load_index(i, *postings[k])
            words, off, ln = sqb.packing.pack_reads(reads, align=4)
            e.push_reads_fixed(words, L, len(reads))
            off_g, tid_g, score_g = e.candidates()
            with pytest.raises(sqb.SketchQuantError):
                e.push_reads_fixed(words[:4], L, len(reads))  # too few words for that many reads
        _, ooff, otid, oscore, _ = port.chain_batch(ks, thr, 0.9, postings, reads)
        assert off_g.tolist() == ooff.tolist()
        assert tid_g.tolist() == otid.tolist() and score_g.tolist() == oscore.tolist()


def test_unsorted_posting_lists(gpu_lib, sqb, port):
    """the reference's index file lists transcripts under a hash in arbitrary order"""
    d = dataset()
    ks = [31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    keys, off, tids = postings[31]
    rng = np.random.default_rng(0)
    shuffled = tids.copy()
    for i in range(len(keys)):
        seg = shuffled[int(off[i]):int(off[i + 1])]
        rng.shuffle(seg)
    off1, tid1, score1, pi1, *_ = _gpu_quant(sqb, d, ks, {31: (keys, off, shuffled)})
    off0, tid0, score0, pi0, *_ = _gpu_quant(sqb, d, ks, postings)
    assert off1.tolist() == off0.tolist() and tid1.tolist() == tid0.tolist() and score1.tolist() == score0.tolist()
    assert pi1.tolist() == pi0.tolist()


def test_edge_cases(gpu_lib, sqb, port):
    thr = port.threshold(SKETCH)
    d = dataset()
    ks = [31]
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    T = len(d["names"])
    # no reads at all: pi = 1/T + pseudocounts with R = 0 -> inf like the reference (0.01f/0 in float)
    with sqb.Engine(ks, T) as e:
        e.load_index(0, *postings[31])
        pi, nr, present, it = e.finish(0, 20, 0.01)
        assert present.sum() == 0 and nr.sum() == 0
    # reads that match nothing still count in R (sparse_chaining.cpp:111)
    junk = [b"A" * 150, b"ACGT" * 40, d["reads"][0]]
    dd = dict(d, reads=junk)
    off, tid, score, pi, nr, present, it, st = _gpu_quant(sqb, dd, ks, postings)
    _, ooff, otid, oscore, R = port.chain_batch(ks, thr, 0.9, postings, junk)
    assert R == 3 and csr_to_lists(off, tid, score) == csr_to_lists(ooff, otid, oscore)
    opi, _ = port.em(ooff, otid, oscore, R, T)
    np.testing.assert_allclose(pi, opi, rtol=RTOL)
    # a k-index without a map contributes nothing (sparse_chaining.cpp:51-53)
    ks2 = [25, 31]
    off, tid, score, *_ = _gpu_quant(sqb, d, ks2, {31: postings[31]})
    _, ooff, otid, oscore, _ = port.chain_batch(ks2, thr, 0.9, {31: postings[31]}, d["reads"])
    assert csr_to_lists(off, tid, score) == csr_to_lists(ooff, otid, oscore)
    # read exactly k long, duplicate reads, read made of one repeated k-mer
    reads = [d["tseqs"][0][:31], d["reads"][1], d["reads"][1], (d["tseqs"][2][:31] * 6)[:150]]
    dd = dict(d, reads=reads)
    off, tid, score, *_ = _gpu_quant(sqb, dd, ks, postings, sketch=1.0)
    p1 = port.postings_from_sequences(d["tseqs"], ks, 0xFFFFFFFF)
    off, tid, score, *_ = _gpu_quant(sqb, dd, ks, p1, sketch=1.0)
    _, ooff, otid, oscore, _ = port.chain_batch(ks, 0xFFFFFFFF, 0.9, p1, reads)
    assert csr_to_lists(off, tid, score) == csr_to_lists(ooff, otid, oscore)


def test_em_from_candidates_and_convergence(gpu_lib, sqb, port):
    """EM/assign kernels on hand-made homologous_segments, including the early-exit branch
    (total_change < tol, isoform_assignment.cpp:62) and heavy transcripts spanning many segments"""
    rng = np.random.default_rng(2)
    T, R = 50, 6000
    ncand = rng.integers(0, 5, R)
    off = np.zeros(R + 1, dtype=np.uint64)
    off[1:] = np.cumsum(ncand)
    tid = np.concatenate([rng.choice(T, n, replace=False) for n in ncand] + [np.zeros(0, int)]).astype(np.uint32)
    hot = rng.random(tid.shape[0]) < 0.5
    score = rng.integers(1, 9, tid.shape[0]).astype(np.int32)
    with sqb.Engine([31], T) as e:
        e.set_option("em_segment", 64)
        e.set_candidates(off, tid, score)
        for iters, tol in ((20, 0.01), (3, 0.01), (20, 1e9), (50, 1.0)):
            pi, nr, present, it = e.finish(R, iters, tol)
            opi, oit = port.em(off, tid, score, R, T, iters, tol)
            onr, opres = port.assign(off, tid, score, T, opi)
            assert it == oit
            np.testing.assert_allclose(pi, opi, rtol=RTOL)
            np.testing.assert_allclose(nr, onr, rtol=RTOL, atol=1e-12)
            assert present.tolist() == opres.tolist()


def test_build_postings_matches_oracle(gpu_lib, sqb, port):
    d = dataset()
    ks = [21, 31]
    thr = port.threshold(SKETCH)
    want = port.postings_from_sequences(d["tseqs"], ks, thr)
    got = sqb.api.build_kmer_to_transcript_map(d["tseqs"], ks)
    for k in ks:
        for a, b in zip(got[k], want[k]):
            assert a.tolist() == b.tolist()


def test_api_mirror(gpu_lib, sqb, port):
    d = dataset()
    thr = port.threshold(SKETCH)
    s = d["tseqs"][0]
    assert sorted(sqb.api.createSketch_FracMinhash_direct(s, 31)) == port.sketch(s, 31, thr).tolist()
    with pytest.raises(ValueError):
        sqb.api.createSketch_FracMinhash_direct("ACGT", 31)


@pytest.mark.skipif(not oracle_py.have_ref(), reason="oracle/_ref not present")
def test_against_reference_code_directly(gpu_lib, sqb, port):
    """same inputs through the reference's own translation units (oracle/_ref/libref_oracle.so)"""
    d = dataset(seed=21, n_reads=300)
    ks = [21, 25, 31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    r = oracle_py.RefOracle(ks)
    r.set_transcripts(d["names"])
    for k in ks:
        r.set_postings(k, *postings[k])
    for i, s in enumerate(d["reads"]):
        r.add_read(b"r%d" % i, s, SKETCH)
    r.chain(0.9)
    r.em(20, 0.01)
    r.assign()
    off, tid, score, pi, nr, present, it, st = _gpu_quant(sqb, d, ks, postings)
    mine = csr_to_lists(off, tid, score)
    for i in range(len(d["reads"])):
        rt, rs = r.read_candidates(b"r%d" % i)
        assert sorted(zip(rt.tolist(), rs.tolist())) == mine[i]
    np.testing.assert_allclose(pi, r.pi(), rtol=RTOL)
    rc, rp = r.counts()
    assert present.tolist() == rp.tolist()
    np.testing.assert_allclose(nr, rc, rtol=RTOL, atol=1e-12)


def test_more_edge_cases(gpu_lib, sqb, port):
    """item boundaries (256/257/512/513 bases), a very long read, reads shorter than some k, eight k values,
    chain fractions 0 / 1 / >1, threshold 0, a single transcript"""
    rng = np.random.default_rng(23)
    d = dataset()
    T = len(d["names"])
    base = d["tseqs"]
    long_t = max(base, key=len)
    reads = [long_t[:n] for n in (255, 256, 257, 511, 512, 513, 1024) if n <= len(long_t)]
    reads += [(long_t * 40)[:100000]]                      # 100 kb of repeats: many items, heavy duplicate removal
    reads += [base[3][:40], base[4][:31], base[5][:20]]    # the last one is shorter than k=21/31: no window at all
    dd = dict(d, reads=reads)
    for ks, frac, sk in (([21, 31], 0.9, SKETCH), ([31], 0.0, SKETCH), ([31], 1.0, SKETCH), ([31], 1.5, SKETCH),
                         ([11, 13, 15, 17, 19, 21, 23, 25], 0.9, SKETCH), ([31], 0.9, 0.0), ([31], 0.9, 1.0)):
        thr = port.threshold(sk)
        kk = sorted(set(ks))
        postings = port.postings_from_sequences(base, kk, thr)
        rr = [r for r in reads if len(r) >= max(ks)]       # admission is the caller's job (main.cpp:136-138)
        dd = dict(d, reads=rr)
        off, tid, score, pi, nr, present, it, st = _gpu_quant(sqb, dd, ks, postings, fraction=frac, sketch=sk)
        _, ooff, otid, oscore, R = port.chain_batch(ks, thr, frac, postings, rr)
        assert off.tolist() == ooff.tolist(), (ks, frac, sk)
        assert tid.tolist() == otid.tolist() and score.tolist() == oscore.tolist(), (ks, frac, sk)
        opi, oit = port.em(ooff, otid, oscore, R, T)
        onr, opres = port.assign(ooff, otid, oscore, T, opi)
        assert it == oit
        np.testing.assert_allclose(pi, opi, rtol=RTOL)
        np.testing.assert_allclose(nr, onr, rtol=RTOL, atol=1e-12)
        assert present.tolist() == opres.tolist()
    # one transcript only
    one = {"tseqs": [base[0]], "names": ["only"], "reads": [base[0][:150], base[0][10:200], base[1][:150]]}
    thr = port.threshold(SKETCH)
    p1 = port.postings_from_sequences(one["tseqs"], [31], thr)
    off, tid, score, pi, nr, present, it, st = _gpu_quant(sqb, one, [31], p1)
    _, ooff, otid, oscore, R = port.chain_batch([31], thr, 0.9, p1, one["reads"])
    assert tid.tolist() == otid.tolist() and score.tolist() == oscore.tolist()
    opi, _ = port.em(ooff, otid, oscore, R, 1)
    np.testing.assert_allclose(pi, opi, rtol=RTOL)


def test_engine_reuse_stream_and_errors(gpu_lib, sqb, port):
    import torch
    d = dataset()
    ks = [31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    T = len(d["names"])
    words, off, ln = sqb.packing.pack_reads(d["reads"])
    with sqb.Engine(ks, T) as e:
        e.load_index(0, *postings[31])
        st = torch.cuda.Stream()
        e.set_stream(st.cuda_stream)
        e.set_profiling(True)
        outs = []
        for _ in range(3):  # same engine, reads forgotten in between: identical answers, bit for bit
            e.reset_reads()
            e.push_reads(words, off, ln)
            outs.append(e.finish(0, 20, 0.01))
        for o in outs[1:]:
            assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])
        s = e.stats()
        assert s["launches"] > 0 and s["ms_vote"] > 0 and s["em_classes"] > 0
        # argument errors come back as codes with a message, never as a crash
        with pytest.raises(sqb.SketchQuantError) as ei:
            e.load_index(5, *postings[31])
        assert ei.value.code == -2
        bad = postings[31][2].copy()
        bad[0] = T + 5
        with pytest.raises(sqb.SketchQuantError):
            e.load_index(0, postings[31][0], postings[31][1], bad)
        with pytest.raises(sqb.SketchQuantError):
            e.set_option("no_such_option", 1)
    with pytest.raises(sqb.SketchQuantError):
        sqb.Engine([], 10)
    with pytest.raises(sqb.SketchQuantError):
        sqb.Engine([31] * 9, 10)
    with pytest.raises(sqb.SketchQuantError):
        sqb.Engine([31], 0)


@pytest.mark.parametrize("ks", [[31], [21, 25, 31]])
def test_scrambled_transcript_ids(gpu_lib, sqb, port, ks):
    """A reference-written index stores its transcripts in unordered_map order (src/data_io.cpp:185-196): the
    engine renumbers them internally, results come back in the caller's numbering whatever it is."""
    d = dataset(n_genes=80, n_reads=2000, seed=31)
    T = len(d["tseqs"])
    perm = np.random.default_rng(4).permutation(T)          # new id of transcript i
    inv = np.argsort(perm)
    d2 = {"tseqs": [d["tseqs"][i] for i in inv], "names": [d["names"][i] for i in inv], "reads": d["reads"]}
    thr = port.threshold(SKETCH)
    p1 = port.postings_from_sequences(d["tseqs"], ks, thr)
    p2 = port.postings_from_sequences(d2["tseqs"], ks, thr)
    off1, tid1, score1, pi1, nr1, pr1, it1, st1 = _gpu_quant(sqb, d, ks, p1)
    off2, tid2, score2, pi2, nr2, pr2, it2, st2 = _gpu_quant(sqb, d2, ks, p2)
    _, ooff, otid, oscore, R = port.chain_batch(ks, thr, 0.9, p2, d2["reads"])
    assert off2.tolist() == ooff.tolist() and tid2.tolist() == otid.tolist() and score2.tolist() == oscore.tolist()
    opi, oit = port.em(ooff, otid, oscore, R, T)
    onr, opres = port.assign(ooff, otid, oscore, T, opi)
    np.testing.assert_allclose(pi2, opi, rtol=RTOL)
    np.testing.assert_allclose(nr2, onr, rtol=RTOL, atol=1e-12)
    assert pr2.tolist() == opres.tolist() and it2 == oit
    # the same quantification under another labelling
    np.testing.assert_allclose(pi2[perm], pi1, rtol=RTOL)
    np.testing.assert_allclose(nr2[perm], nr1, rtol=RTOL, atol=1e-12)
    assert csr_to_lists(off2, perm[tid1].astype(np.uint32), score1) == csr_to_lists(off2, tid2, score2)
    # scrambling must not push reads off the bit-mask kernel
    assert st2["mid_reads"] <= st1["mid_reads"] + 0.02 * len(d["reads"])


def test_repeated_id_in_a_posting_list(gpu_lib, sqb, port):
    """a hand-made index may name a transcript twice under one hash: the reference votes once per posting
    (src/sparse_chaining.cpp:64-69), and so does every vote kernel"""
    d = dataset()
    ks = [31]
    thr = port.threshold(SKETCH)
    keys, off, tids = port.postings_from_sequences(d["tseqs"], ks, thr)[31]
    # double the first transcript of every third list
    new_off, new_tids = [0], []
    for i in range(len(keys)):
        seg = tids[int(off[i]):int(off[i + 1])].tolist()
        if i % 3 == 0:
            seg = [seg[0]] + seg
        new_tids += seg
        new_off.append(len(new_tids))
    post = {31: (keys, np.asarray(new_off, dtype=np.uint64), np.asarray(new_tids, dtype=np.uint32))}
    off_g, tid_g, score_g, pi, nr, *_ = _gpu_quant(sqb, d, ks, post)
    _, ooff, otid, oscore, R = port.chain_batch(ks, thr, 0.9, post, d["reads"])
    assert off_g.tolist() == ooff.tolist() and tid_g.tolist() == otid.tolist() and score_g.tolist() == oscore.tolist()


def test_device_push_alignment_contract(gpu_lib, sqb, port):
    """sq_push_reads_device: packed words on a 16-byte boundary go through (the sketch kernel stages a warp's span
    with one bulk copy, which needs that alignment); any other start is refused, not silently mis-read"""
    import torch
    d = dataset()
    ks = [21, 31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    reads = d["reads"][:300] + [d["tseqs"][2][:2000]]
    words, off, ln = sqb.packing.pack_reads(reads)
    nb = int(sum(len(r) for r in reads))
    with sqb.Engine(ks, len(d["names"])) as e:
        for i, k in enumerate(ks):
            e.load_index(i, *postings[k])
        buf = torch.zeros(words.shape[0] + 8, dtype=torch.int32, device="cuda")
        b = torch.from_numpy(off.astype(np.uint32).view(np.int32)).cuda()
        l = torch.from_numpy(ln.astype(np.uint32).view(np.int32)).cuda()
        for shift in (1, 2, 3):
            w = buf[shift:shift + words.shape[0]]
            assert w.data_ptr() % 16 == 4 * shift
            with pytest.raises(sqb.SketchQuantError):
                e.push_reads_device(w.data_ptr(), w.numel(), b.data_ptr(), l.data_ptr(), l.numel(), nb + 4 * l.numel())
        w = buf[4:4 + words.shape[0]]
        w.copy_(torch.from_numpy(words.view(np.int32)))
        e.push_reads_device(w.data_ptr(), w.numel(), b.data_ptr(), l.data_ptr(), l.numel(), nb + 4 * l.numel())
        off_g, tid_g, score_g = e.candidates()
    _, ooff, otid, oscore, _ = port.chain_batch(ks, thr, 0.9, postings, reads)
    assert off_g.tolist() == ooff.tolist()
    assert tid_g.tolist() == otid.tolist() and score_g.tolist() == oscore.tolist()


def test_store_grows_and_vote_reruns(gpu_lib, sqb, port):
    """The vote writes into the free tail of the candidate store; when a batch has more pairs than the room made for
    it (cand_per_read = 1; a first batch of reads that hit nothing, so "twice the rate so far" is tiny), the store
    grows and the batch's vote is run again: same candidates, same EM result"""
    d = dataset(n_genes=60, n_reads=3000, seed=29)
    ks = [21, 31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    rng = np.random.default_rng(11)
    junk = [bytes(b"ACGT"[c] for c in rng.integers(0, 4, 150)) for _ in range(4200)]
    real = d["reads"]
    assert len(real) >= 1000
    reads = junk + real
    T = len(d["names"])
    with sqb.Engine(ks, T) as e:
        e.set_option("cand_per_read", 1)
        e.set_option("batch_bases", 1 << 18)  # ~1700 reads per batch: junk batches first, then the real ones
        for i, k in enumerate(ks):
            e.load_index(i, *postings[k])
        e.push_reads(*sqb.packing.pack_reads(reads))
        off_g, tid_g, score_g = e.candidates()
        pi, nr, present, it = e.finish(0, 20, 0.01)
        st = e.stats()
    assert st["batches"] >= 4
    _, ooff, otid, oscore, R = port.chain_batch(ks, thr, 0.9, postings, reads)
    assert int(ooff[-1]) > 2 * len(real)  # several pairs per real read: more than one per read was needed
    assert off_g.tolist() == ooff.tolist()
    assert tid_g.tolist() == otid.tolist() and score_g.tolist() == oscore.tolist()
    opi, oit = port.em(ooff, otid, oscore, R, T)
    onr, opres = port.assign(ooff, otid, oscore, T, opi)
    assert it == oit
    np.testing.assert_allclose(pi, opi, rtol=RTOL)
    np.testing.assert_allclose(nr, onr, rtol=RTOL, atol=1e-12)
    assert present.tolist() == opres.tolist()
