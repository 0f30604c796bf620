"""The C restatement (oracle/quant_oracle.c) against the known-answer vectors of SURVEY.md Appendix B:
seeds read out of the reference's own build/test binary, upstream ntHash vectors, and the threshold
arithmetic of src/sketch.cpp:25-26."""
import numpy as np


def test_forward_hash_kats(port):
    kats = [(b"ACATG", 0xc496f4f40beaa773), (b"A" * 31, 0xfffffffeaf928327),
            (b"ACGTACGTACGTACGTACGTACGTACGTACG", 0xa11ab471672ce8d2),
            (b"GATTACAGATTACAGATTACAGATTACAGAT", 0x58bde9cd1b88636f), (b"T" * 21, 0xf2b50b9cbf12562e)]
    for s, want in kats:
        assert port.fwd_hash64(s) == want
        assert int(port.hash32_windows(s, len(s))[0]) == want & 0xFFFFFFFF


def test_rolling_kat_k31(port):
    s = b"CAGATTTTCATATTATGCAGAAAATCTACTTCGCCTGATA"
    want = [0x167048e7, 0x16e8b050, 0xcaf7c298, 0xd7da66b5, 0xa865673b, 0x49d01b35, 0x430f845a, 0x41219dd4,
            0x9b59eeeb, 0xfbddd371]
    assert port.hash32_windows(s, 31).tolist() == want


def test_rolling_equals_direct(port):
    rng = np.random.default_rng(5)
    s = bytes(rng.choice(list(b"ACGT"), 500).tolist())
    for k in (1, 5, 21, 31, 33, 34, 67, 81):
        h = port.hash32_windows(s, k)
        assert len(h) == len(s) - k + 1
        for p in (0, 1, 17, len(s) - k):
            assert int(h[p]) == port.fwd_hash64(s[p:p + k]) & 0xFFFFFFFF


def test_threshold(port):
    # (uint32_t)(UINT32_MAX * (double)0.05f) = 214748367, not 214748364
    assert port.threshold(float(np.float32(0.05))) == 214748367 == 0x0CCCCCCF
    assert port.threshold(float(np.float32(0.01))) == 42949671
    assert port.threshold(float(np.float32(0.1))) == 429496735
    assert port.threshold(float(np.float32(0.2))) == 858993471


def test_invalid_windows_are_skipped(port):
    s = b"ACGTACGTNACGTACGTAC"
    h = port.hash32_windows(s, 4)
    clean = [port.fwd_hash64(s[p:p + 4]) & 0xFFFFFFFF for p in range(len(s) - 3) if b"N" not in s[p:p + 4]]
    assert h.tolist() == clean
    # lower case hashes like upper case (ntHash seed table)
    assert port.fwd_hash64(b"acgtt") == port.fwd_hash64(b"ACGTT")


def test_sketch_is_a_sorted_set(port):
    s = b"ACGT" * 100  # heavy repetition: 4 distinct 31-mers
    sk = port.sketch(s, 31, 0xFFFFFFFF)
    assert len(sk) == 4 and sorted(set(sk.tolist())) == sk.tolist()
    assert not port.lib.orc_is_valid_sequence(b"ACGN", 4) and port.lib.orc_is_valid_sequence(b"ACGT", 4)
    assert not port.lib.orc_is_valid_sequence(b"acgt", 4)


def test_upstream_canonical_known_answer_vector(port):
    """The upstream ntHash/btllib test vector NtHash("ACATGCATGCA", 3, 5): the base (canonical) hash of the first
    three 5-mers.  ntHash2's canonical hash is forward + reverse (mod 2^64) and the reverse-strand hash of a
    k-mer is the forward hash of its reverse complement, so this pins seeds, rotation and orientation of the
    forward hash the reference uses (src/sketch.cpp:33) independently of the survey-derived forward KATs."""
    seq = b"ACATGCATGCA"
    want = [0xf59ecb45f0e22b9c, 0x38cc00f940aebdae, 0x603a48c5a11c794a]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for p, w in enumerate(want):
        kmer = seq[p:p + 5]
        rc = kmer.translate(comp)[::-1]
        assert (port.fwd_hash64(kmer) + port.fwd_hash64(rc)) & 0xFFFFFFFFFFFFFFFF == w, (p, kmer)
