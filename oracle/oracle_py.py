"""TEST INFRASTRUCTURE ONLY (oracle).  ctypes bindings of the two CPU checkers:

  PortOracle  oracle/_build/libquant_oracle.so   plain-C restatement (quant_oracle.c); travels everywhere
  RefOracle   oracle/_ref/libref_oracle.so       the reference's own translation units behind ref_harness.cpp;
                                                 built only where /root/reference exists, travels as a binary

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libquant_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ref_test")


def build(ref=True):
    """compile the C restatement; and the reference harness when /root/reference is present"""
    targets = ["port"]
    if ref and os.path.isdir("/root/reference/src"):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def have_ref():
    return os.path.exists(REF_SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _Postings(C.Structure):
    _fields_ = [("nkeys", C.c_uint64), ("keys", C.c_void_p), ("off", C.c_void_p), ("tids", C.c_void_p)]


class PortOracle:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = L = C.CDLL(PORT_SO)
        L.orc_fwd_hash64.restype = C.c_uint64
        L.orc_fwd_hash64.argtypes = [C.c_char_p, C.c_uint32]
        L.orc_hash32_windows.restype = C.c_uint64
        L.orc_hash32_windows.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_threshold.restype = C.c_uint32
        L.orc_threshold.argtypes = [C.c_double]
        L.orc_sketch.restype = C.c_uint64
        L.orc_sketch.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.orc_is_valid_sequence.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_em.restype = C.c_int
        L.orc_em.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int,
                             C.c_double, C.c_void_p]
        L.orc_assign.restype = None
        L.orc_assign.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.orc_chain_batch.restype = C.c_uint64
        L.orc_chain_batch.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_double, C.c_void_p, C.c_uint64,
                                      C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_uint64]

    def build_postings(self, blob, soff, k, min_len, threshold):
        """inverted map of one k over many sequences, in C (blob: ASCII bytes, soff: uint64 offsets)"""
        L = self.lib
        L.orc_build_postings.restype = None
        L.orc_build_postings.argtypes = [C.c_uint64, C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        soff = np.ascontiguousarray(soff, dtype=np.uint64)
        nk, npost = C.c_uint64(), C.c_uint64()
        L.orc_build_postings(soff.shape[0] - 1, blob, _p(soff), k, min_len, threshold, C.byref(nk), C.byref(npost),
                             None, None, None)
        keys = np.zeros(nk.value, dtype=np.uint32)
        off = np.zeros(nk.value + 1, dtype=np.uint64)
        tids = np.zeros(npost.value, dtype=np.uint32)
        L.orc_build_postings(soff.shape[0] - 1, blob, _p(soff), k, min_len, threshold, C.byref(nk), C.byref(npost),
                             _p(keys), _p(off), _p(tids))
        return keys, off, tids

    def fwd_hash64(self, s):
        return int(self.lib.orc_fwd_hash64(s, len(s)))

    def threshold(self, fraction):
        return int(self.lib.orc_threshold(fraction))

    def hash32_windows(self, s, k):
        n = max(len(s) - k + 1, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        m = self.lib.orc_hash32_windows(s, len(s), k, _p(out), None)
        return out[:m]

    def selected(self, s, k, threshold):
        """multiset (in window order) of hashes <= threshold"""
        h = self.hash32_windows(s, k)
        return h[h <= np.uint32(threshold)]

    def sketch(self, s, k, threshold):
        n = max(len(s) - k + 1, 1)
        out = np.zeros(n, dtype=np.uint32)
        m = self.lib.orc_sketch(s, len(s), k, threshold, _p(out), n)
        return out[:m]

    def postings_from_sequences(self, seqs, ks, threshold):
        """inverted map per k from transcript sequences (dense ids), like build_kmer_to_transcript_map;
        sequences shorter than max(ks) get no sketch (main.cpp:66-75)"""
        out = {}
        for k in ks:
            pairs = []
            for t, s in enumerate(seqs):
                if len(s) < max(ks):
                    continue
                sk = self.sketch(s, k, threshold)
                pairs.append(np.stack([sk.astype(np.uint64), np.full(sk.shape, t, dtype=np.uint64)], 1))
            if pairs:
                pr = np.concatenate(pairs, 0)
            else:
                pr = np.zeros((0, 2), dtype=np.uint64)
            order = np.lexsort((pr[:, 1], pr[:, 0]))
            pr = pr[order]
            keys, start = np.unique(pr[:, 0], return_index=True)
            off = np.concatenate([start, [pr.shape[0]]]).astype(np.uint64)
            out[k] = (keys.astype(np.uint32), off, pr[:, 1].astype(np.uint32))
        return out

    def chain_batch(self, ks, threshold, fraction, postings, seqs):
        """-> (admitted bool[R_in], cand_off u64[R_in+1], tid, score)"""
        nk = len(ks)
        karr = np.asarray(ks, dtype=np.uint32)
        arr = (_Postings * nk)()
        keep = []
        for i, k in enumerate(ks):
            if k in postings and postings[k] is not None:
                keys, off, tids = postings[k]
                keys = np.ascontiguousarray(keys, dtype=np.uint32)
                off = np.ascontiguousarray(off, dtype=np.uint64)
                tids = np.ascontiguousarray(tids, dtype=np.uint32)
                keep += [keys, off, tids]
                arr[i] = _Postings(keys.shape[0], _p(keys).value, _p(off).value, _p(tids).value)
            else:
                arr[i] = _Postings(0, None, None, None)
        R = len(seqs)
        blob = b"".join(seqs)
        roff = np.zeros(R + 1, dtype=np.uint64)
        roff[1:] = np.cumsum([len(s) for s in seqs])
        admitted = np.zeros(max(R, 1), dtype=np.uint8)
        radm = C.c_uint64()
        cand_off = np.zeros(R + 1, dtype=np.uint64)
        cap = max(1024, 64 * R)
        while True:
            tid = np.zeros(cap, dtype=np.uint32)
            score = np.zeros(cap, dtype=np.int32)
            tot = self.lib.orc_chain_batch(nk, _p(karr), threshold, fraction, C.cast(arr, C.c_void_p), R, blob,
                                           _p(roff), _p(admitted), C.byref(radm), _p(cand_off), _p(tid), _p(score),
                                           cap)
            if tot <= cap:
                break
            cap = int(tot)
        return admitted[:R].astype(bool), cand_off, tid[:tot], score[:tot], int(radm.value)

    def em(self, cand_off, tid, score, R, T, iters=20, tol=0.01):
        cand_off = np.ascontiguousarray(cand_off, dtype=np.uint64)
        tid = np.ascontiguousarray(tid, dtype=np.uint32)
        score = np.ascontiguousarray(score, dtype=np.int32)
        pi = np.zeros(T, dtype=np.float64)
        it = self.lib.orc_em(cand_off.shape[0] - 1, _p(cand_off), _p(tid), _p(score), R, T, iters, tol, _p(pi))
        return pi, it

    def assign(self, cand_off, tid, score, T, pi):
        cand_off = np.ascontiguousarray(cand_off, dtype=np.uint64)
        tid = np.ascontiguousarray(tid, dtype=np.uint32)
        score = np.ascontiguousarray(score, dtype=np.int32)
        pi = np.ascontiguousarray(pi, dtype=np.float64)
        nr = np.zeros(T, dtype=np.float64)
        present = np.zeros(T, dtype=np.uint8)
        self.lib.orc_assign(cand_off.shape[0] - 1, _p(cand_off), _p(tid), _p(score), T, _p(pi), _p(nr), _p(present))
        return nr, present


class RefOracle:
    """The reference's own code (unmodified translation units) behind oracle/ref_harness.cpp."""

    def __init__(self, ks):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (built only where /root/reference is present: make -C oracle ref)")
        self.lib = L = C.CDLL(REF_SO)
        L.refq_create.restype = C.c_void_p
        L.refq_create.argtypes = [C.c_int, C.c_void_p]
        L.refq_destroy.argtypes = [C.c_void_p]
        L.refq_set_transcripts.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.refq_set_postings.argtypes = [C.c_void_p, C.c_uint, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refq_load_index_file.argtypes = [C.c_void_p, C.c_char_p]
        L.refq_get_ks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.refq_num_transcripts.restype = C.c_uint64
        L.refq_num_transcripts.argtypes = [C.c_void_p]
        L.refq_transcript_name.restype = C.c_char_p
        L.refq_transcript_name.argtypes = [C.c_void_p, C.c_uint64]
        L.refq_postings_size.restype = C.c_uint64
        L.refq_postings_size.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
        L.refq_get_postings.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refq_sketch.restype = C.c_uint64
        L.refq_sketch.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_double, C.c_void_p, C.c_uint64]
        L.refq_fastq.restype = C.c_uint64
        L.refq_fastq.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.refq_add_read.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint64, C.c_double]
        L.refq_num_reads.restype = C.c_uint64
        L.refq_num_reads.argtypes = [C.c_void_p]
        L.refq_chain.argtypes = [C.c_void_p, C.c_double]
        L.refq_em.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.refq_assign.argtypes = [C.c_void_p]
        L.refq_times.argtypes = [C.c_void_p, C.c_void_p]
        L.refq_read_sketch.restype = C.c_int64
        L.refq_read_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, C.c_void_p, C.c_uint64]
        L.refq_read_candidates.restype = C.c_int64
        L.refq_read_candidates.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.refq_get_pi.argtypes = [C.c_void_p, C.c_void_p]
        L.refq_candidates_csr.restype = C.c_uint64
        L.refq_candidates_csr.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_uint64]
        L.refq_get_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refq_write_csv.argtypes = [C.c_void_p, C.c_char_p]
        karr = np.asarray(ks, dtype=np.uint32)
        self.h = C.c_void_p(L.refq_create(len(ks), _p(karr)))
        self.ks = list(ks)
        self.T = 0

    def close(self):
        if self.h:
            self.lib.refq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def sketch_of(seq, k, fraction):
        L = C.CDLL(REF_SO)
        L.refq_sketch.restype = C.c_uint64
        L.refq_sketch.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_double, C.c_void_p, C.c_uint64]
        n = max(len(seq) - k + 1, 1)
        out = np.zeros(n, dtype=np.uint32)
        m = L.refq_sketch(seq, len(seq), k, fraction, _p(out), n)
        return out[:m]

    def set_transcripts(self, names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self.lib.refq_set_transcripts(self.h, len(names), arr)
        self.T = len(names)

    def set_postings(self, k, keys, off, tids):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        tids = np.ascontiguousarray(tids, dtype=np.uint32)
        self.lib.refq_set_postings(self.h, k, keys.shape[0], _p(keys), _p(off), _p(tids))

    def load_index_file(self, path):
        nk = self.lib.refq_load_index_file(self.h, path.encode())
        out = np.zeros(max(nk, 1), dtype=np.uint32)
        self.lib.refq_get_ks(self.h, _p(out), nk)
        self.ks = [int(x) for x in out[:nk]]
        self.T = int(self.lib.refq_num_transcripts(self.h))
        return self.ks

    def transcript_names(self):
        return [self.lib.refq_transcript_name(self.h, i).decode() for i in range(self.T)]

    def get_postings(self, k):
        npost = C.c_uint64()
        nkeys = self.lib.refq_postings_size(self.h, k, C.byref(npost))
        keys = np.zeros(nkeys, dtype=np.uint32)
        off = np.zeros(nkeys + 1, dtype=np.uint64)
        tids = np.zeros(npost.value, dtype=np.uint32)
        self.lib.refq_get_postings(self.h, k, _p(keys), _p(off), _p(tids))
        return keys, off, tids

    def fastq(self, path, sketch_size):
        return int(self.lib.refq_fastq(self.h, path.encode(), sketch_size))

    def add_read(self, rid, seq, sketch_size):
        return int(self.lib.refq_add_read(self.h, rid, seq, len(seq), sketch_size))

    def num_reads(self):
        return int(self.lib.refq_num_reads(self.h))

    def chain(self, fraction=0.9):
        self.lib.refq_chain(self.h, fraction)

    def em(self, iters=20, tol=0.01):
        self.lib.refq_em(self.h, iters, tol)

    def assign(self):
        self.lib.refq_assign(self.h)

    def times(self):
        out = np.zeros(4, dtype=np.float64)
        self.lib.refq_times(self.h, _p(out))
        return dict(sketch=out[0], chain=out[1], em=out[2], assign=out[3])

    def read_sketch(self, rid, k, cap=1 << 16):
        out = np.zeros(cap, dtype=np.uint32)
        n = self.lib.refq_read_sketch(self.h, rid, k, _p(out), cap)
        return None if n < 0 else out[:n]

    def read_candidates(self, rid, cap=1 << 14):
        tid = np.zeros(cap, dtype=np.uint32)
        sc = np.zeros(cap, dtype=np.int32)
        n = self.lib.refq_read_candidates(self.h, rid, _p(tid), _p(sc), cap)
        return None if n < 0 else (tid[:n], sc[:n])

    def candidates_csr(self, prefix, n):
        """sparse_chain() output of the reads named prefix+str(i), i < n, as CSR ordered (score desc, index asc)"""
        off = np.zeros(n + 1, dtype=np.uint64)
        cap = max(1024, 8 * n)
        while True:
            tid = np.zeros(cap, dtype=np.uint32)
            score = np.zeros(cap, dtype=np.int32)
            tot = int(self.lib.refq_candidates_csr(self.h, prefix.encode(), n, _p(off), _p(tid), _p(score), cap))
            if tot <= cap:
                return off, tid[:tot], score[:tot]
            cap = tot

    def pi(self):
        out = np.zeros(self.T, dtype=np.float64)
        self.lib.refq_get_pi(self.h, _p(out))
        return out

    def counts(self):
        out = np.zeros(self.T, dtype=np.float64)
        pres = np.zeros(self.T, dtype=np.uint8)
        self.lib.refq_get_counts(self.h, _p(out), _p(pres))
        return out, pres

    def write_csv(self, path):
        self.lib.refq_write_csv(self.h, path.encode())
