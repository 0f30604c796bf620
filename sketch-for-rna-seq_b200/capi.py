"""ctypes binding of include/sketchquant.h.  Fails loudly when the CUDA library is missing or there is no GPU:
there is no CPU fallback anywhere in the product path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SQ_OK = 0
SQ_ERR_NO_DEVICE = -6
SQ_MAX_K_COUNT = 8

EXPORTS = [
    "sq_version", "sq_device_count", "sq_last_error", "sq_threshold_from_fraction", "sq_create", "sq_destroy",
    "sq_set_stream", "sq_set_profiling", "sq_set_option", "sq_load_index", "sq_push_reads", "sq_push_reads_fixed", "sq_push_reads_device",
    "sq_sync", "sq_reset_reads", "sq_finish", "sq_sketch", "sq_num_pairs", "sq_get_candidates", "sq_build_postings",
    "sq_nccl_unique_id", "sq_comm_init", "sq_get_stats", "sq_set_candidates", "sq_host_alloc", "sq_host_free",
]


class SketchQuantError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sketchquant error %d: %s" % (code, msg))
        self.code = code


class Stats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("bases", C.c_uint64), ("kmers", C.c_uint64), ("sketch_hashes", C.c_uint64),
                ("pairs", C.c_uint64), ("overflow_reads", C.c_uint64), ("batches", C.c_uint64),
                ("em_iterations", C.c_int32), ("peer_exchange", C.c_int32),
                ("ms_sketch", C.c_float), ("ms_vote", C.c_float), ("ms_compact", C.c_float), ("ms_sort", C.c_float),
                ("ms_em", C.c_float), ("ms_assign", C.c_float), ("launches", C.c_uint64),
                ("queries", C.c_uint64), ("hits", C.c_uint64), ("postings", C.c_uint64), ("ms_items", C.c_float),
                ("sketch_launches", C.c_uint32), ("vote_launches", C.c_uint32), ("slow_reads", C.c_uint64),
                ("mid_reads", C.c_uint64), ("ms_vote_main", C.c_float), ("ms_lookup", C.c_float), ("em_classes", C.c_uint64), ("em_class_pairs", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("reserved")}


def lib_path():
    return os.path.join(_HERE, "libsketchquant.so")


def load_library():
    """dlopen libsketchquant.so (built in-tree by build.py); raises if it is not there."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise SketchQuantError(-100, "CUDA extension %s is missing: run `python __graft_entry__.py` "
                                     "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    u32p, u64p, vp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_void_p
    lib.sq_version.restype = C.c_char_p
    lib.sq_last_error.restype = C.c_char_p
    lib.sq_last_error.argtypes = [vp]
    lib.sq_threshold_from_fraction.restype = C.c_uint32
    lib.sq_threshold_from_fraction.argtypes = [C.c_double]
    lib.sq_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint32, vp, C.c_uint32, C.c_double, C.c_uint64]
    lib.sq_destroy.argtypes = [vp]
    lib.sq_destroy.restype = None
    lib.sq_set_stream.argtypes = [vp, vp]
    lib.sq_set_profiling.argtypes = [vp, C.c_int]
    lib.sq_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.sq_load_index.argtypes = [vp, C.c_uint32, C.c_uint64, vp, vp, vp]
    lib.sq_push_reads.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint32]
    lib.sq_push_reads_fixed.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32]
    lib.sq_push_reads_device.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint32, C.c_uint64]
    lib.sq_sync.argtypes = [vp]
    lib.sq_reset_reads.argtypes = [vp]
    lib.sq_finish.argtypes = [vp, C.c_uint64, C.c_int, C.c_double, vp, vp, vp, C.POINTER(C.c_int)]
    lib.sq_sketch.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    lib.sq_num_pairs.argtypes = [vp, u64p, u64p]
    lib.sq_get_candidates.argtypes = [vp, vp, vp, vp]
    lib.sq_build_postings.argtypes = [vp, C.c_uint32, vp, C.c_uint64, vp, vp, vp, C.c_uint32, u64p, u64p, vp, vp, vp]
    lib.sq_set_candidates.argtypes = [vp, C.c_uint64, vp, vp, vp]
    lib.sq_nccl_unique_id.argtypes = [vp]
    lib.sq_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.sq_get_stats.argtypes = [vp, C.POINTER(Stats)]
    if hasattr(lib, "sq_push_fastq"):
        lib.sq_push_fastq.argtypes = [vp, vp, C.c_uint64, u64p, u64p]
        lib.sq_admitted_reads.argtypes = [vp, u64p]
    _LIB = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Engine:
    """One engine per GPU (sq_engine).  Mirrors the call order of quantification()
    (reference src/main.cpp:165-197): load_index -> push reads -> finish."""

    def __init__(self, ks, n_transcripts, sketch_fraction=float(np.float32(0.05)), chain_fraction=0.9, device=0,
                 threshold=None):
        self.lib = load_library()
        self.ks = [int(k) for k in ks]
        self.nk = len(self.ks)
        self.T = int(n_transcripts)
        self.threshold = int(self.lib.sq_threshold_from_fraction(sketch_fraction)) if threshold is None else int(threshold)
        self._h = C.c_void_p()
        karr = _u32(self.ks)
        rc = self.lib.sq_create(C.byref(self._h), device, self.nk, _ptr(karr), self.threshold, chain_fraction, self.T)
        if rc != SQ_OK:
            raise SketchQuantError(rc, (self.lib.sq_last_error(None) or b"").decode())
        self._keep = []

    def _check(self, rc):
        if rc != SQ_OK:
            raise SketchQuantError(rc, (self.lib.sq_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.sq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- configuration
    def set_stream(self, cuda_stream):
        self._check(self.lib.sq_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_profiling(self, on=True):
        self._check(self.lib.sq_set_profiling(self._h, int(on)))

    def set_option(self, name, value):
        self._check(self.lib.sq_set_option(self._h, name.encode(), int(value)))

    # ---- index
    def load_index(self, kidx, keys, post_off, post_tid):
        keys = _u32(keys)
        post_off = np.ascontiguousarray(post_off, dtype=np.uint64)
        post_tid = _u32(post_tid)
        assert post_off.shape[0] == keys.shape[0] + 1
        self._check(self.lib.sq_load_index(self._h, kidx, keys.shape[0], _ptr(keys), _ptr(post_off), _ptr(post_tid)))

    def build_postings(self, kidx, packed, base_off, length, seq_tid):
        packed, base_off, length, seq_tid = _u32(packed), _u32(base_off), _u32(length), _u32(seq_tid)
        nk, npost = C.c_uint64(), C.c_uint64()
        n = base_off.shape[0]
        self._check(self.lib.sq_build_postings(self._h, kidx, _ptr(packed), packed.shape[0], _ptr(base_off),
                                               _ptr(length), _ptr(seq_tid), n, C.byref(nk), C.byref(npost),
                                               None, None, None))
        keys = np.empty(nk.value, dtype=np.uint32)
        off = np.empty(nk.value + 1, dtype=np.uint64)
        tid = np.empty(npost.value, dtype=np.uint32)
        self._check(self.lib.sq_build_postings(self._h, kidx, _ptr(packed), packed.shape[0], _ptr(base_off),
                                               _ptr(length), _ptr(seq_tid), n, C.byref(nk), C.byref(npost),
                                               _ptr(keys), _ptr(off), _ptr(tid)))
        return keys, off, tid

    # ---- reads
    def push_reads(self, packed, base_off, length):
        """base_off may be None: reads packed back to back on 4-base boundaries (offsets derived on the GPU)"""
        packed, length = _u32(packed), _u32(length)
        base_off = _u32(base_off) if base_off is not None else None
        self._check(self.lib.sq_push_reads(self._h, _ptr(packed), packed.shape[0], _ptr(base_off), _ptr(length),
                                           length.shape[0]))

    def push_reads_ptr(self, packed_ptr, n_words, base_off_ptr, len_ptr, n_reads):
        """host pointers (e.g. pinned torch tensors' data_ptr())"""
        self._check(self.lib.sq_push_reads(self._h, C.c_void_p(packed_ptr), n_words,
                                           C.c_void_p(base_off_ptr) if base_off_ptr else None,
                                           C.c_void_p(len_ptr), n_reads))

    def push_reads_fixed(self, packed, read_len, n_reads):
        """n_reads reads of read_len bases each, packed back to back on 4-base boundaries: only the words are copied"""
        packed = _u32(packed)
        self._check(self.lib.sq_push_reads_fixed(self._h, _ptr(packed), packed.shape[0], read_len, n_reads))

    def push_reads_fixed_ptr(self, packed_ptr, n_words, read_len, n_reads):
        self._check(self.lib.sq_push_reads_fixed(self._h, C.c_void_p(packed_ptr), n_words, read_len, n_reads))

    def push_reads_device(self, d_packed, n_words, d_base_off, d_len, n_reads, n_bases=0):
        self._check(self.lib.sq_push_reads_device(self._h, C.c_void_p(d_packed), n_words, C.c_void_p(d_base_off),
                                                  C.c_void_p(d_len), n_reads, n_bases))

    def push_fastq(self, text):
        """raw FASTQ text (bytes / uint8 array) -> admitted reads, parsed and packed on the GPU"""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else \
            np.ascontiguousarray(text, dtype=np.uint8)
        nrec, nadm = C.c_uint64(), C.c_uint64()
        self._check(self.lib.sq_push_fastq(self._h, _ptr(buf), buf.shape[0], C.byref(nrec), C.byref(nadm)))
        return nrec.value, nadm.value

    def sync(self):
        self._check(self.lib.sq_sync(self._h))

    def reset_reads(self):
        self._check(self.lib.sq_reset_reads(self._h))

    # ---- results
    def finish(self, R_total=0, em_iters=20, em_tol=0.01):
        pi = np.empty(self.T, dtype=np.float64)
        nr = np.empty(self.T, dtype=np.float64)
        present = np.empty(self.T, dtype=np.uint8)
        it = C.c_int()
        self._check(self.lib.sq_finish(self._h, R_total, em_iters, em_tol, _ptr(pi), _ptr(nr), _ptr(present),
                                       C.byref(it)))
        return pi, nr, present, it.value

    def finish_into(self, pi_ptr, nr_ptr, present_ptr, R_total=0, em_iters=20, em_tol=0.01):
        it = C.c_int()
        self._check(self.lib.sq_finish(self._h, R_total, em_iters, em_tol, C.c_void_p(pi_ptr), C.c_void_p(nr_ptr),
                                       C.c_void_p(present_ptr), C.byref(it)))
        return it.value

    # ---- taps
    def sketch(self, packed, base_off, length):
        """per (read, k-index) multiset of k-mer hashes <= threshold: (counts[n, nk], hashes concatenated)"""
        packed, base_off, length = _u32(packed), _u32(base_off), _u32(length)
        n = base_off.shape[0]
        counts = np.zeros((n, self.nk), dtype=np.uint32)
        total = C.c_uint64()
        self._check(self.lib.sq_sketch(self._h, _ptr(packed), packed.shape[0], _ptr(base_off), _ptr(length), n,
                                       _ptr(counts), None, 0, C.byref(total)))
        hashes = np.empty(total.value, dtype=np.uint32)
        self._check(self.lib.sq_sketch(self._h, _ptr(packed), packed.shape[0], _ptr(base_off), _ptr(length), n,
                                       _ptr(counts), _ptr(hashes), total.value, C.byref(total)))
        return counts, hashes

    def num_pairs(self):
        r, p = C.c_uint64(), C.c_uint64()
        self._check(self.lib.sq_num_pairs(self._h, C.byref(r), C.byref(p)))
        return r.value, p.value

    def candidates(self):
        R, P = self.num_pairs()
        off = np.zeros(R + 1, dtype=np.uint64)
        tid = np.empty(P, dtype=np.uint32)
        score = np.empty(P, dtype=np.int32)
        self._check(self.lib.sq_get_candidates(self._h, _ptr(off), _ptr(tid), _ptr(score)))
        return off, tid, score

    def set_candidates(self, read_off, tid, score):
        read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
        tid = _u32(tid)
        score = np.ascontiguousarray(score, dtype=np.int32)
        self._check(self.lib.sq_set_candidates(self._h, read_off.shape[0] - 1, _ptr(read_off), _ptr(tid), _ptr(score)))

    def stats(self):
        st = Stats()
        self._check(self.lib.sq_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    # ---- multi-GPU
    @staticmethod
    def nccl_unique_id():
        lib = load_library()
        buf = np.zeros(128, dtype=np.uint8)
        rc = lib.sq_nccl_unique_id(_ptr(buf))
        if rc != SQ_OK:
            raise SketchQuantError(rc, (lib.sq_last_error(None) or b"").decode())
        return buf

    def comm_init(self, nranks, rank, uid):
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        self._check(self.lib.sq_comm_init(self._h, nranks, rank, _ptr(uid)))
