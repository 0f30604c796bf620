#!/usr/bin/env python
"""Quant hot-path benchmark (sketch + seed lookup + vote + EM/assign).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs all|none]

One JSON line on stdout (rank 0).  The headline (`value`, `e2e`, `roofline`, `cpu_baseline`, `parity`) is
BASELINE.json config 2; the same line carries a `configs` object with config 4 (long reads), the config 5 sweep
(-k 21,25,31 at FracMinHash scales 0.01/0.05/0.1/0.2) and config 1 (the reference program as shipped and at -O2
next to the drop-in command line).  A step is one full quant pass over the workload:
  value  = reads/s with the packed reads already resident in HBM (sq_push_reads_device + sq_finish)
  e2e    = reads/s through the C ABI from pinned HOST buffers (H2D inside the timed region), results copied back
  parity = the first reads of the workload pushed through the same engine and compared with the CPU leg
           (the reference's own code, oracle/_ref; the C port for the large-scale sweep points)
The reference arm (--impl reference) times the reference's own CPU code on a bounded sample of config 2 per step.
Nothing here reads /root/reference at run time.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SCALES = [0.01, 0.05, 0.1, 0.2]  # config 5 sweep: thresholds 42949671 / 214748367 / 429496735 / 858993471


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries write to the C-level stdout behind Python's back (NCCL prints "NCCL version ..." there).  The
# contract is ONE JSON line on stdout, so fd 1 is pointed at stderr for the whole run and the JSON line goes to
# the saved original descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def f32(x):
    return float(np.float32(x))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genes", type=int, default=62500, help="genes of the synthetic transcriptome (~4 isoforms each)")
    ap.add_argument("--fragments", type=int, default=10_000_000, help="short-read fragments per GPU; both mates are emitted")
    ap.add_argument("--long-reads", type=int, default=1_000_000, help="config 4: long reads per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads of the bounded CPU sample (0: per workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (no cpu_baseline, no parity)")
    ap.add_argument("--chunk", type=int, default=1 << 21, help="reads per pushed batch")
    ap.add_argument("--workload", default="short", choices=["short", "long", "multik"],
                    help="headline workload: short = config 2; long = config 4; multik = config 5")
    ap.add_argument("--sketch", type=float, default=0.05, help="FracMinHash scale factor of the headline workload")
    ap.add_argument("--configs", default="all", choices=["all", "none", "long", "multik", "config1"],
                    help="extra BASELINE configs reported in the `configs` object (headline short only)")
    ap.add_argument("--extra-steps", type=int, default=5, help="timed steps of the extra configs (<= --steps)")
    ap.add_argument("--permute-ids", action="store_true",
                    help="load the index with scrambled transcript ids (a reference-written index is in unordered_map order)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- workload
WORKLOAD_TEXT = {
    "short": "config 2: synthetic human-scale transcriptome, 10M simulated 2x150 bp fragments per GPU (both mates as "
             "forward-strand records), k=31, scale %g, chain 0.9, EM 20 iterations",
    "long": "config 4: same transcriptome, ONT-like forward-strand reads 1-10 kb (log-uniform, clipped to the "
            "transcript), 5%% substitutions, k=31, scale %g",
    "multik": "config 5: same transcriptome and short reads, index -k 21,25,31, scale %g",
}


def make_transcriptome(args, device):
    from _sqpkg import sqb
    t0 = time.time()
    tx = sqb.synth.make_transcriptome(args.genes, seed=7, device=device)
    T = tx["t_off"].numel() - 1
    log("[bench] transcriptome: T=%d, %.1f Mbp (%.1fs)" % (T, float(tx["t_off"][-1]) / 1e6, time.time() - t0))
    return tx, T


def make_reads(args, tx, kind, n_reads, device, rank, want_host=True, chunk=None):
    """kind short: 150 bp, 0.5 % substitutions; long: 1-10 kb, 5 %.  -> list of chunk dicts (device + pinned host)"""
    import torch
    from _sqpkg import sqb
    syn = sqb.synth
    chunks = []
    t0 = time.time()
    sim = dict(read_len=150, err=0.005) if kind == "short" else dict(long_reads=(1000, 10000), err=0.05)
    chunk = chunk or (args.chunk if kind == "short" else min(args.chunk, 1 << 16))
    for ch in syn.simulate_reads(tx, n_reads, seed=1000 + rank, chunk=chunk, **sim):
        words, boff, ln = syn.pack_ragged(ch["codes"], ch["r_off"], align=4)
        c = {"words": words, "boff": boff, "len": ln, "n": ln.numel(), "bases": int(ch["r_off"][-1])}
        if want_host:
            c["h_words"] = torch.empty(words.shape, dtype=words.dtype, pin_memory=True).copy_(words)
            c["h_len"] = torch.empty(ln.shape, dtype=ln.dtype, pin_memory=True).copy_(ln)
        chunks.append(c)
        del ch
    if device != "cpu":
        torch.cuda.synchronize()
    log("[bench] %s reads: %d records in %d chunks (%.1fs)" % (kind, n_reads, len(chunks), time.time() - t0))
    return chunks


def build_index_gpu(engine, tx, ks, permute=None):
    """index postings from the transcript sequences with the engine's own sketch kernel (sq_build_postings);
    permute: optional array new_id[old_id] applied to the transcript ids before loading"""
    import torch
    from _sqpkg import sqb
    tlen = tx["t_off"][1:] - tx["t_off"][:-1]
    keep = torch.nonzero(tlen >= max(ks)).flatten()  # main.cpp:67-75: too-short transcripts get no sketch
    words, boff, ln = sqb.synth.pack_ragged(tx["codes"], tx["t_off"], align=4)
    w = sqb.synth.to_u32(words)
    b, l = sqb.synth.to_u32(boff[keep].contiguous()), sqb.synth.to_u32(ln[keep].contiguous())
    tid = keep.cpu().numpy().astype(np.uint32)
    if permute is not None:
        tid = permute[tid].astype(np.uint32)
    out = {}
    for ki, k in enumerate(ks):
        out[k] = engine.build_postings(ki, w, b, l, tid)
        engine.load_index(ki, *out[k])
    return out


def sample_sequences(chunk, n_sample):
    """first n reads of a chunk as ASCII byte strings"""
    from _sqpkg import sqb
    n = min(n_sample, chunk["n"])
    W = sqb.synth.to_u32(chunk["words"])
    boff = chunk["boff"][:n].cpu().numpy().astype(np.int64)
    ln = chunk["len"][:n].cpu().numpy().astype(np.int64)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    if n and int(ln.min()) == int(ln.max()):  # equal lengths: one gather
        idx = boff[:, None] + np.arange(int(ln[0]), dtype=np.int64)[None, :]
        asc = lut[(W[idx >> 4] >> ((idx & 15) * 2).astype(np.uint32)) & 3]
        seqs = [asc[i].tobytes() for i in range(n)]
    else:
        for i in range(n):
            idx = boff[i] + np.arange(ln[i], dtype=np.int64)
            seqs.append(lut[(W[idx >> 4] >> ((idx & 15) * 2).astype(np.uint32)) & 3].tobytes())
    return seqs


def write_fastq(seqs, path, prefix="s"):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@" + prefix.encode() + b"%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")


# ----------------------------------------------------------------------------- CPU leg (baseline + parity reference)
def cpu_leg(kind, ks, scale, names, postings, seqs, repeats=1):
    """Runs the CPU implementation of the path on `seqs` against the full index.
    kind "reference": the reference's own code behind oracle/_ref/libref_oracle.so (falls back to "port" when that
    binary is absent); "port": the C restatement oracle/quant_oracle.c.
    -> dict(kind, reads, times[], stages, off, tid, score, pi, numreads, present)"""
    import oracle_py
    T = len(names)
    times, stages = [], None
    if kind == "reference" and oracle_py.have_ref():
        tmp = tempfile.mkdtemp(prefix="sqbench")
        fq = os.path.join(tmp, "sample.fq")
        write_fastq(seqs, fq)
        r = oracle_py.RefOracle(ks)
        t0 = time.time()
        r.set_transcripts(names)
        for k in ks:
            r.set_postings(k, *postings[k])
        log("[bench] reference index structures built in %.1fs" % (time.time() - t0))
        for _ in range(repeats):
            t0 = time.time()
            n = r.fastq(fq, scale)
            r.chain(0.9)
            r.em(20, 0.01)
            r.assign()
            times.append(time.time() - t0)
            stages = r.times()
        off, tid, score = r.candidates_csr("s", len(seqs))
        pi = r.pi()
        nr, present = r.counts()
        r.close()
        shutil.rmtree(tmp, ignore_errors=True)
    else:
        kind = "port"
        p = oracle_py.PortOracle()
        thr = p.threshold(scale)
        for _ in range(repeats):
            t0 = time.time()
            _, off, tid, score, n = p.chain_batch(ks, thr, 0.9, postings, seqs)
            t1 = time.time()
            pi, _ = p.em(off, tid, score, n, T)
            nr, present = p.assign(off, tid, score, T, pi)
            times.append(time.time() - t0)
            stages = {"sketch+chain": t1 - t0, "em+assign": time.time() - t1}
    return {"kind": kind, "reads": int(n), "times": times, "stages": {k: round(float(v), 3) for k, v in stages.items()},
            "off": np.asarray(off, dtype=np.int64), "tid": tid, "score": score, "pi": pi, "numreads": nr,
            "present": np.asarray(present, dtype=np.uint8)}


def engine_on_sample(eng, chunk, n):
    """the first n reads of a device chunk through the engine: candidates, pi, NumReads, presence"""
    eng.reset_reads()
    nb = int(chunk["boff"][n - 1]) + int(chunk["len"][n - 1])
    eng.push_reads_device(chunk["words"].data_ptr(), chunk["words"].numel(), chunk["boff"].data_ptr(),
                          chunk["len"].data_ptr(), n, nb + 4)
    off, tid, score = eng.candidates()
    pi, nr, present, it = eng.finish(0, 20, 0.01)
    return off.astype(np.int64), tid, score, pi, nr, present, it


def parity_block(ours, cpu):
    """candidate lists as sets of (transcript, score) per read (order among equal scores is unspecified upstream,
    sparse_chaining.cpp:108-109); pi / NumReads relative differences (the reference sums in hash-table order)"""
    off, tid, score, pi, nr, present, it = ours
    same_off = np.array_equal(off, cpu["off"])
    cand_equal = bool(same_off)
    if same_off and off[-1]:
        # canonical order inside a read: (score desc, tid asc) on both sides
        rid = np.repeat(np.arange(len(off) - 1), np.diff(off))
        o1 = np.lexsort((tid, -score.astype(np.int64), rid))
        o2 = np.lexsort((cpu["tid"], -cpu["score"].astype(np.int64), rid))
        cand_equal = bool(np.array_equal(tid[o1], cpu["tid"][o2]) and np.array_equal(score[o1], cpu["score"][o2]))
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0
    m = cpu["present"] > 0
    return {"reads": len(off) - 1, "pairs": int(off[-1]), "against": cpu["kind"], "candidates_equal": cand_equal,
            "pi_max_rel": rel(pi, cpu["pi"]), "numreads_max_rel": rel(nr[m], cpu["numreads"][m]),
            "present_equal": bool(np.array_equal(present > 0, m)), "tolerance": 1e-6,
            "ok": bool(cand_equal and rel(pi, cpu["pi"]) <= 1e-6 and rel(nr[m], cpu["numreads"][m]) <= 1e-6
                       and np.array_equal(present > 0, m))}


def pin_to_gpu_node(local):
    """Bind this process to the CPUs NVML reports as local to its GPU, so that the pinned host buffers (first
    touch) and the copy engine's reads stay on that socket.  Returns a short description for the log."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to gpu %d" % (len(cpus), local)
    except Exception as ex:  # no NVML / not permitted: run unpinned
        return "unpinned (%s)" % type(ex).__name__
    return "unpinned"


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "fallback 6.65 TB/s (of fallback)"


# ----------------------------------------------------------------------------- one workload on the engine
class Ctx:
    pass


def run_workload(cx, kind, ks, scale, chunks, steps, warmup, cpu_kind, cpu_sample, full=False, permute=None):
    """-> dict with value / e2e / roofline / cpu_baseline / parity of one (workload, k list, scale)"""
    import torch
    from _sqpkg import sqb
    args, dist, dev, world, rank, local, T, tx = cx.args, cx.dist, cx.dev, cx.world, cx.rank, cx.local, cx.T, cx.tx
    scale = f32(scale)
    nk = len(ks)
    eng = sqb.Engine(ks, T, sketch_fraction=scale, chain_fraction=0.9, device=local)
    eng.set_option("batch_bases", 1 << 29)
    if os.environ.get("SQ_BENCH_PEER_EXCHANGE"):  # A/B of the EM exchange at N > 1 (0: ncclAllReduce per iteration)
        eng.set_option("peer_exchange", int(os.environ["SQ_BENCH_PEER_EXCHANGE"]))
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    t0 = time.time()
    postings = build_index_gpu(eng, tx, ks, permute)
    log("[bench] %s k=%s s=%g index: %s keys, %s postings (%.1fs)" % (
        kind, ks, scale, [int(postings[k][0].shape[0]) for k in ks], [int(postings[k][2].shape[0]) for k in ks], time.time() - t0))
    if world > 1:
        uid = torch.from_numpy(eng.nccl_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, uid.cpu().numpy())
    n_reads = sum(c["n"] for c in chunks)
    n_bases = sum(c["bases"] for c in chunks)
    pi = torch.empty(T, dtype=torch.float64, pin_memory=True)
    nr = torch.empty(T, dtype=torch.float64, pin_memory=True)
    pres = torch.empty(T, dtype=torch.uint8, pin_memory=True)

    # R of the M-step (isoform_assignment.cpp:54-57) = reads of the whole job: the caller knows it (every rank pushes
    # n_reads), so the engine need not all-reduce its local counts
    R_all = n_reads * world

    def step_device():
        eng.reset_reads()
        for c in chunks:
            eng.push_reads_device(c["words"].data_ptr(), c["words"].numel(), c["boff"].data_ptr(),
                                  c["len"].data_ptr(), c["n"], c["bases"] + 4 * c["n"])
        return eng.finish_into(pi.data_ptr(), nr.data_ptr(), pres.data_ptr(), R_all, 20, 0.01)

    # equal-length reads (the short-read workloads): sq_push_reads_fixed, only the packed words travel;
    # otherwise the lengths travel too (base_off = NULL, offsets derived on the GPU)
    fixed_len = 150 if kind != "long" else 0

    def step_host():
        eng.reset_reads()
        for c in chunks:
            if fixed_len:
                eng.push_reads_fixed_ptr(c["h_words"].data_ptr(), c["h_words"].numel(), fixed_len, c["n"])
            else:
                eng.push_reads_ptr(c["h_words"].data_ptr(), c["h_words"].numel(), 0, c["h_len"].data_ptr(), c["n"])
        return eng.finish_into(pi.data_ptr(), nr.data_ptr(), pres.data_ptr(), R_all, 20, 0.01)

    # the same pinned buffers copied to the device and nothing else: the host->device ceiling of `e2e` on this box
    # (with N ranks: all ranks copy at the same time, as they do in the e2e leg)
    d_sink = torch.empty(max(c["h_words"].numel() for c in chunks), dtype=chunks[0]["h_words"].dtype, device=dev)

    def step_copy_only():
        for c in chunks:
            d_sink[:c["h_words"].numel()].copy_(c["h_words"], non_blocking=True)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        eng.set_profiling(profile)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = eng.stats()["launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage = {}
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
            if profile:
                st = eng.stats()
                for k2, v in st.items():
                    if k2.startswith("ms_") or k2.endswith("_launches"):
                        stage[k2] = stage.get(k2, 0) + v
                stage["last"] = st
        e1.record(stream)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - w0) * 1e3
        ms = e0.elapsed_time(e1)
        launches = eng.stats()["launches"] - l0
        eng.set_profiling(False)
        if dist:
            t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = t.tolist()
        return ms, wall, launches, stage

    clocks = ClockSampler(local)
    clocks.start()
    # `value`: K steps with nothing but the C-ABI calls inside the timed region
    ms, wall, launches, _ = timed(step_device, steps, warmup)
    clk = clocks.stop()
    # per-stage / per-kernel durations: the same K steps again with the engine's CUDA-event profiling on (event
    # pairs around every stage, resolved by sq_get_stats after each step -- host work that would not belong in
    # `value`); `roofline` is computed from these
    ms_prof, _, _, stage = timed(step_device, steps, 1, profile=True)
    ms_e2e, wall_e2e, _, _ = timed(step_host, steps, max(min(warmup, 3), 1))
    ms_copy, _, _, _ = timed(step_copy_only, steps, 1)
    log("[bench] rank %d %s k=%s s=%g: device-resident %.2f ms/step, profiled %.2f, host-buffer e2e %.2f" % (
        rank, kind, ks, scale, ms / steps, ms_prof / steps, ms_e2e / steps))
    total_reads = n_reads * world
    value = total_reads * steps / (ms / 1e3)
    e2e_value = total_reads * steps / (max(ms_e2e, 1e-9) / 1e3)
    st = stage.get("last", eng.stats())
    iters = st["em_iterations"]

    # ---- roofline (algorithmic bytes defined in DESIGN.md section 4)
    peak, peak_src = hbm_peak()
    n_kmers = sum(n_bases - n_reads * (k - 1) for k in ks)
    b_sketch = n_bases / 4 + 12 * n_reads + 4 * st["sketch_hashes"] + 4 * n_reads * nk
    # vote kernels (whatever their internal encoding): per read item_start/cnt/base_off in and soff/cnt out, 4 B per
    # sketch hash, 12 B per probe (hash + table slot: key, list), 4 B per posting of a hit list (4 B x deg;
    # the bit-mask kernels get them as 64-bit masks, the information is the same), 8 B per candidate
    # The lookup kernel: per probe the hash in (4 B), one table slot (8 B: key + descriptor) and the descriptor out
    # (4 B).  The first vote kernel (bit-sliced thread-per-read for short reads, warp-per-read window kernel for long
    # ones): per read item_start/cnt/hoff in and soff/cnt out, 4 B per descriptor, 4 B per posting of a hit list
    # (4 B x deg; nine in ten arrive inside the descriptor, the information is the same), 8 B per candidate,
    # 24 B per read for the class sort key and list fingerprint the kernel folds while it emits the list.
    vote_name = "vote_bits_kernel" if kind != "long" else "vote_long_kernel"
    b_lookup = 16 * st["queries"]
    b_vote = (12 + 6 * nk) * n_reads + 4 * st["queries"] + 4 * st["postings"] + 8 * st["pairs"] + (8 + 24) * n_reads
    # EM: per iteration the class CSR and its transcript-major copy are read once each (12 B + 8 B gathered per
    # class pair, twice) plus the per-class and per-transcript vectors
    # (short reads: scores fit 8 bits, a pair is ONE word in both copies: 4 B + 8 B gathered, twice)
    b_em = iters * ((24 if kind != "long" else 40) * st["em_class_pairs"] + 20 * st["em_classes"] + 16 * T)
    S = steps
    n_batches = max(int(st["batches"]), 1)
    kern = {
        "sketch_kernel": {"ms": stage.get("ms_sketch", 0) / S, "bytes": b_sketch, "launches": stage.get("sketch_launches", 0) // S},
        "lookup_kernel": {"ms": stage.get("ms_lookup", 0) / S, "bytes": b_lookup, "launches": n_batches * nk},
        vote_name: {"ms": stage.get("ms_vote_main", 0) / S, "bytes": b_vote, "launches": stage.get("vote_launches", 0) // S},
        "em_iterations": {"ms": stage.get("ms_em", 0) / S, "bytes": b_em, "launches": iters},
    }
    for kname, kv in kern.items():
        kv["gbs"] = kv["bytes"] / (kv["ms"] / 1e3) / 1e9 if kv["ms"] > 0 else None
        kv["frac"] = kv["gbs"] / peak if kv["gbs"] else None
    dom = max(kern, key=lambda n: kern[n]["ms"])
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and kind == "short" and nk == 1:
        tj = json.load(open(tpath))
        if dom in tj and kern[dom]["launches"]:
            traffic = tj[dom] * (n_reads / kern[dom]["launches"]) / tj["reads_per_launch"]
            traffic_src = "from profile: %s, scaled to this launch size (not measured in this run)" % tj.get("source", "profiles/traffic.json")
    pipes = None
    if os.path.exists(tpath) and kind == "short" and nk == 1:
        pipes = json.load(open(tpath)).get("pipes", {}).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": kern[dom]["ms"] / max(kern[dom]["launches"], 1),
                "algorithmic_bytes_per_launch": kern[dom]["bytes"] / max(kern[dom]["launches"], 1),
                "kernels": {n: {"ms_per_step": round(v["ms"], 4), "GBps": v["gbs"] and round(v["gbs"], 1),
                                "frac": v["frac"] and round(v["frac"], 4)} for n, v in kern.items()},
                "traffic_source": traffic_src,
                # what actually bounds a kernel that is not memory-bound (from the committed ncu capture, not this run)
                "pipes_from_profile": pipes,
                "stage_ms_per_step": {k2: round(v / S, 4) for k2, v in stage.items() if k2.startswith("ms_")},
                "profiled_ms_per_step": ms_prof / steps,
                "sketch_gkmers_per_s": n_kmers / (kern["sketch_kernel"]["ms"] / 1e3) / 1e9
                if kern["sketch_kernel"]["ms"] > 0 else None}
    out = {
        "value": value, "unit": "reads/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
        "gkmers_per_s": n_kmers * world * steps / (ms / 1e3) / 1e9,
        "config": {"workload": WORKLOAD_TEXT[kind] % scale, "transcripts": T,
                   "transcriptome_mbp": round(float(tx["t_off"][-1]) / 1e6, 1), "reads_per_gpu": n_reads,
                   "read_len": 150 if kind != "long" else "1000-10000", "bases_per_gpu": n_bases, "k": ks,
                   "sketch_scale": scale, "threshold": int(eng.threshold),
                   "index_keys": [int(postings[k][0].shape[0]) for k in ks],
                   "index_postings": [int(postings[k][2].shape[0]) for k in ks], "candidate_pairs": int(st["pairs"]),
                   "transcript_ids": "scrambled (--permute-ids)" if permute is not None else "generator order",
                   "l2_policy": "inputs (%.0f MB packed reads per pass) larger than the 126 MB L2; the index tables "
                                "are meant to stay L2-resident" % (n_bases / 4 / 1e6),
                   "parallelism": "reads sharded across %d GPU(s), index replicated; EM sums per iteration: %s; NumReads/presence: one NCCL all-reduce each" % (
                       world, "none (one GPU)" if world == 1 else ("exchanged over peer memory inside the M-step kernel" if st.get("peer_exchange") else "ncclAllReduce"))},
        "e2e": {"value": e2e_value, "unit": "reads/s",
                "h2d_bytes_per_step": int(sum(c["h_words"].numel() * 4 + (0 if fixed_len else 4 * c["n"]) for c in chunks)),
                "d2h_bytes_per_step": int(T * 17), "ms_per_step": ms_e2e / steps, "wall_ms_per_step": wall_e2e / steps,
                "h2d_copy_only_ms_per_step": ms_copy / steps,
                "input": "2-bit packed reads of one length in pinned host memory (sq_push_reads_fixed: lengths and offsets written on the GPU)"
                if fixed_len else "2-bit packed reads + lengths in pinned host memory (sq_push_reads, offsets derived on the GPU)"},
        "gpu_launches": int(launches), "wall_ms_per_step": wall / steps, "clocks": clk, "roofline": roofline,
        "em_iterations": iters,
        "work": {k2: int(st[k2]) for k2 in ("reads", "sketch_hashes", "queries", "hits", "postings", "pairs", "mid_reads",
                                            "slow_reads", "overflow_reads", "batches", "em_classes", "em_class_pairs")},
    }

    # ---- CPU leg on the first reads of the workload: baseline (N = 1) and parity reference (every N), rank 0
    if rank == 0 and not args.no_cpu_baseline and cpu_sample > 0:
        ns = min(cpu_sample, chunks[0]["n"])
        seqs = sample_sequences(chunks[0], ns)
        names = sqb.synth.transcript_names(T)
        cpu = cpu_leg(cpu_kind, ks, scale, names, postings, seqs)
        if world == 1:
            out["cpu_baseline"] = {"value": cpu["reads"] / cpu["times"][0], "unit": "reads/s", "cores": 1, "kind": cpu["kind"],
                                   "sample": "first %d reads of the workload against the full index (T=%d), single thread "
                                             "(the reference has no threads); stages s: %s" % (cpu["reads"], T, cpu["stages"])}
        pe = eng
        if world > 1:  # the benchmarked engine is bound to the communicator: an engine of its own for the sample
            pe = sqb.Engine(ks, T, sketch_fraction=scale, chain_fraction=0.9, device=local)
            for ki, k in enumerate(ks):
                pe.load_index(ki, *postings[k])
        out["parity"] = parity_block(engine_on_sample(pe, chunks[0], ns), cpu)
        if pe is not eng:
            pe.close()
        log("[bench] parity %s k=%s s=%g: %s" % (kind, ks, scale, out["parity"]))
    if dist:
        dist.barrier()
    eng.close()
    return out


# ----------------------------------------------------------------------------- config 1 (programs, files)
def run_program(cmd, markers):
    """runs a quant/index command line; wall seconds and the time each stdout marker line appeared"""
    t0 = time.time()
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    seen = {}
    for line in p.stdout:
        for m in markers:
            if line.startswith(m) and m not in seen:
                seen[m] = time.time() - t0
    p.wait()
    return p.returncode, time.time() - t0, seen


def read_csv(path):
    rows = open(path).read().splitlines()[1:]
    return {r.split(",")[0]: (float(r.split(",")[1]), float(r.split(",")[2])) for r in rows}


def config1(cx):
    """BASELINE config 1: a small transcriptome + FASTQ on disk, `index` then `quant`, with the reference program
    built at -O2 and with its shipped flags (build.sh:23: -g, no -O), next to the drop-in command line."""
    from _sqpkg import sqb
    import oracle_py
    syn = sqb.synth
    tmp = tempfile.mkdtemp(prefix="sqcfg1")
    try:
        tx = syn.make_transcriptome(1000, seed=7)
        T = tx["t_off"].numel() - 1
        names = syn.transcript_names(T, tx["gene"])
        fa, fq = os.path.join(tmp, "tx.fa"), os.path.join(tmp, "reads.fq")
        syn.write_fasta(fa, names, tx, width=70)
        n = 0
        for ch in syn.simulate_reads(tx, 100_000, 150, seed=8, err=0.005):
            n += syn.write_fastq(fq, ch, first_id=n, mode="ab")
        markers = ["Loading index completed", "Loading read completed", "Sparse chaining completed",
                   "EM estimation completed", "Read assignment completed", "Output written to"]
        progs = {"reference_O2": oracle_py.REF_BIN, "reference_asshipped": oracle_py.REF_BIN + "_asshipped",
                 "ours_cli": os.path.join(ROOT, "build", "test")}
        res = {"transcripts": T, "reads": n, "k": [31], "note": "files on local disk; wall clock of the whole program "
               "(`quant` includes reading the index and the FASTQ); reference: single thread"}
        csvs = {}
        for name, exe in progs.items():
            if not os.path.exists(exe):
                res[name] = {"unavailable": os.path.relpath(exe, ROOT) + " not built"}
                continue
            idx, csv = os.path.join(tmp, name + ".idx"), os.path.join(tmp, name + ".csv")
            rc1, t_index, _ = run_program([exe, "-k", "31", "-o", "index", fa, idx], [])
            rc2, t_quant, seen = run_program([exe, "-o", "quant", idx, fq, csv], markers)
            if rc1 or rc2:
                res[name] = {"error": "exit codes %d / %d" % (rc1, rc2)}
                continue
            csvs[name] = read_csv(csv)
            hot = seen.get("Read assignment completed", t_quant) - seen.get("Loading index completed", 0.0)
            res[name] = {"index_s": round(t_index, 3), "quant_s": round(t_quant, 3),
                         "quant_reads_per_s": round(n / t_quant, 1),
                         "hot_path_s": round(hot, 3), "hot_path_reads_per_s": round(n / max(hot, 1e-9), 1),
                         "marker_s": {k: round(v, 3) for k, v in seen.items()}}
        if "ours_cli" in csvs and "reference_O2" in csvs:
            a, b = csvs["ours_cli"], csvs["reference_O2"]
            worst = max([abs(a[k][j] - b[k][j]) / max(abs(b[k][j]), 1e-300) for k in a if k in b for j in (0, 1)] or [0.0])
            res["csv_rows_equal"] = set(a) == set(b)
            res["csv_max_rel"] = worst  # both programs print 6 significant digits
            res["ok"] = bool(set(a) == set(b) and worst <= 2e-5)
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ----------------------------------------------------------------------------- arms
def summary(r):
    """the per-config record kept in `configs`"""
    k = r["roofline"]["kernels"]
    return {"ms_per_step": r["ms_per_step"], "reads_per_s": r["value"], "gkmers_per_s": r["gkmers_per_s"],
            "e2e_ms_per_step": r["e2e"]["ms_per_step"], "e2e_reads_per_s": r["e2e"]["value"],
            "steps": r["steps"], "kernels": k,
            "hashing_frac": k["sketch_kernel"]["frac"],
            "lookup_frac": k["lookup_kernel"]["frac"],
            "lookup_gprobes_per_s": (r["work"]["queries"] / (k["lookup_kernel"]["ms_per_step"] / 1e3) / 1e9) if k["lookup_kernel"]["ms_per_step"] else None,
            "vote_frac": next((v["frac"] for n, v in k.items() if n.startswith("vote")), None),
            "sketch_gkmers_per_s": r["roofline"]["sketch_gkmers_per_s"],
            "stage_ms_per_step": r["roofline"]["stage_ms_per_step"], "work": r["work"],
            "config": {k2: r["config"][k2] for k2 in ("workload", "reads_per_gpu", "k", "sketch_scale", "threshold", "index_keys")},
            "cpu_baseline": r.get("cpu_baseline"), "parity": r.get("parity")}


def run_ours(args):
    import torch
    from _sqpkg import sqb
    cx = Ctx()
    cx.args = args
    cx.rank = rank = int(os.environ.get("RANK", "0"))
    cx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    cx.dev = dev = "cuda:%d" % local
    numa = pin_to_gpu_node(local)  # before any pinned allocation: host buffers land on the GPU's NUMA node
    cx.dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
        cx.dist = dist
    sqb.load_library()  # fails loudly when the CUDA extension is missing
    log("[bench] rank %d: host affinity %s" % (rank, numa))
    cx.tx, cx.T = make_transcriptome(args, dev)
    permute = None
    if args.permute_ids:
        permute = np.random.default_rng(5).permutation(cx.T).astype(np.uint32)

    steps_x = max(1, min(args.steps, args.extra_steps))
    warm_x = max(3, min(args.warmup, 3))
    head_ks = {"short": [31], "long": [31], "multik": [21, 25, 31]}[args.workload]
    n_head = 2 * args.fragments if args.workload != "long" else args.long_reads
    short_chunks = long_chunks = None
    if args.workload != "long":
        short_chunks = make_reads(args, cx.tx, "short", 2 * args.fragments, dev, rank)
    else:
        long_chunks = make_reads(args, cx.tx, "long", args.long_reads, dev, rank)
    samp = args.cpu_sample or {"short": 1_000_000 if world == 1 else 200_000, "long": 20_000, "multik": 100_000}[args.workload]
    head = run_workload(cx, args.workload, head_ks, args.sketch, short_chunks or long_chunks, args.steps, args.warmup,
                        "reference", samp, full=True, permute=permute)
    out = {"metric": "quant reads/sec (sketch+seed+chain+assign)", "value": head["value"], "unit": "reads/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 hash / f64 EM",
           "data": "synthetic"}
    for k in ("config", "gkmers_per_s", "e2e", "gpu_launches", "wall_ms_per_step", "clocks", "roofline", "em_iterations",
              "work", "cpu_baseline", "parity"):
        if k in head:
            out[k] = head[k]

    # ---- the other BASELINE configs, same transcriptome (reported, not the headline)
    configs = {}
    want = args.configs if args.workload == "short" else "none"

    def guarded(name, fn):
        t0 = time.time()
        try:
            configs[name] = fn()
        except Exception as ex:  # an extra config must not take the headline down
            import traceback
            traceback.print_exc()
            configs[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        log("[bench] %s done in %.1fs" % (name, time.time() - t0))

    if want in ("all", "multik"):
        sweep = {}
        for s in SCALES:
            def one(s=s):
                # s = 0.05 against the reference's own code; the other scales against the C port (the reference's
                # string-keyed maps for 3 x 25 M keys do not fit a bounded CPU leg)
                kind = "reference" if abs(s - 0.05) < 1e-9 else "port"
                return summary(run_workload(cx, "multik", [21, 25, 31], s, short_chunks, steps_x, warm_x, kind,
                                            100_000 if kind == "reference" else 50_000))
            guarded("s=%g" % s, one)
            sweep["s=%g" % s] = configs.pop("s=%g" % s)
        configs["config5_multik_sweep"] = {
            "k": [21, 25, 31], "scales": SCALES, "points": sweep,
            "table": [{"scale": s, "k": k, "threshold": int(np.uint32(np.float64(4294967295) * np.float64(np.float32(s)))),
                       "ms_per_step": sweep["s=%g" % s].get("ms_per_step"),
                       "reads_per_s": sweep["s=%g" % s].get("reads_per_s"),
                       "hashing_frac": sweep["s=%g" % s].get("hashing_frac"),
                       "lookup_frac": sweep["s=%g" % s].get("lookup_frac"),
                       "index_keys": (sweep["s=%g" % s].get("config", {}).get("index_keys") or [None] * 3)[i]}
                      for s in SCALES for i, k in enumerate([21, 25, 31])],
            "note": "one fused sketch pass hashes all three k per read, the lookup kernel then runs once per k-index; "
                    "hashing/lookup fractions are measured per scale (all k together), the per-k rows repeat them next to "
                    "that k's index size"}
    if want in ("all", "long"):
        short_chunks = None  # free the short reads before the long ones are made
        import gc
        gc.collect()
        torch.cuda.empty_cache()

        def one_long():
            chunks = make_reads(args, cx.tx, "long", args.long_reads, dev, rank)
            return summary(run_workload(cx, "long", [31], 0.05, chunks, steps_x, warm_x, "reference", 20_000))
        guarded("config4_long", one_long)
    if want in ("all", "config1") and rank == 0 and world == 1 and not args.no_cpu_baseline:
        guarded("config1_programs", lambda: config1(cx))
    if configs:
        out["configs"] = configs
    if rank == 0:
        emit(out)
    if cx.dist:
        cx.dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation on a bounded sample of config 2 per step (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from _sqpkg import sqb
    import oracle_py
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    ks, scale = [31], f32(args.sketch)
    n_sample = args.cpu_sample or 200_000
    tx, T = make_transcriptome(args, dev)
    chunks = make_reads(args, tx, "short", n_sample, dev, 0, want_host=False)
    port = oracle_py.PortOracle()
    t0 = time.time()
    blob = sqb.synth.codes_to_ascii(tx["codes"])
    soff = tx["t_off"].cpu().numpy().astype(np.uint64)
    postings = {k: port.build_postings(blob, soff, k, max(ks), port.threshold(scale)) for k in ks}
    log("[bench] CPU index build: %s keys (%.1fs)" % ([int(postings[k][0].shape[0]) for k in ks], time.time() - t0))
    seqs = sample_sequences(chunks[0], n_sample)
    names = sqb.synth.transcript_names(T)
    cpu = cpu_leg("reference", ks, scale, names, postings, seqs, repeats=args.steps + args.warmup)
    timed = cpu["times"][args.warmup:]
    sec = sum(timed)
    n = cpu["reads"]
    value = n * len(timed) / sec
    cb = {"value": value, "unit": "reads/s", "cores": 1, "kind": cpu["kind"],
          "sample": "%d reads per step against the full index (T=%d), single thread (the reference has no threads)" % (n, T)}
    out = {"impl": "reference", "metric": "quant reads/sec (sketch+seed+chain+assign)", "value": value, "unit": "reads/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / len(timed) * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 hash / f64 EM",
           "data": "synthetic",
           "config": {"workload": "config 2: synthetic human-scale transcriptome, k=31; bounded sample of %d reads per step" % n,
                      "transcripts": T, "read_len": 150, "k": ks},
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "stages_s": cpu["stages"]}
    emit(out)


if __name__ == "__main__":
    a = parse()
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        raise
