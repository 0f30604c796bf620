#!/usr/bin/env python
"""Quant hot-path benchmark (sketch + seed lookup + vote + EM/assign), BASELINE.json config 2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A step is one full quant pass over the synthetic workload:
  value  = reads/s with the packed reads already resident in HBM (sq_push_reads_device + sq_finish)
  e2e    = reads/s through the C ABI from pinned HOST buffers (sq_push_reads: H2D inside the timed region)
           with the result vectors copied back to the host
The reference arm (--impl reference) times the reference's own CPU code (oracle/_ref, else the C port) on a
bounded sample of the same workload.  Nothing here reads /root/reference at run time.
"""
import argparse
import json
import os
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

SKETCH = float(np.float32(0.05))
K_LIST = [31]
READ_LEN = 150


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries write to the C-level stdout behind Python's back (NCCL prints "NCCL version ..." there).  The
# contract is ONE JSON line on stdout, so fd 1 is pointed at stderr for the whole run and the JSON line goes to
# the saved original descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def apply_workload(args):
    """non-default configs of BASELINE.json (parity-test / sweep cases, not the headline line)"""
    global K_LIST, READ_LEN, SKETCH
    SKETCH = float(np.float32(args.sketch))
    if args.workload == "long":
        READ_LEN = None
        if args.fragments == 10_000_000:
            args.fragments = 500_000  # config 4: 1 M long reads
        args.chunk = min(args.chunk, 1 << 16)
    elif args.workload == "multik":
        K_LIST = [21, 25, 31]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genes", type=int, default=62500, help="genes of the synthetic transcriptome (~4 isoforms each)")
    ap.add_argument("--fragments", type=int, default=10_000_000, help="fragments per GPU; both mates are emitted")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="reads of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=1 << 21, help="reads per pushed batch")
    ap.add_argument("--trace", action="store_true", help="print host wall time of every C-ABI call of one step")
    ap.add_argument("--workload", default="short", choices=["short", "long", "multik"],
                    help="short = config 2 (the headline line); long = config 4 (ONT-like 1-10 kb, 5%% error); "
                         "multik = config 5 (-k 21,25,31)")
    ap.add_argument("--sketch", type=float, default=0.05, help="FracMinHash scale factor (config 5 sweep)")
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
def make_workload(args, device, rank, want_host=True):
    import torch
    from _sqpkg import sqb
    syn = sqb.synth
    t0 = time.time()
    tx = syn.make_transcriptome(args.genes, seed=7, device=device)
    T = tx["t_off"].numel() - 1
    tlen = (tx["t_off"][1:] - tx["t_off"][:-1])
    log("[bench] transcriptome: T=%d, %.1f Mbp (%.1fs)" % (T, float(tx["t_off"][-1]) / 1e6, time.time() - t0))
    # both mates of every fragment are separate forward-strand records
    n_reads = 2 * args.fragments
    chunks = []
    t0 = time.time()
    sim = dict(read_len=READ_LEN, err=0.005) if READ_LEN else dict(long_reads=(1000, 10000), err=0.05)
    for ch in syn.simulate_reads(tx, n_reads, seed=1000 + rank, chunk=args.chunk, **sim):
        words, boff, ln = syn.pack_ragged(ch["codes"], ch["r_off"], align=4)
        c = {"words": words, "boff": boff, "len": ln, "n": ln.numel(), "bases": int(ch["r_off"][-1])}
        if want_host:
            c["h_words"] = torch.empty(words.shape, dtype=words.dtype, pin_memory=True).copy_(words)
            c["h_boff"] = torch.empty(boff.shape, dtype=boff.dtype, pin_memory=True).copy_(boff)
            c["h_len"] = torch.empty(ln.shape, dtype=ln.dtype, pin_memory=True).copy_(ln)
        chunks.append(c)
        del ch
    if device != "cpu":
        torch.cuda.synchronize()
    log("[bench] reads: %d records in %d chunks (%.1fs)" % (n_reads, len(chunks), time.time() - t0))
    return tx, T, tlen, chunks


def build_index_gpu(engine, tx, T):
    """index postings from the transcript sequences with the engine's own sketch kernel (sq_build_postings)"""
    import torch
    from _sqpkg import sqb
    tlen = tx["t_off"][1:] - tx["t_off"][:-1]
    keep = torch.nonzero(tlen >= max(K_LIST)).flatten()  # main.cpp:67-75: too-short transcripts get no sketch
    words, boff, ln = sqb.synth.pack_ragged(tx["codes"], tx["t_off"], align=4)
    w = sqb.synth.to_u32(words)
    b, l = sqb.synth.to_u32(boff[keep].contiguous()), sqb.synth.to_u32(ln[keep].contiguous())
    tid = keep.cpu().numpy().astype(np.uint32)
    out = {}
    for ki, k in enumerate(K_LIST):
        out[k] = engine.build_postings(ki, w, b, l, tid)
        engine.load_index(ki, *out[k])
    return out


# ----------------------------------------------------------------------------- CPU reference sample
def cpu_reference_sample(names, postings, fastq_path, n_sample, repeats=1, want_steps=None):
    """Times the reference's CPU code (oracle/_ref harness; C port when that binary is absent) on the sample.
    Returns (reads/s, kind, per-stage seconds, list of per-repeat seconds)."""
    import oracle_py
    times = []
    if oracle_py.have_ref():
        kind = "reference"
        r = oracle_py.RefOracle(K_LIST)
        t0 = time.time()
        r.set_transcripts(names)
        for k in K_LIST:
            r.set_postings(k, *postings[k])
        log("[bench] reference index structures built in %.1fs" % (time.time() - t0))
        stages = None
        for _ in range(repeats):
            t0 = time.time()
            n = r.fastq(fastq_path, SKETCH)
            r.chain(0.9)
            r.em(20, 0.01)
            r.assign()
            times.append(time.time() - t0)
            stages = r.times()
        r.close()
    else:
        kind = "port"
        p = oracle_py.PortOracle()
        seqs = [ln for i, ln in enumerate(open(fastq_path, "rb").read().split(b"\n")) if i % 4 == 1]
        thr = p.threshold(SKETCH)
        stages = {}
        for _ in range(repeats):
            t0 = time.time()
            _, off, tid, score, R = p.chain_batch(K_LIST, thr, 0.9, postings, seqs)
            t1 = time.time()
            pi, _ = p.em(off, tid, score, R, len(names))
            p.assign(off, tid, score, len(names), pi)
            times.append(time.time() - t0)
            stages = {"sketch+chain": t1 - t0, "em+assign": time.time() - t1}
            n = R
    return n, kind, stages, times


def write_sample_fastq(chunk, n_sample, path):
    from _sqpkg import sqb
    n = min(n_sample, chunk["n"])
    W = sqb.synth.to_u32(chunk["words"])
    boff = chunk["boff"][:n].cpu().numpy().astype(np.int64)
    ln = chunk["len"][:n].cpu().numpy().astype(np.int64)
    # unpack n reads of equal stride quickly
    idx = boff[:, None] + np.arange(int(ln.max()), dtype=np.int64)[None, :]
    codes = (W[np.minimum(idx >> 4, len(W) - 1)] >> ((idx & 15) * 2).astype(np.uint32)) & 3  # ragged tails are cut below
    asc = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    with open(path, "wb") as f:
        for i in range(n):
            s = asc[i, :ln[i]].tobytes()
            f.write(b"@s%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    return n


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


# ----------------------------------------------------------------------------- arms
def run_ours(args):
    import torch
    from _sqpkg import sqb
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    numa = pin_to_gpu_node(local)  # before any pinned allocation: host buffers land on the GPU's NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev))
    sqb.load_library()  # fails loudly when the CUDA extension is missing
    log("[bench] rank %d: host affinity %s" % (rank, numa))

    tx, T, tlen, chunks = make_workload(args, dev, rank)
    eng = sqb.Engine(K_LIST, T, sketch_fraction=SKETCH, chain_fraction=0.9, device=local)
    eng.set_option("batch_bases", 1 << 29)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    t0 = time.time()
    postings = build_index_gpu(eng, tx, T)
    log("[bench] index: %s keys, %s postings (%.1fs)" % ([int(postings[k][0].shape[0]) for k in K_LIST],
                                                         [int(postings[k][2].shape[0]) for k in K_LIST], time.time() - t0))
    if world > 1:
        uid = torch.from_numpy(eng.nccl_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).to(dev)
        dist.broadcast(uid, 0)
        eng.comm_init(world, rank, uid.cpu().numpy())
        log("[bench] rank %d: NCCL communicator ready" % rank)
    n_reads = sum(c["n"] for c in chunks)
    n_bases = sum(c["bases"] for c in chunks)
    pi = torch.empty(T, dtype=torch.float64, pin_memory=True)
    nr = torch.empty(T, dtype=torch.float64, pin_memory=True)
    pres = torch.empty(T, dtype=torch.uint8, pin_memory=True)

    def step_device():
        eng.reset_reads()
        for c in chunks:
            eng.push_reads_device(c["words"].data_ptr(), c["words"].numel(), c["boff"].data_ptr(),
                                  c["len"].data_ptr(), c["n"], c["bases"] + 4 * c["n"])
        return eng.finish_into(pi.data_ptr(), nr.data_ptr(), pres.data_ptr(), 0, 20, 0.01)

    # equal-length reads (the short-read workloads): sq_push_reads_fixed, only the packed words travel;
    # otherwise the lengths travel too (base_off = NULL, offsets derived on the GPU)
    fixed_len = READ_LEN if READ_LEN and all(int(c["h_len"].min()) == READ_LEN == int(c["h_len"].max()) for c in chunks) else 0

    def step_host():
        eng.reset_reads()
        for c in chunks:
            if fixed_len:
                eng.push_reads_fixed_ptr(c["h_words"].data_ptr(), c["h_words"].numel(), fixed_len, c["n"])
            else:
                eng.push_reads_ptr(c["h_words"].data_ptr(), c["h_words"].numel(), 0, c["h_len"].data_ptr(), c["n"])
        return eng.finish_into(pi.data_ptr(), nr.data_ptr(), pres.data_ptr(), 0, 20, 0.01)

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

    if args.trace:
        for name, host in (("device", False), ("host", True)):
            for rep in range(2):
                torch.cuda.synchronize()
                t = [time.perf_counter()]
                eng.reset_reads(); t.append(time.perf_counter())
                for c in chunks:
                    if host:
                        eng.push_reads_ptr(c["h_words"].data_ptr(), c["h_words"].numel(), c["h_boff"].data_ptr(),
                                           c["h_len"].data_ptr(), c["n"])
                    else:
                        eng.push_reads_device(c["words"].data_ptr(), c["words"].numel(), c["boff"].data_ptr(),
                                              c["len"].data_ptr(), c["n"], c["bases"] + 4 * c["n"])
                    t.append(time.perf_counter())
                eng.sync(); t.append(time.perf_counter())
                eng.finish_into(pi.data_ptr(), nr.data_ptr(), pres.data_ptr(), 0, 20, 0.01); t.append(time.perf_counter())
                d = [(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])]
                log("[trace %s #%d] reset %.2f | pushes %s | sync %.2f | finish %.2f | total %.2f ms" %
                    (name, rep, d[0], " ".join("%.2f" % x for x in d[1:-2]), d[-2], d[-1], sum(d)))
    clocks = ClockSampler(local)
    clocks.start()
    # `value`: K steps with nothing but the C-ABI calls inside the timed region
    ms, wall, launches, _ = timed(step_device, args.steps, args.warmup)
    clk = clocks.stop()
    log("[bench] rank %d: device-resident %.2f ms/step" % (rank, ms / args.steps))
    # per-stage / per-kernel durations: the same K steps again with the engine's CUDA-event profiling on (event
    # pairs around every stage, resolved by sq_get_stats after each step -- host work that would not belong in
    # `value`); `roofline` is computed from these
    ms_prof, _, _, stage = timed(step_device, args.steps, 1, profile=True)
    log("[bench] rank %d: device-resident, profiled %.2f ms/step" % (rank, ms_prof / args.steps))
    ms_e2e, wall_e2e, _, _ = timed(step_host, args.steps, max(args.warmup, 1))
    log("[bench] rank %d: host-buffer e2e %.2f ms/step" % (rank, ms_e2e / args.steps))
    total_reads = n_reads * world
    value = total_reads * args.steps / (ms / 1e3)
    e2e_value = total_reads * args.steps / (max(ms_e2e, 0.0) / 1e3)
    st = stage.get("last", eng.stats())
    iters = st["em_iterations"]

    # ---- roofline of the dominant kernel (algorithmic bytes defined in DESIGN.md)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (of fallback)"
    nk = len(K_LIST)
    n_kmers = sum(n_bases - n_reads * (k - 1) for k in K_LIST)
    b_sketch = n_bases / 4 + 12 * n_reads + 4 * st["sketch_hashes"] + 4 * n_reads * nk
    # vote kernels (whatever their internal encoding): per read item_start/cnt/base_off in and soff/cnt out, 4 B per
    # sketch hash, 12 B per probe (hash + table slot: key, offset), 4 B per posting of a hit list (SURVEY 8d:
    # 4 B x deg; the bit-mask kernel gets them as 64-bit masks, the information is the same), 8 B per candidate
    if nk == 1 and args.workload == "short":
        vote_name = "vote_bits_kernel"
    else:
        # long reads span several items: the warp-per-read kernel does the work (timed with the whole vote stage)
        vote_name = "vote_kernel" if args.workload == "long" else ("vote_quad_kernel" if nk <= 4 else "vote_fast_kernel")
    b_vote = (12 + 2 * nk) * n_reads + 4 * st["sketch_hashes"] + (4 + 8) * st["queries"] + 4 * st["postings"] \
        + 8 * st["pairs"] + 8 * n_reads
    b_em = iters * (24 * st["pairs"] + 16 * T)
    S = args.steps
    kern = {
        "sketch_kernel": {"ms": stage.get("ms_sketch", 0) / S, "bytes": b_sketch, "launches": stage.get("sketch_launches", 0) // S},
        vote_name: {"ms": stage.get("ms_vote" if args.workload == "long" else "ms_vote_main", 0) / S, "bytes": b_vote, "launches": stage.get("vote_launches", 0) // S},
        "em_iterations": {"ms": stage.get("ms_em", 0) / S, "bytes": b_em, "launches": iters},
    }
    for kname, kv in kern.items():
        kv["gbs"] = kv["bytes"] / (kv["ms"] / 1e3) / 1e9 if kv["ms"] > 0 else None
        kv["frac"] = kv["gbs"] / peak if kv["gbs"] else None
    dom = max(kern, key=lambda n: kern[n]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.workload == "short":
        tj = json.load(open(tpath))
        if dom in tj and kern[dom]["launches"]:
            traffic = tj[dom] * (n_reads / kern[dom]["launches"]) / tj["reads_per_launch"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": kern[dom]["ms"] / max(kern[dom]["launches"], 1),
                "algorithmic_bytes_per_launch": kern[dom]["bytes"] / max(kern[dom]["launches"], 1),
                "kernels": {n: {"ms_per_step": round(v["ms"], 4), "GBps": v["gbs"] and round(v["gbs"], 1),
                                "frac": v["frac"] and round(v["frac"], 4)} for n, v in kern.items()},
                "traffic_source": "profiles/traffic.json (ncu --set full capture, scaled to this launch size)" if traffic else None,
                "concurrency": "launch durations are taken inside the step, where the previous batch's compaction and "
                               "vote follow-up kernels run on a second stream next to this kernel (alone under ncu: "
                               "0.375 ms per 2.1 M reads for vote_bits_kernel)",
                "stage_ms_per_step": {k2: round(v / S, 4) for k2, v in stage.items() if k2.startswith("ms_")},
                "profiled_ms_per_step": ms_prof / args.steps,
                "sketch_gkmers_per_s": n_kmers / (kern["sketch_kernel"]["ms"] / 1e3) / 1e9
                if kern["sketch_kernel"]["ms"] > 0 else None}

    out = {
        "metric": "quant reads/sec (sketch+seed+chain+assign)", "value": value, "unit": "reads/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 hash / f64 EM",
        "data": "synthetic",
        "config": {"workload": {"short": "config 2: synthetic human-scale transcriptome, 10M simulated 2x150 bp fragments per GPU "
                                         "(both mates as forward-strand records), k=31, scale 0.05, chain 0.9, EM 20 iterations",
                                "long": "config 4: same transcriptome, ONT-like forward-strand reads 1-10 kb (log-uniform, clipped to "
                                        "the transcript), 5% substitutions, k=31",
                                "multik": "config 5: same transcriptome and short reads, index -k 21,25,31, scale %g" % args.sketch}[args.workload],
                   "transcripts": T, "transcriptome_mbp": round(float(tx["t_off"][-1]) / 1e6, 1),
                   "reads_per_gpu": n_reads, "read_len": READ_LEN or "1000-10000", "bases_per_gpu": n_bases, "k": K_LIST,
                   "sketch_scale": args.sketch,
                   "index_keys": [int(postings[k][0].shape[0]) for k in K_LIST],
                   "index_postings": [int(postings[k][2].shape[0]) for k in K_LIST], "candidate_pairs": int(st["pairs"]),
                   "l2_policy": "inputs (%.0f MB packed reads + %.0f MB index table) larger than the 126 MB L2"
                                % (n_bases / 4 / 1e6, sum(32 * (1 << int(np.ceil(np.log2(max(postings[k][0].shape[0] / 2, 2))))) for k in K_LIST) / 1e6),
                   "parallelism": "reads sharded across %d GPU(s), index replicated, NCCL all-reduce of T-vectors" % world},
        "gkmers_per_s": n_kmers * world * args.steps / (ms / 1e3) / 1e9,
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(sum(c["h_words"].numel() * 4 + (0 if fixed_len else 4 * c["n"]) for c in chunks)),
                "d2h_bytes_per_step": int(T * 17), "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_e2e / args.steps,
                "input": "2-bit packed reads of one length in pinned host memory (sq_push_reads_fixed: lengths and offsets written on the GPU)"
                if fixed_len else "2-bit packed reads + lengths in pinned host memory (sq_push_reads, offsets derived on the GPU)"},
        "gpu_launches": int(launches), "wall_ms_per_step": wall / args.steps,
        "clocks": clk, "roofline": roofline, "em_iterations": iters,
        "work": {k2: int(st[k2]) for k2 in ("reads", "sketch_hashes", "queries", "hits", "postings", "pairs",
                                            "mid_reads", "slow_reads", "overflow_reads", "batches", "em_classes", "em_class_pairs")},
    }

    if rank == 0 and not args.no_cpu_baseline:
        tmp = tempfile.mkdtemp(prefix="sqbench")
        fq = os.path.join(tmp, "sample.fq")
        ns = write_sample_fastq(chunks[0], args.cpu_sample, fq)
        names = sqb.synth.transcript_names(T)
        n, kind, stages, times = cpu_reference_sample(names, postings, fq, ns)
        out["cpu_baseline"] = {"value": n / times[0], "unit": "reads/s", "cores": 1, "kind": kind,
                               "sample": "first %d reads of the workload against the full index (T=%d), single thread; "
                                         "stages s: %s" % (n, T, {k2: round(float(v), 3) for k2, v in stages.items()})}
    if rank == 0:
        emit(out)
    eng.close()
    if dist:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation on a bounded sample per step (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from _sqpkg import sqb
    import oracle_py
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    args.fragments = max(args.cpu_sample // 2, 1)  # only the sample is needed
    tx, T, tlen, chunks = make_workload(args, dev, 0, want_host=False)
    port = oracle_py.PortOracle()
    t0 = time.time()
    blob = sqb.synth.codes_to_ascii(tx["codes"])
    soff = tx["t_off"].cpu().numpy().astype(np.uint64)
    postings = {k: port.build_postings(blob, soff, k, max(K_LIST), port.threshold(SKETCH)) for k in K_LIST}
    log("[bench] CPU index build: %s keys (%.1fs)" % ([int(postings[k][0].shape[0]) for k in K_LIST], time.time() - t0))
    tmp = tempfile.mkdtemp(prefix="sqbench")
    fq = os.path.join(tmp, "sample.fq")
    ns = write_sample_fastq(chunks[0], args.cpu_sample, fq)
    names = sqb.synth.transcript_names(T)
    n, kind, stages, times = cpu_reference_sample(names, postings, fq, ns, repeats=args.steps + args.warmup)
    timed = times[args.warmup:]
    sec = sum(timed)
    value = n * len(timed) / sec
    cb = {"value": value, "unit": "reads/s", "cores": 1, "kind": kind,
          "sample": "%d reads per step against the full index (T=%d), single thread (the reference has no threads)" % (n, T)}
    out = {"impl": "reference", "metric": "quant reads/sec (sketch+seed+chain+assign)", "value": value, "unit": "reads/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / len(timed) * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 hash / f64 EM",
           "data": "synthetic",
           "config": {"workload": "config 2: synthetic human-scale transcriptome, k=31; bounded sample of %d reads per step" % n,
                      "transcripts": T, "read_len": READ_LEN, "k": K_LIST},
           "cpu_baseline": cb,
           "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "stages_s": {k2: round(float(v), 3) for k2, v in stages.items()}}
    emit(out)


if __name__ == "__main__":
    a = parse()
    apply_workload(a)
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
