"""N>1 host logic on the CPU (gloo, world_size 2): reads are sharded by contiguous record ranges, the index
is replicated, and the only exchange is a sum all-reduce of T-vectors (posterior sums per EM iteration, R once,
NumReads and presence at the end).  The per-rank compute is played by the CPU oracle here; the
GPU engine runs the same schedule with ncclAllReduce in place of the gloo call."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def shard_range(n, rank, world):
    """contiguous record range of a rank (same rule as bench.py / the CLI host)"""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _em_sharded(rank, world, off, tid, score, T, iters, tol):
    """EM of isoform_assignment.cpp:9-68 with the per-iteration all-reduce of posterior sums"""
    R_local = torch.tensor([len(off) - 1], dtype=torch.int64)
    dist.all_reduce(R_local)
    R = int(R_local)
    pi = np.full(T, 1.0 / T)
    done = 0
    for _ in range(iters):
        ps = np.zeros(T)
        for r in range(len(off) - 1):
            b, e = int(off[r]), int(off[r + 1])
            num = pi[tid[b:e]] * score[b:e].astype(np.float64)
            den = 0.0
            for v in num:
                den += v
            if den > 1e-10:
                np.add.at(ps, tid[b:e], num * (1.0 / den))
        t = torch.from_numpy(ps)
        dist.all_reduce(t)  # the one collective of the EM loop
        new = (t.numpy() + float(np.float32(0.01) / np.float32(R))) + float(np.float32(0.01))
        change = np.abs(new - pi).sum()
        pi = new
        done += 1
        if change < tol:
            break
    return pi, R, done


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), HERE]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py
    from datasets import SKETCH, dataset
    p = oracle_py.PortOracle()
    d = dataset()
    ks = [21, 31]
    thr = p.threshold(SKETCH)
    postings = p.postings_from_sequences(d["tseqs"], ks, thr)
    T = len(d["names"])
    lo, hi = shard_range(len(d["reads"]), rank, world)
    _, off, tid, score, R_local = p.chain_batch(ks, thr, 0.9, postings, d["reads"][lo:hi])
    pi, R, iters = _em_sharded(rank, world, off, tid, score, T, 20, 0.01)
    nr, present = p.assign(off, tid, score, T, pi)
    t_nr, t_pr = torch.from_numpy(nr.copy()), torch.from_numpy(present.astype(np.int32))
    dist.all_reduce(t_nr)
    dist.all_reduce(t_pr)
    if rank == 0:
        q.put((pi, t_nr.numpy(), (t_pr.numpy() > 0), R, iters))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                lo, hi = shard_range(n, r, w)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_two_rank_quant_equals_single_process(port):
    from datasets import SKETCH, dataset
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    prt = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, prt, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    pi, nr, present, R, iters = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    d = dataset()
    ks = [21, 31]
    thr = port.threshold(SKETCH)
    postings = port.postings_from_sequences(d["tseqs"], ks, thr)
    T = len(d["names"])
    _, off, tid, score, R1 = port.chain_batch(ks, thr, 0.9, postings, d["reads"])
    pi1, it1 = port.em(off, tid, score, R1, T)
    nr1, pr1 = port.assign(off, tid, score, T, pi1)
    assert R == R1 and iters == it1
    np.testing.assert_allclose(pi, pi1, rtol=1e-9)
    np.testing.assert_allclose(nr, nr1, rtol=1e-9, atol=1e-12)
    assert present.tolist() == pr1.astype(bool).tolist()
