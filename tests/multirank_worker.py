"""Worker of tests/test_gpu_multi.py::test_ranks_equal_one_engine (run under torch.distributed.run, one rank per GPU).

Every rank owns an engine on its GPU with an index replica and an NCCL communicator (sq_comm_init) and pushes a
contiguous shard of the reads; rank 0 then feeds ALL reads to one more engine without a communicator.  Checked
(the reference semantics: R over all shards and one posterior_sums per iteration, src/isoform_assignment.cpp:30-60):
  * candidate CSR of the concatenated shards == the single engine's, bit for bit
  * pi / NumReads within 1e-9 relative of the single engine's, `present` equal, same iteration count
  * pi, NumReads and present bitwise identical on every rank
Writes one JSON object to the path given as argv[1] (rank 0).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    from _sqpkg import sqb
    from datasets import dataset, SKETCH

    out_path = sys.argv[1]
    ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [31]
    n_reads = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    d = dataset(n_genes=300, n_reads=n_reads, seed=23)
    T = len(d["tseqs"])
    postings = sqb.api.build_kmer_to_transcript_map(d["tseqs"], ks, SKETCH, device=local)
    reads = d["reads"]
    per = (len(reads) + world - 1) // world
    mine = reads[rank * per:(rank + 1) * per]

    eng = sqb.Engine(ks, T, sketch_fraction=SKETCH, device=local)
    for ki, k in enumerate(ks):
        eng.load_index(ki, *postings[k])
    uid = torch.from_numpy(eng.nccl_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
    dist.broadcast(uid, 0)
    eng.comm_init(world, rank, uid.cpu().numpy())
    eng.set_option("batch_bases", 1 << 20)  # several batches per shard
    if mine:
        eng.push_reads(*sqb.packing.pack_reads(mine))
    off, tid, score = eng.candidates()
    eng.set_option("peer_exchange", 2)  # at any number of ranks (the default takes it only where it is faster: two)
    pi, nr, present, iters = eng.finish(0, 20, 0.01)
    peer_used = int(eng.stats()["peer_exchange"])
    # the same pass with ncclAllReduce per iteration instead of the exchange inside the M-step kernel
    eng.set_option("peer_exchange", 0)
    pi_n, nr_n, present_n, iters_n = eng.finish(0, 20, 0.01)
    nccl_used = int(eng.stats()["peer_exchange"]) == 0
    eng.close()

    # every rank's vectors must be the same bits
    mine_vec = torch.from_numpy(np.concatenate([pi.view(np.int64), nr.view(np.int64), present.astype(np.int64)])).cuda()
    allv = [torch.empty_like(mine_vec) for _ in range(world)]
    dist.all_gather(allv, mine_vec)
    same_bits = all(bool(torch.equal(allv[0], v)) for v in allv)
    shards = [None] * world
    dist.all_gather_object(shards, (off.astype(np.int64), tid, score, int(iters)))
    res = None
    if rank == 0:
        one = sqb.Engine(ks, T, sketch_fraction=SKETCH, device=local)
        for ki, k in enumerate(ks):
            one.load_index(ki, *postings[k])
        one.push_reads(*sqb.packing.pack_reads(reads))
        off1, tid1, score1 = one.candidates()
        pi1, nr1, present1, iters1 = one.finish(0, 20, 0.01)
        one.close()
        cat_off, base = [np.zeros(1, dtype=np.int64)], 0
        for o, _, _, _ in shards:
            cat_off.append(o[1:] + base)
            base += int(o[-1])
        cat_off = np.concatenate(cat_off)
        cat_tid = np.concatenate([s[1] for s in shards])
        cat_score = np.concatenate([s[2] for s in shards])
        rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0
        res = {
            "world": world, "reads": len(reads), "pairs": int(off1[-1]), "T": T, "k": ks,
            "candidates_equal": bool(np.array_equal(cat_off, off1.astype(np.int64)) and np.array_equal(cat_tid, tid1)
                                     and np.array_equal(cat_score, score1)),
            "pi_max_rel": rel(pi, pi1), "numreads_max_rel": rel(nr[present1 > 0], nr1[present1 > 0]),
            "present_equal": bool(np.array_equal(present, present1)),
            "ranks_bitwise_identical": bool(same_bits),
            "iterations": [int(s[3]) for s in shards], "iterations_one": int(iters1),
            "peer_exchange": peer_used, "nccl_path_ran": bool(nccl_used),
            "nccl_vs_peer_pi_max_rel": rel(pi_n, pi), "nccl_vs_peer_numreads_max_rel": rel(nr_n[present > 0], nr[present > 0]),
            "nccl_vs_peer_same_present_and_iterations": bool(np.array_equal(present_n, present) and iters_n == iters),
        }
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
