"""Two GPUs of one box: reads sharded, index replicated, NCCL all-reduce of the T-vectors.  The CSV must equal
the single-GPU one (same candidate sets; sums differ only by re-association below the 6 printed digits)."""
import os
import subprocess

import pytest

from datasets import dataset
from test_gpu_cli import OURS, assert_csv_equal, read_csv, write_inputs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(OURS), reason="build/test not built")]


def test_two_gpus_equal_one(gpu_lib, tmp_path):
    if gpu_lib.sq_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = dataset(n_genes=60, n_reads=3000, seed=17)
    fa, fq = write_inputs(tmp_path, d)
    idx = str(tmp_path / "i.idx")
    subprocess.run([OURS, "-k", "21,31", "-o", "index", fa, idx], check=True, capture_output=True)
    out = {}
    for g in (1, 2):
        csv = str(tmp_path / ("g%d.csv" % g))
        r = subprocess.run([OURS, "--gpus", str(g), "-o", "quant", idx, fq, csv], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out[g] = read_csv(csv)
    assert_csv_equal(out[2], out[1])


@pytest.mark.parametrize("world,klist", [(2, "31"), (2, "21,31"), (4, "31"), (8, "31")])
def test_ranks_equal_one_engine(gpu_lib, tmp_path, world, klist):
    """C-ABI level: N per-rank engines with sq_comm_init (torchrun, NCCL) against one engine fed all reads"""
    import json
    import sys
    if gpu_lib.sq_device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / "res.json")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multirank_worker.py")
    port = 29500 + (os.getpid() % 500)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), worker, out, klist, "20000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    keep = os.environ.get("SQ_TEST_ARTIFACTS")
    if keep:
        os.makedirs(keep, exist_ok=True)
        json.dump(res, open(os.path.join(keep, "ranks_equal_one_engine_n%d_k%s.json" % (world, klist.replace(",", "_"))), "w"))
    assert res["candidates_equal"], res
    assert res["present_equal"] and res["ranks_bitwise_identical"], res
    assert res["pi_max_rel"] <= 1e-9 and res["numreads_max_rel"] <= 1e-9, res
    assert res["iterations"] == [res["iterations_one"]] * world, res
    # separate processes on one box: the EM sums travel over peer memory inside the M-step kernel; the NCCL path
    # (option peer_exchange = 0) gives the same result up to the order of the N-term sum
    assert res["peer_exchange"] == 1 and res["nccl_path_ran"], res
    assert res["nccl_vs_peer_pi_max_rel"] <= 1e-12 and res["nccl_vs_peer_numreads_max_rel"] <= 1e-12, res
    assert res["nccl_vs_peer_same_present_and_iterations"], res
