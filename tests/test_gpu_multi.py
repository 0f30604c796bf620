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
