"""The C-ABI library loads without a GPU, exports every symbol include/sketchquant.h declares, and refuses
to compute without a device (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sketchquant.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sq_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(sqb):
    lib = sqb.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsketchquant.so does not export %s" % n
    assert sorted(sqb.capi.EXPORTS) == names


def test_threshold_helper_needs_no_gpu(sqb):
    lib = sqb.load_library()
    assert lib.sq_threshold_from_fraction(float(np.float32(0.05))) == 214748367
    assert b"sm_100a" in lib.sq_version()


def test_no_cpu_fallback(sqb):
    lib = sqb.load_library()
    if lib.sq_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(sqb.SketchQuantError) as ei:
        sqb.Engine([31], 10)
    assert ei.value.code == sqb.capi.SQ_ERR_NO_DEVICE
    with pytest.raises(sqb.SketchQuantError):
        sqb.api.createSketch_FracMinhash_direct("ACGT" * 20, 31)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sketch-for-rna-seq_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "quant_oracle" not in txt and "libref_oracle" not in txt, f
