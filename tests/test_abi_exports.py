"""CPU test: the C-ABI library builds, loads and exports every symbol include/abawaca_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "abawaca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "abawaca_b200", "csrc")], check=True)
    from abawaca_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert missing == []
    assert sorted(capi.EXPORTS) == declared


def test_default_params_match_reference_constants():
    from abawaca_b200 import capi
    p = capi.default_params()
    assert (p.cluster_ndps_threshold, p.sensitivity_threshold, p.specificity_threshold, p.product_threshold, p.sum_threshold) == (100, 0.8, 0.8, 0.8, 1.6)
    assert (p.scg_overlap_threshold, p.scg_min_size, p.fraction_dps_in, p.split_scaf_ratio_threshold, p.max_snps, p.window_size) == (0.2, 500000, 0.8, 0.1, 15, 2000)


def test_no_device_means_no_context():
    """Without a GPU the product must fail loudly instead of falling back to the CPU."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from abawaca_b200 import capi
    with pytest.raises(capi.AbwError):
        capi.Context(0)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "abawaca_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("oracle/_ref", "") or f == "synth.py" or "oracle" not in re.sub(r"#.*|//.*", "", text).lower(), f
