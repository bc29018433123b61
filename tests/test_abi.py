"""The C-ABI library loads without a GPU and exports every symbol include/b381.h declares; calls that
would need a device fail loudly (there is no CPU fallback in libb381.so)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b381.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b381_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = _declared()
    for must in ("b381_init", "b381_pairing_batch", "b381_miller_loop_batch", "b381_final_exp_batch",
                 "b381_pairing_product_is_one", "b381_g1_sum", "b381_g2_sum", "b381_g1_msm", "b381_g1_msm_shard_dev",
                 "b381_g1_fold_dev", "b381_verify_aggregate_common_batch_dev"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from bls_b200 import capi
    lib = capi.load()
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_pod_sizes_match_go_structs():
    from bls_b200 import layout as L
    assert (L.G1_AFFINE.itemsize, L.G2_AFFINE.itemsize, L.G1_JAC.itemsize, L.G2_JAC.itemsize, L.FP12.itemsize) == (104, 200, 144, 288, 576)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bls_b200 import capi
    with pytest.raises(capi.B381Error) as e:
        capi.Ctx(0)
    assert e.value.code == -4            # B381_ERR_NO_DEVICE
    lib = capi.load()
    assert lib.b381_pairing_batch(None, None, None, ctypes.c_size_t(1), None) == -1   # B381_ERR_ARG, no crash


def test_product_package_never_uses_the_oracle():
    """oracle/ is test infrastructure: nothing under bls_b200/ may import, link or dlopen it"""
    pkg = os.path.join(ROOT, "bls_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|pyoracle|liboracle|oracle/|oracle\.hpp|orc_")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h", ".cc", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(txt), (dirpath, f)
