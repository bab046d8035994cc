"""CPU tests: the C-ABI library loads and exports every symbol include/saspa_b200.h declares; argument
errors are reported without touching a GPU; the product refuses to run without CUDA."""
import os
import re

import pytest
import torch

from saspa_aug_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "saspa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(saspa_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    names = _declared()
    assert len(names) >= 25
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/saspa_b200.h but missing from libsaspa_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in saspa_aug_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_tuning_hooks_are_not_part_of_the_public_abi():
    """The kernel-selection overrides (process-wide, for tests and A/B timing) are declared in csrc/tuning_hooks.h only."""
    hooks = open(os.path.join(ROOT, "saspa_aug_b200", "csrc", "tuning_hooks.h")).read()
    hooks = sorted(set(re.findall(r"\b(saspa_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hooks, flags=re.S))))
    assert hooks == sorted(_lib.TUNING_HOOKS) and not set(hooks) & set(_declared())
    lib = _lib.load()
    for n in hooks:
        assert hasattr(lib, n)
    assert lib.saspa_attention_impl(-1) == 0 and lib.saspa_conv_impl(-1) == 0 and lib.saspa_gemm_force_ctas(-1) == 0  # the library starts in auto


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.saspa_version() == 100
    assert isinstance(lib.saspa_last_error_string(), bytes)


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    rc = lib.saspa_canny_u8(None, 1, 8, 8, 2, 1, 2, None, 1, None, None, 0, None)
    assert rc == -1 and b"channels" in lib.saspa_last_error_string()
    rc = lib.saspa_gemm_bf16(None, 8, None, 8, None, 8, 4, 4, 4, None, None)
    assert rc == -1
    assert lib.saspa_canny_u8(None, 0, 8, 8, 3, 1, 2, None, 1, None, None, 0, None) == 0  # empty batch is a no-op
    assert lib.saspa_pil_ksize(512, 256, 0) == 5 and lib.saspa_pil_ksize(512, 224, 1) == 11


def test_pil_coeffs_host_match_oracle():
    import ctypes

    import numpy as np

    from oracle import clib

    lib = _lib.load()
    o = clib.lib()
    for in_size, out_size, f in [(512, 256, 0), (512, 224, 1), (704, 308, 1), (100, 250, 1)]:
        ks = lib.saspa_pil_ksize(in_size, out_size, f)
        b = (ctypes.c_int32 * (2 * out_size))()
        c = (ctypes.c_int32 * (ks * out_size))()
        k = ctypes.c_int()
        assert lib.saspa_pil_coeffs_host(in_size, out_size, f, ctypes.byref(k), b, c, ks * out_size) == 0
        bp, kp = ctypes.POINTER(ctypes.c_int)(), ctypes.POINTER(ctypes.c_int32)()
        ks2 = o.oracle_pil_coeffs(in_size, out_size, f, ctypes.byref(bp), ctypes.byref(kp))
        assert ks2 == ks == k.value
        assert np.array_equal(np.ctypeslib.as_array(bp, (2 * out_size,)), np.array(b))
        assert np.array_equal(np.ctypeslib.as_array(kp, (ks * out_size,)), np.array(c))


def test_ops_refuse_cpu_tensors():
    with pytest.raises(_lib.SaspaError):
        ops.canny(torch.zeros((1, 8, 8, 3), dtype=torch.uint8), 1, 2)
    with pytest.raises(_lib.SaspaError):
        ops.gemm(torch.zeros((8, 8), dtype=torch.bfloat16), torch.zeros((8, 8), dtype=torch.bfloat16))
