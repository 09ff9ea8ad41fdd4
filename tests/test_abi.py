"""CPU: the C ABI library loads and exports every entry point include/orbx.h declares, the Python
declarations cover exactly that set, and the product never reaches into the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "orbx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(orb[xmv]_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from orb_slam_fusion_b200 import _abi
    names = header_functions()
    assert len(names) >= 25
    assert names == set(_abi.SIGNATURES), names ^ set(_abi.SIGNATURES)
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    _abi.lib()


def test_no_gpu_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        return
    import orb_slam_fusion_b200 as P
    import numpy as np
    z = np.zeros(2, np.int32)
    for make in (lambda: P.OrbExtractor(1000, 1.2, 8, 20, 7), lambda: P.ORBmatcher(),
                 lambda: P.ORBVocabulary(2, 1, z, np.ones(2, np.uint8), np.zeros((2, 32), np.uint8), np.ones(2))):
        try:
            make()
        except P.OrbxError as e:
            assert e.code == -4
        else:
            raise AssertionError("constructed without a CUDA device")


def test_bad_arguments_are_rejected_before_any_cuda_call():
    from orb_slam_fusion_b200 import _abi
    lib = _abi.lib()
    h = ctypes.c_void_p()
    for p in [_abi.Params(1000, 1.2, 0, 20, 7), _abi.Params(1000, 1.2, 17, 20, 7), _abi.Params(0, 1.2, 8, 20, 7),
              _abi.Params(1000, 1.0, 8, 20, 7)]:
        assert lib.orbx_create(ctypes.byref(p), 0, 1, ctypes.byref(h)) == _abi.E_ARG
    assert lib.orbx_create(ctypes.byref(_abi.Params(1000, 1.2, 8, 20, 7)), 0, 0, ctypes.byref(h)) == _abi.E_ARG
    assert lib.orbx_last_error(None) == b"null handle"
    assert lib.orbx_launch_count(None) == 0


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "orb_slam_fusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".cc")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle/|liborb_oracle", txt, re.M), f
