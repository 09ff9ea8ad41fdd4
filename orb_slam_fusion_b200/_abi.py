"""ctypes declarations of the C ABI in include/orbx.h (liborbx_b200.so).

There is no CPU fallback: if the CUDA library has not been built, importing anything that
computes raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C orb_slam_fusion_b200/csrc`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORBX_LIB") or os.path.join(_HERE, "liborbx_b200.so")  # ORBX_LIB: A/B builds

OK, E_EMPTY, E_ARG, E_CAP, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_DEVICE, MEM_HOST_ASYNC = 0, 1, 2
STAGE_LEVEL, STAGE_BLUR, STAGE_CAND, STAGE_SELECTED = 0, 1, 2, 3
MAX_LEVELS = 16
OPT_CLAIM_SEQUENTIAL = 1
NCCL_ID_BYTES = 128
EDGE = 19

# cv::KeyPoint layout (28 bytes)
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
WQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4")])
WR_DTYPE = np.dtype([("best_dist", "<i4"), ("best_idx", "<i4"), ("best_level", "<i4"),
                     ("best_dist2", "<i4"), ("best_level2", "<i4")])


class Params(C.Structure):
    _fields_ = [("num_feats", C.c_int), ("scale_factor", C.c_float), ("num_levs", C.c_int),
                ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int)]


class GridGeom(C.Structure):
    _fields_ = [("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float), ("inv_h", C.c_float),
                ("cols", C.c_int32), ("rows", C.c_int32)]


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orbx error %d: %s" % (code, msg))
        self.code = code


vp, i32, i64, f32, sz, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64, C.c_double
pi32 = C.POINTER(C.c_int)

# name -> (restype, argtypes); the single source for the loader and for tests/test_abi.py
SIGNATURES = {
    "orbx_create": (i32, [C.POINTER(Params), i32, i32, C.POINTER(vp)]),
    "orbx_destroy": (None, [vp]),
    "orbx_last_error": (C.c_char_p, [vp]),
    "orbx_tables": (i32, [vp, vp, vp, vp, vp, vp]),
    "orbx_max_keypoints": (i32, [vp]),
    "orbx_extract": (i32, [vp, vp, i32, i32, sz, i32, i32, vp, vp, i32, pi32, pi32]),
    "orbx_extract_begin": (i32, [vp, vp, i32, i32, sz, i32, i32]),
    "orbx_extract_end": (i32, [vp, vp, vp, i32, pi32, pi32]),
    "orbx_compute_pyramid": (i32, [vp, vp, i32, i32, sz]),
    "orbx_pyramid_level": (i32, [vp, i32, vp, sz, pi32, pi32]),
    "orbx_extract_batch": (i32, [vp, vp, i32, i32, i32, sz, sz, i32, i32, i32, vp, vp, i32, vp, vp, vp]),
    "orbx_sync": (i32, [vp]),
    "orbx_launch_count": (C.c_longlong, [vp]),
    "orbx_debug_dropped": (i32, [vp, C.POINTER(C.c_longlong), i32]),
    "orbx_set_profiling": (i32, [vp, i32]),
    "orbx_stage_times": (i32, [vp, vp, C.POINTER(C.c_longlong), i32]),
    "orbx_stage_download": (i32, [vp, i32, i32, i32, vp, sz, pi32]),
    "orbx_synth_frames": (i32, [i32, i32, vp, i32, i32, i32, sz, sz, u64, u64, i32, u64, vp]),
    "orbm_create": (i32, [i32, C.POINTER(vp)]),
    "orbm_destroy": (None, [vp]),
    "orbm_last_error": (C.c_char_p, [vp]),
    "orbm_sync": (i32, [vp]),
    "orbm_launch_count": (C.c_longlong, [vp]),
    "orbm_set_option": (i32, [vp, i32, i32]),
    "orbm_nccl_unique_id": (i32, [vp]),
    "orbm_nccl_comm_create": (i32, [vp, i32, i32, i32, C.POINTER(vp)]),
    "orbm_nccl_comm_destroy": (i32, [vp]),
    "orbm_nccl_version": (i32, []),
    "orbm_knn2_sharded": (i32, [vp, vp, vp, i32, vp, i64, i64, dbl, vp, vp, vp, i32, vp]),
    "orbm_hamming_pairs": (i32, [vp, vp, vp, i64, vp, i32, vp]),
    "orbm_knn2": (i32, [vp, vp, i32, vp, i64, i64, vp, vp, i32, vp]),
    "orbm_top2_merge": (i32, [vp, vp, vp, i32, i32, vp, vp, i32, vp]),
    "orbm_ratio_test": (i32, [vp, vp, vp, i32, dbl, vp, i32, vp]),
    "orbm_stereo_rowband": (i32, [vp, vp, vp, i32, vp, vp, i32, vp, i32, i32, f32, f32, vp, vp, i32, vp]),
    "orbm_stereo_refine": (i32, [vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, f32, f32, f32, vp, vp, vp, i32, vp]),
    "orbm_stereo_matches_last": (i32, [vp, vp, vp, f32, f32, vp, vp, i32, pi32]),
    "orbm_distinctive": (i32, [vp, vp, vp, i32, i32, vp, vp, i32, vp]),
    "orbm_window_search_fuse": (i32, [vp, vp, vp, i32, C.POINTER(GridGeom), vp, vp, i32, vp, vp, vp, i32, vp, i32, vp]),
    "orbm_search_by_projection": (i32, [vp, vp, vp, i32, C.POINTER(GridGeom), vp, vp, i32, vp, vp, vp, vp, i32, f32, vp, vp, i32, vp]),
    "orbm_search_by_projection_last": (i32, [vp, vp, vp, i32, C.POINTER(GridGeom), vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, vp, vp, i32, vp]),
    "orbm_search_by_bow": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, i32, vp, vp, i32, vp]),
    "orbm_search_by_bow_kf": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, i32, vp, vp, i32, vp]),
    "orbm_search_for_triangulation": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32,
                                            vp, vp, i32, vp]),
    "orbm_window_search": (i32, [vp, vp, vp, i32, C.POINTER(GridGeom), vp, vp, i32, vp, vp, i32, vp]),
    "orbm_window_search_stereo": (i32, [vp, vp, vp, i32, C.POINTER(GridGeom), vp, vp, i32, vp, vp, vp, vp, vp, i32, vp]),
    "orbm_synth_descriptors": (i32, [i32, vp, i64, i64, u64, vp]),
    "orbm_popc_peak": (i32, [i32, i32, C.POINTER(dbl)]),
    "orbv_create": (i32, [i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(vp)]),
    "orbv_load_text": (i32, [i32, C.c_char_p, C.POINTER(vp)]),
    "orbv_destroy": (None, [vp]),
    "orbv_last_error": (C.c_char_p, [vp]),
    "orbv_info": (i32, [vp, pi32, pi32, pi32, pi32, pi32, pi32]),
    "orbv_sync": (i32, [vp]),
    "orbv_launch_count": (C.c_longlong, [vp]),
    "orbv_max_features": (i32, []),
    "orbv_features": (i32, [vp, vp, i32, i32, vp, vp, vp, i32, vp]),
    "orbv_transform": (i32, [vp, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
}

_LIB = None


def lib():
    """Load liborbx_b200.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is not built: run __graft_entry__.build() (nvcc, sm_100a). "
                              "orb_slam_fusion_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (rt, at) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = rt
            fn.argtypes = at
        _LIB = L
    return _LIB


def ptr(a):
    """Address of a numpy array, a torch tensor, an int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch.Tensor


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the C ABI reads a NULL stream as "the handle's own stream"


def torch_stream(device):
    """cudaStream_t of torch's current stream on `device`; torch's default stream has handle 0, which
    the C ABI would take for "use the handle's stream", so it is passed as cudaStreamLegacy."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream or CUDA_STREAM_LEGACY
