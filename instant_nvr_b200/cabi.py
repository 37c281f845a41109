"""ctypes binding of ``libnvr_b200.so`` (``include/nvr_b200.h``).

The structures below mirror the header field for field; ``tests/test_cabi_load.py`` checks the
sizes against the compiled library's view and that every declared symbol is exported.
There is deliberately no fallback: if the shared library is missing, ``load()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

MAX_LEVELS = 16
NUM_PARTS = 5
ABI_VERSION = 10

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libnvr_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int64_p = C.POINTER(C.c_int64)


class NvrGrid(C.Structure):
    _fields_ = [
        ("dense", C.c_void_p), ("hash", C.c_void_p), ("bounds", C.c_void_p),
        ("n_levels", C.c_int32), ("n_feat", C.c_int32), ("start_hash", C.c_int32), ("sum_features", C.c_int32),
        ("table_size", C.c_int64),
        ("res", C.c_int32 * MAX_LEVELS), ("size", C.c_float * MAX_LEVELS), ("dense_off", C.c_int64 * MAX_LEVELS),
    ]


class NvrLinear(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("in_dim", C.c_int32), ("out_dim", C.c_int32)]


class NvrPart(C.Structure):
    _fields_ = [
        ("grid", NvrGrid), ("occ", NvrLinear * 2), ("rgb", NvrLinear * 3),
        ("n_rgb", C.c_int32), ("n_latent", C.c_int32), ("rgb_latent", C.c_void_p),
    ]


class NvrParams(C.Structure):
    _fields_ = [("part", NvrPart * NUM_PARTS), ("deformer_grid", NvrGrid), ("deformer_mlp", NvrLinear * 3)]


class NvrFrame(C.Structure):
    _fields_ = [
        ("R", C.c_void_p), ("Th", C.c_void_p),
        ("pbw", C.c_void_p), ("pbw_dims", C.c_int32 * 3), ("pbw_channels", C.c_int32), ("pbounds", C.c_void_p),
        ("part_pts", C.c_void_p), ("part_pbw", C.c_void_p), ("lengths2", C.c_void_p),
        ("maxlen", C.c_int32), ("_pad0", C.c_int32),
        ("A", C.c_void_p), ("big_A", C.c_void_p),
        ("tuv", C.c_void_p), ("tuv_dims", C.c_int32 * 3), ("_pad1", C.c_int32), ("tbounds", C.c_void_p),
        ("frame_dim", C.c_void_p), ("latent_index", C.c_void_p), ("topology_key", C.c_int64),
    ]


class NvrConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("smpl_thresh", C.c_float), ("mlp_mode", C.c_int32),
                ("tune", C.c_uint32), ("_pad", C.c_uint32)]


class NvrCounters(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_survivors", C.c_int64), ("n_pairs", C.c_int64 * NUM_PARTS),
                ("n_far_pairs", C.c_int64 * NUM_PARTS), ("kernel_launches", C.c_int64), ("n_passes", C.c_int64)]


class NvrStageProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 7), ("launches", C.c_int64 * 7), ("passes", C.c_int64), ("survivors", C.c_int64),
                ("pairs", C.c_int64 * NUM_PARTS), ("far_pairs", C.c_int64 * NUM_PARTS),
                ("embed_part_ms", C.c_double * NUM_PARTS), ("mlp_part_ms", C.c_double * NUM_PARTS)]


class NvrIpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 64)]


class NvrAdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("step", C.c_int64), ("lr", C.c_double), ("weight_decay", C.c_double)]


class NvrSmplPose(C.Structure):
    _fields_ = [("Rh", C.c_double * 3), ("Th", C.c_double * 3), ("poses", C.c_double * 72), ("big_poses", C.c_double * 72),
                ("joints", C.c_float * 72), ("parents", C.c_int32 * 24)]


class NvrSmplOut(C.Structure):
    _fields_ = [("R", C.c_void_p), ("Th", C.c_void_p), ("A", C.c_void_p), ("big_A", C.c_void_p), ("ppts", C.c_void_p),
                ("part_pts", C.c_void_p), ("pbounds", C.c_void_p), ("wbounds", C.c_void_p)]


STAGE_NAMES = ("prep", "cull", "knn", "warp", "embed", "mlp", "resolve")

# name -> (restype, argtypes); exactly the declarations of include/nvr_b200.h
SYMBOLS = {
    "nvr_abi_version": (C.c_int, []),
    "nvr_create": (C.c_int, [C.POINTER(NvrConfig), C.POINTER(C.c_void_p)]),
    "nvr_destroy": (C.c_int, [C.c_void_p]),
    "nvr_last_error": (C.c_char_p, [C.c_void_p]),
    "nvr_bind_params": (C.c_int, [C.c_void_p, C.POINTER(NvrParams)]),
    "nvr_bind_frame": (C.c_int, [C.c_void_p, C.POINTER(NvrFrame), C.c_void_p]),
    "nvr_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "nvr_query_points": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_render_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_render_rays_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_frame_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(NvrIpcHandle)]),
    "nvr_frame_connect": (C.c_int, [C.c_void_p, C.POINTER(NvrIpcHandle)]),
    "nvr_frame_disconnect": (C.c_int, [C.c_void_p]),
    "nvr_frame_destroy": (C.c_int, [C.c_void_p]),
    "nvr_render_rays_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "nvr_render_rays_frame_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                             C.POINTER(C.c_void_p)]),
    "nvr_allgather_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "nvr_deformer_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nvr_embed_part": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nvr_part_mlp": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    "nvr_query_points_debug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_prepare_inference": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "nvr_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_train_scratch_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "nvr_train_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.POINTER(NvrParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_train_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nvr_distortion_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "nvr_distortion_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                          C.c_void_p]),
    "nvr_deformer_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(NvrParams), C.c_void_p]),
    "nvr_composite_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nvr_composite_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]),
    "nvr_adam_step": (C.c_int, [C.c_void_p, C.POINTER(NvrAdamTensor), C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32,
                                C.c_void_p]),
    "nvr_rays_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "nvr_generate_rays": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_assemble_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "nvr_sq_diff_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nvr_ssim_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_void_p, C.c_void_p]),
    "nvr_smpl_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "nvr_smpl_pose_frame": (C.c_int, [C.c_void_p, C.POINTER(NvrSmplPose), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_float,
                                      C.POINTER(NvrSmplOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nvr_smpl_volume_dims": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p]),
    "nvr_smpl_bweights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                    C.c_void_p, C.c_void_p]),
    "nvr_gather_footprint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.c_void_p]),
    "nvr_profile": (C.c_int, [C.c_void_p, C.c_int32]),
    "nvr_profile_read": (C.c_int, [C.c_void_p, C.POINTER(NvrStageProfile)]),
    "nvr_read_counters": (C.c_int, [C.c_void_p, C.POINTER(NvrCounters), C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library and attach prototypes.  Raises if it was not built
    (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built and instant_nvr_b200 has no CPU fallback. "
            "Run __graft_entry__.build() (nvcc -gencode arch=compute_100a,code=sm_100a).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)            # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    if lib.nvr_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libnvr_b200.so ABI {lib.nvr_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


# ---- descriptor builders (host values only; pointers are whatever the tensors' data_ptr() is) ----
def grid_desc(gp) -> NvrGrid:
    """``gp``: a params.GridParams module."""
    spec = gp.spec
    g = NvrGrid()
    g.dense, g.hash, g.bounds = gp.dense.data_ptr(), gp.hash.data_ptr(), gp.bounds.data_ptr()
    g.n_levels, g.n_feat, g.start_hash = spec.n_levels, spec.n_feat, spec.start_hash
    g.sum_features = int(spec.sum_features)
    g.table_size = spec.T
    size32 = getattr(gp, "_size32_host", None)                 # frozen buffer: one device->host read per module, not per call
    if size32 is None or gp._size32_src != (gp.entries_size.data_ptr(), gp.entries_size._version):
        size32 = gp.entries_size.detach().cpu().tolist()       # already fp32-rounded
        gp._size32_host, gp._size32_src = size32, (gp.entries_size.data_ptr(), gp.entries_size._version)
    for l in range(spec.n_levels):
        g.res[l], g.size[l], g.dense_off[l] = spec.res[l], size32[l], spec.dense_offsets[l]
    for l in range(spec.n_levels, MAX_LEVELS):
        g.res[l], g.size[l], g.dense_off[l] = 2, 1.0, 0
    return g


def linear_desc(lin) -> NvrLinear:
    d = NvrLinear()
    d.weight, d.bias = lin.weight.data_ptr(), lin.bias.data_ptr()
    d.in_dim, d.out_dim = lin.in_features, lin.out_features
    return d
