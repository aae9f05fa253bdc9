"""ctypes binding of libt2n_b200.so (C ABI in include/t2n_b200.h).

The library is the product; there is no Python/CPU fallback.  Importing this module never
fails (so CPU-only hosts can import the package and run the host-side logic), but any attempt
to *use* the kernels without the shared library raises NativeLibraryError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libt2n_b200.so")
ABI_VERSION = 9


class NativeLibraryError(RuntimeError):
    pass


F3 = C.c_float * 3
I3 = C.c_int * 3
P3 = C.c_void_p * 3


class T2NField(C.Structure):
    _fields_ = [
        ("aabb_lo", F3), ("aabb_hi", F3), ("inv_aabb", F3),
        ("grid", I3),
        ("step_size", C.c_float),
        ("near_clip", C.c_float), ("far_clip", C.c_float),
        ("distance_scale", C.c_float), ("density_shift", C.c_float),
        ("weight_thres", C.c_float), ("eval_z_min", C.c_float),
        ("act", C.c_int), ("shading", C.c_int), ("app_dim", C.c_int), ("feature_c", C.c_int),
        ("n_sigma", I3), ("n_app", I3),
        ("mlp_in", C.c_int), ("mlp_in_pad", C.c_int),
        ("fea_pe", C.c_int), ("view_pe", C.c_int),
    ]


class T2NParams(C.Structure):
    _fields_ = [
        ("sigma_plane", P3), ("sigma_line", P3), ("app_plane", P3), ("app_line", P3),
        ("basis", C.c_void_p),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("pair_desc", C.c_void_p), ("col_perm", C.c_void_p),
    ]


class T2NGrads(C.Structure):
    _fields_ = [
        ("sigma_plane", P3), ("sigma_line", P3), ("app_plane", P3), ("app_line", P3),
        ("basis", C.c_void_p),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("app_done_event", C.c_void_p),
    ]


class T2NAlphaMask(C.Structure):
    _fields_ = [("volume", C.c_void_p), ("dims", I3), ("aabb_lo", F3), ("inv_size", F3)]


class T2NBatch(C.Structure):
    _fields_ = [("rays", C.c_void_p), ("jitter", C.c_void_p), ("R", C.c_int), ("S", C.c_int),
                ("is_train", C.c_int), ("white_bg", C.c_int)]


class T2NOutputs(C.Structure):
    _fields_ = [("rgb_map", C.c_void_p), ("depth_map", C.c_void_p), ("z_vals", C.c_void_p),
                ("weight", C.c_void_p)]


class T2NTransGrad(C.Structure):
    _fields_ = [("coef", C.c_void_p), ("depth_gt", C.c_void_p), ("delta", C.c_float)]


class T2NAdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_longlong), ("lr", C.c_float)]


class T2NScratch(C.Structure):
    _fields_ = [("sigma_feat", C.c_void_p), ("trans", C.c_void_p), ("acc", C.c_void_p),
                ("dsum", C.c_void_p), ("ray_start", C.c_void_p), ("ray_count", C.c_void_p),
                ("slots", C.c_void_p), ("app_rgb", C.c_void_p), ("counters", C.c_void_p),
                ("w1_packed", C.c_void_p), ("ray_flags", C.c_void_p), ("w1_grad_packed", C.c_void_p),
                ("mma_pack", C.c_void_p), ("act_h1_img", C.c_void_p), ("act_h2_img", C.c_void_p),
                ("act_feat", C.c_void_p), ("act_rows", C.c_int64), ("bwd_pack", C.c_void_p), ("bwd_img", C.c_void_p)]


# every symbol include/t2n_b200.h declares, with its ctypes signature
SYMBOLS = {
    "t2n_abi_version": (C.c_int, []),
    "t2n_error_string": (C.c_char_p, [C.c_int]),
    "t2n_device_sm_count": (C.c_int, []),
    "t2n_mma_pack_floats": (C.c_size_t, [C.POINTER(T2NField)]),
    "t2n_bwd_pack_floats": (C.c_size_t, [C.POINTER(T2NField)]),
    "t2n_bwd_image_row_bytes": (C.c_size_t, [C.POINTER(T2NField)]),
    "t2n_debug_make_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2n_debug_wgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "t2n_render_forward": (C.c_int, [C.POINTER(T2NField), C.POINTER(T2NParams), C.POINTER(T2NAlphaMask),
                                     C.POINTER(T2NBatch), C.POINTER(T2NOutputs), C.POINTER(T2NScratch),
                                     C.c_void_p]),
    "t2n_render_backward": (C.c_int, [C.POINTER(T2NField), C.POINTER(T2NParams), C.POINTER(T2NAlphaMask),
                                      C.POINTER(T2NBatch), C.POINTER(T2NOutputs), C.POINTER(T2NScratch),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(T2NGrads), C.c_void_p]),
    "t2n_render_backward_tg": (C.c_int, [C.POINTER(T2NField), C.POINTER(T2NParams), C.POINTER(T2NAlphaMask),
                                         C.POINTER(T2NBatch), C.POINTER(T2NOutputs), C.POINTER(T2NScratch),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(T2NTransGrad),
                                         C.POINTER(T2NGrads), C.c_void_p]),
    "t2n_data_loss": (C.c_int, [C.POINTER(T2NOutputs), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "t2n_adam_step": (C.c_int, [C.POINTER(T2NAdamTensor), C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                C.c_void_p]),
    "t2n_tv_blocks": (C.c_int, []),
    "t2n_tv_plane_sums": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2n_tv_plane_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p]),
    "t2n_get_rays": (C.c_int, [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2n_rotate_rays": (C.c_int, [C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "t2n_profile_enable": (C.c_int, [C.c_int]),
    "t2n_debug_trace_read": (C.c_int, [C.POINTER(C.c_longlong)]),
    "t2n_debug_trace_read_n": (C.c_int, [C.POINTER(C.c_longlong), C.c_int]),
    "t2n_debug_chunk_program": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_ubyte), C.c_int]),
    "t2n_debug_v2_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "t2n_debug_mma_recipe": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "t2n_debug_mma_bwd_recipe": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "t2n_profile_read": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int]),
    "t2n_dense_alpha": (C.c_int, [C.POINTER(T2NField), C.POINTER(T2NParams), C.POINTER(T2NAlphaMask),
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "t2n_alpha_pool_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "t2n_filter_rays": (C.c_int, [C.POINTER(T2NField), C.POINTER(T2NAlphaMask), C.c_void_p, C.c_longlong, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p]),
    "t2n_resample_plane": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "t2n_crop_plane": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p]),
    "t2n_forward_warp_scratch_doubles": (C.c_size_t, [C.c_int, C.c_int]),
    "t2n_forward_warp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "t2n_depth_discontinuity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2n_weighted_median": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "t2n_assemble_view": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "t2n_compute_alpha": (C.c_int, [C.POINTER(T2NField), C.POINTER(T2NParams), C.POINTER(T2NAlphaMask),
                                    C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol.  Raises
    NativeLibraryError -- loudly, no fallback -- if it is missing or stale."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("T2N_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise NativeLibraryError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C text2nerf_b200/csrc`).  text2nerf_b200 has no CPU / PyTorch fallback.")
    try:
        lib = C.CDLL(p)
    except OSError as e:  # pragma: no cover
        raise NativeLibraryError(f"cannot load {p}: {e}") from e
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{p} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.t2n_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"{p} has ABI {lib.t2n_abi_version()}, binding expects {ABI_VERSION}; rebuild it")
    if path is None:
        _lib = lib
    return lib


KERNEL_NAMES = {0: "march", 1: "pack_w1", 2: "appearance", 3: "finalize", 4: "app_backward_ffma",
                5: "unpack_w1_grad", 6: "ray_backward", 7: "pack_bwd", 8: "app_backward_mma", 9: "wgrad", 10: "data_loss", 11: "tv_sums", 12: "tv_grad", 13: "adam", 14: "app_scatter"}


def profile_read():
    """[(kernel name, ms)] of the most recent forward or backward call (profiling enabled)."""
    ids = (C.c_int * 16)()
    ms = (C.c_float * 16)()
    n = load().t2n_profile_read(ids, ms, 16)
    return [(KERNEL_NAMES.get(ids[i], str(ids[i])), float(ms[i])) for i in range(n)]


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().t2n_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")
