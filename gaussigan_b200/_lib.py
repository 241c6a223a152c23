"""ctypes binding of ``libgaussigan_sm100.so`` (the C ABI in ``include/gaussigan.h``).

There is NO fallback: if the shared library is missing the import raises, and every entry point that
returns a negative status raises ``GaussiganError`` with ``gg_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

GG_PARAM_STRIDE = 8
GG_MOMENTS = 5
GG_MAX_SCALES = 4
LAYOUT_NHWC = 0
LAYOUT_NCHW = 1

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libgaussigan_sm100.so"


class GaussiganError(RuntimeError):
    """A C-ABI call returned a negative status."""


def _load() -> C.CDLL:
    path = Path(os.environ.get("GAUSSIGAN_LIB", _LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python gaussigan_b200/build.py` "
            "(nvcc, sm_100a). gaussigan_b200 has no CPU or PyTorch fallback.")
    return C.CDLL(str(path))


lib = _load()

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_ip = C.POINTER(C.c_int)
_pp = C.POINTER(C.c_void_p)

#: name -> (restype, argtypes); mirrors include/gaussigan.h one to one
SIGNATURES = {
    "gg_version": (_i, []),
    "gg_last_error": (C.c_char_p, []),
    "gg_set_fast_mask": (_i, [_i]),
    "gg_set_pdl": (_i, [_i]),
    "gg_project_fwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _d, _d, _d, _d, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "gg_project_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _d, _d, _d, _d, _i, _i, _vp, _i, _vp, _vp,
                            _vp, _vp, _vp, _i, _vp]),
    "gg_project_fwd_views": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _d, _d, _d, _d, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "gg_project_bwd_views": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _d, _d, _d, _d, _i, _i, _i, _vp, _i, _vp, _vp,
                                  _vp, _vp, _vp, _i, _vp]),
    "gg_conic_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "gg_conic_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "gg_splat_fwd": (_i, [_vp, _i, _i, _i, _ip, _pp, _pp, _i, _i, _vp]),
    "gg_splat_bwd_tiles": (_i, [_i, _i, _ip, _ip, _ip, _i]),
    "gg_splat_fwd_concat": (_i, [_vp, _i, _i, _i, _ip, _pp, _ip, _ip, _pp, _i, _vp]),
    "gg_splat_bwd_tiles_concat": (_i, [_i, _i, _ip, _ip, _ip, _ip, _ip]),
    "gg_splat_bwd_concat": (_i, [_vp, _i, _i, _i, _ip, _pp, _ip, _ip, _pp, _vp, _i, _i, _vp]),
    "gg_mask_l1_loss_bwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "gg_silhouette_loss_fused": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "gg_sigma_from_frame_fwd": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp]),
    "gg_sigma_from_frame_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    "gg_p2p_buffer_bytes": (C.c_size_t, [_i, _i]),
    "gg_shared_grad_allreduce_p2p": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _i, _vp]),
    "gg_shared_grad_post_p2p": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _vp]),
    "gg_shared_grad_wait_p2p": (_i, [_i, _vp, _i, _i, _vp, _i, _vp]),
    "gg_svd3_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp]),
    "gg_svd3_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "gg_deform_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "gg_deform_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "gg_cycle_tiles": (_i, [_i]),
    "gg_cycle_loss_count": (_i, [_i, _i, _i]),
    "gg_cycle_l1_fused": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp]),
    "gg_colorize": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "gg_target_strips_bound": (C.c_size_t, [_i, _i]),
    "gg_encode_target_strips": (C.c_longlong, [_vp, _i, _i, _vp, C.c_size_t]),
    "gg_silhouette_step_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "gg_silhouette_step_loss_count": (_i, [_i, _i, _i]),
    "gg_silhouette_step_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _d, _d, _d, _i, _i, _i,
                                     _vp, _i, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _i, _vp]),
    "gg_splat_bwd": (_i, [_vp, _i, _i, _i, _ip, _pp, _pp, _i, _vp, _i, _i, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.gg_last_error() or b"").decode()


def check(status: int, what: str) -> int:
    if status < 0:
        raise GaussiganError(f"{what} failed ({status}): {last_error()}")
    return status


def int_array(values):
    return (C.c_int * len(values))(*[int(v) for v in values])


def ptr_array(ptrs):
    """Host array of device pointers; ``None``/0 entries become NULL."""
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) if p else None for p in ptrs])


def set_fast_mask(enable: bool) -> bool:
    """Mask-only launches: forward-differenced strips (True, default) or one exponential per pixel (False).
    Returns the previous setting (``gg_set_fast_mask``)."""
    return bool(lib.gg_set_fast_mask(int(bool(enable))))


def set_pdl(enable: bool) -> bool:
    """Programmatic dependent launch on/off for every kernel of the library (``gg_set_pdl``); returns the previous
    setting."""
    return bool(lib.gg_set_pdl(int(bool(enable))))
