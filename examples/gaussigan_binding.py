"""gaussigan_binding.py -- the stub a maintainer of the TF1 reference would drop next to mask/mask_gen_dynamic.py.

numpy + ctypes only (no PyTorch): numpy arrays in, numpy arrays out, through the host-buffer entry point of
``libgaussigan_sm100.so`` (include/gaussigan.h), which does the H2D/D2H copies itself.  Wrapped in ``tf.py_func`` it
replaces ``get_landmarks`` + the sum composite + ``m_recon_loss_bis`` (mask/mask_gen_dynamic.py:707, 785-787) and, through
``tf.RegisterGradient``, their ``tf.gradients``.  INTEGRATION.md section 3 walks through it;
tests/test_capi.py and tests/test_parity_gpu.py execute it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("GAUSSIGAN_LIB", os.path.join(_HERE, "..", "gaussigan_b200", "lib", "libgaussigan_sm100.so"))
_lib = C.CDLL(_LIB_PATH)
_lib.gg_silhouette_step_workspace_bytes.restype = C.c_size_t
_lib.gg_silhouette_step_workspace_bytes.argtypes = [C.c_int] * 4
_lib.gg_silhouette_step_loss_count.restype = C.c_int
_lib.gg_silhouette_step_loss_count.argtypes = [C.c_int] * 3
_lib.gg_last_error.restype = C.c_char_p
_lib.gg_silhouette_step_host.restype = C.c_int
_lib.gg_silhouette_step_host.argtypes = (
    [C.c_void_p] * 6 + [C.c_int] + [C.c_double] * 4 + [C.c_int] * 3 +
    [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p])

GG_TARGET_U8 = 0
_cudart = None


def _rt():
    """libcudart through ctypes: the stub owns its device workspace (cudaMalloc) like any framework allocator would."""
    global _cudart
    if _cudart is None:
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _cudart = C.CDLL(name)
                break
            except OSError:
                continue
        if _cudart is None:
            raise RuntimeError("libcudart not found")
    return _cudart


def loss_partials_count(B, K, n):
    """Number of float32 loss partials one step writes (depends on K and on the arithmetic mode): ALWAYS size the
    buffer with this call -- it is B * T for K <= 64 (1024 at B=64, K=16, n=256), GG_LOSS_PARTIALS = 296 otherwise."""
    cnt = _lib.gg_silhouette_step_loss_count(B, K, n)
    if cnt < 0:
        raise ValueError(_lib.gg_last_error().decode())
    return cnt


def silhouette_loss_and_grads(mus, sigma, rotx, roty, rotz, target_u8, device=0):
    """numpy in (mus [B,K,3] f32, sigma [B,K,3,3] f64, rot* [B] f32 degrees, target [B,n,n] uint8) ->
    (loss, dmus [B,K,3] f32, dsigma [B,K,3,3] f64, drot [B,3] f32).
    == mask_recon_loss_bis (mask/mask_gen_dynamic.py:785-787) of the render at :707 and its tf.gradients."""
    mus = np.ascontiguousarray(mus, np.float32)
    sigma = np.ascontiguousarray(sigma, np.float64)
    rotx, roty, rotz = (np.ascontiguousarray(r, np.float32).reshape(-1) for r in (rotx, roty, rotz))
    target_u8 = np.ascontiguousarray(target_u8, np.uint8)
    B, K = sigma.shape[:2]
    n = target_u8.shape[1]
    assert target_u8.shape == (B, n, n) and mus.size == B * K * 3 and rotx.size == B
    rt = _rt()
    rt.cudaSetDevice(device)
    ws_bytes = _lib.gg_silhouette_step_workspace_bytes(B, K, n, GG_TARGET_U8)
    ws = C.c_void_p()
    if rt.cudaMalloc(C.byref(ws), C.c_size_t(ws_bytes)) != 0:
        raise MemoryError("cudaMalloc(%d) failed" % ws_bytes)
    try:
        loss_p = np.empty(loss_partials_count(B, K, n), np.float32)
        dmus = np.empty((B, K, 3), np.float32)
        dsig = np.empty((B, K, 3, 3), np.float64)
        drot = np.empty((B, 3), np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = _lib.gg_silhouette_step_host(p(mus), p(sigma), p(rotx), p(roty), p(rotz), p(target_u8), GG_TARGET_U8,
                                          0.0, 0.0, -2.0, 1.0, B, K, n, p(loss_p), loss_p.size, p(dmus), p(dsig),
                                          p(drot), None, ws, ws_bytes, device, None)
        if rc < 0:
            raise RuntimeError(_lib.gg_last_error().decode())
        if rt.cudaStreamSynchronize(None) != 0:
            raise RuntimeError("CUDA error while running the step")
    finally:
        rt.cudaFree(ws)
    return float(loss_p.astype(np.float64).sum() / (B * n * n)), dmus, dsig, drot
