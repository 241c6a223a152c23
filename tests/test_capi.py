"""The C-ABI shared library: loads, exports every symbol include/gaussigan.h declares, validates arguments.
No compute calls here (no GPU needed): only host-side argument checks and the tile planner."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "gaussigan.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"GG_API\s+[\w\s\*]+?\b(gg_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gaussigan_b200 import _lib
    names = _declared()
    assert len(names) >= 9 and "gg_splat_fwd" in names and "gg_project_bwd" in names
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in gaussigan.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_constants():
    from gaussigan_b200 import _lib
    assert _lib.lib.gg_version() == 200
    hdr = open(os.path.join(ROOT, "include", "gaussigan.h")).read()
    assert int(re.search(r"#define GG_PARAM_STRIDE (\d+)", hdr).group(1)) == _lib.GG_PARAM_STRIDE
    assert int(re.search(r"#define GG_MOMENTS (\d+)", hdr).group(1)) == _lib.GG_MOMENTS
    assert int(re.search(r"#define GG_MAX_SCALES (\d+)", hdr).group(1)) == _lib.GG_MAX_SCALES


def test_argument_validation_sets_last_error():
    from gaussigan_b200 import _lib
    lib, ia = _lib.lib, _lib.int_array
    assert lib.gg_splat_bwd_tiles(0, 1, ia([256]), ia([0]), ia([1]), 0) == -2          # K <= 0
    assert "K" in _lib.last_error()
    assert lib.gg_splat_bwd_tiles(16, 0, ia([256]), ia([0]), ia([1]), 0) == -2         # no scales
    assert lib.gg_splat_bwd_tiles(16, 5, ia([8] * 5), ia([0] * 5), ia([1] * 5), 0) == -2
    assert lib.gg_splat_bwd_tiles(16, 1, ia([0]), ia([0]), ia([1]), 0) == -2           # n < 1
    assert lib.gg_splat_bwd_tiles(16, 1, ia([256]), ia([0]), ia([1]), 7) == -2         # layout
    # NULL / misaligned pointers are rejected before any launch (so this is safe without a GPU)
    assert lib.gg_project_fwd(None, None, 1, None, None, None, 0., 0., -2., 1., 4, 8, None, None, None, 0, None) == -1
    assert lib.gg_splat_fwd(None, 4, 8, 1, ia([32]), None, None, 0, 0, None) == -1
    assert lib.gg_splat_fwd(C.c_void_p(4), 4, 8, 1, ia([32]), None, None, 0, 0, None) == -3
    with pytest.raises(_lib.GaussiganError):
        _lib.check(lib.gg_project_fwd(None, None, 1, None, None, None, 0., 0., -2., 1., 4, -1, None, None, None, 0, None), "x")
    # B == 0 is a valid empty batch: no launch, status ok
    assert lib.gg_project_fwd(None, None, 1, None, None, None, 0., 0., -2., 1., 0, 8, None, None, None, 0, None) == 0


def test_tile_planner():
    from gaussigan_b200 import _lib
    lib, ia = _lib.lib, _lib.int_array
    # mask-only gradient at 256: strips of 8 px, 8 rows per pass, GG_RPT_BWD (default 2) passes per tile
    t256 = lib.gg_splat_bwd_tiles(16, 1, ia([256]), ia([0]), ia([1]), 0)
    assert t256 > 0 and 256 % t256 == 0
    # more scales -> more tiles; maps and mask-only scales are planned independently and add up
    t4 = lib.gg_splat_bwd_tiles(16, 4, ia([32, 64, 128, 256]), ia([1, 1, 1, 1]), ia([0, 0, 0, 1]), 0)
    t3 = lib.gg_splat_bwd_tiles(16, 3, ia([32, 64, 128]), ia([1, 1, 1]), ia([0, 0, 0]), 0)
    t1 = lib.gg_splat_bwd_tiles(16, 1, ia([256]), ia([1]), ia([1]), 0)
    assert t4 == t3 + t1
    # K that is not 4*2^m falls back to the strip kernels but still plans
    assert lib.gg_splat_bwd_tiles(6, 1, ia([100]), ia([1]), ia([0]), 0) > 0
    assert lib.gg_splat_bwd_tiles(6, 1, ia([100]), ia([1]), ia([0]), 1) > 0
    # nothing requested -> zero tiles
    assert lib.gg_splat_bwd_tiles(6, 1, ia([100]), ia([0]), ia([0]), 0) == 0
    # NHWC maps with n % 8 == 0: tiles of the persistent bulk-copy backward are sized in bytes (128 KB of upstream
    # gradient by default), whatever K is; the direct kernels (gg_set_fast_mask(0)) plan by CTA passes instead
    for K, n in ((16, 256), (6, 256), (12, 128), (5, 64), (64, 512), (128, 32)):
        t = lib.gg_splat_bwd_tiles(K, 1, ia([n]), ia([1]), ia([0]), 0)
        rows = -(-n // t)                                   # rows per tile
        assert t == -(-n // rows) and rows * n * K * 4 <= 2 * 128 * 1024 or rows == 1, (K, n, t)
    prev = lib.gg_set_fast_mask(0)
    try:
        assert lib.gg_splat_bwd_tiles(16, 1, ia([256]), ia([1]), ia([0]), 0) == 32       # 32 CTA passes of 256 quads
        assert lib.gg_splat_bwd_tiles(6, 1, ia([256]), ia([1]), ia([0]), 0) > 0          # strip kernels
    finally:
        lib.gg_set_fast_mask(prev)


def test_public_api_rejects_cpu_tensors_and_bad_shapes():
    import torch
    import gaussigan_b200 as gg
    z = torch.zeros(2, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        gg.get_landmarks(torch.zeros(2, 4, 1, 3), torch.zeros(2, 4, 3, 3, dtype=torch.float64), z, z, z, 0, 0, -2.)
    with pytest.raises(RuntimeError, match="no CPU path"):
        gg.get_gaussian_maps_2d(torch.zeros(2, 4, 2), torch.zeros(2, 4, 2, 2), [32, 32])
    with pytest.raises(AssertionError):
        gg.get_landmarks(None, torch.zeros(2, 4, 3, 3), z, z, z, 0, 0, -2.)          # reference: assert mus is not None
    with pytest.raises(RuntimeError, match="no CPU path"):
        from gaussigan_b200.plan import SilhouetteStep
        SilhouetteStep(2, 4, 32, "cpu")


def test_missing_library_fails_loudly(tmp_path):
    import subprocess
    import sys
    env = dict(os.environ, GAUSSIGAN_LIB=str(tmp_path / "nope.so"), PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", "import gaussigan_b200"], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU or PyTorch fallback" in r.stderr


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys
    code = "import sys, gaussigan_b200, gaussigan_b200.plan, gaussigan_b200.dist; sys.exit(int(any(m.split('.')[0]=='oracle' for m in sys.modules)))"
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gaussigan_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_reference_side_stub_sizes_its_buffers():
    """INTEGRATION.md section 3 / examples/gaussigan_binding.py: the loss-partials buffer is sized with
    gg_silhouette_step_loss_count (1024 at B=64, K=16, n=256 -- NOT the 296 of the unfused kernel), and the library
    rejects a buffer that is too small before touching anything (the stub used to overflow)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gaussigan_binding", os.path.join(ROOT, "examples", "gaussigan_binding.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    from gaussigan_b200 import _lib
    lib = _lib.lib
    assert stub.loss_partials_count(64, 16, 256) == lib.gg_silhouette_step_loss_count(64, 16, 256) == 64 * 16
    assert stub.loss_partials_count(4, 65, 64) == 296                      # K > 64: the unfused loss kernel
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "np.empty(296" not in doc and "gg_silhouette_step_loss_count" in doc
    # capacity check happens on the host, before any CUDA call: fake (non-NULL, aligned) pointers are never touched
    fake = C.c_void_p(1 << 20)
    ws = lib.gg_silhouette_step_workspace_bytes(64, 16, 256, 0)
    rc = lib.gg_silhouette_step_host(fake, fake, fake, fake, fake, fake, 0, 0., 0., -2., 1., 64, 16, 256,
                                     fake, 296, fake, fake, fake, None, fake, ws, 0, None)
    assert rc == -5 and "1024" in _lib.last_error()
    rc = lib.gg_silhouette_step_host(fake, fake, fake, fake, fake, fake, 0, 0., 0., -2., 1., 64, 16, 256,
                                     fake, 1024, fake, fake, fake, None, fake, ws - 1, 0, None)
    assert rc == -5 and "workspace" in _lib.last_error()


def test_mode_switches_are_safe_to_flip_from_threads():
    """gg_set_fast_mask / gg_set_pdl are atomics; the tile count follows the mode in force, and a stale T is
    rejected by the launch entry points (GG_E_WORKSPACE) instead of being trusted."""
    import threading
    from gaussigan_b200 import _lib
    lib, ia = _lib.lib, _lib.int_array
    t_fast = lib.gg_splat_bwd_tiles(16, 1, ia([256]), ia([0]), ia([1]), 0)
    stop = threading.Event()

    def flipper():
        while not stop.is_set():
            lib.gg_set_fast_mask(0)
            lib.gg_set_fast_mask(1)

    th = threading.Thread(target=flipper)
    th.start()
    try:
        seen = {lib.gg_splat_bwd_tiles(16, 1, ia([256]), ia([0]), ia([1]), 0) for _ in range(20000)}
    finally:
        stop.set()
        th.join()
        lib.gg_set_fast_mask(1)
    assert seen <= {t_fast, 32} and t_fast in seen          # only the two modes' answers, never garbage
    # a launch with a T that does not match the mode in force is refused on the host side
    fake = C.c_void_p(1 << 20)
    rc = lib.gg_splat_bwd(fake, 2, 16, 1, ia([256]), None, _lib.ptr_array([1 << 20]), 0, fake, t_fast + 1, 0, None)
    assert rc == -5


def test_strip_packing_is_lossless_on_the_host():
    """gg_encode_target_strips (host function, no CUDA): decode the buffer in numpy and compare with the image -- two-level,
    anti-aliased and noise masks, n that is not a multiple of 8 and n > 256 (several header words per row)."""
    import numpy as np
    from gaussigan_b200 import _lib
    lib = _lib.lib
    rng = np.random.default_rng(0)
    for n in (1, 5, 8, 37, 100, 256, 300, 520):
        for kind in range(3):
            B = 2
            if kind == 0:
                img = (rng.random((B, n, n)) < 0.5).astype(np.uint8) * 255
                img[:, : n // 2] = 0
                img[:, n // 2:, : n // 3] = 255
            elif kind == 1:
                img = np.clip(rng.normal(128, 200, (B, n, n)), 0, 255).astype(np.uint8)
            else:
                img = rng.integers(0, 256, (B, n, n), dtype=np.uint8)
            cap = lib.gg_target_strips_bound(B, n)
            buf = np.zeros(cap // 8 + 1, np.uint64)
            used = lib.gg_encode_target_strips(img.ctypes.data, B, n, buf.ctypes.data, cap)
            assert 0 < used <= cap
            raw = buf.view(np.uint8)[:used]
            hdr = raw[:16].view(np.uint32)
            assert hdr[0] == 0x54534747 and hdr[1] == n and hdr[2] == B and hdr[3] == used
            offs = raw[16:16 + 4 * B * n].view(np.uint32)
            pay0 = (16 + 4 * B * n + 7) & ~7
            strips = (n + 7) // 8
            W = (strips + 31) // 32
            dec = np.zeros((B * n, strips * 8), np.uint8)
            for r in range(B * n):
                rec = raw[pay0 + offs[r]:]
                M = rec[:4 * W].view(np.uint32)
                V = rec[4 * W:8 * W].view(np.uint32)
                k = 0
                for s in range(strips):
                    if (M[s >> 5] >> (s & 31)) & 1:
                        dec[r, 8 * s:8 * s + 8] = rec[8 * W + 8 * k:8 * W + 8 * k + 8]
                        k += 1
                    elif (V[s >> 5] >> (s & 31)) & 1:
                        dec[r, 8 * s:8 * s + 8] = 255
            assert np.array_equal(dec[:, :n].reshape(B, n, n), img), (n, kind)
    # too small a buffer is refused
    img = np.zeros((1, 64, 64), np.uint8)
    small = np.zeros(8, np.uint64)
    assert lib.gg_encode_target_strips(img.ctypes.data, 1, 64, small.ctypes.data, 64) == -5
