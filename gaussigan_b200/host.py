"""Host-buffer front end: the training-style silhouette step for callers whose data lives in host memory.

``HostSilhouetteStep.submit(inputs)`` is one ``gg_silhouette_step_host`` call (include/gaussigan.h): H2D of
the Gaussians, camera and target mask, projection, splat, ``m_recon_loss_bis`` and its gradient
(mask/mask_gen_dynamic.py:785-787), splat backward, projection backward, D2H of the loss partials and the
gradients.  ``depth`` independent slots (device workspace + pinned result buffers + a CUDA stream each) are
used round-robin, so the copies of step i+1 overlap the kernels of step i.
"""
from __future__ import annotations

import ctypes as C
from typing import List, NamedTuple, Optional

import numpy as np
import torch

from ._lib import check, lib

GG_TARGET_U8, GG_TARGET_F32, GG_TARGET_BITS, GG_TARGET_STRIPS = 0, 1, 2, 3
GG_LOSS_PARTIALS = 296


def encode_target_strips(image_u8: torch.Tensor, pin: bool = True) -> torch.Tensor:
    """Pack 8-bit target masks ``[B,n,n]`` uint8 (host tensor) losslessly by 8-pixel strips (``GG_TARGET_STRIPS``,
    include/gaussigan.h): a strip whose pixels are all 0 or all 255 costs two header bits, any other strip its 8 raw
    bytes.  Silhouette masks are mostly flat, so the result is several times smaller than the image -- it is what
    crosses the host link.  Returns a 1-D uint8 host tensor (pinned by default) holding exactly the bytes in use; the
    step's results are bit-identical to those on the raw image.  One pass at memory speed: meant to be called by the
    data loader when it produces the mask."""
    assert image_u8.dtype == torch.uint8 and image_u8.dim() == 3 and image_u8.shape[1] == image_u8.shape[2]
    assert not image_u8.is_cuda, "encode_target_strips packs HOST images"
    img = image_u8.contiguous()
    B, n = int(img.shape[0]), int(img.shape[1])
    cap = int(lib.gg_target_strips_bound(B, n))
    scratch = torch.empty(cap // 8 + 1, dtype=torch.int64)
    used = int(lib.gg_encode_target_strips(img.data_ptr(), B, n, scratch.data_ptr(), cap))
    check(used, "gg_encode_target_strips")
    out = torch.empty((used + 7) // 8, dtype=torch.int64, pin_memory=pin)
    out.copy_(scratch[:out.numel()])
    return out.view(torch.uint8)[:used]


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(gpu_index: int) -> dict:
    """Bind the calling process to the CPU cores and the memory of the NUMA node GPU ``gpu_index`` hangs off.

    Call it once per rank BEFORE allocating pinned host buffers: the pages of ``cudaHostAlloc`` memory are placed by
    the allocating thread's memory policy, and with all ranks of a node left on the default policy every staging
    buffer ends up on NUMA node 0 -- the host-to-device copies of the ranks whose GPUs sit under the other socket then
    cross the inter-socket link and all eight share one socket's DRAM (round 1: 52.6 GB/s per GPU alone, 27.3 GB/s with
    eight ranks).  Uses sysfs (``/sys/bus/pci/devices/<bdf>/numa_node``, ``local_cpulist``), ``os.sched_setaffinity``
    and ``set_mempolicy(MPOL_PREFERRED)``; every step is best effort (containers may restrict cpusets) and the outcome
    is returned for the caller to log."""
    import os
    info = {"gpu": int(gpu_index), "node": None, "cpus_bound": None, "mempolicy": None}
    try:
        prop = torch.cuda.get_device_properties(gpu_index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read())
        info["node"] = node
        local = _parse_cpulist(open(base + "/local_cpulist").read())
        allowed = os.sched_getaffinity(0)
        use = local & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["cpus_bound"] = len(use)
        else:
            info["cpus_bound"] = 0 if not use else len(use)
        if node >= 0:
            libc = C.CDLL(None, use_errno=True)
            mask = C.c_ulong(1 << node)
            MPOL_PREFERRED, SYS_set_mempolicy = 1, 238          # x86_64
            rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, C.byref(mask), C.c_ulong(C.sizeof(mask) * 8))
            info["mempolicy"] = "preferred node %d" % node if rc == 0 else "errno %d" % C.get_errno()
    except Exception as exc:                      # pragma: no cover - depends on the host
        info["error"] = repr(exc)[:160]
    return info


class HostInputs(NamedTuple):
    """Pinned host tensors: mus [B,K,3] f32, sigma [B,K,3,3] f64, rot* [B] f32 (degrees), target [B,n,n] u8|f32."""
    mus: torch.Tensor
    sigma: torch.Tensor
    rotx: torch.Tensor
    roty: torch.Tensor
    rotz: torch.Tensor
    target: torch.Tensor


class StepResult:
    """Pinned host result buffers of one slot; valid after ``wait()``."""

    def __init__(self, B, K, n, want_mask, nloss):
        pin = dict(pin_memory=True)
        self.loss_partials = torch.zeros(nloss, dtype=torch.float32, **pin)
        self.dmus = torch.zeros((B, K, 3), dtype=torch.float32, **pin)
        self.dsigma = torch.zeros((B, K, 3, 3), dtype=torch.float64, **pin)
        self.drot = torch.zeros((B, 3), dtype=torch.float32, **pin)
        self.mask = torch.zeros((B, n, n), dtype=torch.float32, **pin) if want_mask else None
        self.event: Optional[torch.cuda.Event] = None
        self._npix = B * n * n

    def wait(self) -> "StepResult":
        if self.event is not None:
            self.event.synchronize()
        return self

    @property
    def loss(self) -> float:
        """mean |0.5*(target+1) - mask| (float64 sum of the per-CTA partials)."""
        self.wait()
        return float(self.loss_partials.double().sum().item() / self._npix)


class HostSilhouetteStep:
    #: K <= 64: project_fwd, silhouette_loss_fused, project_bwd; else project_fwd, splat_fwd, mask_l1_loss_bwd,
    #: splat_bwd, project_bwd
    LAUNCHES_PER_STEP = 3

    def __init__(self, B: int, K: int, size: int = 256, device="cuda", depth: int = 2,
                 tx: float = 0.0, ty: float = 0.0, tz: float = -2.0, focal: float = 1.0,
                 target_dtype: torch.dtype = torch.uint8, want_mask: bool = False, target_bits: bool = False,
                 target_strips: bool = False):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("gaussigan_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if target_dtype not in (torch.uint8, torch.float32):
            raise ValueError("target_dtype must be uint8 (raw 0..255 mask) or float32 (mask in [-1,1])")
        self.B, self.K, self.n, self.dev = int(B), int(K), int(size), dev
        self.t = (float(tx), float(ty), float(tz))
        self.focal = float(focal)
        self.kind = GG_TARGET_U8 if target_dtype == torch.uint8 else GG_TARGET_F32
        self.target_dtype = target_dtype
        if target_bits:          # binary masks, 1 bit per pixel (GG_TARGET_BITS): 8x fewer bytes over PCIe
            if target_dtype != torch.uint8 or K > 64:
                raise ValueError("target_bits needs uint8 storage and K <= 64")
            self.kind = GG_TARGET_BITS
        if target_strips:        # any 8-bit mask, losslessly packed by strips (GG_TARGET_STRIPS): ~7x fewer bytes over PCIe
            if target_dtype != torch.uint8 or K > 64 or target_bits:
                raise ValueError("target_strips needs uint8 masks and K <= 64 (and excludes target_bits)")
            self.kind = GG_TARGET_STRIPS
        self.ws_bytes = int(lib.gg_silhouette_step_workspace_bytes(self.B, self.K, self.n, self.kind))
        if self.ws_bytes <= 0:
            raise ValueError("bad B/K/size")
        self.depth = int(depth)
        self.ws = [torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev) for _ in range(self.depth)]
        self.streams = [torch.cuda.Stream(dev) for _ in range(self.depth)]
        self.nloss = check(lib.gg_silhouette_step_loss_count(self.B, self.K, self.n), "gg_silhouette_step_loss_count")
        self.LAUNCHES_PER_STEP = 3 if self.K <= 64 else 5
        self.results = [StepResult(self.B, self.K, self.n, want_mask, self.nloss) for _ in range(self.depth)]
        self._next = 0
        # per-call constants of submit(): the middle of the argument list, every slot's own addresses, one event per slot
        self._call_mid = (self.kind, *self.t, self.focal, self.B, self.K, self.n)
        self._slot_args = [(r.loss_partials.data_ptr(), r.loss_partials.numel(), r.dmus.data_ptr(), r.dsigma.data_ptr(),
                            r.drot.data_ptr(), r.mask.data_ptr() if r.mask is not None else None, w.data_ptr(),
                            self.ws_bytes, dev.index, st.cuda_stream)
                           for r, w, st in zip(self.results, self.ws, self.streams)]
        self._events = [torch.cuda.Event() for _ in range(self.depth)]
        self._hcache = {}
        self.target_numel = B * size * ((size + 7) // 8) if self.kind == GG_TARGET_BITS else B * size * size
        tb = 4 if self.kind == GG_TARGET_F32 else 1
        self.small_h2d_bytes = B * K * 3 * 4 + B * K * 9 * 8 + 3 * B * 4
        #: bytes copied host -> device per step (strips: the size of the LAST submitted target; it is data dependent)
        self.h2d_bytes = self.small_h2d_bytes + self.target_numel * tb
        self.d2h_bytes = (self.nloss * 4 + B * K * 3 * 4 + B * K * 9 * 8 + B * 3 * 4
                          + (B * size * size * 4 if want_mask else 0))

    def make_host_inputs(self, inp: dict, target_pm1: torch.Tensor) -> HostInputs:
        """Pack oracle-style inputs (mus [B,K,1,3], sigma, rot [B,1], target [B,n,n,1] in {-1,+1}) into pinned buffers."""
        B, K, n = self.B, self.K, self.n
        t = target_pm1.reshape(B, n, n)
        if self.kind == GG_TARGET_STRIPS:
            raw = ((t + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
            t = encode_target_strips(raw)
        elif self.kind == GG_TARGET_BITS:
            assert bool(((t == 1) | (t == -1)).all()), "target_bits: the masks must be two-level (-1 / +1)"
            t = torch.from_numpy(np.packbits((t > 0).numpy(), axis=-1, bitorder="little"))      # [B,n,ceil(n/8)]
        elif self.kind == GG_TARGET_U8:
            t = ((t + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
        else:
            t = t.float()
        # the five small arrays in ONE pinned block, spaced like the device workspace (256-byte aligned, rot* back to
        # back): gg_silhouette_step_host then moves them with a single copy
        al = lambda x: (x + 255) & ~255
        o_sig = al(B * K * 3 * 4)
        o_rot = o_sig + al(B * K * 9 * 8)
        block = torch.zeros(o_rot + 3 * B * 4, dtype=torch.uint8).pin_memory()
        mus = block[:B * K * 12].view(torch.float32).view(B, K, 3)
        sigma = block[o_sig:o_sig + B * K * 72].view(torch.float64).view(B, K, 3, 3)
        rots = [block[o_rot + i * B * 4:o_rot + (i + 1) * B * 4].view(torch.float32) for i in range(3)]
        mus.copy_(inp["mus"].reshape(B, K, 3).float())
        sigma.copy_(inp["sigma"].double())
        for r, k in zip(rots, ("rotx", "roty", "rotz")):
            r.copy_(inp[k].reshape(B).float())
        return HostInputs(mus, sigma, rots[0], rots[1], rots[2], t if t.is_pinned() else t.contiguous().pin_memory())

    def _check_inputs(self, h: HostInputs):
        B, K = self.B, self.K
        assert h.mus.dtype == torch.float32 and h.mus.numel() == B * K * 3 and h.mus.is_contiguous()
        assert h.sigma.dtype == torch.float64 and h.sigma.numel() == B * K * 9 and h.sigma.is_contiguous()
        if self.kind == GG_TARGET_STRIPS:
            assert h.target.dtype == torch.uint8 and h.target.is_contiguous() and h.target.data_ptr() % 8 == 0
        else:
            assert h.target.dtype == self.target_dtype and h.target.numel() == self.target_numel and h.target.is_contiguous()
        for r in (h.rotx, h.roty, h.rotz):
            assert r.dtype == torch.float32 and r.numel() == B and r.is_contiguous()
        assert not h.mus.is_cuda and not h.target.is_cuda, "HostSilhouetteStep takes HOST tensors"

    def submit(self, h: HostInputs) -> StepResult:
        """Enqueue one whole step on the next slot; returns that slot's (asynchronously filled) result.

        The host side of a submit is what bounds small batches, so the per-call Python work is kept to a minimum: the
        shape checks and the address look-ups of a ``HostInputs`` object run the first time it is submitted (the object
        must not be re-pointed afterwards; refilling its tensors in place is what a data loader does), the slot's own
        addresses are formed once, and every slot re-records one event."""
        key = id(h)
        hp = self._hcache.get(key)
        if hp is None or hp[0] is not h:
            self._check_inputs(h)
            if len(self._hcache) > 64:
                self._hcache.clear()
            hp = (h, (h.mus.data_ptr(), h.sigma.data_ptr(), h.rotx.data_ptr(), h.roty.data_ptr(), h.rotz.data_ptr(),
                      h.target.data_ptr()), h.target.numel())
            self._hcache[key] = hp
        if self.kind == GG_TARGET_STRIPS:
            self.h2d_bytes = self.small_h2d_bytes + hp[2]
        i = self._next
        self._next = (i + 1) % self.depth
        res = self.results[i]
        if res.event is not None:
            res.event.synchronize()                       # the slot's previous step must have drained
        rc = lib.gg_silhouette_step_host(*hp[1], *self._call_mid, *self._slot_args[i])
        if rc:
            check(rc, "gg_silhouette_step_host")
        ev = self._events[i]
        ev.record(self.streams[i])
        res.event = ev
        return res

    def drain(self):
        """Make the current stream wait for every slot (so CUDA events recorded after it bracket the work)."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            cur.wait_stream(st)
