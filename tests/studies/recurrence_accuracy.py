"""Numerical study (CPU, numpy float32): forward-differenced Gaussians along a pixel strip vs the fp64 oracle.

g(x_j) for x_j = x_0 + j h:  g_{j+1} = g_j r_j,  r_{j+1} = r_j c,  c = exp2(2 nA h^2).
Packed two-lane form used by the kernel: (g_j, g_{j+1}) *= (R_j, R_{j+1});  (R_j, R_{j+1}) *= c^4.
Emulates ex2.approx as correctly-rounded exp2 perturbed by up to `ulps` ulp.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import reference_port as rp

f32 = np.float32
rng = np.random.default_rng(0)

def ex2(x, ulps=2.0):
    y = np.exp2(x.astype(np.float64)).astype(f32)
    if ulps:
        y = (y.astype(np.float64) * (1.0 + rng.uniform(-ulps, ulps, size=y.shape) * 2.0 ** -24)).astype(f32)
    return y

def stage(mu2d, sig2d):
    det = sig2d[..., 0, 0] * sig2d[..., 1, 1] - sig2d[..., 0, 1] * sig2d[..., 1, 0]
    a = sig2d[..., 1, 1] / det; b = -0.5 * (sig2d[..., 0, 1] + sig2d[..., 1, 0]) / det; c = sig2d[..., 0, 0] / det
    L = 1.4426950408889634
    r = b / a; e = c - b * r
    return (mu2d[..., 0].astype(f32), mu2d[..., 1].astype(f32), (-a * L).astype(f32), (-r).astype(f32), (-e * L).astype(f32))

def render(params, n, run, mode):
    mux, muy, nA, nr, nE = [p[:, :, None, None] for p in params]     # [B,K,1,1]
    g = rp.tf_linspace_f32(n)
    h = f32(f32(2.0) / f32(n - 1))
    y = g[None, None, :, None]
    dy = (y - muy).astype(f32)
    sh = (nr * dy + mux).astype(f32)            # fma ~ one rounding
    sh = (nr.astype(np.float64) * dy + mux).astype(f32)
    e = ((nE * dy).astype(f32) * dy).astype(f32)
    B, K = mux.shape[:2]
    out = np.zeros((B, K, n, n), f32)
    if mode == "direct":
        x = g[None, None, None, :]
        t = (x - sh).astype(f32)
        arg = (nA.astype(np.float64) * (t * t).astype(f32) + e).astype(f32)
        return ex2(arg)
    # recurrence over runs of `run` pixels
    nAh2 = (f32(2.0) * nA * h).astype(f32)
    nAhh = ((nA * h).astype(f32) * h).astype(f32)
    c1 = np.exp2(2.0 * nA.astype(np.float64) * float(h) * float(h))     # accurate (double) then rounded
    c2 = (c1 ** 2).astype(f32); c4 = (c1 ** 4).astype(f32); c1 = c1.astype(f32)
    for s0 in range(0, n, run):
        x0 = g[s0]
        t0 = (x0 - sh).astype(f32)
        a0 = (nA.astype(np.float64) * (t0 * t0).astype(f32) + e).astype(f32)
        g0 = ex2(a0)
        r0 = ex2((nAh2.astype(np.float64) * t0 + nAhh).astype(f32))
        if mode == "seq":
            gj, rj = g0, r0
            for j in range(run):
                if s0 + j < n: out[:, :, :, s0 + j] = gj[..., 0]
                gj = (gj * rj).astype(f32); rj = (rj * c1).astype(f32)
        else:   # packed pairs
            g1 = (g0 * r0).astype(f32); r1 = (r0 * c1).astype(f32)
            R0 = (r0 * r1).astype(f32); R1 = (R0 * c2).astype(f32)
            ga, gb, Ra, Rb = g0, g1, R0, R1
            for j in range(0, run, 2):
                if s0 + j < n: out[:, :, :, s0 + j] = ga[..., 0]
                if s0 + j + 1 < n: out[:, :, :, s0 + j + 1] = gb[..., 0]
                ga = (ga * Ra).astype(f32); gb = (gb * Rb).astype(f32)
                Ra = (Ra * c4).astype(f32); Rb = (Rb * c4).astype(f32)
    return out

if __name__ == "__main__":
    B, K = 16, 16
    for n in (256, 128, 64, 32):
        inp = rp.synth_inputs(B, K, seed=56)
        mu2d, sig2d = rp.project(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0., 0., -2.)
        ref = rp.get_gaussian_maps_2d(mu2d, sig2d, [n, n]).permute(0, 3, 1, 2).numpy()      # [B,K,n,n] f64
        params = stage(mu2d.numpy(), sig2d.numpy())
        print(f"n={n}  max nA={np.abs(params[2]).max():.1f}  mask max={ref.sum(1).max():.2f}")
        for mode, run in (("direct", 0), ("seq", 4), ("seq", 8), ("pair", 8), ("pair", 16), ("seq", 16)):
            o = render(params, n, run, mode).astype(np.float64)
            em = np.abs(o - ref).max(); ek = np.abs(o.sum(1) - ref.sum(1)).max() / ref.sum(1).max()
            print(f"   {mode:6s} run={run:2d}  maps max|err| = {em:.2e}   mask max|err|/max = {ek:.2e}")
