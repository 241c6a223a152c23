"""SASS evidence for the hot kernels: per-kernel opcode histogram and the listing of the hot loop (the innermost
backward-branch loop that contains MUFU.EX2), from the shipped library with cuobjdump.  No GPU needed.
    python tools/sass_hist.py > profiles/r2_sass_hot_kernels.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("GAUSSIGAN_LIB") or os.path.join(ROOT, "gaussigan_b200", "lib", "libgaussigan_sm100.so")
WANT = [("splat_bwd_mask_fast_kernel<1,false,8,false>", "_ZN2gg26splat_bwd_mask_fast_kernelILi1ELb0ELi8ELb0EEEvNS_9SplatArgsE"),
        ("splat_fwd_mask_fast_kernel<1,8>", "_ZN2gg26splat_fwd_mask_fast_kernelILi1ELi8EEEvNS_9SplatArgsE"),
        ("splat_bwd_mask_fast_kernel<1,true,8,false> (fused loss)", "_ZN2gg26splat_bwd_mask_fast_kernelILi1ELb1ELi8ELb0EEEvNS_9SplatArgsE")]
PACKED = ("FFMA2", "FMUL2", "FADD2")
def sass(fn):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins
def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]
for name, fn in WANT:
    ins = sass(fn)
    print("=" * 100)
    print(f"{name}: {len(ins)} SASS instructions")
    hist = collections.Counter(opcode(t) for _, t in ins)
    print("  whole kernel:", ", ".join(f"{k} {v}" for k, v in hist.most_common(24)))
    # innermost loop with MUFU.EX2: backward branches whose body contains EX2; take the one with the fewest instructions
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] <= i:
            j = addr[int(m.group(1), 16)]
            body = ins[j:i + 1]
            nex2 = sum("MUFU.EX2" in x for _, x in body)
            if nex2:
                loops.append((len(body), j, i, nex2))
    if not loops:
        continue
    n, j, i, nex2 = min(loops)
    body = ins[j:i + 1]
    h = collections.Counter(opcode(t) for _, t in body)
    packed = sum(h[p] for p in PACKED)
    scalar = sum(h[p] for p in ("FFMA", "FMUL", "FADD"))
    print(f"  hot loop (one Gaussian on 2 rows x 8 pixels; both the recurrence and the guarded direct branch are inside): "
          f"{n} instructions at {ins[j][0]:#x}..{ins[i][0]:#x}")
    print("   ", ", ".join(f"{k} {v}" for k, v in h.most_common()))
    print(f"    packed FP32x2 {packed}, scalar FP32 {scalar}, MUFU.EX2 {nex2} (the direct branch holds 16 of them), "
          f"LDS {h['LDS']}, STS {h['STS']}, LDL/STL {h['LDL'] + h['STL']}")
    # the path an UNGUARDED Gaussian takes: loop start .. the guard's branch, then from the branch target to the loop
    # end, without the transpose flush (BSSY .. BSYNC, executed for every 6th Gaussian only)
    guards = [x for x, (_, t) in enumerate(body) if t.startswith("FSETP")]
    k = guards[0] if guards else None
    if k is not None:
        # per guard: `@!P BRA T` skips the direct-evaluation block that follows it (the block ends with a BRA over the
        # recurrence block); the unguarded path is the loop body without those blocks
        drop = set()
        for x, (a, t) in enumerate(body):            # a direct block = (its guard's conditional BRA, the BRA that ends it]
            if re.match(r"BRA\s+0x", t) and x + 1 < len(body):
                y = next((z for z in range(x - 1, -1, -1) if re.search(r"@!?P\d\s+BRA\s+0x", body[z][1])), None)
                if y is not None and sum("MUFU.EX2" in q for _, q in body[y:x]) >= 8:
                    drop.update(b[0] for b in body[y + 1:x + 1])
        # second layout (Horner backward): `@!P BRA T` jumps forward to the direct block at T, which sits behind the
        # unguarded block; that one ends with `BRA J` right before T, and the direct block is [T, J)
        baddr = {a: x for x, (a, _) in enumerate(body)}
        for x, (a, t) in enumerate(body):
            m = re.search(r"@!?P\d\s+BRA\s+0x([0-9a-f]+)", t)
            if not m:
                continue
            T = int(m.group(1), 16)
            if T not in baddr or baddr[T] <= x or baddr[T] == 0:
                continue
            m2 = re.match(r"BRA\s+0x([0-9a-f]+)", body[baddr[T] - 1][1])
            if not m2 or int(m2.group(1), 16) not in baddr:
                continue
            J = baddr[int(m2.group(1), 16)]
            if J > baddr[T] and sum("MUFU.EX2" in q for _, q in body[baddr[T]:J]) >= 8:
                drop.update(b[0] for b in body[baddr[T]:J])
        path = [b for b in body if b[0] not in drop]
        ngauss = len(guards)
        flush_lo = next((a for a, t in path if t.startswith("BSSY")), None)
        flush_hi = next((a for a, t in path if t.startswith("BSYNC")), None)
        if flush_lo is not None and flush_hi is not None and sum(
                "LDS.128" in t for a, t in path if flush_lo <= a <= flush_hi) >= 4:
            flush = [b for b in path if flush_lo <= b[0] <= flush_hi]
            path = [b for b in path if not (flush_lo <= b[0] <= flush_hi)]
        else:
            flush = []
        hp, hf = collections.Counter(opcode(t) for _, t in path), collections.Counter(opcode(t) for _, t in flush)
        pk, sc = sum(hp[q] for q in PACKED), sum(hp[q] for q in ("FFMA", "FMUL", "FADD"))
        if ngauss > 1:
            print(f"    (the compiler unrolled the loop: {ngauss} Gaussians per iteration; counts below are per iteration)")
        print(f"    recurrence path of {ngauss} unguarded Gaussian(s): {len(path)} issue slots = {pk} packed FP32x2 + {sc} scalar FP32 "
              f"+ {sum('MUFU.EX2' in t for _, t in path)} MUFU.EX2 + {hp['LDS']} LDS + {hp['STS']} STS + {hp['LDL']} LDL + "
              f"{len(path) - pk - sc - sum('MUFU.EX2' in t for _, t in path) - hp['LDS'] - hp['STS'] - hp['LDL']} other "
              f"-> FP32-pipe lane-ops per (pixel, Gaussian) = (2 x {pk} + {sc}) / {16 * ngauss} = {(2 * pk + sc) / (16 * ngauss):.2f}")
        if flush:
            print(f"    transpose flush (every 6th Gaussian): {len(flush)} instructions, {hf['LDS']} LDS, {hf['FADD']} FADD")
    for a, t in body:
        print(f"      {a:04x}  {t}")
