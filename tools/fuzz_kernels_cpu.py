"""Fuzzes the fused kernels' CUDA sources on the CPU (tests/emu/cuda_on_cpu.h): random batches through k_classify + the glyph
kernel, both CTA shapes of k_path and its striped form, in varying thread orders, compared byte for byte with the CPU emulation of
the kernels' arithmetic (tests/emu, fixed = True).  No GPU needed.

    python tools/fuzz_kernels_cpu.py [--gen mixed|extreme|tall] [--seed N] [--seconds S]

mixed    paths of 1-60 commands (lines, curves, Conics, extra Moves, Closes, repeated points), 6-400 px, random similarity
         transforms, a third on a half-integer grid (ties, tile-boundary cases)
extreme  sub-pixel shapes, shapes exactly one tile wide, coordinates near +-31 500, lines to the origin from there
tall     thin paths with lines of 2 500 - 30 000 pixels (bands, stripes, the DDA's overshoot of long lines)

Round 2: `extreme` found the striped form dropping the tiles a 31 000-pixel line leaves past its end pixel."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import emu
import emu_glyphs as EK
from test_glyph_kernel_cpu import CLOSE, CMD, CONIC, CUBIC, LINE, MOVE, QUAD, batch, mkpath
from test_kpath_cpu import compare

NPTS = {LINE: 1, MOVE: 1, QUAD: 2, CUBIC: 3, CONIC: 2, CLOSE: 0}


def gen_path(rng, extreme, conics):
    n = int(rng.integers(1, 60))
    if extreme:
        scale = float(rng.choice([0.4, 2.0, 7.9, 8.0, 16.0, 64.0, 900.0]))
        org = rng.choice([0.0, -8.0, 8.0, 31000.0, -31500.0, 4095.5], 2) + (rng.uniform(-3, 3, 2) if rng.random() < 0.5 else 0.0)
    else:
        scale = float(rng.choice([6.0, 20.0, 45.0, 120.0, 400.0]))
        org = rng.uniform(-200, 200, 2) if rng.random() < 0.5 else np.zeros(2)
    snap = rng.random() < 0.3

    def pts(k):
        v = rng.uniform(0, scale, (k, 2)) + org
        if snap:
            v = np.round(v * 2) / 2
        return v.astype(np.float32).reshape(-1)

    a = np.zeros(n + 1, CMD)
    a[0]["tag"] = MOVE if rng.random() < 0.9 else LINE
    a[0]["v"][:2] = pts(1)
    probs = [0.4, 0.2, 0.15, 0.08, 0.12, 0.05] if conics else [0.5, 0.25, 0.1, 0.05, 0.1, 0.0]
    for i, t in enumerate(rng.choice([LINE, QUAD, CUBIC, CLOSE, MOVE, CONIC], n, p=probs), 1):
        a[i]["tag"] = t
        k = NPTS[int(t)]
        if k:
            a[i]["v"][: 2 * k] = pts(k)
        if t == CONIC:
            a[i]["v"][4] = float(rng.choice([0.3, 0.7071, 1.0, 2.0]))
        if rng.random() < 0.05 and i > 1:
            a[i]["v"][:2] = a[i - 1]["v"][:2]  # a repeated point (the origin after a Close): degenerate or very long lines
    return a


def gen_tall(rng):
    n = int(rng.integers(2, 9))
    hx, hy = float(rng.choice([6.0, 12.0, 30.0])), float(rng.choice([2500.0, 6000.0, 9000.0, 14000.0, 30000.0]))
    horiz = rng.random() < 0.3
    p = []
    for _ in range(n):
        x, y = rng.uniform(0, hx), rng.uniform(0, hy)
        if rng.random() < 0.3:
            x, y = round(x), round(y)
        p.append((y, x) if horiz else (x, y))
    cmds = [(MOVE, *p[0])] + [(LINE, *q) for q in p[1:]]
    if rng.random() < 0.5:
        cmds.append((CLOSE,))
    return mkpath(*cmds)


def one(seed, gen):
    rng = np.random.default_rng(seed)
    if gen == "tall":
        paths, xfs = [gen_tall(rng) for _ in range(6)], None
    else:
        paths = [gen_path(rng, gen == "extreme", bool(seed % 2)) for _ in range(int(rng.integers(20, 90)))]
        xfs = []
        for _ in paths:
            s, th = rng.uniform(0.2, 1.5), rng.uniform(0, 6.28)
            if gen == "extreme" or rng.random() < 0.3:
                xfs.append([1, 0, 0, 1, 0, 0])
            else:
                xfs.append([s * np.cos(th), -s * np.sin(th), s * np.sin(th), s * np.cos(th), rng.uniform(-30, 30), rng.uniform(-30, 30)])
    cmds, off, xf = batch(paths, xfs)
    ref = emu.rasterize(cmds, off, xf, fixed=True)
    order = int(seed % 3) if seed % 3 < 2 else int(seed)
    n = len(paths)
    if gen != "tall":
        g = EK.run(cmds, off, xf, order=order, grid=int(rng.integers(1, 4)))
        assert g.status[0] == 0 and g.status[2] == 0
        handed = set(int(p) for p in g.handed_over)
        compare(g, ref, [int(p) for p in g.small if int(p) not in handed])
    for shape in ("pkl",) if gen == "tall" else ("pkl", "pks"):
        r = EK.run_kpath(cmds, off, xf, shape=shape, order=order, grid=2)
        assert r.status[0] == 0, (shape, r.status)
        ho = set(int(p) for p in r.handed_over)
        compare(r, ref, [p for p in range(n) if p not in ho])
        if shape == "pkl" and ho:
            r2 = EK.run_kpath(cmds, off, xf, shape="pkl", striped=True, paths=r.handed_over, order=order, grid=2, prev=r)
            ho2 = set(int(p) for p in r2.handed_over)
            compare(r2, ref, [p for p in range(n) if p not in ho2])
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gen", default="mixed", choices=["mixed", "extreme", "tall"])
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--seconds", type=float, default=120.0)
    args = ap.parse_args()
    t0, seed, paths = time.time(), args.seed, 0
    while time.time() - t0 < args.seconds:
        try:
            paths += one(seed, args.gen)
        except AssertionError as e:
            print(f"MISMATCH gen={args.gen} seed={seed}: {e}")
            sys.exit(1)
        seed += 1
    print(f"ok: gen={args.gen}, seeds {args.seed}..{seed - 1}, {paths} paths, every kernel byte-identical to the emulated arithmetic")


if __name__ == "__main__":
    main()
