#!/usr/bin/env python
"""Generates tests/golden/svg_*.npz: the rasteriser INPUTS of BASELINE config 2.

The reference's examples/svg.rs walks a usvg tree and, per path node, builds a `Transform` from
the node transform (svg.rs:112-116: row-major [a c; b d], offset (e, f)) and a `Vec<PathCmd>` from
the segments (svg.rs:118-138: MoveTo/LineTo/CurveTo/ClosePath -> Move/Line/Cubic/Close, f64 -> f32),
then rasterises the fill (svg.rs:140-147) and the stroke (svg.rs:149-156) with one fresh
`Rasterizer` each.  usvg 0.13 is a third-party crate that is not vendored and cannot run here, so
this script restates the subset of its behaviour the three bundled files need:
elements svg/g/path; attributes transform (matrix / translate / scale lists), fill, stroke,
stroke-width inherited through <g>; path letters M m L l H h V v C c S s Z z; ancestor transforms
accumulated into the path transform; computation in f64, cast to f32 at the end.  The SVG front
end's parity with real usvg is NOT pinned by any reference test ("parity unpinned"); config 2's
rasteriser input is DEFINED as the arrays this script writes, and parity is GPU vs oracle on them.

Run in the build container (reads /root/reference/examples/res, which does not exist on the GPU box):
    python tools/svg_fixtures.py
"""
from __future__ import annotations

import os
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = "/root/reference/examples/res"
OUT = os.path.join(ROOT, "tests", "golden")
MOVE, LINE, QUADRATIC, CUBIC, CONIC, CLOSE = range(6)

NUM = re.compile(r"[-+]?(?:\d*\.\d+|\d+\.?)(?:[eE][-+]?\d+)?")


def mat_mul(a, b):
    """(a then-applied-after b): p -> a(b(p)); matrices as (a, b, c, d, e, f) of the SVG matrix() form."""
    a0, a1, a2, a3, a4, a5 = a
    b0, b1, b2, b3, b4, b5 = b
    return (a0 * b0 + a2 * b1, a1 * b0 + a3 * b1, a0 * b2 + a2 * b3, a1 * b2 + a3 * b3, a0 * b4 + a2 * b5 + a4, a1 * b4 + a3 * b5 + a5)


def parse_transform(s):
    m = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)
    for name, args in re.findall(r"(\w+)\s*\(([^)]*)\)", s or ""):
        v = [float(x) for x in NUM.findall(args)]
        if name == "matrix":
            t = tuple(v[:6])
        elif name == "translate":
            t = (1.0, 0.0, 0.0, 1.0, v[0], v[1] if len(v) > 1 else 0.0)
        elif name == "scale":
            t = (v[0], 0.0, 0.0, v[1] if len(v) > 1 else v[0], 0.0, 0.0)
        else:
            raise ValueError(f"unsupported transform {name}")
        m = mat_mul(m, t)  # list composed left to right: the rightmost is applied to the point first
    return m


def parse_path(d):
    """-> list of (tag, x1, y1, x2, y2, x3, y3) absolute segments, as usvg would hand them over."""
    toks = re.findall(r"[MmLlHhVvCcSsZz]|" + NUM.pattern, d)
    out = []
    i = 0
    cx = cy = sx = sy = 0.0
    pcx = pcy = None  # second control point of the previous curve (for S/s)
    cmd = None
    while i < len(toks):
        t = toks[i]
        if re.fullmatch(r"[A-Za-z]", t):
            cmd = t
            i += 1
            if cmd in "Zz":
                out.append((CLOSE, 0, 0, 0, 0, 0, 0))
                cx, cy = sx, sy
                pcx = pcy = None
                continue
        if cmd is None:
            raise ValueError("path data does not start with a command")

        def take(n):
            nonlocal i
            v = [float(x) for x in toks[i:i + n]]
            if len(v) != n:
                raise ValueError("truncated path data")
            i += n
            return v

        rel = cmd.islower()
        c = cmd.upper()
        if c == "M":
            x, y = take(2)
            if rel:
                x, y = cx + x, cy + y
            out.append((MOVE, x, y, 0, 0, 0, 0))
            cx, cy, sx, sy = x, y, x, y
            cmd = "l" if rel else "L"  # implicit line-to for the following pairs
            pcx = pcy = None
        elif c == "L":
            x, y = take(2)
            if rel:
                x, y = cx + x, cy + y
            out.append((LINE, x, y, 0, 0, 0, 0))
            cx, cy = x, y
            pcx = pcy = None
        elif c == "H":
            (x,) = take(1)
            if rel:
                x = cx + x
            out.append((LINE, x, cy, 0, 0, 0, 0))
            cx = x
            pcx = pcy = None
        elif c == "V":
            (y,) = take(1)
            if rel:
                y = cy + y
            out.append((LINE, cx, y, 0, 0, 0, 0))
            cy = y
            pcx = pcy = None
        elif c == "C":
            x1, y1, x2, y2, x, y = take(6)
            if rel:
                x1, y1, x2, y2, x, y = cx + x1, cy + y1, cx + x2, cy + y2, cx + x, cy + y
            out.append((CUBIC, x1, y1, x2, y2, x, y))
            pcx, pcy = x2, y2
            cx, cy = x, y
        elif c == "S":
            x2, y2, x, y = take(4)
            if rel:
                x2, y2, x, y = cx + x2, cy + y2, cx + x, cy + y
            x1, y1 = (2 * cx - pcx, 2 * cy - pcy) if pcx is not None else (cx, cy)
            out.append((CUBIC, x1, y1, x2, y2, x, y))
            pcx, pcy = x2, y2
            cx, cy = x, y
        else:
            raise ValueError(f"unsupported path command {cmd}")
    return out


def walk(node, style, xf, paints):
    tag = node.tag.split("}")[-1]
    if tag not in ("svg", "g", "path"):
        return
    style = dict(style)
    for k in ("fill", "stroke", "stroke-width"):
        if node.get(k) is not None:
            style[k] = node.get(k)
    xf = mat_mul(xf, parse_transform(node.get("transform")))
    if tag == "path":
        segs = parse_path(node.get("d", ""))
        if not segs:
            return
        if style["fill"] != "none":
            paints.append((segs, xf, 0, 0.0))
        if style["stroke"] != "none":
            paints.append((segs, xf, 1, float(NUM.findall(style["stroke-width"])[0])))
        return
    for ch in node:
        walk(ch, style, xf, paints)


def convert(name):
    tree = ET.parse(os.path.join(RES, name))
    paints = []
    walk(tree.getroot(), {"fill": "black", "stroke": "none", "stroke-width": "1"}, (1.0, 0.0, 0.0, 1.0, 0.0, 0.0), paints)
    tags, vals, off, xfs, kind, width = [], [], [0], [], [], []
    for segs, m, k, w in paints:
        for s in segs:
            tags.append(s[0])
            vals.append(s[1:])
        off.append(len(tags))
        a, b, c, d, e, f = m
        xfs.append([a, c, b, d, e, f])  # svg.rs:112-116: Mat2x2::new(a, c, b, d), Vec2::new(e, f)
        kind.append(k)
        width.append(w)
    return dict(tag=np.array(tags, np.uint32), v=np.array(vals, np.float64).astype(np.float32).reshape(-1, 6),
                cmd_off=np.array(off, np.uint32), xf=np.array(xfs, np.float64).astype(np.float32).reshape(-1, 6),
                kind=np.array(kind, np.uint8), width=np.array(width, np.float64).astype(np.float32))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, short in (("Ghostscript_Tiger.svg", "tiger"), ("lorem-ipsum.svg", "lorem_ipsum"), ("calabi-yau.svg", "calabi_yau")):
        d = convert(name)
        path = os.path.join(OUT, f"svg_{short}.npz")
        np.savez_compressed(path, **d)
        n = len(d["cmd_off"]) - 1
        hist = np.bincount(d["tag"], minlength=6)
        print(f"{short}: {n} paints ({int((d['kind'] == 0).sum())} fills, {int((d['kind'] == 1).sum())} strokes), {len(d['tag'])} cmds "
              f"(M{hist[0]} L{hist[1]} C{hist[3]} Z{hist[5]}), {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    sys.exit(main())
