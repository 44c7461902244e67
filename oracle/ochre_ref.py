"""Second, independent CPU restatement of the ochre rasteriser in numpy float32 scalars.

TEST INFRASTRUCTURE ONLY (same rule as ochre_oracle.c): used by tests/ to
cross-check the C oracle on small inputs and to regenerate tests/golden/.
Pure-Python loops -- small cases only.

Follows the reference's Rust directly (not the C file):
  src/geom.rs:50-52 (lerp), :170-178 + :260-262 (Transform::apply)
  src/path.rs:16-37 (transform), :41-109 (flatten), :114-144 (free flatten), :152-274 (stroke)
  src/rasterizer.rs:61-69 (move_to), :72-140 (line_to), :145-165 (command/fill),
                    :169-171 (stroke), :180-268 (finish)

PARITY STATUS: "parity unpinned" -- the reference has no tests or golden vectors
and cannot be compiled here (no Rust toolchain); see oracle/ochre_oracle.c.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
ONE = F(1.0)
ZERO = F(0.0)
INF = F(np.inf)
TILE = 8
TOL = F(0.1)

MOVE, LINE, QUAD, CUBIC, CONIC, CLOSE = range(6)


def _min(a, b):  # Rust f32::min ignores a NaN operand
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def _max(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def _i16(f) -> int:  # `as i16`: saturating, NaN -> 0
    if f != f:
        return 0
    if f <= -32768.0:
        return -32768
    if f >= 32767.0:
        return 32767
    return int(f)


def _wrap16(i: int) -> int:
    return ((i + 32768) & 0xFFFF) - 32768


def _u8(f) -> int:
    if f != f or f <= 0.0:
        return 0
    if f >= 255.0:
        return 255
    return int(f)


def _signum(f):
    if f != f:
        return f
    return F(math.copysign(1.0, float(f)))


class V:
    __slots__ = ("x", "y")
    __array_ufunc__ = None  # make `np.float32 * V` defer to V.__rmul__

    def __init__(self, x, y):
        self.x = F(x)
        self.y = F(y)

    def __add__(self, o):
        return V(self.x + o.x, self.y + o.y)

    def __sub__(self, o):
        return V(self.x - o.x, self.y - o.y)

    def __rmul__(self, s):  # f32 * Vec2
        s = F(s)
        return V(s * self.x, s * self.y)

    def mulr(self, s):  # Vec2 * f32
        s = F(s)
        return V(self.x * s, self.y * s)

    def __eq__(self, o):
        return bool(self.x == o.x and self.y == o.y)

    def __ne__(self, o):
        return not self.__eq__(o)

    def dot(self, o):
        return self.x * o.x + self.y * o.y

    def length(self):
        return np.sqrt(self.dot(self))

    def t(self):
        return (float(self.x), float(self.y))


def lerp(t, a, b):
    return (ONE - t) * a + t * b


class Xf:
    def __init__(self, m=(1, 0, 0, 1), off=(0, 0)):
        self.m = [F(v) for v in m]
        self.o = V(*off)

    def apply(self, v: V) -> V:
        r = V(self.m[0] * v.x + self.m[1] * v.y, self.m[2] * v.x + self.m[3] * v.y)
        return r + self.o


def transform_cmd(cmd, xf: Xf):
    tag = cmd[0]
    if tag in (MOVE, LINE):
        return (tag, xf.apply(cmd[1]))
    if tag == QUAD:
        return (tag, xf.apply(cmd[1]), xf.apply(cmd[2]))
    if tag == CUBIC:
        return (tag, xf.apply(cmd[1]), xf.apply(cmd[2]), xf.apply(cmd[3]))
    if tag == CONIC:
        return (tag, xf.apply(cmd[1]), xf.apply(cmd[2]), F(cmd[3]))
    return (CLOSE,)


def flatten_cmd(cmd, last: V, tol, emit):
    tag = cmd[0]
    with np.errstate(all="ignore"):
        if tag == MOVE:
            emit((MOVE, cmd[1]))
        elif tag == LINE:
            emit((LINE, cmd[1]))
        elif tag == QUAD:
            c, p = cmd[1], cmd[2]
            dt = np.sqrt((F(4.0) * tol) / (last - F(2.0) * c + p).length())
            t = ZERO
            while t < ONE:
                t = _min(t + dt, ONE)
                p01 = lerp(t, last, c)
                p12 = lerp(t, c, p)
                emit((LINE, lerp(t, p01, p12)))
        elif tag == CUBIC:
            c1, c2, p = cmd[1], cmd[2], cmd[3]
            a = F(-1.0) * last + F(3.0) * c1 - F(3.0) * c2 + p
            b = F(3.0) * (last - F(2.0) * c1 + c2)
            conc = _max(b.length(), (a + b).length())
            dt = np.sqrt((np.sqrt(F(8.0)) * tol) / conc)
            t = ZERO
            while t < ONE:
                t = _min(t + dt, ONE)
                p01 = lerp(t, last, c1)
                p12 = lerp(t, c1, c2)
                p23 = lerp(t, c2, p)
                p012 = lerp(t, p01, p12)
                p123 = lerp(t, p12, p23)
                emit((LINE, lerp(t, p012, p123)))
        elif tag == CONIC:
            c, p, w = cmd[1], cmd[2], F(cmd[3])

            def rec(t0, t1, p0, p1):
                t = F(0.5) * (t0 + t1)
                p01 = lerp(t, last, w * c)
                p12 = lerp(t, w * c, p)
                denom = (ONE - t) * (ONE - t) + F(2.0) * t * (ONE - t) * w + t * t
                mid = (ONE / denom) * lerp(t, p01, p12)
                err = (mid - F(0.5) * (p0 + p1)).length()
                if err > tol:
                    rec(t0, t, p0, mid)
                    rec(t, t1, mid, p1)
                else:
                    emit((LINE, mid))
                    emit((LINE, p1))

            rec(ZERO, ONE, last, p)
        else:
            emit((CLOSE,))


def flatten_path(path, tol=TOL):
    last = V(0, 0)
    out = []
    for cmd in path:
        flatten_cmd(cmd, last, tol, out.append)
        if cmd[0] in (MOVE, LINE):
            last = cmd[1]
        elif cmd[0] in (QUAD, CONIC):
            last = cmd[2]
        elif cmd[0] == CUBIC:
            last = cmd[3]
    return out


def stroke_path(polygon, width):
    width = F(width)
    out = []

    def pt(c):
        assert c[0] in (MOVE, LINE)
        return c[1]

    def join(prev_n, next_n, point):
        with np.errstate(all="ignore"):
            off = ONE / (ONE + prev_n.dot(next_n))
        if abs(off) > 2.0:
            out.append((LINE, point + (F(0.5) * width) * prev_n))
            out.append((LINE, point + (F(0.5) * width) * next_n))
        else:
            out.append((LINE, point + (F(0.5) * width * off) * (prev_n + next_n)))

    def offset(contour, closed, reverse):
        n = len(contour)
        first_point = pt(contour[0]) if closed == reverse else pt(contour[-1])
        prev_point = first_point
        prev_normal = V(0, 0)
        i = 0
        while True:
            if i < n:
                next_point = pt(contour[n - i - 1]) if reverse else pt(contour[i])
            else:
                next_point = first_point
            if next_point != prev_point or i == n:
                tan = next_point - prev_point
                nn = V(-tan.y, tan.x)
                ln = nn.length()
                nn = V(0, 0) if ln == 0.0 else nn.mulr(ONE / ln)
                join(prev_normal, nn, prev_point)
                prev_point = next_point
                prev_normal = nn
            i += 1
            if i > n:
                break

    cs = ce = 0
    closed = False
    it = iter(polygon)
    while True:
        cmd = next(it, None)
        if cmd is not None and cmd[0] == CLOSE:
            closed = True
        if cmd is None or cmd[0] in (MOVE, CLOSE):
            if cs != ce:
                contour = polygon[cs:ce]
                base = len(out)
                offset(contour, closed, False)
                out[base] = (MOVE, pt(out[base]))
                if closed:
                    out.append((CLOSE,))
                base = len(out)
                offset(contour, closed, True)
                if closed:
                    out[base] = (MOVE, pt(out[base]))
                out.append((CLOSE,))
        if cmd is None:
            break
        if cmd[0] == MOVE:
            cs = ce
            ce = cs + 1
        elif cmd[0] == LINE:
            ce += 1
        elif cmd[0] == CLOSE:
            cs = ce + 1
            ce = cs
            closed = True
        else:
            raise ValueError("stroke: path is not piecewise-linear")
    return out


class Rasterizer:
    def __init__(self):
        self.incs = []  # (x, y, area, height)
        self.tincs = []  # (tile_x, tile_y, sign)
        self.lines = []
        self.first = V(0, 0)
        self.last = V(0, 0)
        self.tile_y_prev = 0

    def move_to(self, p: V):
        if self.last != self.first:
            self.line_to(self.first)
        self.first = p
        self.last = p
        self.tile_y_prev = _i16(np.floor(p.y)) >> 3

    def line_to(self, point: V):
        last = self.last
        if point != last:
            self.lines.append((last.t(), point.t()))
            with np.errstate(all="ignore"):
                x_dir = _i16(_signum(point.x - last.x))
                y_dir = _i16(_signum(point.y - last.y))
                dtdx = ONE / (point.x - last.x)
                dtdy = ONE / (point.y - last.y)
                x = _i16(np.floor(last.x))
                y = _i16(np.floor(last.y))
                row_t0 = ZERO
                col_t0 = ZERO
                if last.y == point.y:
                    row_t1 = INF
                else:
                    next_y = F(_wrap16(y + 1)) if point.y > last.y else F(y)
                    row_t1 = _min(dtdy * (next_y - last.y), ONE)
                if last.x == point.x:
                    col_t1 = INF
                else:
                    next_x = F(_wrap16(x + 1)) if point.x > last.x else F(x)
                    col_t1 = _min(dtdx * (next_x - last.x), ONE)
                x_step = abs(dtdx)
                y_step = abs(dtdy)
                while True:
                    t0 = _max(row_t0, col_t0)
                    t1 = _min(row_t1, col_t1)
                    p0 = (ONE - t0) * last + t0 * point
                    p1 = (ONE - t1) * last + t1 * point
                    height = p1.y - p0.y
                    right = F(_wrap16(x + 1))
                    area = F(0.5) * height * ((right - p0.x) + (right - p1.x))
                    self.incs.append((x, y, area, height))
                    if row_t1 < col_t1:
                        row_t0 = row_t1
                        row_t1 = _min(row_t1 + y_step, ONE)
                        y = _wrap16(y + y_dir)
                    else:
                        col_t0 = col_t1
                        col_t1 = _min(col_t1 + x_step, ONE)
                        x = _wrap16(x + x_dir)
                    done = bool(row_t0 == ONE or col_t0 == ONE)
                    if done:
                        x = _i16(np.floor(point.x))
                        y = _i16(np.floor(point.y))
                    tile_y = y >> 3
                    if tile_y != self.tile_y_prev:
                        d = tile_y - self.tile_y_prev
                        sign = ((d + 128) & 0xFF) - 128
                        self.tincs.append((x >> 3, min(self.tile_y_prev, tile_y), sign))
                        self.tile_y_prev = tile_y
                    if done:
                        break
        self.last = point

    def command(self, cmd):
        def emit(c):
            if c[0] == MOVE:
                self.move_to(c[1])
            elif c[0] == LINE:
                self.line_to(c[1])

        flatten_cmd(cmd, self.last, TOL, emit)

    def fill(self, path, xf: Xf):
        for cmd in path:
            self.command(transform_cmd(cmd, xf))

    def stroke(self, path, width, xf: Xf):
        self.fill(stroke_path(flatten_path(path, TOL), width), xf)

    def finish(self):
        """Returns the call sequence: ('tile', x, y, bytes64) / ('span', x, y, w)."""
        if self.last != self.first:
            self.line_to(self.first)
        incs = self.incs
        bins = []
        cur = [0, 0, 0, 0]  # tile_x, tile_y, start, end
        if incs:
            cur[0] = incs[0][0] >> 3
            cur[1] = incs[0][1] >> 3
        for i, (x, y, _, _) in enumerate(incs):
            tx, ty = x >> 3, y >> 3
            if tx != cur[0] or ty != cur[1]:
                bins.append(tuple(cur))
                cur = [tx, ty, i, i]
            cur[3] += 1
        bins.append(tuple(cur))
        bins.sort(key=lambda b: (b[1], b[0]))  # stable
        tincs = sorted(self.tincs, key=lambda t: (t[1], t[0]))

        calls = []
        areas = [ZERO] * 64
        heights = [ZERO] * 64
        prev = [ZERO] * 8
        ti = 0
        winding = 0
        nb = len(bins)
        for i, (tx, ty, s, e) in enumerate(bins):
            for (x, y, a, h) in incs[s:e]:
                k = (y & 7) * 8 + (x & 7)
                areas[k] = areas[k] + a
                heights[k] = heights[k] + h
            last_of_tile = i + 1 == nb or bins[i + 1][0] != tx or bins[i + 1][1] != ty
            if last_of_tile:
                tile = bytearray(64)
                nxt = [ZERO] * 8
                for y in range(8):
                    acc = prev[y]
                    for x in range(8):
                        tile[y * 8 + x] = _u8(_min(abs(acc + areas[y * 8 + x]) * F(256.0), F(255.0)))
                        acc = acc + heights[y * 8 + x]
                    nxt[y] = acc
                calls.append(("tile", _wrap16(tx * 8), _wrap16(ty * 8), bytes(tile)))
                areas = [ZERO] * 64
                heights = [ZERO] * 64
                same_row = i + 1 < nb and bins[i + 1][1] == ty
                prev = nxt if same_row else [ZERO] * 8
                if same_row and bins[i + 1][0] > tx + 1:
                    while ti < len(tincs):
                        t = tincs[ti]
                        if (t[1], t[0]) > (ty, tx):
                            break
                        winding += t[2]
                        ti += 1
                    if winding != 0:
                        w = bins[i + 1][0] - tx - 1
                        calls.append(("span", _wrap16((tx + 1) * 8), _wrap16(ty * 8), ((w & 0xFFFF) * 8) & 0xFFFF))
        return calls


def cmds_from_array(arr):
    """arr: iterable of (tag, v0..v5) rows -> tuple commands."""
    out = []
    for row in arr:
        tag = int(row[0])
        v = [F(x) for x in row[1:7]]
        if tag in (MOVE, LINE):
            out.append((tag, V(v[0], v[1])))
        elif tag == QUAD:
            out.append((tag, V(v[0], v[1]), V(v[2], v[3])))
        elif tag == CUBIC:
            out.append((tag, V(v[0], v[1]), V(v[2], v[3]), V(v[4], v[5])))
        elif tag == CONIC:
            out.append((tag, V(v[0], v[1]), V(v[2], v[3]), v[4]))
        else:
            out.append((CLOSE,))
    return out
