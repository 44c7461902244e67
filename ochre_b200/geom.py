"""Host-side geometry and path types, mirroring the reference's flat `ochre::` namespace.

Reference: src/geom.rs (`Vec2` :5, `Mat2x2` :133, `Transform` :204-262) and
src/path.rs (`PathCmd` :5-12).  These are plain host PODs; the arithmetic that
matters for parity (Transform::apply on path points) runs on the device
(csrc/raster_core.cuh `xf_apply`).  Transform composition (`then`, `scale`,
`rotate`, ...) is evaluated here in float32, operation for operation as
geom.rs:145-200 and :236-257 write it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

TILE_SIZE = 8  # rasterizer.rs:4

F = np.float32

# tags of `OchreCmd` (include/ochre_b200.h); order = declaration order of PathCmd, path.rs:5-12
MOVE, LINE, QUADRATIC, CUBIC, CONIC, CLOSE = range(6)

#: numpy view of `OchreCmd` (28 bytes)
CMD_DTYPE = np.dtype([("tag", "<u4"), ("v", "<f4", (6,))])
#: numpy view of `OchreSpan` (8 bytes)
SPAN_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<u2"), ("pad", "<u2")])


@dataclass(frozen=True)
class Vec2:
    x: float
    y: float

    @staticmethod
    def new(x: float, y: float) -> "Vec2":
        return Vec2(float(F(x)), float(F(y)))


@dataclass(frozen=True)
class Mat2x2:
    """Row-major 2x2 (geom.rs:133)."""

    m: tuple

    @staticmethod
    def new(a, b, c, d) -> "Mat2x2":
        return Mat2x2((float(F(a)), float(F(b)), float(F(c)), float(F(d))))

    @staticmethod
    def id() -> "Mat2x2":
        return Mat2x2((1.0, 0.0, 0.0, 1.0))

    @staticmethod
    def scale(s) -> "Mat2x2":
        s = float(F(s))
        return Mat2x2((s, 0.0, 0.0, s))

    @staticmethod
    def rotate(angle) -> "Mat2x2":  # geom.rs:152-154
        c = float(F(math.cos(float(F(angle)))))
        s = float(F(math.sin(float(F(angle)))))
        return Mat2x2((c, s, -s, c))

    def mul(self, rhs: "Mat2x2") -> "Mat2x2":  # geom.rs:157-168
        a = [F(v) for v in self.m]
        b = [F(v) for v in rhs.m]
        return Mat2x2(
            (
                float(a[0] * b[0] + a[1] * b[2]),
                float(a[0] * b[1] + a[1] * b[3]),
                float(a[2] * b[0] + a[3] * b[2]),
                float(a[2] * b[1] + a[3] * b[3]),
            )
        )

    def mul_vec(self, v: Vec2) -> Vec2:  # geom.rs:170-178
        a = [F(x) for x in self.m]
        x, y = F(v.x), F(v.y)
        return Vec2(float(a[0] * x + a[1] * y), float(a[2] * x + a[3] * y))


@dataclass(frozen=True)
class Transform:
    matrix: Mat2x2
    offset: Vec2

    @staticmethod
    def new(matrix: Mat2x2, offset: Vec2) -> "Transform":
        return Transform(matrix, offset)

    @staticmethod
    def id() -> "Transform":
        return Transform(Mat2x2.id(), Vec2(0.0, 0.0))

    @staticmethod
    def translate(x, y) -> "Transform":
        return Transform(Mat2x2.id(), Vec2.new(x, y))

    @staticmethod
    def scale(s) -> "Transform":
        return Transform(Mat2x2.scale(s), Vec2(0.0, 0.0))

    @staticmethod
    def rotate(angle) -> "Transform":
        return Transform(Mat2x2.rotate(angle), Vec2(0.0, 0.0))

    def then(self, t: "Transform") -> "Transform":  # geom.rs:252-257
        mo = t.matrix.mul_vec(self.offset)
        return Transform(
            t.matrix.mul(self.matrix),
            Vec2(float(F(mo.x) + F(t.offset.x)), float(F(mo.y) + F(t.offset.y))),
        )

    def apply(self, v: Vec2) -> Vec2:  # geom.rs:260-262
        r = self.matrix.mul_vec(v)
        return Vec2(float(F(r.x) + F(self.offset.x)), float(F(r.y) + F(self.offset.y)))

    def as_row(self) -> np.ndarray:
        """`OchreTransform` layout: m[4], ox, oy."""
        return np.array([*self.matrix.m, self.offset.x, self.offset.y], dtype=np.float32)

    @staticmethod
    def from_row(row) -> "Transform":
        r = [float(v) for v in row]
        return Transform(Mat2x2.new(r[0], r[1], r[2], r[3]), Vec2.new(r[4], r[5]))


class PathCmd:
    """Constructors named as the reference's enum variants (path.rs:5-12).

    A command is stored as one `CMD_DTYPE` record.
    """

    __slots__ = ("tag", "v")

    def __init__(self, tag: int, v: Sequence[float] = ()):
        self.tag = tag
        vv = [0.0] * 6
        for i, x in enumerate(v):
            vv[i] = float(x)
        self.v = tuple(vv)

    @staticmethod
    def Move(p: Vec2) -> "PathCmd":
        return PathCmd(MOVE, (p.x, p.y))

    @staticmethod
    def Line(p: Vec2) -> "PathCmd":
        return PathCmd(LINE, (p.x, p.y))

    @staticmethod
    def Quadratic(c: Vec2, p: Vec2) -> "PathCmd":
        return PathCmd(QUADRATIC, (c.x, c.y, p.x, p.y))

    @staticmethod
    def Cubic(c1: Vec2, c2: Vec2, p: Vec2) -> "PathCmd":
        return PathCmd(CUBIC, (c1.x, c1.y, c2.x, c2.y, p.x, p.y))

    @staticmethod
    def Conic(c: Vec2, p: Vec2, weight: float) -> "PathCmd":
        return PathCmd(CONIC, (c.x, c.y, p.x, p.y, weight))

    Close: "PathCmd"  # set below

    def __repr__(self):
        names = ["Move", "Line", "Quadratic", "Cubic", "Conic", "Close"]
        return f"PathCmd.{names[self.tag]}{self.v}"


PathCmd.Close = PathCmd(CLOSE)


def cmds_to_array(path: Iterable) -> np.ndarray:
    """Sequence of PathCmd (or an existing CMD_DTYPE array) -> CMD_DTYPE array."""
    if isinstance(path, np.ndarray):
        if path.dtype != CMD_DTYPE:
            raise TypeError("expected an array of CMD_DTYPE")
        return np.ascontiguousarray(path)
    path = list(path)
    arr = np.zeros(len(path), dtype=CMD_DTYPE)
    for i, c in enumerate(path):
        arr[i]["tag"] = c.tag
        arr[i]["v"] = c.v
    return arr


def make_cmds(rows: Iterable[Sequence[float]]) -> np.ndarray:
    """rows of (tag, v0, ...) -> CMD_DTYPE array (missing values are 0)."""
    rows = list(rows)
    arr = np.zeros(len(rows), dtype=CMD_DTYPE)
    for i, r in enumerate(rows):
        arr[i]["tag"] = int(r[0])
        for j, x in enumerate(r[1:7]):
            arr[i]["v"][j] = x
    return arr


IDENTITY_ROW = np.array([1, 0, 0, 1, 0, 0], dtype=np.float32)
