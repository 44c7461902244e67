"""ochre_b200 -- B200-native drop-in for ochre's path rasteriser hot path.

Python mirror of the reference's flat `ochre::` namespace (src/lib.rs:47-53) for the path
`Rasterizer::fill` -> `finish`; the arithmetic runs in hand-written sm_100a CUDA kernels
behind the C ABI of include/ochre_b200.h.  There is no CPU fallback.
"""
from .geom import (CLOSE, CMD_DTYPE, CONIC, CUBIC, LINE, MOVE, QUADRATIC, SPAN_DTYPE, TILE_SIZE, Mat2x2, PathCmd, Transform,
                   Vec2, cmds_to_array, make_cmds)
from .api import BatchResult, Context, Rasterizer, TileBuilder, default_context, finish_batch, flatten, stroke_to_fill

__all__ = [
    "TILE_SIZE", "Vec2", "Mat2x2", "Transform", "PathCmd", "Rasterizer", "TileBuilder", "Context", "BatchResult",
    "finish_batch", "flatten", "stroke_to_fill", "default_context", "cmds_to_array", "make_cmds", "CMD_DTYPE", "SPAN_DTYPE",
    "MOVE", "LINE", "QUADRATIC", "CUBIC", "CONIC", "CLOSE",
]
