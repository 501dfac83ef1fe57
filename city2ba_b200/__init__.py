"""city2ba_b200 — B200 (sm_100a) implementation of city2ba's visibility / observation-generation
hot path and elementwise noise pass, behind the reference's library surface.

    from city2ba_b200 import generate, synthetic, noise, BAProblem, SnavelyCamera

All compute goes through libcity2ba_cuda.so (include/city2ba_cuda.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import C2BError, Context, MultiContext, context  # noqa: F401
from . import generate, synthetic, noise, baproblem  # noqa: F401
from .baproblem import BAProblem, SnavelyCamera  # noqa: F401
from .generate import MultiScene, Scene, VisGraph, visibility_graph, visibility_graph_multi  # noqa: F401
