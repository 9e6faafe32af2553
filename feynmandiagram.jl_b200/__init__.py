"""fdgraph-b200: Blackwell-native back end for FeynmanDiagram.jl's computational-graph evaluator.

Only the hot path lives here: the Graph node model the evaluator consumes
(``graph``), its flattening (``program``), the ``Compilers.compile`` / ``eval_graph!`` mirror
(``compilers``) and the native library (``csrc`` -> libfdgraph.so, bound in ``_capi``).
"""
from . import graph, program  # noqa: F401
from .graph import (FeynmanGraph, Graph, Power, Prod, Sum, Unitary, constant_graph,  # noqa: F401
                    linear_combination, multi_product, uidreset)
from .program import RawGraph, flatten  # noqa: F401
from . import compilers as Compilers  # noqa: F401  (the reference's module name)
from .compilers import Evaluator, LeafGenerator, compile, compile_file, compile_raw  # noqa: F401
from .sharding import shard_range  # noqa: F401

__all__ = ["Graph", "FeynmanGraph", "Sum", "Prod", "Power", "Unitary", "constant_graph", "linear_combination",
           "multi_product", "uidreset", "RawGraph", "flatten", "Compilers", "Evaluator", "LeafGenerator", "compile", "compile_file", "compile_raw", "shard_range"]
