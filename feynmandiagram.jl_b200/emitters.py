"""Source-text emitters of the reference's ``Compilers`` module, kept for users who export a graph to other tools.

    to_julia_str / to_Cstr / to_python_str          src/backend/static.jl:98-133, :155-197, compiler_python.jl:9-52
    compile_Julia / compile_C / compile_Python      static.jl:244-279, compiler_python.jl:53-60 (files are APPENDED to)

They produce the same text as the reference (statement order = post-order DFS with first visit of an id winning, leaf
numbering, `root[...] = g<ID>` right after the node, `* factor` only where the factor is not 1, Julia's printing of
Float64 literals) and the same ``leafmap``.  Nothing here evaluates anything: evaluation is ``compile`` ->
libfdgraph.so.  The emitted C text doubles as an independent check of the lowering (tests compile it with gcc and
compare it bit for bit with the oracle and the GPU).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from .graph import Graph, Power, Prod, Sum


def julia_float(x: float) -> str:
    """`string(x::Float64)` in Julia: shortest round-trip digits, fixed notation for 1e-5 < |x| < 1e6."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    r = repr(x)
    mant, _, exp = r.partition("e")
    sign = "-" if mant.startswith("-") else ""
    mant = mant.lstrip("-")
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if exp:
        e10 = int(exp) + len(ip) - 1
    elif ip.strip("0"):
        e10 = len(ip.lstrip("0")) - 1
    else:
        e10 = -(len(fp) - len(fp.lstrip("0")) + 1) if fp.strip("0") else 0
    digits = digits.rstrip("0") or "0"
    if x == 0:
        return sign + "0.0"
    if -5 < e10 < 6:
        if e10 >= 0:
            whole = digits[: e10 + 1].ljust(e10 + 1, "0")
            frac = digits[e10 + 1:] or "0"
        else:
            whole = "0"
            frac = "0" * (-e10 - 1) + digits
        return f"{sign}{whole}.{frac}"
    return f"{sign}{digits[0]}.{digits[1:] or '0'}e{e10}"


def _factor(f: float) -> str:
    return "" if f == 1 else f" * {julia_float(f)}"


def to_static(operator, subgraphs: Sequence[Graph], factors: Sequence[float], lang: str = "julia") -> str:
    """static.jl:13-80."""
    if isinstance(operator, (Sum, Prod)) and not isinstance(operator, Power):
        if len(subgraphs) == 1:
            return f"(g{subgraphs[0].id}{_factor(factors[0])})"
        terms = [f"g{g.id}{_factor(f)}" for g, f in zip(subgraphs, factors)]
        return "(" + (" + " if isinstance(operator, Sum) else " * ").join(terms) + ")"
    if isinstance(operator, Power):
        if lang == "c":
            return f"pow(g{subgraphs[0].id}, {operator.N}){_factor(factors[0])}"
        if lang not in ("julia", "python"):
            raise ValueError("Unsupported language")
        return f"((g{subgraphs[0].id}){'^' if lang == 'julia' else '**'}{operator.N}{_factor(factors[0])})"
    raise NotImplementedError(
        f"Static representation for computational graph nodes with operator {operator!r} not yet implemented!")


def _walk(graphs: Sequence[Graph]):
    """PostOrderDFS over every graph with the emitter's skip rules (static.jl:106-129): yields (node, is_leaf) for the
    first visit of every id; sub-trees of an id already emitted are not entered again (their nodes were all emitted)."""
    seen_leaf, seen_node, walked = set(), set(), set()
    for graph in graphs:
        stack = [(graph, 0)]
        while stack:
            node, i = stack.pop()
            if i == 0:
                if id(node) in walked:  # the same OBJECT again: everything below it has been emitted
                    continue
                walked.add(id(node))
            if i < len(node.subgraphs):
                stack.append((node, i + 1))
                stack.append((node.subgraphs[i], 0))
                continue
            if not node.subgraphs:
                if node.id in seen_leaf:
                    continue
                seen_leaf.add(node.id)
                yield node, True
            else:
                if node.id in seen_node:
                    continue
                seen_node.add(node.id)
                yield node, False


def _emit(graphs, root, leaf_fmt, stmt_fmt, root_fmt, lang, first_leaf, declare=None):
    graphs = list(graphs)
    root = [g.id for g in graphs] if root is None else list(root)
    pos = {}
    for k, r in enumerate(root):
        pos.setdefault(r, k)  # findfirst
    body: List[str] = []
    leafmap: Dict[int, Graph] = {}
    idx = first_leaf
    for g, is_leaf in _walk(graphs):
        if is_leaf:
            body.append(leaf_fmt.format(g=g.id, k=idx))
            leafmap[idx if first_leaf == 1 else idx + 1] = g
            idx += 1
        else:
            body.append(stmt_fmt.format(g=g.id, e=to_static(g.operator, g.subgraphs, g.subgraph_factors, lang)))
        if declare is not None:
            declare.append(f" g{g.id},")
        if g.id in pos:
            body.append(root_fmt.format(r=pos[g.id] + (1 if lang == "julia" else 0), g=g.id))
    return "".join(body), leafmap


def to_julia_str(graphs: Sequence[Graph], root: Optional[Sequence[int]] = None, name: str = "eval_graph!") -> Tuple[str, Dict[int, Graph]]:
    """static.jl:98-133.  leafmap keys are Julia's 1-based indices of leafVal."""
    body, leafmap = _emit(graphs, root, "    g{g} = leafVal[{k}]\n", "    g{g} = {e}\n", "    root[{r}] = g{g}\n", "julia", 1)
    return f"\nfunction {name}(root::AbstractVector, leafVal::AbstractVector)\n" + body + "end", leafmap


_C_TYPES = {"Float64": "double ", "Float32": "float ", "Int64": "long long ", "Int32": "int ", "ComplexF32": "complex float ",
            "ComplexF64": "complex double "}


def julia_to_C_typestr(datatype: str) -> str:
    """static.jl:135-153."""
    if datatype not in _C_TYPES:
        raise TypeError("Unsupported type")
    return _C_TYPES[datatype]


def to_Cstr(graphs: Sequence[Graph], root: Optional[Sequence[int]] = None, datatype: str = "Float64", name: str = "eval_graph") -> Tuple[str, Dict[int, Graph]]:
    """static.jl:155-197.  0-based leafVal / root indices in the text; leafmap keys stay 1-based like the reference's."""
    ctype = julia_to_C_typestr(datatype)
    declare: List[str] = []
    body, leafmap = _emit(graphs, root, "    g{g} = leafVal[{k}];\n", "    g{g} = {e};\n", "    root[{r}] = g{g};\n", "c", 0, declare)
    decl = ("    " + ctype + "".join(declare))[:-1] + ";\n"  # chop(declare) * ";\n"
    return f"\nvoid {name}({ctype}*root, {ctype}*leafVal)\n{{\n" + decl + body + "}", leafmap


def to_python_str(graphs: Sequence[Graph], root: Optional[Sequence[int]] = None, name: str = "eval_graph", in_place: bool = False) -> Tuple[str, Dict[int, Graph]]:
    """compiler_python.jl:9-52."""
    graphs = list(graphs)
    body, leafmap = _emit(graphs, root, "    g{g} = leafVal[:, {k}]\n", "    g{g} = {e}\n", "    root[:, {r}] = g{g}\n", "python", 0)
    if in_place:
        head = f"def {name}(root, leafVal):\n"
    else:
        head = ("import torch\n" + f"def {name}(leafVal):\n" +
                f"    root = torch.empty(leafVal.shape[0], {len(graphs)}, dtype=leafVal.dtype, device=leafVal.device)\n")
    return head + body + "    return root\n\n", leafmap


def compile_Julia(graphs, filename: str, root=None, func_name: str = "eval_graph!") -> Dict[int, Graph]:
    """static.jl:244-251 (append mode)."""
    text, leafmap = to_julia_str(graphs, root, func_name)
    with open(filename, "a") as fh:
        fh.write(text)
    return leafmap


def compile_C(graphs, filename: str, datatype: str = "Float64", root=None, func_name: str = "eval_graph") -> Dict[int, Graph]:
    """static.jl:269-279 (append mode; `#include <math.h>` when the file is new)."""
    text, leafmap = to_Cstr(graphs, root, datatype, func_name)
    with open(filename, "a") as fh:
        if fh.tell() == 0:
            fh.write("#include <math.h>\n")
        fh.write(text)
    return leafmap


def compile_Python(graphs, filename: str, root=None, func_name: str = "eval_graph") -> Dict[int, Graph]:
    """compiler_python.jl:53-60 (append mode)."""
    text, leafmap = to_python_str(graphs, root, func_name)
    with open(filename, "a") as fh:
        fh.write(text)
    return leafmap
