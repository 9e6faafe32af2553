"""The source-text emitters (Compilers.to_julia_str / to_Cstr / to_python_str, compile_Julia / _C / _Python):
text identical to what the reference writes, and -- compiled / executed -- the same bits as the oracle."""
import os

import numpy as np
import pytest

import fdgraph_b200 as fd
import graphgen
from fdgraph_b200 import emitters
from oracle import emit_c
from oracle import oracle as O


def _example():
    fd.uidreset()
    v1, v2 = fd.FeynmanGraph([]), fd.FeynmanGraph([])
    return fd.FeynmanGraph([v1, v2], factor=1.5), v1, v2


def test_text_of_the_reference_example():
    # the graph of test/compiler.jl:4-15, formats of static.jl:100-131, :159-196 and compiler_python.jl:23-51
    g, v1, v2 = _example()
    js, leafmap = fd.Compilers.to_julia_str([g], name="eval_graph!")
    assert js == ("\nfunction eval_graph!(root::AbstractVector, leafVal::AbstractVector)\n"
                  "    g1 = leafVal[1]\n    g2 = leafVal[2]\n    g3 = (g1 + g2)\n    g4 = (g3 * 1.5)\n    root[1] = g4\nend")
    assert leafmap == {1: v1, 2: v2}
    cs, leafmap_c = fd.Compilers.to_Cstr([g])
    assert cs == ("\nvoid eval_graph(double *root, double *leafVal)\n{\n    double  g1, g2, g3, g4;\n"
                  "    g1 = leafVal[0];\n    g2 = leafVal[1];\n    g3 = (g1 + g2);\n    g4 = (g3 * 1.5);\n    root[0] = g4;\n}")
    assert leafmap_c == {1: v1, 2: v2}
    ps, _ = fd.Compilers.to_python_str([g])
    assert ps == ("import torch\ndef eval_graph(leafVal):\n"
                  "    root = torch.empty(leafVal.shape[0], 1, dtype=leafVal.dtype, device=leafVal.device)\n"
                  "    g1 = leafVal[:, 0]\n    g2 = leafVal[:, 1]\n    g3 = (g1 + g2)\n    g4 = (g3 * 1.5)\n    root[:, 0] = g4\n    return root\n\n")
    assert fd.Compilers.to_python_str([g], in_place=True)[0].startswith("def eval_graph(root, leafVal):\n    g1 = ")
    assert fd.Compilers.julia_to_C_typestr("ComplexF64") == "complex double "
    with pytest.raises(TypeError):
        fd.Compilers.julia_to_C_typestr("BigFloat")  # static.jl:151 error("Unsupported type")


def test_operators_factors_and_root_placement():
    fd.uidreset()
    a, b = fd.Graph([]), fd.Graph([])
    s = fd.Graph([a, b], operator=fd.Sum(), subgraph_factors=[1.0, -0.5])
    p = fd.Graph([s, a, s], operator=fd.Prod(), subgraph_factors=[2.0, 1.0, 1.0])
    q = fd.Graph([p], operator=fd.Power(3), subgraph_factors=[1.0e-5])
    js, _ = fd.Compilers.to_julia_str([q, s], root=[s.id, 77, q.id, s.id])
    assert "    g3 = (g1 + g2 * -0.5)\n    root[1] = g3\n" in js          # root[findfirst] right after the node
    assert "    g4 = (g3 * 2.0 * g1 * g3)\n" in js
    assert "    g5 = ((g4)^3 * 1.0e-5)\n    root[3] = g5\n" in js          # Julia prints 1.0e-5
    assert js.count("root[") == 2                                             # the second visit of s is skipped
    cs, _ = fd.Compilers.to_Cstr([q, s], root=[s.id, 77, q.id, s.id])
    assert "    g5 = pow(g4, 3) * 1.0e-5;\n    root[2] = g5;\n" in cs
    assert "((g4)**3 * 1.0e-5)" in fd.Compilers.to_python_str([q, s])[0]
    for x, want in [(1.5, "1.5"), (-1.0, "-1.0"), (1e-5, "1.0e-5"), (1e-4, "0.0001"), (1e6, "1.0e6"), (999999.0, "999999.0"),
                    (1234567.0, "1.234567e6"), (1 / 3, "0.3333333333333333"), (1e20, "1.0e20"), (0.0, "0.0"), (2.0, "2.0")]:
        assert emitters.julia_float(x) == want


def test_file_emitters_append(tmp_path):
    g, _, _ = _example()
    cfile = tmp_path / "f.c"
    fd.Compilers.compile_C([g], str(cfile))
    fd.Compilers.compile_C([g], str(cfile), func_name="second")
    text = cfile.read_text()
    assert text.startswith("#include <math.h>\n\nvoid eval_graph(") and text.count("#include") == 1 and "void second(" in text
    jfile = tmp_path / "f.jl"
    lm = fd.Compilers.compile_Julia([g], str(jfile))
    fd.Compilers.compile_Julia([g], str(jfile), func_name="again!")
    assert jfile.read_text().count("function ") == 2 and sorted(lm) == [1, 2]
    pfile = tmp_path / "f.py"
    fd.Compilers.compile_Python([g], str(pfile))
    assert pfile.read_text().startswith("import torch\ndef eval_graph(leafVal):")


@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_emitted_c_compiled_by_gcc_has_the_oracles_bits(seed, dtype):
    roots = graphgen.random_dag(40 + seed, n_leaves=9, n_inner=70, n_roots=3, p_power=0.25, max_pow=6)
    raw, nodes = fd.flatten(roots)
    em = emit_c.Emitted(raw, dtype)
    npdt = np.float64 if dtype == "f64" else np.complex128
    leaf = np.ascontiguousarray(graphgen.leaf_values(2, em.orc.n_leaves, 300, dtype=npdt, signed=True).T)
    assert em.eval(leaf).tobytes() == em.orc.eval(leaf, mode="emitter", layout="sample").tobytes()
    # leaf numbering of the text == leaf numbering of the lowering (the reference's leafmap)
    _, leafmap = fd.Compilers.to_Cstr(emit_c.graphs_from_raw(raw), root=[int(r) for r in raw.root_id])
    assert [leafmap[k + 1].id for k in range(em.orc.n_leaves)] == [int(raw.node_id[i]) for i in em.orc.leaf_nodes]


def test_emitted_python_runs_under_torch_with_the_same_bits():
    torch = pytest.importorskip("torch")
    roots = graphgen.random_dag(7, n_leaves=6, n_inner=40, n_roots=2, p_power=0.0)
    raw, _ = fd.flatten(roots)
    text, _ = fd.Compilers.to_python_str(roots)
    ns = {}
    exec(text, ns)
    orc = O.Oracle(raw)
    leaf = np.ascontiguousarray(graphgen.leaf_values(3, orc.n_leaves, 257, signed=True).T)
    got = ns["eval_graph"](torch.from_numpy(leaf)).numpy()
    assert got.tobytes() == orc.eval(leaf, mode="emitter", layout="sample").tobytes()


def test_real_workload_through_the_emitted_c():
    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", "parquet_ver4_o2.npz"))
    em = emit_c.Emitted(raw)
    leaf = 0.5 + np.random.default_rng(5).random((64, em.orc.n_leaves))
    assert em.eval(leaf, nthreads=2).tobytes() == em.orc.eval(leaf, mode="emitter", layout="sample").tobytes()
