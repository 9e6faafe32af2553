"""Random computational graphs for parity tests (test infrastructure)."""
from __future__ import annotations

import random
from typing import List

import numpy as np

import fdgraph_b200 as fd

FACTORS = [1.0, 1.0, 1.0, -1.0, 2.0, 0.5, -0.25, 3.0, 1.5, -1.0 / 3.0]


def random_dag(seed: int, n_leaves: int = 8, n_inner: int = 40, n_roots: int = 3, max_fan: int = 5,
               p_power: float = 0.08, max_pow: int = 3, deep: bool = False) -> List[fd.Graph]:
    """A random DAG over Sum / Prod / Power with shared sub-graphs and mixed factors.  Built with the raw
    Graph constructor (no algebraic merging), so unary Sum/Prod chains and duplicate operands occur."""
    rng = random.Random(seed)
    pool = [fd.Graph([]) for _ in range(n_leaves)]
    if rng.random() < 0.5:
        pool.append(fd.constant_graph())  # a Unitary leaf
    for _ in range(n_inner):
        r = rng.random()
        if r < p_power:
            c = rng.choice(pool)
            g = fd.Graph([c], operator=fd.Power(rng.randint(2, max_pow)), subgraph_factors=[rng.choice(FACTORS)])
        else:
            k = rng.randint(1, max_fan)
            if deep:  # bias towards recently created nodes: long dependency chains, deep nesting
                kids = [pool[max(0, len(pool) - 1 - int(rng.expovariate(0.6)))] for _ in range(k)]
            else:
                kids = [rng.choice(pool) for _ in range(k)]
            op = fd.Sum() if rng.random() < 0.5 else fd.Prod()
            g = fd.Graph(kids, operator=op, subgraph_factors=[rng.choice(FACTORS) for _ in kids])
        pool.append(g)
    inner = pool[-n_inner:]
    roots = [inner[-1]] + [rng.choice(inner) for _ in range(n_roots - 1)]
    return roots


def sum_of_products(seed: int, n_leaves: int, n_terms: int, term_len: int, n_roots: int = 1) -> List[fd.Graph]:
    """The dominant shape of diagram graphs: root = Sum_k f_k * Prod(leaves) (readfile.jl:393-406,564-584)."""
    rng = random.Random(seed)
    leaves = [fd.Graph([]) for _ in range(n_leaves)]
    roots = []
    for _ in range(n_roots):
        terms, facs = [], []
        for _ in range(n_terms):
            ls = rng.sample(leaves, min(term_len, n_leaves))
            terms.append(fd.Graph(ls, operator=fd.Prod()))
            facs.append(rng.choice([1.0, -1.0, 2.0, -2.0, 0.5, 4.0]))
        roots.append(fd.Graph(terms, operator=fd.Sum(), subgraph_factors=facs))
    return roots


def leaf_values(seed: int, n_leaves: int, batch: int, dtype=np.float64, signed: bool = False, ld: int = 0) -> np.ndarray:
    """(L, ld) array, values 0.5 + U[0,1) (optionally random sign) -- bounded away from 0."""
    rng = np.random.default_rng(seed)
    ld = max(ld, batch)

    def one():
        x = 0.5 + rng.random((n_leaves, ld))
        if signed:
            x *= rng.choice([-1.0, 1.0], size=x.shape)
        return x

    if np.dtype(dtype) == np.complex128:
        out = np.empty((n_leaves, ld), np.complex128)
        out.real = one()
        out.imag = one()
        return out
    return one()


def random_tree(seed: int, depth: int = 7, n_leaves: int = 6, max_fan: int = 3) -> List[fd.Graph]:
    """A random expression TREE (every inner node used once): nothing is materialised, everything nests, so
    the lowering needs the accumulator registers and -- beyond depth 4 -- spilled accumulators."""
    rng = random.Random(seed)
    leaves = [fd.Graph([]) for _ in range(n_leaves)]

    def build(d):
        if d == 0 or rng.random() < 0.15:
            return rng.choice(leaves)
        if rng.random() < 0.1:
            return fd.Graph([build(d - 1)], operator=fd.Power(rng.randint(2, 3)), subgraph_factors=[rng.choice(FACTORS)])
        k = rng.randint(2, max_fan)
        kids = [build(d - 1) for _ in range(k)]
        op = fd.Sum() if rng.random() < 0.6 else fd.Prod()
        return fd.Graph(kids, operator=op, subgraph_factors=[rng.choice(FACTORS) for _ in kids])

    return [build(depth), build(depth - 2)]
