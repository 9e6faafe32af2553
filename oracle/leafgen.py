"""CPU restatement of the integrand's leaf values -- TEST INFRASTRUCTURE ONLY (same rules as oracle.py).

Reference: example/benchmark.jl:44-81 (the loop over leaves), :93-111 (`green_derive`, order 0 only: higher orders call
Lehmann.Spectral.kernelFermiT_dw*, a dependency that is not vendored), :113-127 (`green`), :70-73 (interaction);
src/frontend/pool.jl:69-78 (`update`: loops = K[:, 1:loopNum] * basis).  Vectorised over samples with numpy; the order of
the floating-point operations inside one sample is the one written in the reference (sum over loop momenta in index
order, then dot(kq, kq) over components in order).  `exp` is the C library's: the GPU agrees to ~1e-15 relative.
"""
from __future__ import annotations

import numpy as np

TAU_CUTOFF = 1e-10
EIGHT_PI = 8 * np.pi  # Julia: 8π == 8 * Float64(π)


def green(tau, w, beta):
    """benchmark.jl:113-127 on arrays."""
    tau = np.where(tau == 0.0, -TAU_CUTOFF, tau)
    with np.errstate(over="ignore", invalid="ignore"):
        pos_tau = np.where(w > 0.0, np.exp(-w * tau) / (1 + np.exp(-w * beta)), np.exp(w * (beta - tau)) / (1 + np.exp(w * beta)))
        neg_tau = np.where(w > 0.0, -np.exp(-w * (tau + beta)) / (1 + np.exp(-w * beta)), -np.exp(-w * tau) / (1 + np.exp(w * beta)))
    return np.where(tau > 0.0, pos_tau, neg_tau)


def green_derive(tau, w, beta, order: int):
    """example/benchmark.jl:93-111: (-1)^n / n! times the n-th omega-derivative of the fermionic kernel (`green` is n = 0).

    The reference takes the derivatives from Lehmann.Spectral.kernelFermiT_dw^n, a dependency that is not vendored, so
    its BITS are unpinned.  The function itself is unambiguous -- d^n/dw^n of `green(tau, w, beta)` -- and is evaluated
    here in closed form: green = exp(L(w)) with L' = D = -tau' + beta n_F(w), so the n-th derivative is green times the
    complete Bell polynomial Y_n(D, D', ..., D^(n-1)), and the derivatives of the Fermi function n_F are polynomials in
    n_F.  Checked against 50-digit numerical differentiation (mpmath) in tests/test_leafgen.py."""
    if order == 0:
        return green(tau, w, beta)
    if order > 5:
        raise NotImplementedError("not implemented!")  # benchmark.jl:107
    tau = np.where(tau == 0.0, -TAU_CUTOFF, tau)
    g0 = green(tau, w, beta)
    tp = np.where(tau > 0.0, tau, tau + beta)  # green(tau < 0) = -green(tau + beta): same w-dependence with tau' = tau + beta
    with np.errstate(over="ignore", invalid="ignore"):
        n = np.where(w > 0.0, np.exp(-w * beta) / (1 + np.exp(-w * beta)), 1 / (1 + np.exp(w * beta)))  # Fermi function
    m = n * (1 - n)
    D = -tp + beta * n
    D1 = -beta ** 2 * m
    D2 = beta ** 3 * m * (1 - 2 * n)
    D3 = -beta ** 4 * m * (1 - 6 * n + 6 * n * n)
    D4 = beta ** 5 * m * (1 - 14 * n + 36 * n * n - 24 * n * n * n)
    P2 = D * D
    P3 = P2 * D
    P4 = P3 * D
    P5 = P4 * D
    Y = [None, D, P2 + D1, P3 + 3 * D * D1 + D2, P4 + 6 * P2 * D1 + 4 * D * D2 + 3 * D1 * D1 + D3,
         P5 + 10 * P3 * D1 + 10 * P2 * D2 + 15 * D * D1 * D1 + 5 * D * D3 + 10 * D1 * D2 + D4][order]
    coef = {1: -1.0, 2: 1.0 / 2.0, 3: -1.0 / 6.0, 4: 1.0 / 24.0, 5: -1.0 / 120.0}[order]
    return coef * (g0 * Y)


def _pow_int(x, n: int):
    """Julia ^(x::Float64, n::Integer) for the small orders that occur (0..3 exact forms; larger n by repeated squaring
    with the same products as x*x*... is NOT Julia's compensated pow_body -- only used up to 3 here)."""
    if n == 0:
        return np.ones_like(x)
    if n == 1:
        return x
    if n == 2:
        return x * x
    if n == 3:
        return x * x * x
    raise NotImplementedError("interaction derivative orders above 3 are not restated")


def leaf_values(meta: dict, K: np.ndarray, T: np.ndarray, kF: float, beta: float, lam: float) -> np.ndarray:
    """K: (dim, n_loops, B), T: (n_tau, B) -> leaf (L, B)."""
    L = len(meta["leaf_type"])
    dim, n_loops, B = K.shape
    out = np.ones((L, B))
    for l in range(L):
        t = int(meta["leaf_type"][l])
        if t == 0:
            continue
        basis = meta["loop_basis"][int(meta["loop_index"][l])]
        q2 = np.zeros(B)
        for c in range(dim):
            kq = np.zeros(B)
            for j in range(n_loops):
                kq = kq + K[c, j] * basis[j]
            q2 = q2 + kq * kq
        if t == 1:
            tau = T[int(meta["tau_out"][l])] - T[int(meta["tau_in"][l])]
            out[l] = green_derive(tau, q2 - kF * kF, beta, int(meta["leaf_order"][l][0]))
        else:
            invK = 1.0 / (q2 + lam)
            out[l] = EIGHT_PI / invK * _pow_int(lam * invK, int(meta["leaf_order"][l][1]))
    return out
