// fdg_vm.cuh -- the sm_100a graph-evaluation kernel: a packet VM, one Monte-Carlo sample (or two)
// per thread.  Replaces the reference's generated straight-line function
//   eval_graph!(root, leafVal)            src/backend/static.jl:100,117,123,127,131
// and its batched torch form              src/backend/compiler_python.jl:23-49
// (one elementwise launch per node, all intermediates through HBM) with ONE launch that keeps every
// intermediate on chip:
//   * the leaf matrix is batch-major, so the 32 lanes of a warp read 32 (or 64) consecutive samples
//     of one leaf: coalesced 8/16-byte accesses, staged global -> shared with cp.async (no register
//     round trip, latency hidden by the prefetch distance chosen at lowering time);
//   * the partial folds of the nodes being evaluated live in four accumulator registers per sample;
//   * values used more than once (and staged leaves) live in a shared-memory slot file laid out
//     slot-major [slot][thread] so that a warp access is conflict free;
//   * control flow is warp-uniform: every thread executes the same packet stream (fdg_isa.h).
// Arithmetic uses __dmul_rn / __dadd_rn only (never contracted to FMA), in the reference's fold
// order, so results are bit-identical to the emitted Julia / C function.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fdg_isa.h"

namespace fdg {

struct VmArgs {
    const uint4 *prog;
    const void *leaf;
    long long ld_leaf;
    void *root;
    long long ld_root;
    long long batch;
    void *scratch;          // [n_scratch][gridDim.x * blockDim.x] values
    double *partial;        // accumulate mode: [gridDim.x * warps][R * width] per-warp sums
    long long n_tiles;
    int n_roots;
    int n_slots;
};

// ---- value types ---------------------------------------------------------------------------------
template <int S>
struct VReal {
    double x[S];
    static constexpr int kSamples = S;
    static constexpr int kWidth = 1;  // doubles per sample
};
struct VCplx {
    double x[2];  // re, im
    static constexpr int kSamples = 1;
    static constexpr int kWidth = 2;
};

template <int S>
__device__ __forceinline__ VReal<S> vmul(const VReal<S> &a, const VReal<S> &b) {
    VReal<S> r;
#pragma unroll
    for (int i = 0; i < S; ++i) r.x[i] = __dmul_rn(a.x[i], b.x[i]);
    return r;
}
template <int S>
__device__ __forceinline__ VReal<S> vadd(const VReal<S> &a, const VReal<S> &b) {
    VReal<S> r;
#pragma unroll
    for (int i = 0; i < S; ++i) r.x[i] = __dadd_rn(a.x[i], b.x[i]);
    return r;
}
template <int S>
__device__ __forceinline__ VReal<S> vscale(const VReal<S> &a, double f) {
    VReal<S> r;
#pragma unroll
    for (int i = 0; i < S; ++i) r.x[i] = __dmul_rn(a.x[i], f);
    return r;
}
// Julia: *(z::Complex, w::Complex) = Complex(re(z)re(w) - im(z)im(w), re(z)im(w) + im(z)re(w))
__device__ __forceinline__ VCplx vmul(const VCplx &a, const VCplx &b) {
    VCplx r;
    r.x[0] = __dsub_rn(__dmul_rn(a.x[0], b.x[0]), __dmul_rn(a.x[1], b.x[1]));
    r.x[1] = __dadd_rn(__dmul_rn(a.x[0], b.x[1]), __dmul_rn(a.x[1], b.x[0]));
    return r;
}
__device__ __forceinline__ VCplx vadd(const VCplx &a, const VCplx &b) {
    VCplx r;
    r.x[0] = __dadd_rn(a.x[0], b.x[0]);
    r.x[1] = __dadd_rn(a.x[1], b.x[1]);
    return r;
}
// Julia: *(z::Complex, x::Real) = Complex(re(z)*x, im(z)*x)
__device__ __forceinline__ VCplx vscale(const VCplx &a, double f) {
    VCplx r;
    r.x[0] = __dmul_rn(a.x[0], f);
    r.x[1] = __dmul_rn(a.x[1], f);
    return r;
}

// x^n for n >= 4, Float64: Julia >= 1.9 `pow_body` (base/math.jl), compensated square-and-multiply.
__device__ __forceinline__ double pow_body_f64(double x, unsigned n) {
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
    while (n > 1) {
        if (n & 1u) {
            const double err = __fma_rn(y, xnlo, __dmul_rn(x, ynlo));
            const double p = __dmul_rn(x, y);
            ynlo = __dadd_rn(__fma_rn(x, y, -p), err);
            y = p;
        }
        const double err = __dmul_rn(__dmul_rn(x, 2.0), xnlo);
        const double p = __dmul_rn(x, x);
        xnlo = __dadd_rn(__fma_rn(x, x, -p), err);
        x = p;
        n >>= 1;
    }
    const double err = __fma_rn(y, xnlo, __dmul_rn(x, ynlo));
    const bool fin = isfinite(x) && isfinite(err);
    return fin ? __fma_rn(x, y, err) : __dmul_rn(x, y);
}
template <int S>
__device__ __forceinline__ VReal<S> vpow(const VReal<S> &a, unsigned n) {
    if (n == 2) return vmul(a, a);
    if (n == 3) return vmul(vmul(a, a), a);  // literal_pow: x*x*x
    VReal<S> r;
#pragma unroll
    for (int i = 0; i < S; ++i) r.x[i] = pow_body_f64(a.x[i], n);
    return r;
}
// Complex: literal_pow for 2, 3; Base.power_by_squaring for n >= 4 (base/intfuncs.jl)
__device__ __forceinline__ VCplx vpow(const VCplx &a, unsigned n) {
    if (n == 2) return vmul(a, a);
    if (n == 3) return vmul(vmul(a, a), a);
    VCplx x = a;
    int t = __ffs(n);  // trailing_zeros(n) + 1
    n >>= t;
    while (--t > 0) x = vmul(x, x);
    VCplx y = x;
    while (n > 0) {
        t = __ffs(n);
        n >>= t;
        while (--t >= 0) x = vmul(x, x);
        y = vmul(y, x);
    }
    return y;
}

// ---- memory helpers --------------------------------------------------------------------------------
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t smem_addr, const void *g) {
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait(unsigned n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}

template <class V>
struct alignas(sizeof(V)) Packed {
    V v;
};

// ---- the kernel --------------------------------------------------------------------------------------
// V: VReal<1>, VReal<2> or VCplx.  ACC: false = write root per sample, true = per-warp running sums.
template <class V, bool ACC>
__global__ void __launch_bounds__(256) fdg_vm_kernel(const VmArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int S = V::kSamples;
    constexpr int W = V::kWidth;
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const uint32_t stride = (uint32_t)T * (uint32_t)sizeof(V);
    unsigned char *const my = smem + (size_t)tid * sizeof(V);
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
    const long long gthread = (long long)blockIdx.x * T + tid;
    const long long gstride = (long long)gridDim.x * T;
    double *racc = nullptr;
    const int RW = a.n_roots * W;
    // dynamic smem: [slot file: n_slots * T values | accumulate mode: per-warp running sums racc[warp][root * W]]
    const uint32_t slot_file_bytes = (uint32_t)a.n_slots * stride;
    if constexpr (ACC) {
        racc = reinterpret_cast<double *>(smem + slot_file_bytes) + (size_t)(tid >> 5) * RW;
        for (int r = (tid & 31); r < RW; r += 32) racc[r] = 0.0;
        __syncwarp();
    }
    auto slot_ptr = [&](uint32_t s) -> V * { return reinterpret_cast<V *>(my + s * stride); };
    auto ld = [&](uint32_t s) -> V { return reinterpret_cast<const Packed<V> *>(my + s * stride)->v; };

    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const long long b0 = (tile * T + tid) * S;
        const bool active = b0 < a.batch;
        // inactive threads re-read a valid (aligned) sample so that every address stays in bounds
        long long bl = b0;
        if (!active) bl = S == 2 ? ((a.batch - 1) & ~1LL) : (a.batch - 1);
        const unsigned char *const leaf_b = static_cast<const unsigned char *>(a.leaf) + (size_t)bl * (8 * W);
        const size_t leaf_stride = (size_t)a.ld_leaf * (8 * W);

        V acc0, acc1, acc2, acc3;
#pragma unroll
        for (int i = 0; i < S * W; ++i) acc0.x[i] = acc1.x[i] = acc2.x[i] = acc3.x[i] = 0.0;

        const uint4 *pc = a.prog;
        uint4 pk = __ldg(pc);
        for (;;) {
            const uint4 nx = __ldg(pc + 1);  // the program is padded with a trailing END packet
            ++pc;
            const uint32_t hdr = pk.x;
            const uint32_t op = FDG_HDR_OP(hdr);
            const uint32_t n = FDG_HDR_N(hdr);
            const uint32_t arg = FDG_HDR_ARG(hdr);
            const double f = __hiloint2double((int)pk.w, (int)pk.z);
            if (op == FDG_OP_END) break;
            switch (op) {
                case FDG_OP_LDL: {
                    const uint32_t w[3] = {pk.y, pk.z, pk.w};
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if (i < (int)n) {
                            const uint32_t s = w[i] & (FDG_MAX_SLOTS - 1);
                            const uint32_t l = w[i] >> FDG_LDL_SLOT_BITS;
                            cp_async<sizeof(V)>(my_s + s * stride, leaf_b + (size_t)l * leaf_stride);
                        }
                    }
                    cp_async_commit();
                } break;
                case FDG_OP_WAIT: cp_async_wait(arg); break;
                case FDG_OP_SPILL: {
                    V *g = static_cast<V *>(a.scratch) + ((size_t)arg * gstride + gthread);
                    *reinterpret_cast<Packed<V> *>(g) = *reinterpret_cast<const Packed<V> *>(slot_ptr(pk.y));
                } break;
                case FDG_OP_FILL: {
                    const V *g = static_cast<const V *>(a.scratch) + ((size_t)arg * gstride + gthread);
                    *reinterpret_cast<Packed<V> *>(slot_ptr(pk.y)) = *reinterpret_cast<const Packed<V> *>(g);
                } break;

#define FDG_CASE(BASE, D, A, P, ...)     \
    case FDG_REGOP(BASE, D): {           \
        V &A_ = A;                       \
        V &P_ = P;                       \
        (void)P_;                        \
        __VA_ARGS__                      \
    } break;
#define FDG_CASE4(BASE, ...)                    \
    FDG_CASE(BASE, 0, acc0, acc0, __VA_ARGS__)  \
    FDG_CASE(BASE, 1, acc1, acc0, __VA_ARGS__)  \
    FDG_CASE(BASE, 2, acc2, acc1, __VA_ARGS__)  \
    FDG_CASE(BASE, 3, acc3, acc2, __VA_ARGS__)

                    FDG_CASE4(FDG_R_MOV, {
                        if (n == 1) {
                            A_ = ld(pk.y);
                        } else if (n == 2) {
                            const V v1 = ld(pk.y), v2 = ld(pk.z);
                            A_ = vmul(v1, v2);
                        } else {
                            const V v1 = ld(pk.y), v2 = ld(pk.z), v3 = ld(pk.w);
                            A_ = vmul(vmul(v1, v2), v3);
                        }
                    })
                    FDG_CASE4(FDG_R_MUL, {
                        if (n == 1) {
                            A_ = vmul(A_, ld(pk.y));
                        } else if (n == 2) {
                            const V v1 = ld(pk.y), v2 = ld(pk.z);
                            A_ = vmul(vmul(A_, v1), v2);
                        } else {
                            const V v1 = ld(pk.y), v2 = ld(pk.z), v3 = ld(pk.w);
                            A_ = vmul(vmul(vmul(A_, v1), v2), v3);
                        }
                    })
                    FDG_CASE4(FDG_R_ADD, {
                        if (n == 1) {
                            A_ = vadd(A_, ld(pk.y));
                        } else if (n == 2) {
                            const V v1 = ld(pk.y), v2 = ld(pk.z);
                            A_ = vadd(vadd(A_, v1), v2);
                        } else {
                            const V v1 = ld(pk.y), v2 = ld(pk.z), v3 = ld(pk.w);
                            A_ = vadd(vadd(vadd(A_, v1), v2), v3);
                        }
                    })
                    FDG_CASE4(FDG_R_MOVF, { A_ = vscale(ld(pk.y), f); })
                    FDG_CASE4(FDG_R_MULF, { A_ = vscale(vmul(A_, ld(pk.y)), f); })
                    FDG_CASE4(FDG_R_ADDF, { A_ = vadd(A_, vscale(ld(pk.y), f)); })
                    FDG_CASE4(FDG_R_SCALE, { A_ = vscale(A_, f); })
                    FDG_CASE4(FDG_R_RADDF, { P_ = vadd(P_, vscale(A_, f)); })
                    FDG_CASE4(FDG_R_RMULF, { P_ = vscale(vmul(P_, A_), f); })
                    FDG_CASE4(FDG_R_XADDF, { A_ = vadd(ld(pk.y), vscale(A_, f)); })
                    FDG_CASE4(FDG_R_XMULF, { A_ = vscale(vmul(ld(pk.y), A_), f); })
                    FDG_CASE4(FDG_R_POW, { A_ = vpow(A_, arg); })
                    FDG_CASE4(FDG_R_ST, { *reinterpret_cast<Packed<V> *>(slot_ptr(arg)) = *reinterpret_cast<Packed<V> *>(&A_); })
                    FDG_CASE4(FDG_R_ROOT, {
                        if constexpr (!ACC) {
                            if (active) {
                                double *o = static_cast<double *>(a.root) + ((size_t)arg * a.ld_root + b0) * W;
                                if constexpr (S == 2) {
                                    if (b0 + 1 < a.batch)
                                        *reinterpret_cast<double2 *>(o) = make_double2(A_.x[0], A_.x[1]);
                                    else
                                        o[0] = A_.x[0];
                                } else if constexpr (W == 2) {
                                    *reinterpret_cast<double2 *>(o) = make_double2(A_.x[0], A_.x[1]);
                                } else {
                                    o[0] = A_.x[0];
                                }
                            }
                        } else {
                            // fixed-shape reduction: samples of the thread, then the xor tree over lanes
                            double s[W];
                            if constexpr (W == 2) {
                                s[0] = active ? A_.x[0] : 0.0;
                                s[1] = active ? A_.x[1] : 0.0;
                            } else {
                                s[0] = active ? A_.x[0] : 0.0;
                                if constexpr (S == 2) s[0] = __dadd_rn(s[0], (b0 + 1 < a.batch) ? A_.x[1] : 0.0);
                            }
#pragma unroll
                            for (int k = 0; k < W; ++k) {
#pragma unroll
                                for (int m = 16; m >= 1; m >>= 1) s[k] = __dadd_rn(s[k], __shfl_xor_sync(0xffffffffu, s[k], m));
                            }
                            if ((tid & 31) == 0) {
#pragma unroll
                                for (int k = 0; k < W; ++k) racc[arg * W + k] = __dadd_rn(racc[arg * W + k], s[k]);
                            }
                        }
                    })
#undef FDG_CASE4
#undef FDG_CASE
                default: break;
            }
            pk = nx;
        }
        cp_async_wait(0);
    }
    if constexpr (ACC) {
        __syncwarp();
        const long long warp_row = (long long)blockIdx.x * (T >> 5) + (tid >> 5);
        for (int r = (tid & 31); r < RW; r += 32) a.partial[warp_row * RW + r] = racc[r];
    }
}

// acc[r] += sum over rows of partial[row][r], rows added in index order (deterministic)
__global__ void fdg_reduce_partials(const double *__restrict__ partial, long long rows, int rw, double *acc) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rw) return;
    double s = 0.0;
    for (long long i = 0; i < rows; ++i) s = __dadd_rn(s, partial[i * rw + r]);
    acc[r] = __dadd_rn(acc[r], s);
}

}  // namespace fdg
