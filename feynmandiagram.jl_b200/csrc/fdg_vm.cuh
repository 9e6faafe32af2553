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
//
// Dynamic shared memory:  [ slot file: n_slots x T values | per-warp program buffers: 2 x 32 packets |
//                           accumulate mode: per-warp running sums racc[warp][root * W] ]
// Program fetch: the 32 lanes of a warp fetch the NEXT chunk of 32 packets with one coalesced 16-byte load
// each while the current chunk executes out of shared memory (broadcast LDS.128 per packet), so the global
// latency of the instruction stream is hidden behind a whole chunk of work.
template <class V, bool ACC>
__global__ void __launch_bounds__(256) fdg_vm_kernel(const VmArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int S = V::kSamples;
    constexpr int W = V::kWidth;
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t stride = (uint32_t)T * (uint32_t)sizeof(V);
    unsigned char *const my = smem + (size_t)tid * sizeof(V);
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
    const long long gthread = (long long)blockIdx.x * T + tid;
    const long long gstride = (long long)gridDim.x * T;
    const uint32_t slot_file_bytes = (uint32_t)a.n_slots * stride;
    uint4 *const wbuf = reinterpret_cast<uint4 *>(smem + slot_file_bytes) + warp * (2 * FDG_CHUNK);
    const int RW = a.n_roots * W;
    double *racc = nullptr;
    if constexpr (ACC) {
        racc = reinterpret_cast<double *>(smem + slot_file_bytes + (size_t)(T >> 5) * 2 * FDG_CHUNK * sizeof(uint4)) +
               (size_t)warp * RW;
        for (int r = lane; r < RW; r += 32) racc[r] = 0.0;
        __syncwarp();
    }
    auto slot_ptr = [&](uint32_t s) -> Packed<V> * { return reinterpret_cast<Packed<V> *>(my + s * stride); };
    auto ld = [&](uint32_t s) -> V { return reinterpret_cast<const Packed<V> *>(my + s * stride)->v; };

    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const long long b0 = (tile * T + tid) * S;
        const bool active = b0 < a.batch;
        // inactive threads re-read a valid (aligned) sample so that every address stays in bounds
        long long bl = b0;
        if (!active) bl = S == 2 ? ((a.batch - 1) & ~1LL) : (a.batch - 1);
        const unsigned char *const leaf_b = static_cast<const unsigned char *>(a.leaf) + (size_t)bl * (8 * W);
        const size_t leaf_stride = (size_t)a.ld_leaf * (8 * W);

        V A, R1, R2, R3;
#pragma unroll
        for (int i = 0; i < S * W; ++i) A.x[i] = R1.x[i] = R2.x[i] = R3.x[i] = 0.0;

        // prime the program buffers: chunk 0 into buffer 0, chunk 1 in flight (the program is padded on upload)
        const uint4 *gp = a.prog;
        __syncwarp();
        wbuf[lane] = __ldg(gp + lane);
        uint4 nxt = __ldg(gp + FDG_CHUNK + lane);
        __syncwarp();
        int cur = 0;
        bool done = false;
        while (!done) {
            const uint4 *pb = wbuf + cur * FDG_CHUNK;
            uint4 pk = pb[0];
#pragma unroll 1
            for (int i = 0; i < FDG_CHUNK; ++i) {
                uint4 pn = pb[(i + 1) & (FDG_CHUNK - 1)];  // next packet (wraps harmlessly on the last one)
                const uint32_t hdr = pk.x;
                const uint32_t op = FDG_HDR_OP(hdr);
                const uint32_t k = FDG_HDR_K(hdr);
                if (hdr & (7u << 10)) cp_async_wait(FDG_HDR_WAIT(hdr) - 1u);
                if (op == FDG_OP_TERM) {
                    const double f = __hiloint2double((int)pk.w, (int)pk.z);
                    V t = ld(hdr >> 20);
                    if (k > 1) {
                        const V u = ld(pk.y & 0xffffu);
                        if (k > 2) {
                            const V w = ld(pk.y >> 16);
                            t = vmul(vmul(t, u), w);
                            if (k > 3) {
                                const uint4 e = pn;  // extension packet: s3..s10
                                ++i;
                                pn = pb[(i + 1) & (FDG_CHUNK - 1)];
                                t = vmul(t, ld(e.x & 0xffffu));
                                if (k > 4) {
                                    t = vmul(t, ld(e.x >> 16));
                                    if (k > 5) {
                                        t = vmul(t, ld(e.y & 0xffffu));
                                        if (k > 6) t = vmul(t, ld(e.y >> 16));
                                        if (k > 7) t = vmul(t, ld(e.z & 0xffffu));
                                        if (k > 8) t = vmul(t, ld(e.z >> 16));
                                        if (k > 9) t = vmul(t, ld(e.w & 0xffffu));
                                        if (k > 10) t = vmul(t, ld(e.w >> 16));
                                    }
                                }
                            }
                        } else {
                            t = vmul(t, u);
                        }
                    }
                    t = vscale(t, f);
                    if (hdr & (1u << 14)) {  // first term of a fold
                        if (hdr & (1u << 13)) {
                            R3 = R2;
                            R2 = R1;
                            R1 = A;
                        }
                        A = t;
                    } else {
                        A = vadd(A, t);
                    }
                } else if (op == FDG_OP_MUL) {
                    if (k == 1) {
                        A = vmul(A, ld(pk.y));
                    } else if (k == 2) {
                        const V v1 = ld(pk.y), v2 = ld(pk.z);
                        A = vmul(vmul(A, v1), v2);
                    } else {
                        const V v1 = ld(pk.y), v2 = ld(pk.z), v3 = ld(pk.w);
                        A = vmul(vmul(vmul(A, v1), v2), v3);
                    }
                } else if (op == FDG_OP_MOV) {
                    if (hdr & (1u << 13)) {
                        R3 = R2;
                        R2 = R1;
                        R1 = A;
                    }
                    if (k == 1) {
                        A = ld(pk.y);
                    } else if (k == 2) {
                        const V v1 = ld(pk.y), v2 = ld(pk.z);
                        A = vmul(v1, v2);
                    } else {
                        const V v1 = ld(pk.y), v2 = ld(pk.z), v3 = ld(pk.w);
                        A = vmul(vmul(v1, v2), v3);
                    }
                } else if (op == FDG_OP_LDL) {
                    const uint32_t w[3] = {pk.y, pk.z, pk.w};
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        if (q < (int)k) {
                            const uint32_t s = w[q] & (FDG_MAX_SLOTS - 1);
                            const uint32_t l = w[q] >> FDG_LDL_SLOT_BITS;
                            cp_async<sizeof(V)>(my_s + s * stride, leaf_b + (size_t)l * leaf_stride);
                        }
                    }
                    cp_async_commit();
                } else if (op == FDG_OP_RADDF) {
                    const double f = __hiloint2double((int)pk.w, (int)pk.z);
                    A = vadd(R1, vscale(A, f));
                    R1 = R2;
                    R2 = R3;
                } else if (op == FDG_OP_RMULF) {
                    const double f = __hiloint2double((int)pk.w, (int)pk.z);
                    A = vscale(vmul(R1, A), f);
                    R1 = R2;
                    R2 = R3;
                } else if (op == FDG_OP_ST) {
                    *slot_ptr(pk.y) = *reinterpret_cast<Packed<V> *>(&A);
                } else {
                    const double f = __hiloint2double((int)pk.w, (int)pk.z);
                    switch (op) {
                        case FDG_OP_END: done = true; break;
                        case FDG_OP_ADD:
                            if (k == 1) {
                                A = vadd(A, ld(pk.y));
                            } else if (k == 2) {
                                const V v1 = ld(pk.y), v2 = ld(pk.z);
                                A = vadd(vadd(A, v1), v2);
                            } else {
                                const V v1 = ld(pk.y), v2 = ld(pk.z), v3 = ld(pk.w);
                                A = vadd(vadd(vadd(A, v1), v2), v3);
                            }
                            break;
                        case FDG_OP_MULF: A = vscale(vmul(A, ld(pk.y)), f); break;
                        case FDG_OP_SCALE: A = vscale(A, f); break;
                        case FDG_OP_XADDF: A = vadd(ld(pk.y), vscale(A, f)); break;
                        case FDG_OP_XMULF: A = vscale(vmul(ld(pk.y), A), f); break;
                        case FDG_OP_POW: A = vpow(A, pk.y); break;
                        case FDG_OP_SPILL: {
                            V *g = static_cast<V *>(a.scratch) + ((size_t)pk.z * gstride + gthread);
                            *reinterpret_cast<Packed<V> *>(g) = *slot_ptr(pk.y);
                        } break;
                        case FDG_OP_FILL: {
                            const V *g = static_cast<const V *>(a.scratch) + ((size_t)pk.z * gstride + gthread);
                            *slot_ptr(pk.y) = *reinterpret_cast<const Packed<V> *>(g);
                        } break;
                        case FDG_OP_ROOT: {
                            const uint32_t r = pk.y;
                            if constexpr (!ACC) {
                                if (active) {
                                    double *o = static_cast<double *>(a.root) + ((size_t)r * a.ld_root + b0) * W;
                                    if constexpr (S == 2) {
                                        if (b0 + 1 < a.batch)
                                            *reinterpret_cast<double2 *>(o) = make_double2(A.x[0], A.x[1]);
                                        else
                                            o[0] = A.x[0];
                                    } else if constexpr (W == 2) {
                                        *reinterpret_cast<double2 *>(o) = make_double2(A.x[0], A.x[1]);
                                    } else {
                                        o[0] = A.x[0];
                                    }
                                }
                            } else {
                                // fixed-shape reduction: samples of the thread, then the xor tree over lanes
                                double s0 = active ? A.x[0] : 0.0, s1 = 0.0;
                                if constexpr (W == 2) s1 = active ? A.x[1] : 0.0;
                                if constexpr (S == 2) s0 = __dadd_rn(s0, (b0 + 1 < a.batch) ? A.x[1] : 0.0);
#pragma unroll
                                for (int m = 16; m >= 1; m >>= 1) {
                                    s0 = __dadd_rn(s0, __shfl_xor_sync(0xffffffffu, s0, m));
                                    if constexpr (W == 2) s1 = __dadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, m));
                                }
                                if (lane == 0) {
                                    racc[r * W] = __dadd_rn(racc[r * W], s0);
                                    if constexpr (W == 2) racc[r * W + 1] = __dadd_rn(racc[r * W + 1], s1);
                                }
                            }
                        } break;
                        default: break;  // NOP
                    }
                    if (done) break;
                }
                pk = pn;
            }
            if (done) break;
            // hand over to the prefetched chunk and start fetching the one after it
            cur ^= 1;
            wbuf[cur * FDG_CHUNK + lane] = nxt;
            gp += FDG_CHUNK;
            nxt = __ldg(gp + FDG_CHUNK + lane);
            __syncwarp();
        }
        cp_async_wait(0);
    }
    if constexpr (ACC) {
        __syncwarp();
        const long long warp_row = (long long)blockIdx.x * (T >> 5) + warp;
        for (int r = lane; r < RW; r += 32) a.partial[warp_row * RW + r] = racc[r];
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
