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

// ---- shared-memory access by 32-bit shared address ------------------------------------------------------
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
// A value of 4 samples is stored as two 16-byte halves `plane` bytes apart ([slot][half][thread]) so that each
// LDS.128 / STS.128 of a warp covers 512 contiguous bytes (conflict free); smaller values are contiguous.
template <class V>
__device__ __forceinline__ V lds_val(uint32_t addr, uint32_t plane) {
    V v;
    constexpr int N = sizeof(V) / 8;
    if constexpr (N == 1) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v.x[0]) : "r"(addr));
    } else {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x[0]), "=d"(v.x[1]) : "r"(addr));
        if constexpr (N == 4)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x[2]), "=d"(v.x[3]) : "r"(addr + plane));
    }
    return v;
}
template <class V>
__device__ __forceinline__ void sts_val(uint32_t addr, uint32_t plane, const V &v) {
    constexpr int N = sizeof(V) / 8;
    if constexpr (N == 1) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v.x[0]) : "memory");
    } else {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x[0]), "d"(v.x[1]) : "memory");
        if constexpr (N == 4)
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr + plane), "d"(v.x[2]), "d"(v.x[3]) : "memory");
    }
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}

// One block of `n` term records with K operands each (FDG_OP_TERM): the dispatch-free inner loop.
//   A = A + (v[s0] * v[s1] * ... * v[s(K-1)]) * f     for every record, in order.
template <class V, int K>
__device__ __forceinline__ void term_block(uint32_t &pp, uint32_t n, bool first, V &A, const uint32_t my_s, const uint32_t stride,
                                           const uint32_t plane) {
    constexpr uint32_t REC = K > 4 ? 32 : 16;
#pragma unroll 1
    for (uint32_t j = 0; j < n; ++j) {
        const uint4 r = lds_u4(pp);
        uint4 e = make_uint4(0, 0, 0, 0);
        if constexpr (K > 4) e = lds_u4(pp + 16);
        pp += REC;
        const uint32_t s[12] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16, e.x & 0xffffu, e.x >> 16,
                                e.y & 0xffffu, e.y >> 16, e.z & 0xffffu, e.z >> 16, e.w & 0xffffu, e.w >> 16};
        V v[K];
#pragma unroll
        for (int q = 0; q < K; ++q) v[q] = lds_val<V>(my_s + s[q] * stride, plane);
        V t = v[0];
#pragma unroll
        for (int q = 1; q < K; ++q) t = vmul(t, v[q]);
        t = vscale(t, __hiloint2double((int)r.w, (int)r.z));
        if (first) {
            A = t;
            first = false;
        } else {
            A = vadd(A, t);
        }
    }
}

// ---- the kernel --------------------------------------------------------------------------------------
// V: VReal<1>, VReal<2>, VReal<4> or VCplx.  ACC: false = write root per sample, true = per-warp running sums.
//
// Dynamic shared memory:  [ slot file: n_slots x T values | per-warp program buffers: 2 x 32 packets |
//                           accumulate mode: per-warp running sums racc[warp][root * W] ]
// Program fetch: the 32 lanes of a warp fetch the NEXT chunk of 32 packets with one coalesced 16-byte load
// each while the current chunk executes out of shared memory (broadcast LDS.128 per packet), so the global
// latency of the instruction stream is hidden behind a whole chunk of work.
template <class V, bool ACC>
__global__ void __launch_bounds__(256, 2) fdg_vm_kernel(const VmArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int S = V::kSamples;
    constexpr int W = V::kWidth;
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t stride = (uint32_t)T * (uint32_t)sizeof(V);
    const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem);
    // 4 samples/thread: [slot][half][thread] planes of T * 16 bytes; otherwise [slot][thread]
    const uint32_t plane = (uint32_t)T * 16u;
    const uint32_t my_s = smem_s + (uint32_t)tid * (sizeof(V) == 32 ? 16u : (uint32_t)sizeof(V));
    const long long gthread = (long long)blockIdx.x * T + tid;
    const long long gstride = (long long)gridDim.x * T;
    const uint32_t slot_file_bytes = (uint32_t)a.n_slots * stride;
    const uint32_t wbuf_s = smem_s + slot_file_bytes + (uint32_t)warp * (2 * FDG_CHUNK * 16);
    uint4 *const wbuf = reinterpret_cast<uint4 *>(smem + slot_file_bytes) + warp * (2 * FDG_CHUNK);
    const int RW = a.n_roots * W;
    double *racc = nullptr;
    if constexpr (ACC) {
        racc = reinterpret_cast<double *>(smem + slot_file_bytes + (size_t)(T >> 5) * 2 * FDG_CHUNK * sizeof(uint4)) +
               (size_t)warp * RW;
        for (int r = lane; r < RW; r += 32) racc[r] = 0.0;
        __syncwarp();
    }

    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const long long b0 = (tile * T + tid) * S;
        const bool active = b0 < a.batch;
        // inactive threads re-read a valid (aligned) sample group so that every address stays in bounds
        long long bl = b0;
        if (!active) bl = (a.batch - 1) & ~(long long)(S - 1);
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
        uint32_t cur = 0;
        uint32_t pp = wbuf_s;                    // shared address of the next packet
        uint32_t pend = wbuf_s + FDG_CHUNK * 16;  // end of the current chunk
        for (;;) {
            if (pp == pend) {
                // hand over to the prefetched chunk and start fetching the one after it
                cur ^= 1u;
                wbuf[cur * FDG_CHUNK + lane] = nxt;
                gp += FDG_CHUNK;
                nxt = __ldg(gp + FDG_CHUNK + lane);
                __syncwarp();
                pp = wbuf_s + cur * (FDG_CHUNK * 16);
                pend = pp + FDG_CHUNK * 16;
            }
            const uint4 pk = lds_u4(pp);
            pp += 16;
            const uint32_t hdr = pk.x;
            const uint32_t op = FDG_HDR_OP(hdr);
            const uint32_t k = FDG_HDR_K(hdr);
            if (hdr & (7u << 10)) cp_async_wait(FDG_HDR_WAIT(hdr) - 1u);
            if (hdr & (1u << 13)) {  // push: this packet starts a nested fold (rare: keep it a real branch)
                asm volatile("" ::: "memory");
                R3 = R2;
                R2 = R1;
                R1 = A;
            }
            if (op == FDG_OP_TERM) {
                const bool first = (hdr >> 14) & 1u;
                const uint32_t n = pk.y;
                switch (k) {
                    case 1: term_block<V, 1>(pp, n, first, A, my_s, stride, plane); break;
                    case 2: term_block<V, 2>(pp, n, first, A, my_s, stride, plane); break;
                    case 3: term_block<V, 3>(pp, n, first, A, my_s, stride, plane); break;
                    case 4: term_block<V, 4>(pp, n, first, A, my_s, stride, plane); break;
                    case 5: term_block<V, 5>(pp, n, first, A, my_s, stride, plane); break;
                    case 6: term_block<V, 6>(pp, n, first, A, my_s, stride, plane); break;
                    case 7: term_block<V, 7>(pp, n, first, A, my_s, stride, plane); break;
                    case 8: term_block<V, 8>(pp, n, first, A, my_s, stride, plane); break;
                    case 9: term_block<V, 9>(pp, n, first, A, my_s, stride, plane); break;
                    case 10: term_block<V, 10>(pp, n, first, A, my_s, stride, plane); break;
                    default: term_block<V, 11>(pp, n, first, A, my_s, stride, plane); break;
                }
                continue;
            }
            const double f = __hiloint2double((int)pk.w, (int)pk.z);
            switch (op) {
                case FDG_OP_END: goto program_done;
                case FDG_OP_LDL: {
                    // the k load words sit right behind the header word (w1..w3, then the next packet)
                    uint32_t wp = pp - 12;
#pragma unroll 1
                    for (uint32_t q = 0; q < k; ++q, wp += 4) {
                        const uint32_t w = lds_u32(wp);
                        const uint32_t s = w & (FDG_MAX_SLOTS - 1);
                        const uint32_t l = w >> FDG_LDL_SLOT_BITS;
                        const unsigned char *src = leaf_b + (size_t)l * leaf_stride;
                        if constexpr (sizeof(V) == 32) {
                            cp_async<16>(my_s + s * stride, src);
                            cp_async<16>(my_s + s * stride + plane, src + 16);
                        } else {
                            cp_async<sizeof(V)>(my_s + s * stride, src);
                        }
                    }
                    if (k > 3) pp += 16;
                    cp_async_commit();
                } break;
                case FDG_OP_MOV:
                    if (k == 1) {
                        A = lds_val<V>(my_s + pk.y * stride, plane);
                    } else if (k == 2) {
                        const V v1 = lds_val<V>(my_s + pk.y * stride, plane), v2 = lds_val<V>(my_s + pk.z * stride, plane);
                        A = vmul(v1, v2);
                    } else {
                        const V v1 = lds_val<V>(my_s + pk.y * stride, plane), v2 = lds_val<V>(my_s + pk.z * stride, plane),
                                v3 = lds_val<V>(my_s + pk.w * stride, plane);
                        A = vmul(vmul(v1, v2), v3);
                    }
                    break;
                case FDG_OP_MUL:
                    if (k == 1) {
                        A = vmul(A, lds_val<V>(my_s + pk.y * stride, plane));
                    } else if (k == 2) {
                        const V v1 = lds_val<V>(my_s + pk.y * stride, plane), v2 = lds_val<V>(my_s + pk.z * stride, plane);
                        A = vmul(vmul(A, v1), v2);
                    } else {
                        const V v1 = lds_val<V>(my_s + pk.y * stride, plane), v2 = lds_val<V>(my_s + pk.z * stride, plane),
                                v3 = lds_val<V>(my_s + pk.w * stride, plane);
                        A = vmul(vmul(vmul(A, v1), v2), v3);
                    }
                    break;
                case FDG_OP_ADD:
                    if (k == 1) {
                        A = vadd(A, lds_val<V>(my_s + pk.y * stride, plane));
                    } else if (k == 2) {
                        const V v1 = lds_val<V>(my_s + pk.y * stride, plane), v2 = lds_val<V>(my_s + pk.z * stride, plane);
                        A = vadd(vadd(A, v1), v2);
                    } else {
                        const V v1 = lds_val<V>(my_s + pk.y * stride, plane), v2 = lds_val<V>(my_s + pk.z * stride, plane),
                                v3 = lds_val<V>(my_s + pk.w * stride, plane);
                        A = vadd(vadd(vadd(A, v1), v2), v3);
                    }
                    break;
                case FDG_OP_RADDF:
                    A = vadd(R1, vscale(A, f));
                    R1 = R2;
                    R2 = R3;
                    break;
                case FDG_OP_RMULF:
                    A = vscale(vmul(R1, A), f);
                    R1 = R2;
                    R2 = R3;
                    break;
                case FDG_OP_ST: sts_val<V>(my_s + pk.y * stride, plane, A); break;
                case FDG_OP_MULF: A = vscale(vmul(A, lds_val<V>(my_s + pk.y * stride, plane)), f); break;
                case FDG_OP_SCALE: A = vscale(A, f); break;
                case FDG_OP_XADDF: A = vadd(lds_val<V>(my_s + pk.y * stride, plane), vscale(A, f)); break;
                case FDG_OP_XMULF: A = vscale(vmul(lds_val<V>(my_s + pk.y * stride, plane), A), f); break;
                case FDG_OP_POW: A = vpow(A, pk.y); break;
                case FDG_OP_SPILL: {
                    V *g = static_cast<V *>(a.scratch) + ((size_t)pk.z * gstride + gthread);
                    const V v = lds_val<V>(my_s + pk.y * stride, plane);
#pragma unroll
                    for (int i = 0; i < S * W; ++i) g->x[i] = v.x[i];
                } break;
                case FDG_OP_FILL: {
                    const V *g = static_cast<const V *>(a.scratch) + ((size_t)pk.z * gstride + gthread);
                    V v;
#pragma unroll
                    for (int i = 0; i < S * W; ++i) v.x[i] = g->x[i];
                    sts_val<V>(my_s + pk.y * stride, plane, v);
                } break;
                case FDG_OP_ROOT: {
                    const uint32_t r = pk.y;
                    if constexpr (!ACC) {
                        if (active) {
                            double *o = static_cast<double *>(a.root) + ((size_t)r * a.ld_root + b0) * W;
                            if constexpr (W == 2) {
                                *reinterpret_cast<double2 *>(o) = make_double2(A.x[0], A.x[1]);
                            } else if (b0 + S <= a.batch) {
                                if constexpr (S == 1) o[0] = A.x[0];
                                if constexpr (S >= 2) *reinterpret_cast<double2 *>(o) = make_double2(A.x[0], A.x[1]);
                                if constexpr (S == 4) *reinterpret_cast<double2 *>(o + 2) = make_double2(A.x[2], A.x[3]);
                            } else {
#pragma unroll
                                for (int i = 0; i < S; ++i)
                                    if (b0 + i < a.batch) o[i] = A.x[i];
                            }
                        }
                    } else {
                        // fixed-shape reduction: samples of the thread in order, then the xor tree over lanes
                        double s0 = 0.0, s1 = 0.0;
                        if constexpr (W == 2) {
                            s0 = active ? A.x[0] : 0.0;
                            s1 = active ? A.x[1] : 0.0;
                        } else {
                            s0 = active ? A.x[0] : 0.0;
#pragma unroll
                            for (int i = 1; i < S; ++i) s0 = __dadd_rn(s0, (b0 + i < a.batch) ? A.x[i] : 0.0);
                        }
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
        }
    program_done:
        cp_async_wait(0);
    }
    if constexpr (ACC) {
        __syncwarp();
        const long long warp_row = (long long)blockIdx.x * (T >> 5) + warp;
        for (int r = lane; r < RW; r += 32) a.partial[warp_row * RW + r] = racc[r];
    }
}

// acc[r] += sum over rows of partial[row][r].  One block per column r; thread t adds rows t, t+256, ... in order, then a
// fixed binary tree over the 256 per-thread sums: the order of additions depends only on `rows` (deterministic).
__global__ void __launch_bounds__(256) fdg_reduce_partials(const double *__restrict__ partial, long long rows, int rw, double *acc) {
    __shared__ double sh[256];
    const int r = blockIdx.x;
    double s = 0.0;
    for (long long i = threadIdx.x; i < rows; i += 256) s = __dadd_rn(s, partial[i * rw + r]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int m = 128; m >= 1; m >>= 1) {
        if ((int)threadIdx.x < m) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + m]);
        __syncthreads();
    }
    if (threadIdx.x == 0) acc[r] = __dadd_rn(acc[r], sh[0]);
}

}  // namespace fdg
