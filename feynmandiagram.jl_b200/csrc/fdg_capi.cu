// fdg_capi.cu -- C ABI of libfdgraph.so (include/fdgraph.h): handles, launches, host pipeline, NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/fdgraph.h"
#include "fdg_jit.h"
#include "fdg_lower.h"
#include "fdg_vm.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    // a missing driver / device is its own status: there is no CPU fallback behind this library
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError)
        return fail(FDG_ERR_NO_DEVICE, std::string(what) + ": " + cudaGetErrorString(e) +
                                           " (libfdgraph has no CPU fallback; a CUDA device is required)");
    return fail(FDG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
// nothing may be thrown across the C ABI: every entry point that can allocate runs inside this
template <class F>
int guarded(F &&f) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return fail(FDG_ERR_BAD_ARG, "out of memory");
    } catch (const std::exception &e) {
        return fail(FDG_ERR_BAD_ARG, std::string("internal error: ") + e.what());
    } catch (...) {
        return fail(FDG_ERR_BAD_ARG, "internal error");
    }
}
#define CUDA_TRY(x)                                    \
    do {                                               \
        cudaError_t e_ = (x);                          \
        if (e_ != cudaSuccess) return cuda_fail(e_, #x); \
    } while (0)

// Work buffers of one launch sequence.  They are kept PER STREAM: two sequences of the same handle issued on different
// streams (fdg_eval_host alternates two) may run at the same time and must not share them.
struct StreamScratch {
    void *scratch = nullptr;  // packet VM: spilled values
    size_t scratch_bytes = 0;
    double *partial = nullptr;  // accumulate mode: per-warp partial sums
    size_t partial_bytes = 0;
    void *cross = nullptr;  // specialised back end: values crossing kernel boundaries, [row][thread * spt]
    size_t cross_bytes = 0;
    uint32_t *progress = nullptr;  // pipeline form: stages passed by every tile of 32 samples
    size_t progress_bytes = 0;
    unsigned long long *pipe_stats = nullptr;  // pipeline form: [0] stall flag, then (busy, waiting) clocks per stage
    size_t pipe_stats_bytes = 0;
};

struct DeviceState {
    uint4 *d_prog = nullptr;
    std::map<cudaStream_t, StreamScratch> per_stream;
    int sm_count = 0;
    int max_smem_optin = 0;
    std::vector<unsigned> smids;  // pipeline form: the SM ids blocks land on (one block per SM), sorted; probed once
    std::vector<std::vector<unsigned>> groups;  // pipeline form: SMs that share an instruction cache (measured once per device)
    unsigned *d_smtab = nullptr;  // pipeline form: SM id -> (stage of the pass, block index within the stage, blocks of the stage, 0)
    int smtab_stages = 0;         // stages per pass the table on the device describes
    // host pipeline (fdg_eval_host)
    cudaStream_t streams[2] = {nullptr, nullptr};
    void *d_leaf[2] = {nullptr, nullptr};
    void *d_root[2] = {nullptr, nullptr};
    size_t d_leaf_bytes = 0, d_root_bytes = 0;
};

}  // namespace

struct JitVariant {
    fdg::JitPlan plan;
    bool compiled = false;
    std::map<int, std::vector<cudaKernel_t>> pipe_kernel;  // per device: the linked pipeline kernel of every pass
    std::vector<unsigned long long> last_stats;  // pipeline form: statistics of the last launch that was read back
    std::map<int, std::vector<cudaKernel_t>> kernels;  // per device
    std::map<int, cudaLibrary_t> libs_first;           // (libraries are kept alive with the handle)
    std::map<int, std::vector<cudaLibrary_t>> libs;
};

struct fdg_program {
    fdg::Lowered low;
    fdg::Lowered low_cse;  // fdg_options.cse == 0 (automatic): the same program with common sub-expressions merged ...
    bool has_cse = false;  // ... if that removes anything; the specialised back end plans both and keeps the faster plan
    int backend = FDG_BACKEND_AUTO;
    int jit_segment = 0;
    bool fma = false;  // fdg_options.fma: contraction allowed in the specialised kernels (opt-in, not bit-identical)
    std::map<int, JitVariant> jit;  // key = spt * 2 + accumulate
    int threads = 128;
    int spt = 0;  // samples per thread: 0 auto
    int blocks_per_sm = 0;
    std::map<int, DeviceState> dev;
    std::atomic<long long> launches{0};
    int last_jit_key = -1;  // the variant of the specialised back end the last launch ran (fdg_jit_info with samples_per_thread 0)
    std::mutex mu;
    std::mutex host_mu;  // fdg_eval_host: its staging buffers and streams belong to one caller at a time
};

struct fdg_comm {
    void *nccl_comm = nullptr;
};

namespace {

template <class V, bool ACC>
int launch_variant(fdg_program *h, DeviceState &ds, fdg::VmArgs &args, long long batch, cudaStream_t stream) {
    const fdg::Lowered &low = h->low;
    StreamScratch &ss = ds.per_stream[stream];
    constexpr int S = V::kSamples;
    constexpr int W = V::kWidth;
    auto kern = fdg::fdg_vm_kernel<V, ACC>;
    int T = h->threads;
    auto smem_for = [&](int t) -> size_t {
        size_t b = (size_t)low.n_slots * t * sizeof(V);
        b += (size_t)(t / 32) * 2 * FDG_CHUNK * 16;  // per-warp program buffers
        if (ACC) b += (size_t)(t / 32) * low.R * W * sizeof(double);
        return b;
    };
    while (T > 32 && smem_for(T) > (size_t)ds.max_smem_optin) T /= 2;
    const size_t smem = smem_for(T);
    if (smem > (size_t)ds.max_smem_optin)
        return fail(FDG_ERR_CAPACITY, "slot file does not fit shared memory: " + std::to_string(smem) + " bytes");
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    if (occ < 1) return fail(FDG_ERR_CAPACITY, "kernel cannot be resident with this slot file");
    if (h->blocks_per_sm > 0) occ = std::min(occ, h->blocks_per_sm);
    const long long per_tile = (long long)T * S;
    const long long n_tiles = (batch + per_tile - 1) / per_tile;
    const long long grid = std::min<long long>(n_tiles, (long long)ds.sm_count * occ);
    args.n_tiles = n_tiles;
    args.n_roots = (int)low.R;
    args.n_slots = low.n_slots;
    if (low.n_scratch > 0) {
        const size_t need = (size_t)low.n_scratch * grid * T * sizeof(V);
        if (need > ss.scratch_bytes) {
            if (ss.scratch) CUDA_TRY(cudaFree(ss.scratch));
            ss.scratch = nullptr;
            ss.scratch_bytes = 0;
            CUDA_TRY(cudaMalloc(&ss.scratch, need));
            ss.scratch_bytes = need;
        }
        args.scratch = ss.scratch;
    }
    long long rows = 0;
    if (ACC) {
        rows = grid * (T / 32);
        const size_t need = (size_t)rows * low.R * W * sizeof(double);
        if (need > ss.partial_bytes) {
            if (ss.partial) CUDA_TRY(cudaFree(ss.partial));
            ss.partial = nullptr;
            ss.partial_bytes = 0;
            CUDA_TRY(cudaMalloc((void **)&ss.partial, std::max<size_t>(need, 256)));
            ss.partial_bytes = std::max<size_t>(need, 256);
        }
        args.partial = ss.partial;
    }
    kern<<<(unsigned)grid, T, smem, stream>>>(args);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    if (ACC) {
        const int rw = (int)low.R * W;
        if (rw > 0) {
            fdg::fdg_reduce_partials<<<rw, 256, 0, stream>>>(ss.partial, rows, rw, static_cast<double *>(args.root));
            CUDA_TRY(cudaGetLastError());
            h->launches++;
        }
    }
    return FDG_OK;
}

int get_device_state(fdg_program *h, DeviceState **out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    auto it = h->dev.find(dev);
    if (it == h->dev.end()) {
        DeviceState ds;
        CUDA_TRY(cudaDeviceGetAttribute(&ds.sm_count, cudaDevAttrMultiProcessorCount, dev));
        CUDA_TRY(cudaDeviceGetAttribute(&ds.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        // the program, padded with END packets to whole chunks plus two (the kernel prefetches two chunks ahead)
        std::vector<uint32_t> w = h->low.words;
        const size_t chunk_words = 4 * FDG_CHUNK;
        w.resize(((w.size() + chunk_words - 1) / chunk_words + 2) * chunk_words, 0u);
        CUDA_TRY(cudaMalloc((void **)&ds.d_prog, w.size() * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpy(ds.d_prog, w.data(), w.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        it = h->dev.emplace(dev, ds).first;
    }
    *out = &it->second;
    return FDG_OK;
}

// ---- specialised back end ----------------------------------------------------------------------------------------
int jit_get(fdg_program *h, int spt, bool acc, JitVariant **out, bool wide = false, const std::vector<int> *groups = nullptr, bool bulk = false) {
    int key = spt * 2 + (acc ? 1 : 0) + (wide ? 64 : 0) + (bulk ? 32 : 0);
    if (groups) {
        unsigned hash = 2166136261u;
        for (const int g : *groups) hash = (hash ^ (unsigned)g) * 16777619u;
        key += 128 + 1024 * (int)(hash % 1000003u);
    }
    JitVariant &v = h->jit[key];
    if (!v.compiled && groups) {
        std::string err;
        fdg::PipeOptions po;
        po.groups = *groups;
        if (const char *e = getenv("FDG_PIPE_THREADS")) po.threads = std::max(32, std::min(256, atoi(e) / 32 * 32));
        int budget = h->jit_segment > 0 ? h->jit_segment : 4800;
        if (const char *e = getenv("FDG_PIPE_BUDGET")) budget = std::max(64, atoi(e));
        // profile-guided re-cut (experiments): "start0,start1,...,startS" and "w0,...,w(S-1)" of a previous plan
        auto parse = [](const char *txt, auto &out) {
            std::string t(txt);
            size_t pos = 0;
            while (pos < t.size()) {
                size_t q = t.find(',', pos);
                if (q == std::string::npos) q = t.size();
                out.push_back((typename std::remove_reference<decltype(out)>::type::value_type)atof(t.substr(pos, q - pos).c_str()));
                pos = q + 1;
            }
        };
        if (const char *e = getenv("FDG_PIPE_PREV")) parse(e, po.prev_start);
        if (const char *e = getenv("FDG_PIPE_MEASURED")) parse(e, po.measured);
        if (po.prev_start.size() != po.measured.size() + 1) po.prev_start.clear(), po.measured.clear();
        int rc = fdg::jit_plan(h->low, spt, acc, budget, wide, h->fma, v.plan, err, &po);
        if (rc == FDG_OK) rc = fdg::jit_compile(v.plan, err);
        if (rc != FDG_OK) {
            h->jit.erase(key);
            return fail(rc, err);
        }
        v.compiled = true;
    }
    if (!v.compiled) {
        std::string err;
        // The instruction budget of a kernel is an estimate; what counts is the machine code ptxas makes of it: a kernel
        // beyond the 128 KB instruction cache runs 20-50 % slower.  If the largest kernel of a
        // multi-kernel plan comes out above 120 KB the plan is redone with a proportionally smaller budget (at most twice).
        int budget = h->jit_segment > 0 ? h->jit_segment : 4000;
        if (bulk && h->jit_segment <= 0) {
            budget = 4400;  // the bulk form spends one instruction per input row where the ring spends five
            if (const char *e = getenv("FDG_JIT_BULK_BUDGET")) budget = std::max(64, atoi(e));
        }
        int rc = FDG_OK;
        fdg::jit_set_root_leaf_bias(0.0, 4);
        // automatic CSE: plan the program with and without merged sub-expressions (planning is cheap, assembling is not) and
        // keep the plan whose modelled time is lower: fewer operations against more values crossing kernel boundaries
        const fdg::Lowered *lowp = &h->low, *mergedp = nullptr;
        int scoped_reach = 0;  // 0: the default reach of scoped merging (two thirds of a kernel)
        // Last step, once the variant is chosen and assembled: the root order with a leaf bias (see order_roots) is tried on
        // paper for a few weights; a plan that moves at least 1.5 % fewer rows is assembled and kept if it does not spill more
        // (headline graph: 3186 -> 3072 rows per sample, 212 -> 217 M samples/s)
        auto refine_root_order = [&]() {
            const int64_t work = (h->low.muls_vv + h->low.muls_vf + h->low.adds_vv + h->low.pow_muls) * (h->low.dtype == FDG_C128 ? 4 : 1);
            if (h->low.R < 3 || v.plan.seg.size() < 2 || work < 2 * (int64_t)budget || getenv("FDG_JIT_ROOT_LEAF_WEIGHT")) return;
            auto rows = [](const fdg::JitPlan &p) { return (double)(p.leaf_loads + p.cross_loads + p.cross_stores) + (double)p.refetch_loads / 6.0; };
            const double ws[4] = {0.02, 0.03, 0.02, 0.1};
            const long wins[4] = {1, 1, 2, 8};
            std::vector<fdg::JitPlan> cands;
            const double limit = rows(v.plan) * 0.985;
            for (int i = 0; i < 4; ++i) {
                fdg::jit_set_root_leaf_bias(ws[i], wins[i]);
                fdg::JitPlan trial;
                std::string e0;
                if (fdg::jit_plan(*lowp, spt, acc, budget, wide, h->fma, trial, e0, nullptr, mergedp, scoped_reach, bulk) == FDG_OK &&
                    trial.seg.size() <= v.plan.seg.size() && rows(trial) < limit)
                    cands.push_back(std::move(trial));
            }
            fdg::jit_set_root_leaf_bias(0.0, 4);
            std::sort(cands.begin(), cands.end(), [&](const fdg::JitPlan &a, const fdg::JitPlan &b) { return rows(a) < rows(b); });
            for (size_t i = 0; i < cands.size() && i < 2; ++i) {  // (assembling is what costs time: the best two at most)
                std::string e1;
                if (fdg::jit_compile(cands[i], e1) == FDG_OK && cands[i].max_code_bytes <= 120 * 1024 &&
                    fdg::jit_spill_bytes(cands[i]) <= fdg::jit_spill_bytes(v.plan) + 512) {
                    cands[i].uses_cse = v.plan.uses_cse;
                    v.plan = std::move(cands[i]);
                    break;
                }
            }
        };
        if (h->has_cse) {
            // three ways to evaluate the same bits: as emitted; with equal sub-expressions merged everywhere (fewest
            // operations, but the shared values travel between kernels); merged only where the copies sit close together
            // ("scoped": fewer operations at about the traffic of the plain plan)
            fdg::JitPlan plain, full, scoped;
            std::string e1, e2, e3;
            const int es = h->low.dtype == FDG_C128 ? 16 : 8;
            int mode = -1;  // FDG_CSE_MODE: 0 plain, 1 merged, 2 scoped (experiments); default: the model decides
            if (const char *e = getenv("FDG_CSE_MODE")) mode = atoi(e);
            if (fdg::jit_plan(h->low, spt, acc, budget, wide, h->fma, plain, e1, nullptr, nullptr, 0, bulk) == FDG_OK &&
                fdg::jit_plan(h->low_cse, spt, acc, budget, wide, h->fma, full, e2, nullptr, nullptr, 0, bulk) == FDG_OK &&
                fdg::jit_plan(h->low, spt, acc, budget, wide, h->fma, scoped, e3, nullptr, &h->low_cse, 0, bulk) == FDG_OK) {
                const double tp = fdg::jit_model_ns(plain, es), tf = fdg::jit_model_ns(full, es), ts = fdg::jit_model_ns(scoped, es);
                fdg::JitPlan *cand = tf <= ts ? &full : &scoped;
                if (mode == 1) cand = &full;
                if (mode == 2) cand = &scoped;
                if (mode != 0 && (mode > 0 || std::min(tf, ts) < 0.95 * tp)) {
                    // at least 5 % faster on paper.  More shared values also mean more registers held: assemble both and let
                    // the spills ptxas reports have their say
                    if (fdg::jit_compile(plain, e1) == FDG_OK && fdg::jit_compile(*cand, e2) == FDG_OK) {
                        // A scoped plan that spills gets a second chance with a quarter of the reach: values merged over a
                        // shorter distance are held in registers for a shorter time (the headline graph: 5.7 KB of spill
                        // stores per sample at two thirds of a kernel, 1.9 KB at a quarter; 194 -> 209 M samples/s)
                        fdg::JitPlan near;
                        if (cand == &scoped && !getenv("FDG_CSE_SCOPE") && fdg::jit_spill_bytes(scoped) > 1024 &&
                            fdg::jit_plan(h->low, spt, acc, budget, wide, h->fma, near, e3, nullptr, &h->low_cse, std::max(64, budget / 4), bulk) == FDG_OK &&
                            fdg::jit_compile(near, e3) == FDG_OK && fdg::jit_model_ns(near, es) < fdg::jit_model_ns(scoped, es)) {
                            scoped = std::move(near);
                            scoped_reach = std::max(64, budget / 4);
                        }
                        const bool take = mode > 0 || fdg::jit_model_ns(*cand, es) < 0.95 * fdg::jit_model_ns(plain, es);
                        fdg::JitPlan &pick = take ? *cand : plain;
                        if (take && cand == &full) lowp = &h->low_cse;
                        if (take && cand == &scoped) mergedp = &h->low_cse;
                        if (pick.seg.size() < 2 || pick.max_code_bytes <= 120 * 1024) {  // assembled already and within the cache budget
                            v.plan = std::move(pick);
                            v.plan.uses_cse = take;
                            refine_root_order();
                            v.compiled = true;
                            *out = &v;
                            return FDG_OK;
                        }
                    }
                }
            }
        }
        for (int attempt = 0; attempt < 3; ++attempt) {
            rc = fdg::jit_plan(*lowp, spt, acc, budget, wide, h->fma, v.plan, err, nullptr, mergedp, scoped_reach, bulk);
            v.plan.uses_cse = lowp == &h->low_cse || mergedp != nullptr || h->low.cse_removed > 0;
            if (rc == FDG_OK) rc = fdg::jit_compile(v.plan, err);
            if (rc != FDG_OK || v.plan.seg.size() < 2 || v.plan.max_code_bytes <= 120 * 1024 || getenv("FDG_JIT_NO_REFIT")) break;
            const int smaller = (int)((double)budget * 112.0 * 1024.0 / (double)v.plan.max_code_bytes);
            if (smaller >= budget || smaller < 64) break;
            budget = smaller;
        }
        if (rc != FDG_OK) {
            h->jit.erase(key);
            return fail(rc, err);
        }
        refine_root_order();
        v.compiled = true;
    }
    *out = &v;
    return FDG_OK;
}

// FP64 rate of the device, measured: 8 independent chains per thread, either one DMUL and one DADD per step (what the
// bit-exact kernels may issue: nothing is contracted) or one DFMA.  The roofline's FP64 denominator (bench.py).
template <bool FMA>
__global__ void __launch_bounds__(256) fdg_fp64_probe(long long iters, double *sink) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double a = 1.0000000001, c = 1e-12;
    for (long long k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (FMA) {
                x[i] = __fma_rn(x[i], a, c);
                x[i] = __fma_rn(x[i], a, c);
            } else {
                x[i] = __dmul_rn(x[i], a);
                x[i] = __dadd_rn(x[i], c);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) sink[0] = s;  // never true: keeps the chains alive
}

// One block per SM (the launch asks for more than half of an SM's shared memory); every block reports the SM it runs on and
// stays until all have reported, so that no SM is counted twice.
__global__ void fdg_probe_smids(unsigned *out, unsigned *counter) {
    extern __shared__ char fdg_probe_smem[];
    if (threadIdx.x == 0) {
        unsigned s;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
        out[blockIdx.x] = s;
        __threadfence();
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        while (atomicAdd(counter, 0u) < gridDim.x && clock64() - t0 < 2000000000ll) __nanosleep(200);
    }
}

constexpr int FDG_SMTAB_ENTRIES = 4096;

int probe_smids(DeviceState &ds) {
    if (!ds.smids.empty()) return FDG_OK;
    unsigned *d = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d, ((size_t)ds.sm_count + 1) * 4));
    CUDA_TRY(cudaMemset(d, 0, ((size_t)ds.sm_count + 1) * 4));
    const size_t smem = (size_t)ds.max_smem_optin / 2 + 1024;
    CUDA_TRY(cudaFuncSetAttribute(fdg_probe_smids, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned *d_out = d, *d_cnt = d + ds.sm_count;
    void *args[] = {&d_out, &d_cnt};
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ds.sm_count);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelExC(&cfg, (const void *)fdg_probe_smids, args));
    std::vector<unsigned> ids((size_t)ds.sm_count);
    CUDA_TRY(cudaMemcpy(ids.data(), d, ids.size() * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaFree(d));
    std::sort(ids.begin(), ids.end());
    if (std::adjacent_find(ids.begin(), ids.end()) != ids.end() || ids.back() >= (unsigned)FDG_SMTAB_ENTRIES)
        return fail(FDG_ERR_CAPACITY, "could not place one block on every SM: the pipeline form is not available on this device");
    ds.smids = ids;
    return FDG_OK;
}

// Which SMs share an instruction cache.  Measured, because it depends on which TPCs of the die are enabled: SM i runs 96 KB
// of straight-line code alone, then together with SM j running other 96 KB; the SMs j that slow it down by more than 15 % sit
// behind the same 128 KB cache (a GPC: 12 to 20 SMs on the B200s seen so far, tools/icache_probe.py).  One-time cost ~0.5 s.
int probe_icache_groups(DeviceState &ds) {
    if (!ds.groups.empty()) return FDG_OK;
    if (const char *e = getenv("FDG_PIPE_GROUP_SIZE")) {  // experiments: consecutive SM ids in groups of this size
        const int gsz = std::max(1, atoi(e));
        for (size_t i = 0; i < ds.smids.size(); i += (size_t)gsz)
            ds.groups.emplace_back(ds.smids.begin() + (long)i, ds.smids.begin() + (long)std::min(ds.smids.size(), i + (size_t)gsz));
        return FDG_OK;
    }
    std::string ptx = ".version 8.7\n.target sm_100a\n.address_size 64\n\n";
    const int n_instr = 96 * 1024 / 16;
    for (int f = 0; f < 2; ++f) {
        ptx += ".func fdg_body" + std::to_string(f) + "(.param .b64 a_iters, .param .b64 a_out)\n{\n\t.reg .f64 %fd<12>;\n\t.reg .b64 %rd<4>;\n\t.reg .pred %p<2>;\n"
               "\tld.param.u64 %rd0, [a_iters];\n\tld.param.u64 %rd1, [a_out];\n";
        for (int i = 0; i < 8; ++i) ptx += "\tmov.f64 %fd" + std::to_string(i) + ", 0d3FF000000000" + std::to_string(f) + std::to_string(i) + "00;\n";
        ptx += "\tmov.f64 %fd8, 0d3FEFFFFF00000000;\n\tmov.f64 %fd9, 0d3F50624DD2F1A9FC;\n\tmov.f64 %fd10, 0d3FEFFFFE00000000;\nL" + std::to_string(f) + ":\n";
        for (int k = 0; k < n_instr; ++k) {
            const std::string x = "%fd" + std::to_string(k % 8);
            ptx += "\tfma.rn.f64 " + x + ", " + x + ", %fd" + ((k / 8 + f) % 3 ? "8" : "10") + ", %fd9;\n";
        }
        ptx += "\tsub.u64 %rd0, %rd0, 1;\n\tsetp.ne.u64 %p0, %rd0, 0;\n\t@%p0 bra L" + std::to_string(f) + ";\n";
        for (int i = 1; i < 8; ++i) ptx += "\tadd.rn.f64 %fd0, %fd0, %fd" + std::to_string(i) + ";\n";
        ptx += "\tst.global.f64 [%rd1], %fd0;\n\tret;\n}\n";
    }
    ptx += ".visible .entry fdg_icache_probe(.param .u64 p_tab, .param .u64 p_clk, .param .u64 p_sink, .param .u64 p_iters)\n.maxntid 256, 1, 1\n{\n"
           "\t.reg .b64 %rd<12>;\n\t.reg .b32 %r<6>;\n\t.reg .pred %p<3>;\n"
           "\tld.param.u64 %rd0, [p_tab];\n\tcvta.to.global.u64 %rd0, %rd0;\n\tld.param.u64 %rd1, [p_clk];\n\tcvta.to.global.u64 %rd1, %rd1;\n"
           "\tld.param.u64 %rd2, [p_sink];\n\tcvta.to.global.u64 %rd2, %rd2;\n\tld.param.u64 %rd3, [p_iters];\n"
           "\tmov.u32 %r0, %smid;\n\tmul.wide.u32 %rd4, %r0, 4;\n\tadd.u64 %rd5, %rd0, %rd4;\n\tld.global.u32 %r1, [%rd5];\n"
           "\tsetp.eq.u32 %p0, %r1, 0;\n\t@%p0 bra DONE;\n"
           "\tmov.u32 %r2, %tid.x;\n\tmul.wide.u32 %rd6, %r2, 8;\n\tmul.wide.u32 %rd7, %r0, 2048;\n\tadd.u64 %rd6, %rd6, %rd7;\n\tadd.u64 %rd6, %rd2, %rd6;\n"
           "\tbar.sync 0;\n\tmov.u64 %rd8, %clock64;\n\tsetp.eq.u32 %p1, %r1, 1;\n\t@%p1 bra C0;\n"
           "\t{\n\t.param .b64 q0;\n\t.param .b64 q1;\n\tst.param.b64 [q0], %rd3;\n\tst.param.b64 [q1], %rd6;\n\tcall.uni fdg_body1, (q0, q1);\n\t}\n\tbra FIN;\n"
           "C0:\n\t{\n\t.param .b64 q0;\n\t.param .b64 q1;\n\tst.param.b64 [q0], %rd3;\n\tst.param.b64 [q1], %rd6;\n\tcall.uni fdg_body0, (q0, q1);\n\t}\n"
           "FIN:\n\tbar.sync 0;\n\tmov.u64 %rd9, %clock64;\n\tsub.u64 %rd9, %rd9, %rd8;\n"
           "\tsetp.eq.u32 %p2, %r2, 0;\n\tmul.wide.u32 %rd4, %r0, 8;\n\tadd.u64 %rd10, %rd1, %rd4;\n\t@%p2 st.global.u64 [%rd10], %rd9;\n"
           "DONE:\n\tret;\n}\n";
    std::vector<char> cubin;
    std::string err;
    int rc = fdg::jit_assemble(ptx, 1, cubin, err);
    if (rc != FDG_OK) return fail(rc, err);
    cudaLibrary_t lib;
    CUDA_TRY(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaKernel_t k;
    CUDA_TRY(cudaLibraryGetKernel(&k, lib, "fdg_icache_probe"));
    const size_t smem = (size_t)ds.max_smem_optin / 2 + 1024;
    CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t NS = FDG_SMTAB_ENTRIES;
    unsigned *d_tab = nullptr;
    unsigned long long *d_clk = nullptr;
    double *d_sink = nullptr;
    CUDA_TRY(cudaMalloc((void **)&d_tab, NS * 4));
    CUDA_TRY(cudaMalloc((void **)&d_clk, NS * 8));
    CUDA_TRY(cudaMalloc((void **)&d_sink, NS * 256 * 8));
    std::vector<unsigned> tab(NS, 0);
    std::vector<unsigned long long> clk(NS, 0);
    long long iters = 6;
    auto run = [&](unsigned i, long j, unsigned long long *out) -> int {
        std::fill(tab.begin(), tab.end(), 0u);
        tab[i] = 1;
        if (j >= 0) tab[(size_t)j] = 2;
        CUDA_TRY(cudaMemcpy(d_tab, tab.data(), NS * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemset(d_clk, 0, NS * 8));
        void *args[] = {&d_tab, &d_clk, &d_sink, &iters};
        for (int rep = 0; rep < 2; ++rep)  // the first run warms the caches
            CUDA_TRY(cudaLaunchKernel((const void *)k, dim3((unsigned)ds.sm_count), dim3(256), args, smem, nullptr));
        CUDA_TRY(cudaMemcpy(out, d_clk + i, 8, cudaMemcpyDeviceToHost));
        return FDG_OK;
    };
    std::vector<unsigned> todo = ds.smids;
    std::vector<std::vector<unsigned>> groups;
    rc = FDG_OK;
    while (!todo.empty() && rc == FDG_OK) {
        const unsigned i = todo[0];
        unsigned long long alone = 0, t = 0;
        rc = run(i, -1, &alone);
        std::vector<unsigned> grp{i};
        for (size_t q = 1; q < todo.size() && rc == FDG_OK; ++q) {
            rc = run(i, (long)todo[q], &t);
            if ((double)t > 1.15 * (double)alone) grp.push_back(todo[q]);
        }
        std::vector<unsigned> rest;
        for (const unsigned x : todo)
            if (std::find(grp.begin(), grp.end(), x) == grp.end()) rest.push_back(x);
        todo.swap(rest);
        groups.push_back(grp);
        if (groups.size() > 64) break;  // no structure found (every SM on its own): not a device this form was made for
    }
    cudaFree(d_tab);
    cudaFree(d_clk);
    cudaFree(d_sink);
    cudaLibraryUnload(lib);
    if (rc != FDG_OK) return rc;
    if (groups.size() > 64) return fail(FDG_ERR_CAPACITY, "instruction-cache groups could not be measured on this device");
    ds.groups = groups;
    if (getenv("FDG_PIPE_TRACE")) {
        std::fprintf(stderr, "instruction-cache groups:");
        for (auto &g : groups) std::fprintf(stderr, " %zu", g.size());
        std::fprintf(stderr, "\n");
    }
    return FDG_OK;
}

// SM id -> stage table: stage j of a pass runs on the groups [j * G / gs, (j + 1) * G / gs)
int upload_smtab(DeviceState &ds, int stages_per_pass, cudaStream_t stream) {
    if (ds.d_smtab && ds.smtab_stages == stages_per_pass) return FDG_OK;
    std::vector<unsigned> tab((size_t)FDG_SMTAB_ENTRIES * 4, 0xffffffffu);
    const int G = (int)ds.groups.size(), gs = stages_per_pass;
    std::vector<std::vector<unsigned>> sm_of((size_t)gs);
    for (int g = 0; g < G; ++g)
        for (const unsigned s : ds.groups[(size_t)g]) sm_of[(size_t)((int64_t)g * gs / G)].push_back(s);
    for (int j = 0; j < gs; ++j)
        for (size_t q = 0; q < sm_of[(size_t)j].size(); ++q) {
            unsigned *e = &tab[(size_t)sm_of[(size_t)j][q] * 4];
            e[0] = (unsigned)j, e[1] = (unsigned)q, e[2] = (unsigned)sm_of[(size_t)j].size(), e[3] = 0;
        }
    if (!ds.d_smtab) CUDA_TRY(cudaMalloc((void **)&ds.d_smtab, tab.size() * 4));
    CUDA_TRY(cudaStreamSynchronize(stream));  // the table may still be read by a launch in flight on this stream
    CUDA_TRY(cudaMemcpy(ds.d_smtab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    ds.smtab_stages = stages_per_pass;
    return FDG_OK;
}

template <class P>
int grow(P *&ptr, size_t &have, size_t need) {
    if (need <= have) return FDG_OK;
    if (ptr) CUDA_TRY(cudaFree(ptr));
    ptr = nullptr;
    have = 0;
    CUDA_TRY(cudaMalloc((void **)&ptr, need));
    have = need;
    return FDG_OK;
}

// The pipeline form of the specialised back end (fdg_jit.h): per pass one cooperative launch, one block per SM; the blocks of
// an instruction-cache group run one stage.  Cross rows of a pass live in a ring of `window` tile slots and stay in L2; values
// that go from one pass to a later one travel through a [row][sample] buffer in HBM, which bounds a launch sequence.
int jit_launch_pipeline(fdg_program *h, DeviceState &ds, int dev, bool acc, const void *leaf, int64_t ld_leaf, void *root, int64_t ld_root,
                        int64_t batch, cudaStream_t stream) {
    JitVariant *v = nullptr;
    StreamScratch &ss = ds.per_stream[stream];
    const bool wide = (uint64_t)ld_leaf * (h->low.dtype == FDG_C128 ? 16 : 8) >= (1ull << 32);
    int rc = probe_smids(ds);
    if (rc != FDG_OK) return rc;
    rc = probe_icache_groups(ds);
    if (rc != FDG_OK) return rc;
    std::vector<int> gsz;
    for (auto &g : ds.groups) gsz.push_back((int)g.size());
    rc = jit_get(h, 1, acc, &v, wide, &gsz);
    if (rc != FDG_OK) return rc;
    const fdg::JitPlan &pl = v->plan;
    const fdg::Lowered &low = h->low;
    const int W = low.dtype == FDG_C128 ? 2 : 1;
    const size_t es = 8 * (size_t)W;
    int T = 256;
    if (const char *e = getenv("FDG_PIPE_THREADS")) T = std::max(32, std::min(256, atoi(e) / 32 * 32));
    // one block per SM: ask for more than half of the shared memory of an SM
    const size_t smem = std::max<size_t>((size_t)pl.ring_bytes, (size_t)ds.max_smem_optin / 2 + 1024);
    auto kit = v->pipe_kernel.find(dev);
    if (kit == v->pipe_kernel.end()) {
        std::vector<cudaKernel_t> ks;
        for (int ps = 0; ps < pl.n_pass; ++ps) {
            cudaLibrary_t lib;
            CUDA_TRY(cudaLibraryLoadData(&lib, pl.linked[(size_t)ps].data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
            v->libs[dev].push_back(lib);
            cudaKernel_t k;
            CUDA_TRY(cudaLibraryGetKernel(&k, lib, ("fdg_pipe" + std::to_string(ps)).c_str()));
            CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)k, T, smem));
            if (occ < 1) return fail(FDG_ERR_CAPACITY, "the pipeline kernel cannot be resident");
            ks.push_back(k);
        }
        kit = v->pipe_kernel.emplace(dev, ks).first;
    }
    rc = upload_smtab(ds, pl.stages_per_pass, stream);
    if (rc != FDG_OK) return rc;
    const int warps = T / 32;
    int64_t window = (int64_t)pl.n_sm * warps * 3 / 2;  // tiles in flight: every warp has one, half as many queue between the stages
    if (const char *e = getenv("FDG_PIPE_WINDOW")) window = std::max<int64_t>(atoll(e), (int64_t)pl.n_sm * warps);
    window = std::max<int64_t>(window, 1);
    const int64_t ld_cross = window * 32;
    if (pl.n_cross > 0) {
        rc = grow(ss.cross, ss.cross_bytes, (size_t)pl.n_cross * (size_t)ld_cross * es);
        if (rc != FDG_OK) return rc;
    }
    // launch sequences: the pass-boundary buffer holds n_boundary rows of `sub` samples (8 GiB at most, rows below 4 GiB)
    int64_t sub = (batch + 31) / 32 * 32;
    if (pl.n_boundary > 0) {
        double gb = 8.0;
        if (const char *e = getenv("FDG_JIT_CROSS_GB")) gb = atof(e);
        const int64_t cap = std::max<int64_t>(window * 32 * 4, (int64_t)(gb * (double)(1 << 30)) / ((int64_t)es * (int64_t)pl.n_boundary));
        sub = std::min<int64_t>(sub, cap / 32 * 32);
    }
    sub = std::min<int64_t>(sub, ((int64_t)((1ull << 32) / es) - 32) / 32 * 32);
    void *boundary = nullptr;
    if (pl.n_boundary > 0) {
        rc = grow(ss.scratch, ss.scratch_bytes, (size_t)pl.n_boundary * (size_t)sub * es);  // (the packet VM's spill buffer is free here)
        if (rc != FDG_OK) return rc;
        boundary = ss.scratch;
    }
    const int64_t tiles_sub = sub / 32;
    rc = grow(ss.progress, ss.progress_bytes, std::max<size_t>((size_t)tiles_sub * 4, 256));
    if (rc != FDG_OK) return rc;
    const size_t stats_bytes = 16 + 16 * pl.seg.size();
    rc = grow(ss.pipe_stats, ss.pipe_stats_bytes, stats_bytes);
    if (rc != FDG_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(ss.pipe_stats, 0, stats_bytes, stream));
    const long long rows = (long long)pl.n_sm * warps * pl.n_pass;  // every (pass, warp) has its own row of partial sums
    void *out = root;
    if (acc) {
        const size_t need = std::max<size_t>((size_t)rows * low.R * W * sizeof(double), 256);
        rc = grow(ss.partial, ss.partial_bytes, need);
        if (rc != FDG_OK) return rc;
        out = ss.partial;
    }
    for (int64_t b0 = 0; b0 < batch; b0 += sub) {
        const int64_t nb = std::min<int64_t>(sub, batch - b0);
        CUDA_TRY(cudaMemsetAsync(ss.progress, 0, (size_t)((nb + 31) / 32) * 4, stream));
        if (acc && b0 == 0) CUDA_TRY(cudaMemsetAsync(ss.partial, 0, (size_t)rows * low.R * W * sizeof(double), stream));
        for (int ps = 0; ps < pl.n_pass; ++ps) {
            const void *p_leaf = static_cast<const char *>(leaf) + (size_t)b0 * es;
            // accumulate: the partial rows of pass ps; a later launch sequence adds to the same rows through the running sums
            void *p_out = acc ? static_cast<void *>(static_cast<char *>(out) + (size_t)ps * pl.n_sm * warps * low.R * W * sizeof(double))
                              : static_cast<void *>(static_cast<char *>(root) + (size_t)b0 * es);
            void *p_cross = ss.cross, *p_progress = ss.progress, *p_stats = ss.pipe_stats, *p_smtab = ds.d_smtab, *p_boundary = boundary;
            long long a_ld_leaf = ld_leaf, a_ld_cross = ld_cross, a_ld_root = ld_root, a_batch = nb, a_nroots = low.R * W, a_ntiles = (nb + 31) / 32,
                      a_window = window, a_ld_boundary = sub;
            void *args[] = {(void *)&p_leaf, &a_ld_leaf,  &p_cross,  &a_ld_cross, &p_out,   &a_ld_root,  &a_batch,       &a_nroots,
                            &p_progress,     &a_ntiles,   &a_window, &p_stats,    &p_smtab, &p_boundary, &a_ld_boundary};
            cudaLaunchConfig_t cfg;
            std::memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)pl.n_sm);
            cfg.blockDim = dim3((unsigned)T);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeCooperative;  // all blocks resident at once, or the launch fails: the stages wait for each other
            attr[0].val.cooperative = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelExC(&cfg, (const void *)kit->second[(size_t)ps], args));
            h->launches++;
        }
        if (acc && low.R > 0 && (b0 + sub < batch)) {
            // more launch sequences follow: fold this one's partial rows into the result now (the rows are overwritten, not
            // added to, by the next sequence)
            fdg::fdg_reduce_partials<<<(int)low.R * W, 256, 0, stream>>>(ss.partial, rows, (int)low.R * W, static_cast<double *>(root));
            CUDA_TRY(cudaGetLastError());
            h->launches++;
            CUDA_TRY(cudaMemsetAsync(ss.partial, 0, (size_t)rows * low.R * W * sizeof(double), stream));
        }
    }
    if (acc && low.R > 0) {
        fdg::fdg_reduce_partials<<<(int)low.R * W, 256, 0, stream>>>(ss.partial, rows, (int)low.R * W, static_cast<double *>(root));
        CUDA_TRY(cudaGetLastError());
        h->launches++;
    }
    return FDG_OK;
}

// 1 when the pipeline form should run this call: big program, big batch (FDG_JIT_PIPE = 0 / 1 overrides)
bool pipeline_wanted(const fdg_program *h, int spt, int64_t batch) {
    const fdg::Lowered &low = h->low;
    const int64_t work = (low.muls_vv + low.adds_vv + low.muls_vf + low.pow_muls) * (low.dtype == FDG_C128 ? 4 : 1);
    const int budget = h->jit_segment > 0 ? h->jit_segment : 4000;
    bool want = false;  // default until measured: the classic launch sequence
    if (const char *e = getenv("FDG_JIT_PIPE")) want = atoi(e) != 0;
    int64_t min_batch = 1 << 18;
    if (const char *e = getenv("FDG_PIPE_MIN_BATCH")) min_batch = atoll(e);
    return want && (spt == 1 || low.dtype == FDG_C128) && work >= 3 * (int64_t)budget && batch >= min_batch;
}

int jit_launch(fdg_program *h, DeviceState &ds, int dev, int spt, bool acc, const void *leaf, int64_t ld_leaf, void *root,
               int64_t ld_root, int64_t batch, cudaStream_t stream) {
    if (pipeline_wanted(h, spt, batch)) return jit_launch_pipeline(h, ds, dev, acc, leaf, ld_leaf, root, ld_root, batch, stream);
    JitVariant *v = nullptr;
    StreamScratch &ss = ds.per_stream[stream];
    // row offsets are formed with one 32-bit multiply-add unless a leading dimension reaches 4 GiB
    const bool wide = (uint64_t)ld_leaf * (h->low.dtype == FDG_C128 ? 16 : 8) >= (1ull << 32);
    // Bulk form (persistent warp-specialised kernels fed by cp.async.bulk, DESIGN.md section 4b'): for programs of several
    // kernels on batches that give every SM at least eight tiles (a launch of these kernels costs ~10 us: at four tiles per SM
    // the ring form is still ahead, 168 vs 162 M samples/s on the headline graph); bulk copies move whole 16-byte units from 16-byte aligned addresses.
    // FDG_JIT_BULK = 0 never, 1 whenever the buffers allow it (tests), unset: the rule above.
    bool bulk = false, bulk_forced = false;
    {
        const bool cplx = h->low.dtype == FDG_C128;
        const bool aligned = cplx || (((uintptr_t)leaf % 16 == 0) && (ld_leaf % 2 == 0));
        const int64_t work = (h->low.muls_vv + h->low.muls_vf + h->low.adds_vv + h->low.pow_muls) * (cplx ? 4 : 1);  // ~instructions
        int mode = -1;
        if (const char *e = getenv("FDG_JIT_BULK")) mode = atoi(e);
        const int budget = h->jit_segment > 0 ? h->jit_segment : 4000;
        bulk = aligned && (spt == 1 || cplx) && mode != 0 && (mode > 0 || (work >= 2 * (int64_t)budget && batch >= (int64_t)256 * ds.sm_count * 8));
        bulk_forced = mode > 0;
    }
    int rc = jit_get(h, spt, acc, &v, wide, nullptr, bulk);
    if (rc != FDG_OK) return rc;
    if (bulk && !bulk_forced && h->low.dtype == FDG_C128 && fdg::jit_spill_bytes(v->plan) > 4096) {
        // ComplexF64 has two registers per value: where the 240-register consumers of the bulk form spill, the ring form's
        // 255 registers do better (Taylor-AD sigma order 4: 14 KB of spill stores per sample, 114 vs 129 M samples/s);
        // where they do not, the bulk form wins here too (Taylor-AD sigma order 3: 1.84e9 vs 1.66e9)
        bulk = false;
        rc = jit_get(h, spt, acc, &v, wide, nullptr, false);
        if (rc != FDG_OK) return rc;
    }
    bulk = v->plan.bulk;
    h->last_jit_key = spt * 2 + (acc ? 1 : 0) + (wide ? 64 : 0) + (bulk ? 32 : 0);
    auto &kern = v->kernels[dev];
    if (kern.empty()) {
        for (auto &sg : v->plan.seg) {
            cudaLibrary_t lib;
            CUDA_TRY(cudaLibraryLoadData(&lib, sg.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
            v->libs[dev].push_back(lib);
            cudaKernel_t k;
            CUDA_TRY(cudaLibraryGetKernel(&k, lib, sg.name.c_str()));
            if (bulk) CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, v->plan.bulk_smem));
            kern.push_back(k);
        }
    }
    const fdg::Lowered &low = h->low;
    const int W = low.dtype == FDG_C128 ? 2 : 1;
    const size_t es = 8 * (size_t)W;
    int T = 128;
    if (const char *e = getenv("FDG_JIT_THREADS")) T = std::max(32, std::min(128, atoi(e) / 32 * 32));
    if (bulk) T = 256;  // samples of a tile (the block has 128 more threads: the producer's warpgroup)
    const int64_t per_block = bulk ? 256 : (int64_t)T * spt;
    // sub-batches keep the cross buffer bounded (8 GiB unless FDG_JIT_CROSS_GB says otherwise)
    int64_t sub = batch;
    if (v->plan.n_cross > 0) {
        double cross_gb = 8.0;
        if (const char *e = getenv("FDG_JIT_CROSS_GB")) cross_gb = atof(e);
        const int64_t cap = std::max<int64_t>(per_block * ds.sm_count * 4,
                                              (int64_t)(cross_gb * (double)(1 << 30)) / ((int64_t)es * (int64_t)v->plan.n_cross));
        sub = std::min<int64_t>(batch, cap / per_block * per_block);
    }
    if (const char *e = getenv("FDG_JIT_SUB")) {  // experiment: samples per launch sequence (L2 blocking)
        const int64_t want = atoll(e);
        if (want > 0 && v->plan.seg.size() > 1) sub = std::min<int64_t>(batch, std::max<int64_t>(per_block, want / per_block * per_block));
    }
    // ld_cross bytes < 4 GiB (row offsets are 32-bit multiplies); a plan without a cross buffer has no such limit
    if (v->plan.n_cross > 0) sub = std::min<int64_t>(sub, ((int64_t)((1ull << 32) / es) - per_block) / per_block * per_block);
    if (const char *e = getenv("FDG_JIT_MAX_SUB")) {  // tests: force several launch sequences whatever the plan
        const int64_t want = atoll(e);
        if (want > 0) sub = std::min<int64_t>(sub, std::max<int64_t>(per_block, want / per_block * per_block));
    }
    int64_t max_grid = (sub + per_block - 1) / per_block;
    if (v->plan.persistent) max_grid = std::min<int64_t>(max_grid, (int64_t)ds.sm_count * 16);
    const int64_t ld_cross = max_grid * per_block;
    if (bulk) max_grid = std::min<int64_t>(max_grid, ds.sm_count);  // one block per SM walks the tiles
    if (bulk) {  // experiment (tools/exp_lanes.py): fewer blocks than SMs, so that launch sequences on two streams share the device
        if (const char *e = getenv("FDG_JIT_BULK_GRID")) max_grid = std::max<int64_t>(1, std::min<int64_t>(max_grid, atoll(e)));
    }
    if (v->plan.n_cross > 0) {
        const size_t need = (size_t)v->plan.n_cross * ld_cross * es;
        if (need > ss.cross_bytes) {
            if (ss.cross) CUDA_TRY(cudaFree(ss.cross));
            ss.cross = nullptr;
            ss.cross_bytes = 0;
            CUDA_TRY(cudaMalloc(&ss.cross, need));
            ss.cross_bytes = need;
        }
    }
    long long rows = 0;
    void *out = root;
    if (acc) {
        rows = max_grid * (T / 32);
        const size_t need = std::max<size_t>((size_t)rows * low.R * W * sizeof(double), 256);
        if (need > ss.partial_bytes) {
            if (ss.partial) CUDA_TRY(cudaFree(ss.partial));
            ss.partial = nullptr;
            ss.partial_bytes = 0;
            CUDA_TRY(cudaMalloc((void **)&ss.partial, need));
            ss.partial_bytes = need;
        }
        if (!v->plan.persistent) CUDA_TRY(cudaMemsetAsync(ss.partial, 0, (size_t)rows * low.R * W * sizeof(double), stream));
        out = ss.partial;
    }
    for (int64_t b0 = 0; b0 < batch; b0 += sub) {
        const int64_t nb = std::min<int64_t>(sub, batch - b0);
        const unsigned grid = (unsigned)std::min<int64_t>((nb + per_block - 1) / per_block, max_grid);
        const void *p_leaf = static_cast<const char *>(leaf) + (size_t)b0 * es;  // (sub-batches are whole tiles: stays 16-byte aligned)
        void *p_out = acc ? out : static_cast<void *>(static_cast<char *>(root) + (size_t)b0 * es);
        void *p_cross = ss.cross;
        long long a_ld_leaf = ld_leaf, a_ld_cross = ld_cross, a_ld_root = ld_root, a_batch = nb, a_nroots = low.R * W;
        void *args[] = {(void *)&p_leaf, &a_ld_leaf, &p_cross, &a_ld_cross, &p_out, &a_ld_root, &a_batch, &a_nroots};
        for (size_t sg = 0; sg < kern.size(); ++sg) {
            if (bulk) CUDA_TRY(cudaLaunchKernel((const void *)kern[sg], dim3(grid), dim3(T + 128), args, (size_t)v->plan.bulk_smem, stream));
            else CUDA_TRY(cudaLaunchKernel((const void *)kern[sg], dim3(grid), dim3(T), args, 0, stream));
            h->launches++;
        }
        if (acc && v->plan.persistent && low.R > 0 && b0 + sub < batch) {
            // The grid-stride kernel STORES its warps' sums (the rows are not zeroed): with more launch sequences to come
            // this one's rows -- exactly the rows its grid wrote -- are folded into the result before they are overwritten.
            fdg::fdg_reduce_partials<<<(int)low.R * W, 256, 0, stream>>>(ss.partial, (long long)grid * (T / 32), (int)low.R * W, static_cast<double *>(root));
            CUDA_TRY(cudaGetLastError());
            h->launches++;
        }
        if (acc && v->plan.persistent) rows = (long long)grid * (T / 32);  // rows the last launch wrote
    }
    if (acc && low.R > 0) {
        fdg::fdg_reduce_partials<<<(int)low.R * W, 256, 0, stream>>>(ss.partial, rows, (int)low.R * W, static_cast<double *>(root));
        CUDA_TRY(cudaGetLastError());
        h->launches++;
    }
    return FDG_OK;
}

int do_eval(fdg_program *h, const void *leaf, int64_t ld_leaf, void *root, int64_t ld_root, int64_t batch,
            void *stream, bool accumulate) {
    if (!h) return fail(FDG_ERR_BAD_ARG, "null handle");
    if (batch < 0) return fail(FDG_ERR_BAD_ARG, "negative batch");
    if (batch == 0) return FDG_OK;
    const fdg::Lowered &low = h->low;
    if (low.L > 0 && !leaf) return fail(FDG_ERR_BAD_ARG, "null leaf pointer");
    if (low.R > 0 && !root) return fail(FDG_ERR_BAD_ARG, "null root pointer");
    if (low.L > 0 && ld_leaf < batch) return fail(FDG_ERR_BAD_ARG, "ld_leaf < batch");
    if (!accumulate && low.R > 0 && ld_root < batch) return fail(FDG_ERR_BAD_ARG, "ld_root < batch");
    const bool cplx = low.dtype == FDG_C128;
    const size_t esize = cplx ? 16 : 8;
    if (((uintptr_t)leaf % esize) || ((uintptr_t)root % 8) || (!accumulate && (uintptr_t)root % esize))
        return fail(FDG_ERR_BAD_ARG, "leaf/root pointer not aligned to the element size");
    std::lock_guard<std::mutex> lock(h->mu);
    DeviceState *ds = nullptr;
    int rc = get_device_state(h, &ds);
    if (rc != FDG_OK) return rc;
    fdg::VmArgs args{};
    args.prog = ds->d_prog;
    args.leaf = leaf;
    args.ld_leaf = ld_leaf;
    args.root = root;
    args.ld_root = ld_root;
    args.batch = batch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int backend = h->backend;
    if (const char *e = getenv("FDG_BACKEND")) backend = atoi(e);
    if (cplx) {
        if (backend != FDG_BACKEND_VM && low.N + low.R > 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            rc = jit_launch(h, *ds, dev, 1, accumulate, leaf, ld_leaf, root, ld_root, batch, st);
            if (rc == FDG_OK || backend == FDG_BACKEND_JIT) return rc;
        }
        return accumulate ? launch_variant<fdg::VCplx, true>(h, *ds, args, batch, st)
                          : launch_variant<fdg::VCplx, false>(h, *ds, args, batch, st);
    }
    // S samples per thread need 16-byte aligned sample groups that stay inside the allocation:
    // even leading dimensions and the rounded-up batch within ld_leaf
    auto fits = [&](int64_t s) {
        const int64_t padded = (batch + s - 1) / s * s;
        const bool leaf_ok = low.L == 0 || (((uintptr_t)leaf % 16 == 0) && (ld_leaf % 2 == 0) && padded <= ld_leaf);
        const bool root_ok = accumulate || low.R == 0 || (((uintptr_t)root % 16 == 0) && (ld_root % 2 == 0));
        return leaf_ok && root_ok;
    };
    if (backend != FDG_BACKEND_VM && low.N + low.R > 0) {
        // two samples per thread pay off only for small programs (everything stays in registers); big ones want
        // the registers for the program's own live values
        int jspt = h->spt == 0 ? (low.n_operands <= 400 ? 2 : 1) : (h->spt >= 2 ? 2 : 1);
        if (jspt == 2 && !fits(2)) jspt = 1;
        int dev = 0;
        cudaGetDevice(&dev);
        rc = jit_launch(h, *ds, dev, jspt, accumulate, leaf, ld_leaf, root, ld_root, batch, st);
        if (rc == FDG_OK || backend == FDG_BACKEND_JIT) return rc;
        // AUTO: the specialised path did not build (e.g. PTX compiler unavailable) -> the VM runs the program
    }
    int spt = h->spt;
    if (spt == 0) spt = (batch >= (1 << 16) && fits(4)) ? 4 : (fits(2) ? 2 : 1);  // auto: widest that fits
    if (spt > 1 && !fits(spt))
        return fail(FDG_ERR_BAD_ARG, "several samples per thread need 16-byte aligned buffers, even leading dimensions "
                                     "and the batch rounded up to the group size within ld_leaf");
    if (spt == 4)
        return accumulate ? launch_variant<fdg::VReal<4>, true>(h, *ds, args, batch, st)
                          : launch_variant<fdg::VReal<4>, false>(h, *ds, args, batch, st);
    if (spt == 2)
        return accumulate ? launch_variant<fdg::VReal<2>, true>(h, *ds, args, batch, st)
                          : launch_variant<fdg::VReal<2>, false>(h, *ds, args, batch, st);
    return accumulate ? launch_variant<fdg::VReal<1>, true>(h, *ds, args, batch, st)
                      : launch_variant<fdg::VReal<1>, false>(h, *ds, args, batch, st);
}

}  // namespace

extern "C" {

int fdg_abi_version(void) { return FDG_ABI_VERSION; }
const char *fdg_last_error(void) { return g_err.c_str(); }

static int fdg_compile_impl(const fdg_graph_desc *graph, const fdg_options *opts, fdg_handle *out) {
    if (!graph || !out) return fail(FDG_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    fdg_options o;
    std::memset(&o, 0, sizeof(o));
    if (opts) o = *opts;
    if (o.fma != 0 && o.fma != 1) return fail(FDG_ERR_BAD_ARG, "fdg_options.fma must be 0 or 1");
    if (o.fma == 1 && o.backend == FDG_BACKEND_VM) return fail(FDG_ERR_UNSUPPORTED, "fma applies to the specialised kernels only");
    if (o.backend < FDG_BACKEND_AUTO || o.backend > FDG_BACKEND_JIT) return fail(FDG_ERR_BAD_ARG, "unknown backend");
    if (o.cse < -1 || o.cse > 1) return fail(FDG_ERR_BAD_ARG, "fdg_options.cse must be -1 (never), 0 (automatic) or 1 (always)");
    const int cse_mode = o.cse;
    o.cse = cse_mode == 1 ? 1 : 0;  // what lower() sees: merge or not
    fdg_program *p = new (std::nothrow) fdg_program();
    if (!p) return fail(FDG_ERR_BAD_ARG, "out of memory");
    std::string err;
    int rc;
    try {
        rc = fdg::lower(*graph, o, p->low, err);
    } catch (const std::exception &e) {
        rc = FDG_ERR_BAD_ARG;
        err = std::string("lowering failed: ") + e.what();
    }
    if (rc != FDG_OK) {
        delete p;
        return fail(rc, err);
    }
    if (cse_mode == 0 && o.backend != FDG_BACKEND_VM && p->low.N >= 16) {
        fdg_options oc = o;
        oc.cse = 1;
        std::string e2;
        try {
            if (fdg::lower(*graph, oc, p->low_cse, e2) == FDG_OK && p->low_cse.cse_removed > 0) p->has_cse = true;
        } catch (const std::exception &) {
        }
        if (!p->has_cse) p->low_cse = fdg::Lowered();
    }
    p->backend = o.backend;
    p->jit_segment = o.jit_segment;
    p->fma = o.fma == 1;
    if (p->fma) p->backend = FDG_BACKEND_JIT;  // the packet VM only has the bit-exact arithmetic
    *out = p;
    return FDG_OK;
}

static int fdg_graph_write_impl(const fdg_graph_desc *g, const char *path) {
    if (!g || !path) return fail(FDG_ERR_BAD_ARG, "null argument");
    if (g->n_nodes < 0 || g->n_edges < 0 || g->n_graphs < 0 || g->n_roots < 0) return fail(FDG_ERR_BAD_ARG, "negative size");
    FILE *fp = std::fopen(path, "wb");
    if (!fp) return fail(FDG_ERR_BAD_ARG, std::string("cannot open ") + path + " for writing");
    const int64_t hdr[4] = {g->n_nodes, g->n_edges, g->n_graphs, g->n_roots};
    bool ok = std::fwrite("FDGRAPH\1", 1, 8, fp) == 8 && std::fwrite(hdr, 8, 4, fp) == 4;
    auto put = [&](const void *p, size_t size, int64_t n) {
        if (ok && n > 0) ok = p && std::fwrite(p, size, (size_t)n, fp) == (size_t)n;
    };
    put(g->node_id, 8, g->n_nodes);
    put(g->node_op, 4, g->n_nodes);
    put(g->node_pow, 4, g->n_nodes);
    put(g->child_ptr, 8, g->n_nodes + 1);
    put(g->child_node, 4, g->n_edges);
    put(g->child_factor, 8, g->n_edges);
    put(g->graphs, 4, g->n_graphs);
    put(g->root_id, 8, g->n_roots);
    ok = (std::fclose(fp) == 0) && ok;
    return ok ? FDG_OK : fail(FDG_ERR_BAD_ARG, std::string("writing ") + path + " failed");
}

static int fdg_compile_file_impl(const char *path, const fdg_options *opts, fdg_handle *out) {
    if (!path || !out) return fail(FDG_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    FILE *fp = std::fopen(path, "rb");
    if (!fp) return fail(FDG_ERR_BAD_ARG, std::string("cannot open ") + path);
    char magic[8];
    int64_t hdr[4] = {0, 0, 0, 0};
    bool ok = std::fread(magic, 1, 8, fp) == 8 && std::memcmp(magic, "FDGRAPH\1", 8) == 0 && std::fread(hdr, 8, 4, fp) == 4;
    const int64_t lim = (int64_t)1 << 31;
    ok = ok && hdr[0] >= 0 && hdr[1] >= 0 && hdr[2] >= 0 && hdr[3] >= 0 && hdr[0] < lim && hdr[1] < lim && hdr[2] < lim && hdr[3] < lim;
    if (!ok) {
        std::fclose(fp);
        return fail(FDG_ERR_BAD_GRAPH, std::string(path) + " is not an FDGRAPH file");
    }
    {
        // the header must agree with the length of the file BEFORE anything is sized from it (a corrupt count would
        // otherwise ask for gigabytes)
        const long long want = 40 + hdr[0] * (8 + 4 + 4) + (hdr[0] + 1) * 8 + hdr[1] * (4 + 8) + hdr[2] * 4 + hdr[3] * 8;
        long long have = -1;
        if (std::fseek(fp, 0, SEEK_END) == 0) have = std::ftell(fp);
        if (have != want || std::fseek(fp, 40, SEEK_SET) != 0) {
            std::fclose(fp);
            return fail(FDG_ERR_BAD_GRAPH, std::string(path) + " is truncated or has trailing bytes");
        }
    }
    std::vector<int64_t> node_id, child_ptr, root_id;
    std::vector<int32_t> node_op, node_pow, child_node, graphs;
    std::vector<double> child_factor;
    try {
        node_id.resize((size_t)hdr[0]), child_ptr.resize((size_t)hdr[0] + 1), root_id.resize((size_t)hdr[3]);
        node_op.resize((size_t)hdr[0]), node_pow.resize((size_t)hdr[0]), child_node.resize((size_t)hdr[1]), graphs.resize((size_t)hdr[2]);
        child_factor.resize((size_t)hdr[1]);
    } catch (const std::exception &) {
        std::fclose(fp);
        return fail(FDG_ERR_BAD_ARG, "out of memory reading " + std::string(path));
    }
    auto get = [&](void *p, size_t size, size_t n) {
        if (ok && n > 0) ok = std::fread(p, size, n, fp) == n;
    };
    get(node_id.data(), 8, node_id.size());
    get(node_op.data(), 4, node_op.size());
    get(node_pow.data(), 4, node_pow.size());
    get(child_ptr.data(), 8, child_ptr.size());
    get(child_node.data(), 4, child_node.size());
    get(child_factor.data(), 8, child_factor.size());
    get(graphs.data(), 4, graphs.size());
    get(root_id.data(), 8, root_id.size());
    ok = ok && std::fgetc(fp) == EOF;  // nothing may follow
    std::fclose(fp);
    if (!ok) return fail(FDG_ERR_BAD_GRAPH, std::string(path) + " is truncated or has trailing bytes");
    fdg_graph_desc d;
    d.n_nodes = hdr[0], d.n_edges = hdr[1], d.n_graphs = hdr[2], d.n_roots = hdr[3];
    d.node_id = node_id.data(), d.node_op = node_op.data(), d.node_pow = node_pow.data(), d.child_ptr = child_ptr.data();
    d.child_node = child_node.data(), d.child_factor = child_factor.data(), d.graphs = graphs.data(), d.root_id = root_id.data();
    return fdg_compile(&d, opts, out);
}

// fdg_jit_prepare / _info / _ptx look at the bulk form of the kernels when FDG_JIT_BULK >= 1 (tools, tests)
static bool inspect_bulk(const fdg_program *h, int spt) {
    const char *e = getenv("FDG_JIT_BULK");
    return e && atoi(e) > 0 && (spt == 1 || h->low.dtype == FDG_C128);
}

static int fdg_jit_prepare_impl(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int32_t *n_kernels, int32_t *n_cross,
                    int64_t *cubin_bytes) {
    if (!h) return fail(FDG_ERR_BAD_ARG, "null handle");
    if (samples_per_thread != 1 && samples_per_thread != 2) return fail(FDG_ERR_BAD_ARG, "samples_per_thread must be 1 or 2");
    std::lock_guard<std::mutex> lock(h->mu);
    JitVariant *v = nullptr;
    int rc = jit_get(h, samples_per_thread, accumulate != 0, &v, false, nullptr, inspect_bulk(h, samples_per_thread));
    if (rc != FDG_OK) return rc;
    if (n_kernels) *n_kernels = (int32_t)v->plan.seg.size();
    if (n_cross) *n_cross = v->plan.n_cross;
    if (cubin_bytes) {
        *cubin_bytes = 0;
        for (auto &s : v->plan.seg) *cubin_bytes += (int64_t)s.cubin.size();
    }
    return FDG_OK;
}

static int fdg_jit_info_impl(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int64_t *out, int32_t n_out) {
    if (!h || !out || n_out < 0) return fail(FDG_ERR_BAD_ARG, "null argument");
    std::lock_guard<std::mutex> lock(h->mu);
    // samples_per_thread == 0: the variant the last launch of this handle ran
    auto it = h->jit.find(samples_per_thread == 0 ? h->last_jit_key
                                                  : samples_per_thread * 2 + (accumulate ? 1 : 0) + (inspect_bulk(h, samples_per_thread) ? 32 : 0));
    if (it == h->jit.end() || !it->second.compiled) return fail(FDG_ERR_BAD_ARG, "variant not prepared");
    const fdg::JitPlan &pl = it->second.plan;
    int64_t ops = 0;
    for (auto &sg : pl.seg) ops += sg.n_stmts;
    const int64_t vals[15] = {(int64_t)pl.seg.size(), pl.n_cross, pl.n_cross_values, pl.leaf_loads, pl.cross_loads, pl.cross_stores, ops,
                              pl.persistent ? 1 : 0, pl.max_code_bytes, pl.uses_cse ? 1 : 0, pl.fp64_instr,
                              (int64_t)(1000.0 * fdg::jit_model_ns(pl, h->low.dtype == FDG_C128 ? 16 : 8)), pl.bulk ? 1 : 0, pl.bulk_smem, pl.refetch_loads};
    for (int32_t i = 0; i < n_out && i < 15; ++i) out[i] = vals[i];
    return FDG_OK;
}

static int fdg_pipeline_prepare_impl(fdg_handle h, int32_t accumulate, int32_t n_sm, int32_t what, int64_t *out, int32_t n_out) {
    if (!h || n_sm < 1 || n_out < 0 || (!out && n_out > 0)) return fail(FDG_ERR_BAD_ARG, "bad argument");
    std::lock_guard<std::mutex> lock(h->mu);
    JitVariant *v = nullptr;
    // the instruction-cache groups of the device: the B200 layout measured by tools/icache_probe.py for 148 SMs, otherwise
    // groups of at most 20 SMs (a launch measures the real ones)
    std::vector<int> groups = {12, 18, 18, 20, 20, 20, 20, 20};
    if (n_sm != 148) {
        const int G = (n_sm + 19) / 20;
        groups.assign((size_t)G, n_sm / G);
        for (int g = 0; g < n_sm % G; ++g) groups[(size_t)g] += 1;
    }
    int rc = jit_get(h, 1, accumulate != 0, &v, false, &groups);
    if (rc != FDG_OK) return rc;
    const fdg::JitPlan &pl = v->plan;
    std::vector<int64_t> vals;
    if (what == 0) {
        int64_t ops = 0, linked = 0;
        for (auto &sg : pl.seg) ops += sg.n_stmts;
        for (auto &l : pl.linked) linked += (int64_t)l.size();
        vals = {(int64_t)pl.seg.size(), pl.n_cross, pl.n_cross_values, pl.leaf_loads, pl.cross_loads, pl.cross_stores, ops,
                pl.max_code_bytes, linked, pl.ring_bytes, pl.n_pass, pl.n_boundary};
    } else if (what == 1) {
        vals.assign(pl.stage_blocks.begin(), pl.stage_blocks.end());
    } else if (what == 2) {
        vals.assign(pl.stage_cost.begin(), pl.stage_cost.end());
    } else if (what == 3) {
        vals.assign(pl.stage_start.begin(), pl.stage_start.end());
    } else {
        return fail(FDG_ERR_BAD_ARG, "unknown query");
    }
    for (int32_t i = 0; i < n_out; ++i) out[i] = i < (int32_t)vals.size() ? vals[(size_t)i] : 0;
    return FDG_OK;
}

static int fdg_pipeline_stats_impl(fdg_handle h, void *stream, int64_t *out, int32_t n_out) {
    if (!h || n_out < 0 || (!out && n_out > 0)) return fail(FDG_ERR_BAD_ARG, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaStreamSynchronize(st));
    std::lock_guard<std::mutex> lock(h->mu);
    DeviceState *ds = nullptr;
    int rc = get_device_state(h, &ds);
    if (rc != FDG_OK) return rc;
    auto it = ds->per_stream.find(st);
    if (it == ds->per_stream.end() || !it->second.pipe_stats) return fail(FDG_ERR_BAD_ARG, "no pipeline launch on this stream yet");
    std::vector<unsigned long long> buf(it->second.pipe_stats_bytes / 8);
    CUDA_TRY(cudaMemcpy(buf.data(), it->second.pipe_stats, buf.size() * 8, cudaMemcpyDeviceToHost));
    // out[0] = stall flag, then (busy, waiting) clock sums of stage 0, 1, ...
    for (int32_t i = 0; i < n_out; ++i) out[i] = i == 0 ? (int64_t)(buf[0] & 0xffffffffu) : (i + 1 < (int32_t)buf.size() ? (int64_t)buf[(size_t)i + 1] : 0);
    return FDG_OK;
}

static int fdg_jit_ptx_impl(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int32_t index, const char **ptx,
                const char **ptxas_log) {
    if (!h || !ptx) return fail(FDG_ERR_BAD_ARG, "null argument");
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->jit.find(samples_per_thread == 0 ? h->last_jit_key
                                                  : samples_per_thread * 2 + (accumulate ? 1 : 0) + (inspect_bulk(h, samples_per_thread) ? 32 : 0));
    if (it == h->jit.end() || !it->second.compiled) return fail(FDG_ERR_BAD_ARG, "variant not prepared");
    if (index < 0 || index >= (int32_t)it->second.plan.seg.size()) return fail(FDG_ERR_BAD_ARG, "kernel index out of range");
    *ptx = it->second.plan.seg[(size_t)index].ptx.c_str();
    if (ptxas_log) *ptxas_log = it->second.plan.seg[(size_t)index].info.c_str();
    return FDG_OK;
}

static int fdg_destroy_impl(fdg_handle h) {
    if (!h) return FDG_OK;
    for (auto &kv : h->dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) break;
        cudaSetDevice(kv.first);
        DeviceState &ds = kv.second;
        cudaFree(ds.d_prog);
        cudaFree(ds.d_smtab);
        for (auto &ps : ds.per_stream) {
            cudaFree(ps.second.scratch);
            cudaFree(ps.second.partial);
            cudaFree(ps.second.cross);
            cudaFree(ps.second.progress);
            cudaFree(ps.second.pipe_stats);
        }
        for (auto &kv2 : h->jit) {
            auto it = kv2.second.libs.find(kv.first);
            if (it != kv2.second.libs.end())
                for (auto lib : it->second) cudaLibraryUnload(lib);
        }
        for (int i = 0; i < 2; ++i) {
            if (ds.streams[i]) cudaStreamDestroy(ds.streams[i]);
            cudaFree(ds.d_leaf[i]);
            cudaFree(ds.d_root[i]);
        }
        cudaSetDevice(cur);
    }
    delete h;
    return FDG_OK;
}

static int fdg_stats_impl(fdg_handle h, fdg_stats_t *out) {
    if (!h || !out) return fail(FDG_ERR_BAD_ARG, "null argument");
    std::memset(out, 0, sizeof(*out));
    const fdg::Lowered &l = h->low;
    const bool c = l.dtype == FDG_C128;
    out->n_leaves = l.L;
    out->n_inner = l.N;
    out->n_roots = l.R;
    out->n_operands = l.n_operands;
    out->n_packets = (int64_t)l.words.size() / 4;
    out->n_slots = l.n_slots;
    out->n_scratch = l.n_scratch;
    out->leaf_loads = l.leaf_loads;
    // complex*complex = 4 mul + 2 add, complex*real = 2 mul, complex+complex = 2 add
    out->flops_mul = c ? 4 * (l.muls_vv + l.pow_muls) + 2 * l.muls_vf : l.muls_vv + l.pow_muls + l.muls_vf;
    out->flops_add = c ? 2 * (l.muls_vv + l.pow_muls) + 2 * l.adds_vv : l.adds_vv;
    out->bytes_in = (c ? 16 : 8) * l.L;
    out->bytes_out = (c ? 16 : 8) * l.R;
    out->max_depth = l.max_depth;
    out->cse_removed = l.cse_removed;
    return FDG_OK;
}

static int fdg_leafmap_impl(fdg_handle h, int32_t *leaf_node) {
    if (!h || (!leaf_node && h->low.L > 0)) return fail(FDG_ERR_BAD_ARG, "null argument");
    std::copy(h->low.leaf_node.begin(), h->low.leaf_node.end(), leaf_node);
    return FDG_OK;
}

int fdg_last_root(fdg_handle h, int32_t *out) {
    if (!h || !out) return fail(FDG_ERR_BAD_ARG, "null argument");
    *out = h->low.last_root;
    return FDG_OK;
}

int fdg_program_words(fdg_handle h, const uint32_t **words, int64_t *n_words) {
    if (!h || !words || !n_words) return fail(FDG_ERR_BAD_ARG, "null argument");
    *words = h->low.words.data();
    *n_words = (int64_t)h->low.words.size();
    return FDG_OK;
}

static int fdg_eval_impl(fdg_handle h, const void *leaf, int64_t ld_leaf, void *root, int64_t ld_root, int64_t batch,
             void *stream) {
    return do_eval(h, leaf, ld_leaf, root, ld_root, batch, stream, false);
}

static int fdg_eval_accumulate_impl(fdg_handle h, const void *leaf, int64_t ld_leaf, int64_t batch, double *acc, void *stream) {
    return do_eval(h, leaf, ld_leaf, acc, 0, batch, stream, true);
}

static int fdg_eval_host_impl(fdg_handle h, const void *leaf_host, int64_t ld_leaf, void *root_host, int64_t ld_root,
                  int64_t batch) {
    if (!h) return fail(FDG_ERR_BAD_ARG, "null handle");
    if (batch < 0) return fail(FDG_ERR_BAD_ARG, "negative batch");
    if (batch == 0) return FDG_OK;
    const fdg::Lowered &low = h->low;
    if ((low.L > 0 && !leaf_host) || (low.R > 0 && !root_host)) return fail(FDG_ERR_BAD_ARG, "null host buffer");
    if ((low.L > 0 && ld_leaf < batch) || (low.R > 0 && ld_root < batch))
        return fail(FDG_ERR_BAD_ARG, "leading dimension < batch");
    const size_t es = low.dtype == FDG_C128 ? 16 : 8;
    // one host call at a time per handle: the staging buffers and the two streams are the handle's (callers on other
    // threads wait here; device-pointer calls on their own streams are not affected)
    std::lock_guard<std::mutex> host_lock(h->host_mu);
    DeviceState *ds = nullptr;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        int rc = get_device_state(h, &ds);
        if (rc != FDG_OK) return rc;
    }
    // chunk so that one chunk of leaves is about 64 MiB; an even number of samples keeps pairs aligned
    const size_t row = std::max<size_t>((size_t)std::max<int64_t>(low.L, 1) * es, 1);
    int64_t chunk = (int64_t)((64ull << 20) / row);
    chunk = std::max<int64_t>(1024, chunk) & ~(int64_t)1023;
    chunk = std::min<int64_t>(chunk, (batch + 1) & ~(int64_t)1);
    const size_t need_leaf = (size_t)chunk * std::max<int64_t>(low.L, 1) * es;
    const size_t need_root = (size_t)chunk * std::max<int64_t>(low.R, 1) * es;
    for (int i = 0; i < 2; ++i) {
        if (!ds->streams[i]) CUDA_TRY(cudaStreamCreateWithFlags(&ds->streams[i], cudaStreamNonBlocking));
    }
    if (need_leaf > ds->d_leaf_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (ds->d_leaf[i]) CUDA_TRY(cudaFree(ds->d_leaf[i]));
            ds->d_leaf[i] = nullptr;
            CUDA_TRY(cudaMalloc(&ds->d_leaf[i], need_leaf));
        }
        ds->d_leaf_bytes = need_leaf;
    }
    if (need_root > ds->d_root_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (ds->d_root[i]) CUDA_TRY(cudaFree(ds->d_root[i]));
            ds->d_root[i] = nullptr;
            CUDA_TRY(cudaMalloc(&ds->d_root[i], need_root));
        }
        ds->d_root_bytes = need_root;
    }
    int k = 0;
    for (int64_t c0 = 0; c0 < batch; c0 += chunk, k ^= 1) {
        const int64_t nb = std::min<int64_t>(chunk, batch - c0);
        cudaStream_t st = ds->streams[k];
        if (low.L > 0)
            CUDA_TRY(cudaMemcpy2DAsync(ds->d_leaf[k], (size_t)chunk * es,
                                       static_cast<const char *>(leaf_host) + (size_t)c0 * es, (size_t)ld_leaf * es,
                                       (size_t)nb * es, (size_t)low.L, cudaMemcpyHostToDevice, st));
        int rc = do_eval(h, ds->d_leaf[k], chunk, ds->d_root[k], chunk, nb, st, false);
        if (rc != FDG_OK) {
            const std::string msg = g_err;  // copies still in flight read and write the caller's arrays: let them finish
            cudaStreamSynchronize(ds->streams[0]);
            cudaStreamSynchronize(ds->streams[1]);
            return fail(rc, msg);
        }
        if (low.R > 0)
            CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(root_host) + (size_t)c0 * es, (size_t)ld_root * es,
                                       ds->d_root[k], (size_t)chunk * es, (size_t)nb * es, (size_t)low.R,
                                       cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(ds->streams[0]));
    CUDA_TRY(cudaStreamSynchronize(ds->streams[1]));
    return FDG_OK;
}

static int fdg_set_launch_impl(fdg_handle h, int32_t threads, int32_t samples_per_thread, int32_t blocks_per_sm) {
    if (!h) return fail(FDG_ERR_BAD_ARG, "null handle");
    if (threads != 0 && (threads < 32 || threads > 256 || (threads & (threads - 1))))
        return fail(FDG_ERR_BAD_ARG, "threads must be 32, 64, 128 or 256");
    if (samples_per_thread < 0 || samples_per_thread > 4 || samples_per_thread == 3)
        return fail(FDG_ERR_BAD_ARG, "samples_per_thread must be 0 (auto), 1, 2 or 4");
    if (blocks_per_sm < 0) return fail(FDG_ERR_BAD_ARG, "blocks_per_sm must be >= 0");
    std::lock_guard<std::mutex> lock(h->mu);
    if (threads) h->threads = threads;
    h->spt = samples_per_thread;
    h->blocks_per_sm = blocks_per_sm;
    return FDG_OK;
}

int fdg_launch_count(fdg_handle h, int64_t *out) {
    if (!h || !out) return fail(FDG_ERR_BAD_ARG, "null argument");
    *out = h->launches.load();
    return FDG_OK;
}

// ---- NCCL, resolved at run time so that single-GPU users need no NCCL at all ---------------------------
namespace {
struct Id128 {  // ncclUniqueId is passed by value: 128 bytes
    char b[128];
};
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;
int load_nccl() {
    std::call_once(g_nccl_once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (!g_nccl.lib) return;
        g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(g_nccl.lib, "ncclGetUniqueId"));
        g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(g_nccl.lib, "ncclCommInitRank"));
        g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(g_nccl.lib, "ncclCommDestroy"));
        g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(g_nccl.lib, "ncclAllReduce"));
        g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.lib, "ncclGetErrorString"));
    });
    if (!g_nccl.lib || !g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
        return fail(FDG_ERR_NCCL, "libnccl.so.2 could not be loaded");
    return FDG_OK;
}
int nccl_fail(int rc, const char *what) {
    return fail(FDG_ERR_NCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"));
}
}  // namespace

static int fdg_comm_unique_id_impl(void *id128) {
    if (!id128) return fail(FDG_ERR_BAD_ARG, "null argument");
    int rc = load_nccl();
    if (rc != FDG_OK) return rc;
    int n = g_nccl.GetUniqueId(id128);
    return n == 0 ? FDG_OK : nccl_fail(n, "ncclGetUniqueId");
}

static int fdg_comm_init_impl(fdg_comm_t *out, int32_t nranks, int32_t rank, const void *id128) {
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(FDG_ERR_BAD_ARG, "bad argument");
    int rc = load_nccl();
    if (rc != FDG_OK) return rc;
    Id128 id;
    std::memcpy(id.b, id128, 128);
    fdg_comm *c = new (std::nothrow) fdg_comm();
    if (!c) return fail(FDG_ERR_BAD_ARG, "out of memory");
    int n = g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (n != 0) {
        delete c;
        return nccl_fail(n, "ncclCommInitRank");
    }
    *out = c;
    return FDG_OK;
}

static int fdg_comm_destroy_impl(fdg_comm_t c) {
    if (!c) return FDG_OK;
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    delete c;
    return FDG_OK;
}

static int fdg_allreduce_impl(fdg_comm_t c, double *acc, int64_t n, void *stream) {
    if (!c || !c->nccl_comm || (!acc && n > 0) || n < 0) return fail(FDG_ERR_BAD_ARG, "bad argument");
    if (n == 0) return FDG_OK;
    // ncclFloat64 = 8, ncclSum = 0 (nccl.h)
    int rc = g_nccl.AllReduce(acc, acc, (size_t)n, 8, 0, c->nccl_comm, static_cast<cudaStream_t>(stream));
    return rc == 0 ? FDG_OK : nccl_fail(rc, "ncclAllReduce");
}

}  // extern "C"

// =====================================================================================================================
// N1: leaf values computed on the device from the Monte-Carlo variables (include/fdgraph.h; example/benchmark.jl:44-127)
// =====================================================================================================================
namespace {

constexpr int FDG_LG_MAXLOOPS = 8;
constexpr int FDG_LG_THREADS = 128;

struct LeafMeta {  // one per leaf, read with uniform (broadcast) loads; sorted by loop-basis vector
    int32_t type, order, tau_in, tau_out;
    int32_t out, basis_id;  // column of leafVal this entry writes; index of its loop-basis vector
    double basis[FDG_LG_MAXLOOPS];
};

// x^n for a run-time integer n >= 0: Julia ^(x::Float64, n::Integer) (base/math.jl): 0 -> 1, 1 -> x, 2 -> x*x,
// 3 -> x*x*x, else pow_body (compensated power by squaring)
__device__ __forceinline__ double lg_pow(double x, int n) {
    if (n == 0) return 1.0;
    if (n == 1) return x;
    if (n == 2) return __dmul_rn(x, x);
    if (n == 3) return __dmul_rn(__dmul_rn(x, x), x);
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
    while (n > 1) {
        if (n & 1) {
            const double err = __fma_rn(y, xnlo, __dmul_rn(x, ynlo));
            const double pr = __dmul_rn(x, y);
            ynlo = __dadd_rn(__fma_rn(x, y, -pr), err);
            y = pr;
        }
        const double err = __dmul_rn(__dmul_rn(x, 2.0), xnlo);
        const double pr = __dmul_rn(x, x);
        xnlo = __dadd_rn(__fma_rn(x, x, -pr), err);
        x = pr;
        n >>= 1;
    }
    const double err = __fma_rn(y, xnlo, __dmul_rn(x, ynlo));
    return (isfinite(x) && isfinite(err)) ? __fma_rn(x, y, err) : __dmul_rn(x, y);
}

// green(tau, omega, beta), example/benchmark.jl:113-127 (TAU_CUTOFF = 1e-10; `tau ≈ 0.0` is `tau == 0` for the default
// tolerances of isapprox against an exact zero).  The denominator 1 + exp(-|omega| beta) depends on the momentum only:
// it is computed once per loop-basis vector and divided into every leaf that carries that momentum -- the same
// operations on the same operands as the formula evaluated leaf by leaf, hence the same bits.
__device__ __forceinline__ double lg_green_ebeta(double w, double beta) { return w > 0.0 ? exp(-w * beta) : exp(w * beta); }
__device__ __forceinline__ double lg_green(double tau, double w, double beta, double den) {
    if (tau == 0.0) tau = -1e-10;
    if (tau > 0.0) return w > 0.0 ? exp(-w * tau) / den : exp(w * (beta - tau)) / den;
    return w > 0.0 ? -exp(-w * (tau + beta)) / den : -exp(-w * tau) / den;
}

// (-1)^n / n! d^n/dw^n green(tau, w, beta), n = 1..5 (green_derive, example/benchmark.jl:93-111).  The reference takes
// these from Lehmann.Spectral.kernelFermiT_dw^n (not vendored: its bits are unpinned); here the derivative is written in
// closed form: green = exp(L(w)), L' = D = -tau' + beta n_F, so d^n green = green * Y_n(D, D', ..., D^(n-1)) with Y_n the
// complete Bell polynomial and the derivatives of the Fermi function n_F polynomials in n_F (oracle/leafgen.py, checked
// against 60-digit differentiation).  `ebeta` = exp(-|w| beta), `den` = 1 + ebeta.
__device__ __forceinline__ double lg_green_derive(double tau, double w, double beta, double ebeta, double den, int order) {
    if (tau == 0.0) tau = -1e-10;
    const double g0 = lg_green(tau, w, beta, den);
    const double tp = tau > 0.0 ? tau : tau + beta;
    const double n = w > 0.0 ? ebeta / den : 1.0 / den;
    const double m = n * (1.0 - n);
    const double b2 = beta * beta;
    const double D = -tp + beta * n;
    const double D1 = -b2 * m;
    double Y = D;
    if (order >= 2) {
        const double P2 = D * D;
        if (order == 2) {
            Y = P2 + D1;
        } else {
            const double D2 = b2 * beta * m * (1.0 - 2.0 * n);
            const double P3 = P2 * D;
            if (order == 3) {
                Y = P3 + 3.0 * D * D1 + D2;
            } else {
                const double D3 = -b2 * b2 * m * (1.0 - 6.0 * n + 6.0 * n * n);
                const double P4 = P3 * D;
                if (order == 4) {
                    Y = P4 + 6.0 * P2 * D1 + 4.0 * D * D2 + 3.0 * D1 * D1 + D3;
                } else {
                    const double D4 = b2 * b2 * beta * m * (1.0 - 14.0 * n + 36.0 * n * n - 24.0 * n * n * n);
                    Y = P4 * D + 10.0 * P3 * D1 + 10.0 * P2 * D2 + 15.0 * D * D1 * D1 + 5.0 * D * D3 + 10.0 * D1 * D2 + D4;
                }
            }
        }
    }
    const double coef = order == 1 ? -1.0 : order == 2 ? 0.5 : order == 3 ? -1.0 / 6.0 : order == 4 ? 1.0 / 24.0 : -1.0 / 120.0;
    return coef * (g0 * Y);
}

// ---- the generation kernel proper -------------------------------------------------------------------------------------
// Leaves are grouped by loop-basis vector.  A block of 128 threads (one sample each) stages the sample's variables -- the
// loop momenta K and the times T -- in shared memory ([row][thread]: conflict-free and indexable by the table), then walks
// the basis vectors of its chunk: |K . basis|^2 from the NON-ZERO coefficients of the vector only (3 to 5 of the 7 or 8
// loops of the order-4 graphs), the momentum-only factors exp(-|w| beta) and 1 / (1 + exp(-|w| beta)) once per vector, then
// one exp and one multiply per Green's-function leaf, no exp at all per interaction leaf.  All table reads are uniform over
// the block (broadcast).  Against the reference's formulas evaluated leaf by leaf this changes the last bits (fused
// multiply-adds in the dot products, a multiply by the reciprocal instead of a division): leaf values agree to ~1e-15
// relative (tests: <= 2e-14), which is the bar for this off-path producer (its bits are unpinned in the reference: BLAS
// `mul!` and Lehmann's kernels); the graph evaluation on top of the leaves stays bit-exact.
// exp(x) for x <= 0 with a 32-entry table: n = round(x 32 / ln 2) = 32 m + j, r = x - n ln 2 / 32 in two pieces
// (|r| <= ln 2 / 64), exp(x) = 2^m 2^(j/32) (1 + q(r)) with q of degree 6 (remainder 3e-18 relative).  Eleven FP64
// instructions against the twenty-five of lg_exp_neg; error below 0.9 ulp (checked against 60-digit arithmetic over
// [-746, 0], tools/check_exp_table.py).  `tab` = 2^(j/32), j = 0..31, in shared memory.
__constant__ double lg_exp2_32[32] = {0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, 0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, 0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};  // correctly rounded
// (constants as constant-bank operands of the DFMAs: immediates would cost two moves each per use)
__constant__ double lg_exp_c[10] = {46.16624130844683, 6755399441055744.0, -0.021660849364707246, -2.7791044496520866e-11,
                                    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
// results in the denormal range (x < -708): 2^m as two normal factors.  Kept out of line: the propagators almost never get there.
__device__ __noinline__ double lg_exp_scale_slow(double p, int m) {
    if (m < -1100) return 0.0;
    const int h = m >> 1;
    return p * __hiloint2double((h + 1023) << 20, 0) * __hiloint2double((m - h + 1023) << 20, 0);
}
__device__ __forceinline__ double lg_exp_tab(double x, const double *__restrict__ tab) {
    const double magic = lg_exp_c[1];  // 1.5 * 2^52: the sum below holds n in its low mantissa bits
    const double t = fma(x, lg_exp_c[0], magic);
    const int n = __double2loint(t);
    const double nd = t - magic;
    double r = fma(nd, lg_exp_c[2], x);  // ln 2 / 32, upper 30 bits (n * hi is exact)
    r = fma(nd, lg_exp_c[3], r);
    double q = lg_exp_c[4];
    q = fma(q, r, lg_exp_c[5]);
    q = fma(q, r, lg_exp_c[6]);
    q = fma(q, r, lg_exp_c[7]);
    q = fma(q, r, lg_exp_c[8]);
    q = fma(q, r, lg_exp_c[9]);
    q = q * r;
    const double tj = tab[n & 31];
    const double p = fma(tj, q, tj);
    const int m = n >> 5;
    // x < -1e6 (n no longer fits the trick above) gives 0 like every x < -746
    if (__builtin_expect(m < -1021 || x < -1.0e6, 0)) return x < -1.0e6 ? 0.0 : lg_exp_scale_slow(p, m);
    return __hiloint2double(__double2hiint(p) + (m << 20), __double2loint(p));  // p in [1, 2): the result is normal
}

// 1 / d for d in [1, 2]: the hardware's 20-bit estimate and two Newton steps (about 1 ulp; five instructions against the
// ~20 of a correctly rounded division)
__device__ __forceinline__ double lg_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

struct LgTerm {  // one non-zero coefficient of a loop-basis vector and the first row of its loop momentum in the staged variables
    double coef;
    int32_t off, pad;
};
struct alignas(16) LgBasis {
    // leaves of this momentum by kind: order-0 propagators g0[g0_first ..), leaves whose value does not depend on the times
    // (order-0 interactions, constants) cst[c_first ..), everything else leaves[leaf0 ..)
    int32_t nnz, g0_first, n_g0, c_first, n_c, leaf0, n_leaves, pad;
    LgTerm term[FDG_LG_MAXLOOPS];
};
struct LgLeaf {
    int32_t type, order, tau_in, tau_out, out, pad[3];
};
struct LgG0 {  // an order-0 Green's function leaf: the two times (tau_in | tau_out << 16) and the row it fills ...
    int32_t taus, out;
    long long off;  // ... as an element offset, out * ld_leaf (filled in per leading dimension when the table is uploaded)
};
struct LgCst {  // a leaf that is 1 (kind 0) or the order-0 interaction 8 pi (q^2 + lambda) (kind 1); row as an element offset
    long long off;
    int32_t kind, out;
};

template <int DIM, int S>
__global__ void __launch_bounds__(FDG_LG_THREADS, S == 1 ? 7 : 4)
fdg_leafgen2_kernel(const LgBasis *__restrict__ bases, int n_bases, int bases_per_block, const LgLeaf *__restrict__ leaves, const LgG0 *__restrict__ g0,
                    const LgCst *__restrict__ cst, int n_loops, int n_tau, const double *__restrict__ K, const double *__restrict__ T, long long ld_var, long long batch,
                    double *__restrict__ leaf, long long ld_leaf, double kF2, double beta, double lambda) {
    constexpr int COLS = FDG_LG_THREADS * S;
    extern __shared__ double lg_var[];  // [(n_loops * DIM + n_tau)][COLS], then the table of lg_exp_tab
    const int tid = threadIdx.x;
    const long long base = (long long)blockIdx.x * COLS;
    const int kr = n_loops * DIM;
    double *tab = lg_var + (size_t)(kr + n_tau) * COLS;
    if (tid < 32) tab[tid] = lg_exp2_32[tid];
    bool valid[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const long long b = base + s * FDG_LG_THREADS + tid;  // the block's samples as S coalesced rows of 128
        valid[s] = b < batch;
        const long long bs = valid[s] ? b : 0;
        for (int r = 0; r < kr; ++r) lg_var[r * COLS + s * FDG_LG_THREADS + tid] = K[(long long)r * ld_var + bs];
        for (int r = 0; r < n_tau; ++r) lg_var[(kr + r) * COLS + s * FDG_LG_THREADS + tid] = T[(long long)r * ld_var + bs];
    }
    __syncthreads();  // (the variables are read back by their own thread only; the table by everybody)
    if (S == 1 && !valid[0]) return;  // no barrier below: threads past the end of the batch are done
    double *const lp = leaf + base + tid;  // this thread's column of the leaf matrix
    const double *kv = lg_var + tid;
    const double *tv = lg_var + kr * COLS + tid;
    const int b0 = blockIdx.y * bases_per_block, b1 = min(n_bases, b0 + bases_per_block);
    // The tables are read one record ahead (every table ends with a spare record): a record's load is in flight while the
    // previous one is being worked on -- the kernel is bound by the latency of exactly these loads.
    int4 nh0 = *reinterpret_cast<const int4 *>(&bases[b0].nnz), nh1 = *reinterpret_cast<const int4 *>(&bases[b0].n_c);
    for (int ib = b0; ib < b1; ++ib) {
        const LgBasis *B = bases + ib;  // (uniform loads, 16 bytes at a time)
        const int4 h0 = nh0, h1 = nh1;
        nh0 = *reinterpret_cast<const int4 *>(&B[1].nnz), nh1 = *reinterpret_cast<const int4 *>(&B[1].n_c);
        const int nnz = h0.x, g0_first = h0.y, n_g0 = h0.z, c_first = h0.w, n_c = h1.x, leaf0 = h1.y, n_leaves = h1.z;
        LgG0 mg = g0[g0_first];                                              // first records of this momentum's lists
        int4 mc = *reinterpret_cast<const int4 *>(cst + c_first);
        double kq[S][3];
#pragma unroll
        for (int s = 0; s < S; ++s) kq[s][0] = kq[s][1] = kq[s][2] = 0.0;
        for (int n = 0; n < nnz; ++n) {
            const int4 tm = *reinterpret_cast<const int4 *>(&B->term[n]);
            const double cf = __hiloint2double(tm.y, tm.x);
            const double *kp = kv + tm.z * COLS;
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int c = 0; c < DIM; ++c) kq[s][c] = fma(kp[c * COLS + s * FDG_LG_THREADS], cf, kq[s][c]);
        }
        double q2[S], w[S], aw[S], ebeta[S], den[S], inv_den[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            q2[s] = 0.0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) q2[s] = fma(kq[s][c], kq[s][c], q2[s]);
            w[s] = q2[s] - kF2, aw[s] = fabs(w[s]);
            ebeta[s] = 0.0, den[s] = inv_den[s] = 1.0;
        }
        bool have_den = false;
        if (n_g0 > 0) {
            // ---- order-0 propagators of this momentum: green(tau, w, beta) = s exp(-|w| x) / (1 + exp(-|w| beta)), x in (0, beta],
            //      s = sign(tau).  With st = tau sign(w):  x = st if st > 0, else st + beta -- the four cases of the reference's
            //      formula (example/benchmark.jl:113-127), the same operands in the same operations.
            double sw[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                ebeta[s] = lg_exp_tab(-aw[s] * beta, tab);
                den[s] = 1.0 + ebeta[s];
                inv_den[s] = lg_rcp(den[s]);
                sw[s] = w[s] > 0.0 ? 1.0 : -1.0;
            }
            have_den = true;
            for (int i = g0_first; i < g0_first + n_g0; ++i) {
                const LgG0 m = mg;
                mg = g0[i + 1];
                const double *t_out = tv + (m.taus >> 16) * COLS, *t_in = tv + (m.taus & 0xffff) * COLS;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    double tau = t_out[s * FDG_LG_THREADS] - t_in[s * FDG_LG_THREADS];
                    if (tau == 0.0) tau = -1e-10;
                    const double st = __dmul_rn(tau, sw[s]);
                    const double x = st > 0.0 ? st : st + beta;
                    const double e = lg_exp_tab(-aw[s] * x, tab) * inv_den[s];
                    // sign(tau) copied onto e (e >= 0)
                    const double v = __hiloint2double(__double2hiint(e) | (__double2hiint(tau) & 0x80000000), __double2loint(e));
                    if (S == 1 || valid[s]) lp[m.off + s * FDG_LG_THREADS] = v;
                }
            }
        }
        if (n_c > 0) {
            // ---- leaves that do not depend on the times: 1, or the order-0 interaction 8 pi / invK with invK = 1 / (q^2 + lambda)
            double wv[S];
#pragma unroll
            for (int s = 0; s < S; ++s) wv[s] = 25.132741228718345 * (q2[s] + lambda);
            for (int i = c_first; i < c_first + n_c; ++i) {
                const int4 m = mc;
                mc = *reinterpret_cast<const int4 *>(cst + i + 1);
                const long long off = ((long long)m.y << 32) | (unsigned)m.x;
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if (S == 1 || valid[s]) lp[off + s * FDG_LG_THREADS] = m.z ? wv[s] : 1.0;
            }
        }
        for (int il = leaf0; il < leaf0 + n_leaves; ++il) {
            const LgLeaf m = leaves[il];
            if (m.type == 1 && !have_den) {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    ebeta[s] = lg_exp_tab(-aw[s] * beta, tab);
                    den[s] = 1.0 + ebeta[s];
                    inv_den[s] = lg_rcp(den[s]);
                }
                have_den = true;
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                double v = 1.0;
                if (m.type == 1) {
                    double tau = tv[m.tau_out * COLS + s * FDG_LG_THREADS] - tv[m.tau_in * COLS + s * FDG_LG_THREADS];
                    if (m.order == 0) {
                        if (tau == 0.0) tau = -1e-10;
                        const double x = tau > 0.0 ? (w[s] > 0.0 ? tau : beta - tau) : (w[s] > 0.0 ? tau + beta : -tau);
                        const double e = lg_exp_tab(-aw[s] * x, tab) * inv_den[s];
                        v = tau > 0.0 ? e : -e;
                    } else {
                        v = lg_green_derive(tau, w[s], beta, ebeta[s], den[s], m.order);
                    }
                } else if (m.type == 2) {
                    // 8 pi / invK * (lambda invK)^order with invK = 1 / (q2 + lambda)
                    const double sm = q2[s] + lambda;
                    v = m.order == 0 ? 25.132741228718345 * sm : (25.132741228718345 * sm) * lg_pow(lambda / sm, m.order);
                }
                if (S == 1 || valid[s]) lp[(long long)m.out * ld_leaf + s * FDG_LG_THREADS] = v;
            }
        }
    }
}

}  // namespace

struct fdg_leafgen {
    std::vector<LeafMeta> meta;
    std::vector<LgBasis> bases;  // leaves grouped by loop-basis vector (non-zero coefficients only)
    std::vector<LgLeaf> leaves;
    std::vector<LgG0> g0;        // the order-0 propagators, grouped like `bases`
    std::vector<LgCst> cst;      // leaves that do not depend on the times, grouped like `bases`
    // the generator specialised for this graph (fdg_lgjit.cpp): what it covers, its kernels per (device, wide offsets), and the
    // tables of the leaves it does not cover (counter-term orders >= 1), which the table-driven kernel then fills
    std::vector<fdg::LgJitBasis> jit_bases;
    std::vector<LgBasis> rest_bases;
    std::map<std::pair<int, int>, std::vector<cudaKernel_t>> jit_kernels;
    std::map<std::pair<int, int>, std::vector<cudaLibrary_t>> jit_libs;
    std::map<int, LgBasis *> d_rest;
    bool jit_failed = false;
    std::map<int, std::pair<LgBasis *, LgLeaf *>> d_tab;  // per device
    std::map<std::pair<int, long long>, std::pair<LgG0 *, LgCst *>> d_g0;  // per device and leading dimension of the leaf matrix
    int n_loops = 0, dim = 3, n_tau = 0;
    double kF = 0, beta = 0, lambda = 0;
    std::map<std::pair<int, cudaStream_t>, std::pair<double *, size_t>> d_leaf;  // per device and stream: sub-batch leaf matrix of the fused path
    std::map<int, std::pair<double *, size_t>> d_var;   // per device: staging of (K, T) chunks for the host path
    cudaStream_t streams[2] = {nullptr, nullptr};       // host path: copy stream, run stream (created on first use)
    cudaEvent_t events[4] = {nullptr, nullptr, nullptr, nullptr};  // [k] chunk k copied, [2 + k] buffer k consumed
    std::mutex mu;
    std::mutex host_mu;  // fdg_eval_generated_host: staging buffer, streams and events belong to one caller at a time
};

namespace {
// PTX of the generator specialised for this graph, assembled for sm_100a (host only)
int lg_jit_assemble(fdg_leafgen *g, bool wide, std::vector<fdg::JitSegment> &segs, std::string &err) {
    int budget = 0;
    if (const char *e = getenv("FDG_LG_JIT_BUDGET")) budget = atoi(e);
    int rc = fdg::lgjit_build(g->jit_bases, g->n_loops, g->dim, g->n_tau, g->kF * g->kF, g->beta, g->lambda, wide, budget, segs, err);
    if (rc != FDG_OK) return rc;
    std::vector<std::string> errs(segs.size());
    std::vector<int> rcs(segs.size(), FDG_OK);
    std::vector<std::thread> th;
    for (size_t q = 0; q < segs.size(); ++q)
        th.emplace_back([&, q] { rcs[q] = fdg::jit_assemble(segs[q].ptx, 3, segs[q].cubin, errs[q]); });
    for (auto &t : th) t.join();
    for (size_t q = 0; q < segs.size(); ++q)
        if (rcs[q] != FDG_OK) rc = rcs[q], err = errs[q];
    return rc;
}

int leafgen_launch(fdg_leafgen *g, const double *K, const double *T, int64_t ld_var, int64_t batch, double *leaf,
                   int64_t ld_leaf, cudaStream_t st) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    const int L = (int)g->meta.size();
    if (L == 0 || batch == 0) return FDG_OK;
    auto it = g->d_tab.find(dev);
    if (it == g->d_tab.end()) {
        LgBasis *db = nullptr;
        LgLeaf *dl = nullptr;
        CUDA_TRY(cudaMalloc((void **)&db, (g->bases.size() + 1) * sizeof(LgBasis)));  // (+ 1: the kernel reads one record ahead)
        CUDA_TRY(cudaMemset(db, 0, (g->bases.size() + 1) * sizeof(LgBasis)));
        CUDA_TRY(cudaMalloc((void **)&dl, g->leaves.size() * sizeof(LgLeaf)));
        CUDA_TRY(cudaMemcpy(db, g->bases.data(), g->bases.size() * sizeof(LgBasis), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dl, g->leaves.data(), g->leaves.size() * sizeof(LgLeaf), cudaMemcpyHostToDevice));
        it = g->d_tab.emplace(dev, std::make_pair(db, dl)).first;
    }
    // the table of the order-0 propagators carries row offsets: one copy per device and leading dimension
    auto ig = g->d_g0.find(std::make_pair(dev, (long long)ld_leaf));
    if (ig == g->d_g0.end()) {
        std::vector<LgG0> tab = g->g0;
        for (LgG0 &r : tab) r.off = (long long)r.out * (long long)ld_leaf;
        std::vector<LgCst> tc = g->cst;
        for (LgCst &r : tc) r.off = (long long)r.out * (long long)ld_leaf;
        LgG0 *dg = nullptr;
        LgCst *dc = nullptr;
        CUDA_TRY(cudaMalloc((void **)&dg, (tab.size() + 1) * sizeof(LgG0)));
        CUDA_TRY(cudaMalloc((void **)&dc, (tc.size() + 1) * sizeof(LgCst)));
        CUDA_TRY(cudaMemset(dg, 0, (tab.size() + 1) * sizeof(LgG0)));
        CUDA_TRY(cudaMemset(dc, 0, (tc.size() + 1) * sizeof(LgCst)));
        CUDA_TRY(cudaMemcpy(dg, tab.data(), tab.size() * sizeof(LgG0), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dc, tc.data(), tc.size() * sizeof(LgCst), cudaMemcpyHostToDevice));
        ig = g->d_g0.emplace(std::make_pair(dev, (long long)ld_leaf), std::make_pair(dg, dc)).first;
    }
    const LgG0 *dg0 = ig->second.first;
    const LgCst *dcst = ig->second.second;
    // ---- the generator specialised for this graph, where it can be built; the table-driven kernel below then only fills the
    //      leaves it does not cover
    bool use_jit = !g->jit_failed;
    if (const char *e = getenv("FDG_LG_JIT")) use_jit = use_jit && atoi(e) != 0;
    const LgBasis *d_bases = it->second.first;
    int nb = (int)g->bases.size();
    if (use_jit) {
        const int wide = (uint64_t)ld_leaf * 8 >= (1ull << 32) ? 1 : 0;
        auto jk = g->jit_kernels.find(std::make_pair(dev, wide));
        if (jk == g->jit_kernels.end()) {
            std::vector<fdg::JitSegment> segs;
            std::string err;
            int rc = lg_jit_assemble(g, wide != 0, segs, err);
            if (rc != FDG_OK) {
                g->jit_failed = true;  // (e.g. no PTX compiler library): the table-driven kernel does everything
                use_jit = false;
                if (getenv("FDG_LG_JIT_VERBOSE")) std::fprintf(stderr, "fdg_leafgen: specialised generator not built: %s\n", err.c_str());
            } else {
                std::vector<cudaKernel_t> ks;
                for (auto &sg : segs) {
                    cudaLibrary_t lib;
                    CUDA_TRY(cudaLibraryLoadData(&lib, sg.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
                    g->jit_libs[std::make_pair(dev, wide)].push_back(lib);
                    cudaKernel_t k;
                    CUDA_TRY(cudaLibraryGetKernel(&k, lib, sg.name.c_str()));
                    ks.push_back(k);
                }
                jk = g->jit_kernels.emplace(std::make_pair(dev, wide), std::move(ks)).first;
            }
        }
        if (use_jit) {
            const unsigned grid1 = (unsigned)((batch + 127) / 128);
            long long a_ld_var = ld_var, a_batch = batch, a_ld_leaf = ld_leaf;
            void *args[] = {(void *)&K, (void *)&T, &a_ld_var, &a_batch, (void *)&leaf, &a_ld_leaf};
            for (const cudaKernel_t k : jk->second) CUDA_TRY(cudaLaunchKernel((const void *)k, dim3(grid1), dim3(128), args, 0, st));
            if (g->rest_bases.empty()) return FDG_OK;
            auto ir = g->d_rest.find(dev);
            if (ir == g->d_rest.end()) {
                LgBasis *db = nullptr;
                CUDA_TRY(cudaMalloc((void **)&db, (g->rest_bases.size() + 1) * sizeof(LgBasis)));
                CUDA_TRY(cudaMemset(db, 0, (g->rest_bases.size() + 1) * sizeof(LgBasis)));
                CUDA_TRY(cudaMemcpy(db, g->rest_bases.data(), g->rest_bases.size() * sizeof(LgBasis), cudaMemcpyHostToDevice));
                ir = g->d_rest.emplace(dev, db).first;
            }
            d_bases = ir->second;
            nb = (int)g->rest_bases.size();
        }
    }
    // samples per thread (two share the decoding of the tables, but measure the same: 169 vs 170 M samples/s)
    int spt = 1;
    if (const char *e = getenv("FDG_LG_SPT")) spt = atoi(e) >= 2 ? 2 : 1;
    const int64_t cols = (int64_t)FDG_LG_THREADS * spt;
    // one block walks all the basis vectors of its samples: (K, T) are read once (FDG_LG_BASES_PER_BLOCK splits the walk
    // over several blocks -- experiments)
    int per_block = nb;
    if (const char *e = getenv("FDG_LG_BASES_PER_BLOCK")) per_block = std::max(1, atoi(e));
    dim3 grid((unsigned)((batch + cols - 1) / cols), (unsigned)((nb + per_block - 1) / per_block));
    const double kF2 = g->kF * g->kF;
    const size_t smem = ((size_t)(g->n_loops * g->dim + g->n_tau) * cols + 32) * sizeof(double);
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // (experiments: giving L1 more room for the tables at the price of resident blocks only loses -- 25 % shared memory
        // 79 M samples/s, 50 % 110 M, 100 % 174 M on Parquet vertex4 order 4: the kernel lives on the number of warps in flight)
        int carve = -1;
        if (const char *e = getenv("FDG_LG_CARVEOUT")) carve = atoi(e);
        if (carve >= 0) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        kern<<<grid, FDG_LG_THREADS, smem, st>>>(d_bases, nb, per_block, it->second.second, dg0, dcst, g->n_loops, g->n_tau, K, T, ld_var, batch, leaf,
                                                ld_leaf, kF2, g->beta, g->lambda);
        return FDG_OK;
    };
    int rc;
    if (g->dim == 3) rc = spt == 2 ? launch(fdg_leafgen2_kernel<3, 2>) : launch(fdg_leafgen2_kernel<3, 1>);
    else rc = spt == 2 ? launch(fdg_leafgen2_kernel<2, 2>) : launch(fdg_leafgen2_kernel<2, 1>);
    if (rc != FDG_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    return FDG_OK;
}
}  // namespace

extern "C" {

static int fdg_leafgen_create_impl(const fdg_leafgen_desc *d, fdg_leafgen_t *out) {
    if (!d || !out) return fail(FDG_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (d->n_leaves < 0 || d->n_basis < 0 || d->n_tau < 0) return fail(FDG_ERR_BAD_ARG, "negative size");
    if (d->n_loops < 0 || d->n_loops > FDG_LG_MAXLOOPS) return fail(FDG_ERR_UNSUPPORTED, "n_loops must be <= 8");
    if (d->dim != 2 && d->dim != 3) return fail(FDG_ERR_UNSUPPORTED, "dim must be 2 or 3");
    if (d->n_leaves > 0 && (!d->leaf_type || !d->leaf_order || !d->tau_in || !d->tau_out || !d->loop_index))
        return fail(FDG_ERR_BAD_ARG, "null leaf metadata");
    if (d->n_basis > 0 && !d->loop_basis) return fail(FDG_ERR_BAD_ARG, "null loop basis");
    fdg_leafgen *g = new (std::nothrow) fdg_leafgen();
    if (!g) return fail(FDG_ERR_BAD_ARG, "out of memory");
    g->n_loops = (int)d->n_loops, g->dim = (int)d->dim, g->n_tau = (int)d->n_tau;
    g->kF = d->kF, g->beta = d->beta, g->lambda = d->lambda;
    g->meta.resize((size_t)d->n_leaves);
    for (int64_t l = 0; l < d->n_leaves; ++l) {
        LeafMeta &m = g->meta[(size_t)l];
        std::memset(&m, 0, sizeof(m));
        m.type = d->leaf_type[l];
        if (m.type < 0 || m.type > 2) {
            delete g;
            return fail(FDG_ERR_UNSUPPORTED, "leaf type " + std::to_string(m.type) + " not implemented (example/benchmark.jl:76)");
        }
        if (m.type == 0) continue;
        m.order = d->leaf_order[2 * l + (m.type == 1 ? 0 : 1)];
        if (m.type == 1 && m.order > 5) {
            delete g;
            return fail(FDG_ERR_UNSUPPORTED, "Green's function derivative order > 5: \"not implemented!\" (example/benchmark.jl:107)");
        }
        if (m.order < 0) {
            delete g;
            return fail(FDG_ERR_BAD_ARG, "negative derivative order");
        }
        m.tau_in = d->tau_in[l], m.tau_out = d->tau_out[l];
        const int32_t bi = d->loop_index[l];
        if (m.tau_in < 0 || m.tau_in >= d->n_tau || m.tau_out < 0 || m.tau_out >= d->n_tau || bi < 0 || bi >= d->n_basis) {
            delete g;
            return fail(FDG_ERR_BAD_ARG, "leaf " + std::to_string(l) + ": time or loop-basis index out of range");
        }
        m.basis_id = bi;
        for (int64_t j = 0; j < d->n_loops; ++j) m.basis[j] = d->loop_basis[(size_t)bi * (size_t)d->n_loops + (size_t)j];
    }
    for (int64_t l = 0; l < d->n_leaves; ++l) {
        g->meta[(size_t)l].out = (int32_t)l;
        if (g->meta[(size_t)l].type == 0) g->meta[(size_t)l].basis_id = -1;
    }
    // leaves that carry the same momentum are handled back to back: |K . basis|^2 and the momentum-only factors are
    // computed once per basis vector (404 of them for the 984 leaves of Parquet vertex4 order 4)
    std::stable_sort(g->meta.begin(), g->meta.end(), [](const LeafMeta &a, const LeafMeta &b) { return a.basis_id < b.basis_id; });
    for (size_t i = 0; i < g->meta.size();) {
        size_t j = i;
        while (j < g->meta.size() && g->meta[j].basis_id == g->meta[i].basis_id) ++j;
        LgBasis B;
        std::memset(&B, 0, sizeof(B));
        B.leaf0 = (int32_t)g->leaves.size();
        B.n_leaves = (int32_t)(j - i);
        if (g->meta[i].basis_id >= 0)
            for (int q = 0; q < g->n_loops; ++q)
                if (g->meta[i].basis[q] != 0.0) {
                    B.term[B.nnz].off = q * g->dim;  // first of the DIM rows of loop momentum q in the staged variables
                    B.term[B.nnz].coef = g->meta[i].basis[q];
                    B.nnz++;
                }
        B.g0_first = (int32_t)g->g0.size();
        B.c_first = (int32_t)g->cst.size();
        for (size_t q = i; q < j; ++q) {
            if (g->meta[q].type == 0 || (g->meta[q].type == 2 && g->meta[q].order == 0)) {
                g->cst.push_back({0, g->meta[q].type == 2 ? 1 : 0, g->meta[q].out});
                B.n_c++;
                B.n_leaves--;
                continue;
            }
            if (g->meta[q].type == 1 && g->meta[q].order == 0 && g->meta[q].tau_in < 65536 && g->meta[q].tau_out < 32768) {
                g->g0.push_back({g->meta[q].tau_in | (g->meta[q].tau_out << 16), g->meta[q].out, 0});  // the tight loop of the kernel
                B.n_g0++;
                B.n_leaves--;
                continue;
            }
            LgLeaf lf;
            std::memset(&lf, 0, sizeof(lf));
            lf.type = g->meta[q].type, lf.order = g->meta[q].order, lf.tau_in = g->meta[q].tau_in, lf.tau_out = g->meta[q].tau_out;
            lf.out = g->meta[q].out;
            g->leaves.push_back(lf);
        }
        g->bases.push_back(B);
        {
            fdg::LgJitBasis J;
            J.nnz = B.nnz;
            for (int q = 0; q < B.nnz; ++q) J.idx[q] = B.term[q].off / std::max(g->dim, 1), J.coef[q] = B.term[q].coef;
            for (int q = 0; q < B.n_g0; ++q) {
                const LgG0 &r = g->g0[(size_t)(B.g0_first + q)];
                J.g0.push_back({r.taus & 0xffff, r.taus >> 16, r.out});
            }
            for (int q = 0; q < B.n_c; ++q) {
                const LgCst &r = g->cst[(size_t)(B.c_first + q)];
                (r.kind ? J.w0_out : J.one_out).push_back(r.out);
            }
            g->jit_bases.push_back(std::move(J));
            if (B.n_leaves > 0) {
                LgBasis Rb = B;
                Rb.n_g0 = Rb.n_c = 0;
                g->rest_bases.push_back(Rb);
            }
        }
        i = j;
    }
    *out = g;
    return FDG_OK;
}

static int fdg_leafgen_jit_prepare_impl(fdg_leafgen_t g, int32_t wide, int32_t index, int64_t *out, int32_t n_out, const char **ptx) {
    if (!g || n_out < 0 || (!out && n_out > 0)) return fail(FDG_ERR_BAD_ARG, "bad argument");
    std::lock_guard<std::mutex> lock(g->mu);
    static thread_local std::string keep;
    std::vector<fdg::JitSegment> segs;
    std::string err;
    int rc = lg_jit_assemble(g, wide != 0, segs, err);
    if (rc != FDG_OK) return fail(rc, err);
    int64_t code = 0, lines = 0, biggest = 0;
    for (auto &sg : segs) code += (int64_t)sg.cubin.size(), lines += sg.n_stmts, biggest = std::max<int64_t>(biggest, (int64_t)sg.cubin.size());
    int64_t covered = 0;
    for (auto &b : g->jit_bases) covered += (int64_t)(b.g0.size() + b.w0_out.size() + b.one_out.size());
    const int64_t vals[5] = {(int64_t)segs.size(), code, lines, covered, biggest};
    for (int32_t i = 0; i < n_out && i < 5; ++i) out[i] = vals[i];
    if (ptx) {
        if (index < 0 || index >= (int32_t)segs.size()) return fail(FDG_ERR_BAD_ARG, "kernel index out of range");
        keep = segs[(size_t)index].ptx;
        *ptx = keep.c_str();
    }
    return FDG_OK;
}

static int fdg_leafgen_destroy_impl(fdg_leafgen_t g) {
    if (!g) return FDG_OK;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess) {
        for (auto &kv : g->d_leaf) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second.first);
        }
        for (auto &kv : g->d_tab) {
            cudaSetDevice(kv.first);
            cudaFree(kv.second.first);
            cudaFree(kv.second.second);
        }
        for (auto &kv : g->d_var) {
            cudaSetDevice(kv.first);
            cudaFree(kv.second.first);
        }
        for (auto &kv : g->d_rest) {
            cudaSetDevice(kv.first);
            cudaFree(kv.second);
        }
        for (auto &kv : g->jit_libs) {
            cudaSetDevice(kv.first.first);
            for (auto lib : kv.second) cudaLibraryUnload(lib);
        }
        for (auto &kv : g->d_g0) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second.first);
            cudaFree(kv.second.second);
        }
        cudaSetDevice(cur);
        for (int i = 0; i < 2; ++i)
            if (g->streams[i]) cudaStreamDestroy(g->streams[i]);
        for (int i = 0; i < 4; ++i)
            if (g->events[i]) cudaEventDestroy(g->events[i]);
    }
    delete g;
    return FDG_OK;
}

static int fdg_leafgen_fill_impl(fdg_leafgen_t g, const double *K, const double *T, int64_t ld_var, int64_t batch, double *leaf,
                     int64_t ld_leaf, void *stream) {
    if (!g) return fail(FDG_ERR_BAD_ARG, "null handle");
    if (batch < 0) return fail(FDG_ERR_BAD_ARG, "negative batch");
    if (batch == 0 || g->meta.empty()) return FDG_OK;
    if (!K || !T || !leaf) return fail(FDG_ERR_BAD_ARG, "null pointer");
    if (ld_var < batch || ld_leaf < batch) return fail(FDG_ERR_BAD_ARG, "leading dimension < batch");
    std::lock_guard<std::mutex> lock(g->mu);
    return leafgen_launch(g, K, T, ld_var, batch, leaf, ld_leaf, static_cast<cudaStream_t>(stream));
}

static int fdg_eval_generated_accumulate_impl(fdg_handle h, fdg_leafgen_t g, const double *K, const double *T, int64_t ld_var,
                                  int64_t batch, double *acc, void *stream) {
    if (!h || !g) return fail(FDG_ERR_BAD_ARG, "null handle");
    if (h->low.dtype != FDG_F64) return fail(FDG_ERR_UNSUPPORTED, "generated leaves are Float64");
    if ((int64_t)g->meta.size() != h->low.L) return fail(FDG_ERR_BAD_ARG, "leaf generator and program have different numbers of leaves");
    if (batch < 0) return fail(FDG_ERR_BAD_ARG, "negative batch");
    if (batch == 0) return FDG_OK;
    if (!K || !T || (!acc && h->low.R > 0)) return fail(FDG_ERR_BAD_ARG, "null pointer");
    if (ld_var < batch) return fail(FDG_ERR_BAD_ARG, "ld_var < batch");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // sub-batches: the leaf matrix of a sub-batch lives in a scratch buffer of about 8 GiB (>> L2), never more; big sub-batches
    // are whole waves of 256-sample tiles (one per SM) so that the persistent graph kernels end together
    const int64_t L = std::max<int64_t>(h->low.L, 1);
    double gb = 8.0;
    if (const char *e = getenv("FDG_LEAFGEN_GB")) gb = atof(e);
    int64_t sub = std::max<int64_t>(4096, (int64_t)(gb * (double)(1 << 30)) / (8 * L)) / 1024 * 1024;
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t wave = (int64_t)256 * std::max(sms, 1);
    if (sub >= 4 * wave) sub = sub / wave * wave;
    sub = std::min<int64_t>(sub, (batch + 1023) / 1024 * 1024);
    double *buf = nullptr;
    {
        std::lock_guard<std::mutex> lock(g->mu);
        auto &slot = g->d_leaf[std::make_pair(dev, st)];  // calls on different streams may run at the same time
        const size_t need = (size_t)L * (size_t)sub * sizeof(double);
        if (need > slot.second) {
            if (slot.first) CUDA_TRY(cudaFree(slot.first));
            slot = {nullptr, 0};
            CUDA_TRY(cudaMalloc((void **)&slot.first, need));
            slot.second = need;
        }
        buf = slot.first;
    }
    for (int64_t b0 = 0; b0 < batch; b0 += sub) {
        const int64_t nb = std::min<int64_t>(sub, batch - b0);
        int rc;
        {
            std::lock_guard<std::mutex> lock(g->mu);
            rc = leafgen_launch(g, K + b0, T + b0, ld_var, nb, buf, sub, st);
        }
        if (rc != FDG_OK) return rc;
        h->launches++;
        rc = do_eval(h, buf, sub, acc, 0, nb, st, true);
        if (rc != FDG_OK) return rc;
    }
    return FDG_OK;
}

static int fdg_eval_generated_host_impl(fdg_handle h, fdg_leafgen_t g, const double *K_host, const double *T_host, int64_t ld_var,
                            int64_t batch, double *acc_host) {
    if (!h || !g) return fail(FDG_ERR_BAD_ARG, "null handle");
    if (batch < 0) return fail(FDG_ERR_BAD_ARG, "negative batch");
    if (!K_host || !T_host || (!acc_host && h->low.R > 0)) return fail(FDG_ERR_BAD_ARG, "null pointer");
    if (ld_var < batch) return fail(FDG_ERR_BAD_ARG, "ld_var < batch");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    const int64_t kr = (int64_t)g->n_loops * g->dim, rows = kr + g->n_tau;
    const int64_t R = h->low.R;
    // chunks of (K, T) are copied on one stream while the previous chunk is generated and evaluated on another
    // (about half a million samples: thirteen whole waves of 256-sample tiles on a 148-SM device)
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t wave = (int64_t)256 * std::max(sms, 1);
    const int64_t chunk = std::min<int64_t>(std::max<int64_t>(batch, 1), std::max<int64_t>(1, (1 << 19) / wave) * wave);
    std::lock_guard<std::mutex> host_lock(g->host_mu);
    double *d_var = nullptr;
    cudaStream_t s_copy = nullptr, s_run = nullptr;
    {
        std::lock_guard<std::mutex> lock(g->mu);
        auto &slot = g->d_var[dev];
        const size_t need = (2 * (size_t)rows * (size_t)chunk + (size_t)std::max<int64_t>(R, 1)) * sizeof(double);
        if (need > slot.second) {
            if (slot.first) CUDA_TRY(cudaFree(slot.first));
            slot = {nullptr, 0};
            CUDA_TRY(cudaMalloc((void **)&slot.first, need));
            slot.second = need;
        }
        d_var = slot.first;
        for (int i = 0; i < 2; ++i)
            if (!g->streams[i]) CUDA_TRY(cudaStreamCreateWithFlags(&g->streams[i], cudaStreamNonBlocking));
        for (int i = 0; i < 4; ++i)
            if (!g->events[i]) CUDA_TRY(cudaEventCreateWithFlags(&g->events[i], cudaEventDisableTiming));
        s_copy = g->streams[0], s_run = g->streams[1];
    }
    double *d_acc = d_var + 2 * (size_t)rows * (size_t)chunk;
    CUDA_TRY(cudaMemsetAsync(d_acc, 0, (size_t)std::max<int64_t>(R, 1) * 8, s_run));
    int k = 0;
    for (int64_t c0 = 0; c0 < batch; c0 += chunk, k ^= 1) {
        const int64_t nb = std::min<int64_t>(chunk, batch - c0);
        double *buf = d_var + (size_t)k * (size_t)rows * (size_t)chunk;
        // buffer k is free once the evaluation that read it two chunks ago has finished
        CUDA_TRY(cudaStreamWaitEvent(s_copy, g->events[2 + k], 0));
        CUDA_TRY(cudaMemcpy2DAsync(buf, (size_t)chunk * 8, K_host + c0, (size_t)ld_var * 8, (size_t)nb * 8, (size_t)kr, cudaMemcpyHostToDevice, s_copy));
        CUDA_TRY(cudaMemcpy2DAsync(buf + (size_t)kr * (size_t)chunk, (size_t)chunk * 8, T_host + c0, (size_t)ld_var * 8, (size_t)nb * 8,
                                   (size_t)g->n_tau, cudaMemcpyHostToDevice, s_copy));
        CUDA_TRY(cudaEventRecord(g->events[k], s_copy));
        CUDA_TRY(cudaStreamWaitEvent(s_run, g->events[k], 0));
        int rc = fdg_eval_generated_accumulate(h, g, buf, buf + (size_t)kr * (size_t)chunk, chunk, nb, d_acc, s_run);
        if (rc != FDG_OK) {
            const std::string msg = g_err;
            cudaStreamSynchronize(s_run);
            cudaStreamSynchronize(s_copy);
            return fail(rc, msg);
        }
        CUDA_TRY(cudaEventRecord(g->events[2 + k], s_run));
    }
    if (R > 0) CUDA_TRY(cudaMemcpyAsync(acc_host, d_acc, (size_t)R * 8, cudaMemcpyDeviceToHost, s_run));
    CUDA_TRY(cudaStreamSynchronize(s_run));
    CUDA_TRY(cudaStreamSynchronize(s_copy));
    return FDG_OK;
}

}  // extern "C"

// ---- the exported entry points: the implementations above behind an exception barrier ----------------------------------
extern "C" {
int fdg_compile(const fdg_graph_desc *graph, const fdg_options *opts, fdg_handle *out) {
    return guarded([&] { return fdg_compile_impl(graph, opts, out); });
}
int fdg_graph_write(const fdg_graph_desc *g, const char *path) {
    return guarded([&] { return fdg_graph_write_impl(g, path); });
}
int fdg_compile_file(const char *path, const fdg_options *opts, fdg_handle *out) {
    return guarded([&] { return fdg_compile_file_impl(path, opts, out); });
}
int fdg_jit_prepare(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int32_t *n_kernels, int32_t *n_cross, int64_t *cubin_bytes) {
    return guarded([&] { return fdg_jit_prepare_impl(h, samples_per_thread, accumulate, n_kernels, n_cross, cubin_bytes); });
}
int fdg_jit_info(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int64_t *out, int32_t n_out) {
    return guarded([&] { return fdg_jit_info_impl(h, samples_per_thread, accumulate, out, n_out); });
}
int fdg_pipeline_prepare(fdg_handle h, int32_t accumulate, int32_t n_sm, int32_t what, int64_t *out, int32_t n_out) {
    return guarded([&] { return fdg_pipeline_prepare_impl(h, accumulate, n_sm, what, out, n_out); });
}
int fdg_pipeline_stats(fdg_handle h, void *stream, int64_t *out, int32_t n_out) {
    return guarded([&] { return fdg_pipeline_stats_impl(h, stream, out, n_out); });
}
int fdg_jit_ptx(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int32_t index, const char **ptx, const char **ptxas_log) {
    return guarded([&] { return fdg_jit_ptx_impl(h, samples_per_thread, accumulate, index, ptx, ptxas_log); });
}
int fdg_destroy(fdg_handle h) {
    return guarded([&] { return fdg_destroy_impl(h); });
}
int fdg_stats(fdg_handle h, fdg_stats_t *out) {
    return guarded([&] { return fdg_stats_impl(h, out); });
}
int fdg_leafmap(fdg_handle h, int32_t *leaf_node) {
    return guarded([&] { return fdg_leafmap_impl(h, leaf_node); });
}
int fdg_eval(fdg_handle h, const void *leaf, int64_t ld_leaf, void *root, int64_t ld_root, int64_t batch, void *stream) {
    return guarded([&] { return fdg_eval_impl(h, leaf, ld_leaf, root, ld_root, batch, stream); });
}
int fdg_eval_accumulate(fdg_handle h, const void *leaf, int64_t ld_leaf, int64_t batch, double *acc, void *stream) {
    return guarded([&] { return fdg_eval_accumulate_impl(h, leaf, ld_leaf, batch, acc, stream); });
}
int fdg_eval_host(fdg_handle h, const void *leaf_host, int64_t ld_leaf, void *root_host, int64_t ld_root, int64_t batch) {
    return guarded([&] { return fdg_eval_host_impl(h, leaf_host, ld_leaf, root_host, ld_root, batch); });
}
int fdg_set_launch(fdg_handle h, int32_t threads, int32_t samples_per_thread, int32_t blocks_per_sm) {
    return guarded([&] { return fdg_set_launch_impl(h, threads, samples_per_thread, blocks_per_sm); });
}
int fdg_comm_unique_id(void *id128) {
    return guarded([&] { return fdg_comm_unique_id_impl(id128); });
}
int fdg_comm_init(fdg_comm_t *out, int32_t nranks, int32_t rank, const void *id128) {
    return guarded([&] { return fdg_comm_init_impl(out, nranks, rank, id128); });
}
int fdg_comm_destroy(fdg_comm_t c) {
    return guarded([&] { return fdg_comm_destroy_impl(c); });
}
int fdg_allreduce(fdg_comm_t c, double *acc, int64_t n, void *stream) {
    return guarded([&] { return fdg_allreduce_impl(c, acc, n, stream); });
}
int fdg_leafgen_create(const fdg_leafgen_desc *d, fdg_leafgen_t *out) {
    return guarded([&] { return fdg_leafgen_create_impl(d, out); });
}
int fdg_leafgen_jit_prepare(fdg_leafgen_t g, int32_t wide, int32_t index, int64_t *out, int32_t n_out, const char **ptx) {
    return guarded([&] { return fdg_leafgen_jit_prepare_impl(g, wide, index, out, n_out, ptx); });
}
int fdg_leafgen_destroy(fdg_leafgen_t g) {
    return guarded([&] { return fdg_leafgen_destroy_impl(g); });
}
int fdg_leafgen_fill(fdg_leafgen_t g, const double *K, const double *T, int64_t ld_var, int64_t batch, double *leaf, int64_t ld_leaf, void *stream) {
    return guarded([&] { return fdg_leafgen_fill_impl(g, K, T, ld_var, batch, leaf, ld_leaf, stream); });
}
int fdg_eval_generated_accumulate(fdg_handle h, fdg_leafgen_t g, const double *K, const double *T, int64_t ld_var, int64_t batch, double *acc, void *stream) {
    return guarded([&] { return fdg_eval_generated_accumulate_impl(h, g, K, T, ld_var, batch, acc, stream); });
}
int fdg_eval_generated_host(fdg_handle h, fdg_leafgen_t g, const double *K_host, const double *T_host, int64_t ld_var, int64_t batch, double *acc_host) {
    return guarded([&] { return fdg_eval_generated_host_impl(h, g, K_host, T_host, ld_var, batch, acc_host); });
}
int fdg_probe_fp64(int32_t fma, int64_t iters, void *stream, double *sink_device, int64_t *ops_per_launch) {
    return guarded([&] {
        if (iters < 1 || !sink_device) return fail(FDG_ERR_BAD_ARG, "bad argument");
        int dev = 0, sms = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int blocks = sms * 8;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (fma) fdg_fp64_probe<true><<<blocks, 256, 0, st>>>((long long)iters, sink_device);
        else fdg_fp64_probe<false><<<blocks, 256, 0, st>>>((long long)iters, sink_device);
        CUDA_TRY(cudaGetLastError());
        if (ops_per_launch) *ops_per_launch = (int64_t)blocks * 256 * iters * 16;  // instructions x lanes (an FMA counts once)
        return (int)FDG_OK;
    });
}
}  // extern "C"
