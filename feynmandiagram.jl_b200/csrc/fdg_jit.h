// fdg_jit.h -- specialised back end: statements -> PTX -> sm_100a cubin (see fdg_jit.cpp)
#ifndef FDG_JIT_H
#define FDG_JIT_H
#include <string>
#include <vector>

#include "fdg_lower.h"

namespace fdg {

struct JitSegment {
    std::string name;         // kernel entry name
    std::string ptx;          // generated PTX (one .entry)
    std::vector<char> cubin;  // assembled for sm_100a
    std::string info;         // ptxas verbose log (registers, spills)
    int n_stmts = 0;
};

struct JitPlan {
    int spt = 2;        // samples per thread
    bool acc = false;   // accumulate (per-warp partial sums) instead of per-sample roots
    bool fma = false;   // opt-in: multiplies may be contracted into the adds that read them (not bit-identical)
    int32_t n_cross = 0;  // rows of the cross-segment buffer (rows are reused once their last reader has run)
    int32_t n_cross_values = 0;  // values that cross a kernel boundary
    int64_t leaf_loads = 0, cross_loads = 0, cross_stores = 0;  // global loads / stores per sample over all kernels
    int64_t max_code_bytes = 0;  // machine code of the largest kernel (the instruction cache holds 128 KB)
    int64_t refetch_loads = 0;   // rows fetched a second time within a kernel (served by L2; not in leaf_loads / cross_loads)
    int64_t fp64_instr = 0;      // FP64 arithmetic instructions per sample the kernels execute (folded negations are none)
    bool uses_cse = false;       // planned from the program with common sub-expressions merged
    bool persistent = false;  // single accumulate kernel run as a grid-stride loop (per-thread running sums)
    // ---- bulk form (DESIGN.md section 4b'): persistent warp-specialised kernels.  Blocks of 384 threads, one per SM: eight
    // consumer warps (256 samples per tile, 240 registers after setmaxnreg) run the straight-line arithmetic and read their
    // input rows from a shared-memory ring; the four warps of a producer warpgroup (24 registers) fill the ring with
    // cp.async.bulk row copies (2 KB per row and tile) signalled through mbarriers, following a row table.  No address
    // arithmetic, no cp.async and no wait-group bookkeeping is left in the consumers' instruction stream.
    bool bulk = false;
    int bulk_smem = 0;   // dynamic shared memory of the kernels (barriers + ring + running sums of the roots)
    std::vector<JitSegment> seg;
    // ---- pipeline form: ONE kernel, one resident block per SM; the blocks of stage k run only
    // the code of segment k (it stays in that SM's instruction cache) and tiles of 32 samples flow from stage to stage
    // through L2: cross rows live in a ring of `window` tile slots, progress[tile] counts the stages a tile has passed.
    bool pipeline = false;
    int n_sm = 0;                    // blocks of the kernel (= SMs of the device the plan was made for)
    int n_pass = 0;                  // kernels run one after the other; stage s belongs to pass s / stages_per_pass
    int stages_per_pass = 0;         // = instruction-cache groups of the device
    std::vector<int> stage_blocks;   // SMs of the group each stage runs on
    std::vector<int64_t> stage_cost; // issue-cycle estimate of one tile in each stage (what the split is based on)
    std::vector<int64_t> stage_estimate; // the same as the cuts saw it (per-operation estimate, weighted)
    std::vector<int32_t> stage_start; // first operation of each stage, and the total as the last entry
    int32_t n_boundary = 0;          // rows of the buffer for values that cross from one pass to a later one ([row][sample])
    std::vector<std::string> dispatch_ptx;        // per pass: the entry kernel, SM -> stage function
    std::vector<std::vector<char>> linked;        // per pass: stages + entry linked into one cubin (nvJitLink)
    int ring_bytes = 0;              // dynamic shared memory of the kernel
};

struct PipeOptions {
    // SMs that share an instruction cache (a GPC on B200: 12 to 20 SMs, measured by tools/icache_probe.py) must run the same
    // stage: 128 KB of code per group is all there is.  `groups[g]` = SMs of group g; a pass of the pipeline has one stage per
    // group, a program that needs more code than the groups hold runs as several passes (kernels) one after the other.
    std::vector<int> groups = {12, 18, 18, 20, 20, 20, 20, 20};
    int threads = 256;     // threads per block (one block per SM)
    // profile-guided re-cut: the stage boundaries of a previous plan (operation indices, n + 1 entries) and the measured
    // time per estimated cost of each of its stages; the cost of an operation is scaled by the weight of the stage it was in
    std::vector<int32_t> prev_start;
    std::vector<double> weight;
    // ... or the measured time of each stage of that plan (any unit): the weights are then derived from it
    std::vector<double> measured;
    int n_sm() const {
        int n = 0;
        for (const int g : groups) n += g;
        return n;
    }
};

// Modelled time of one sample on a B200, ns: a smooth maximum of the plan's memory traffic at the rate such kernels reach and
// of its FP64 instructions (plus the cost of spilled registers, once the kernels are assembled) at the rate they reach;
// constants and calibration in fdg_jit.cpp.  Used to choose between plans, not reported as a result.
double jit_model_ns(const JitPlan &plan, int bytes_per_element);
// bytes of spill stores + spill loads per thread over all kernels of an assembled plan (from the ptxas logs)
double jit_spill_bytes(const JitPlan &plan);

// linearise the emitted function in fold order, cut it into segments of `seg_ops` operations, write their PTX
// wide_strides: row offsets need 64 bits (a leading dimension of 4 GiB or more)
// pipe != nullptr: pipeline form (stage functions + entry kernel, linked by jit_compile).
// merged != nullptr: SCOPED merging of common sub-expressions.  `merged` is the same program lowered with merging (its
// `canon` says which statements are copies of which, its sharing decides the order of the roots); `low` is evaluated, but a
// copy reads the value of its class instead of recomputing it when that value was computed at most `scope` operations
// earlier -- close enough to sit in the same kernel, so that merging saves arithmetic without sending more values through
// memory (scope <= 0: about two thirds of a kernel).
int jit_plan(const Lowered &low, int spt, bool acc, int seg_ops, bool wide_strides, bool fma, JitPlan &plan, std::string &err,
             const PipeOptions *pipe = nullptr, const Lowered *merged = nullptr, int scope = 0, bool bulk = false);
// ---- leaf generator specialised per graph (fdg_lgjit.cpp) ----
struct LgJitBasis {  // the leaves that carry one loop-basis vector (momentum)
    struct G0 {
        int32_t tau_in, tau_out, out;
    };
    int32_t nnz = 0;
    int32_t idx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double coef[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<G0> g0;            // order-0 Green's functions: times and row of the leaf matrix
    std::vector<int32_t> w0_out;   // order-0 interactions: rows
    std::vector<int32_t> one_out;  // constant leaves (value 1): rows
};
// straight-line PTX kernels (fdg_lg0, fdg_lg1, ...) that fill the rows of the leaves above for one sample per thread;
// params of every kernel: K, T, ld_var, batch, leaf, ld_leaf.  wide: ld_leaf * 8 does not fit 32 bits.
int lgjit_build(const std::vector<LgJitBasis> &bases, int n_loops, int dim, int n_tau, double kF2, double beta, double lambda, bool wide,
                int budget, std::vector<JitSegment> &out, std::string &err);

// bias of the root ordering towards roots over recently read leaves, for the plans made by this thread from now on
// (0 = none; see order_roots)
void jit_set_root_leaf_bias(double w, long window);
// PTX text -> sm_100a cubin with the PTX compiler library (no GPU needed)
int jit_assemble(const std::string &ptx, int opt_level, std::vector<char> &cubin, std::string &err);
// assemble every segment with the PTX compiler library (no GPU needed), segments in parallel
int jit_compile(JitPlan &plan, std::string &err);

}  // namespace fdg
#endif
