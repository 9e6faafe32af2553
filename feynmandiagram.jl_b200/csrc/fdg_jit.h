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
    bool persistent = false;  // single accumulate kernel run as a grid-stride loop (per-thread running sums)
    std::vector<JitSegment> seg;
};

// linearise the emitted function in fold order, cut it into segments of `seg_ops` operations, write their PTX
// wide_strides: row offsets need 64 bits (a leading dimension of 4 GiB or more)
int jit_plan(const Lowered &low, int spt, bool acc, int seg_ops, bool wide_strides, bool fma, JitPlan &plan, std::string &err);
// assemble every segment with the PTX compiler library (no GPU needed), segments in parallel
int jit_compile(JitPlan &plan, std::string &err);

}  // namespace fdg
#endif
