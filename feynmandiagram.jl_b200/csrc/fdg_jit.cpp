// fdg_jit.cpp -- the specialised back end: the emitted function as hand-scheduled-by-ptxas straight-line sm_100a code.
//
// Where the VM (fdg_vm.cuh) interprets packets, this back end writes the statements of the reference's emitted
// function (src/backend/static.jl:98-133, one `g<ID> = ...` per node) directly as PTX -- `mul.rn.f64` / `add.rn.f64`
// only, in the emitter's left-fold order, never fused -- and assembles it for sm_100a with the PTX compiler
// library shipped in the CUDA toolkit (libnvptxcompiler_static).  It is the direct analogue of the reference's
// emit-source design (to_julia_str -> RuntimeGeneratedFunction), with the sample index as the thread index:
//   * every value of a kernel lives in a register (ptxas allocates; 255 registers per thread), nothing is decoded
//     at run time;
//   * big programs are cut into kernels of about 4000 machine instructions (what the instruction cache holds), the
//     roots ordered so that few values are live across a cut, each cut placed where the fewest are; a value defined
//     in one kernel and read in a later one travels through a per-launch "cross" buffer laid out [row][thread]
//     (coalesced), rows reused once their last reader has run;
//   * inputs (leaf rows of the batch-major leaf matrix, cross rows) are streamed global -> shared: per thread by
//     cp.async, a ring of 32 rows ahead of the arithmetic with compile-time wait counts ("ring form"), or -- big
//     batches -- per block by cp.async.bulk from a producer warpgroup into an mbarrier-signalled ring that persistent
//     consumer warps read ("bulk form", DESIGN.md section 4b'); x * (-1.0) is a folded negation;
//   * equal sub-expressions are merged where the copies sit within one kernel (scoped merging); in the bulk form rows
//     whose next use is far away are fetched again (from L2) instead of being held in registers;
//   * roots are stored per sample (eval) or shuffle-reduced per warp into per-warp partial sums (accumulate);
//     a single small accumulate kernel runs as a grid-stride loop with per-thread running sums.
// Every one of these decisions is bit-neutral: each statement keeps its own left fold.
#include "fdg_jit.h"

#include <nvJitLink.h>
#include <nvPTXCompiler.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <sstream>
#include <thread>
#include <unordered_map>

namespace fdg {
namespace {

std::string dimm(double f) {  // exact double immediate
    uint64_t u;
    std::memcpy(&u, &f, 8);
    char buf[32];
    std::snprintf(buf, sizeof(buf), "0d%016llX", (unsigned long long)u);
    return buf;
}

struct Emitter {
    const Lowered &low;
    const int S;      // f64 registers per value: samples per thread (1 or 2), or 2 = (re, im) of ONE ComplexF64 sample
    const bool acc;   // accumulate mode
    bool cplx = false;
    const char *rnd = ".rn";  // "" in the opt-in contraction mode: ptxas may then fuse a multiply into the add that reads it
    bool persistent = false;  // accumulate mode, single kernel: grid-stride loop with per-thread running sums
    int racc0 = -1;           // first register of the per-root running sums (persistent mode)
    std::ostringstream os;
    // pipeline form: per-thread running sums of the roots in shared memory, [entry][thread] with a padded row, %r16 = the
    // thread's column; `sacc_pos[j]` = column of the partial-sum row (root, or 2 root + re/im) that entry j belongs to
    int sacc_cap = 0, sacc_stride = 0;
    std::vector<int32_t> sacc_pos;
    int nfd = 0, nrd = 32, np = 8, nr = 24;
    Emitter(const Lowered &l, int s, bool a) : low(l), S(s), acc(a) {}

    int new_val() {
        const int r = nfd;
        nfd += S;
        return r;
    }
    std::string fd(int r, int i) const { return "%fd" + std::to_string(r + i); }

    void binop(const char *op, int dst, int a, int b) {
        if (cplx && op[0] == 'm') {
            // Julia *(z::Complex, w::Complex) = Complex(re(z)re(w) - im(z)im(w), re(z)im(w) + im(z)re(w)), nothing fused
            const int t = nfd;
            nfd += 4;
            os << "\tmul" << rnd << ".f64 %fd" << t << ", " << fd(a, 0) << ", " << fd(b, 0) << ";\n";
            os << "\tmul" << rnd << ".f64 %fd" << t + 1 << ", " << fd(a, 1) << ", " << fd(b, 1) << ";\n";
            os << "\tmul" << rnd << ".f64 %fd" << t + 2 << ", " << fd(a, 0) << ", " << fd(b, 1) << ";\n";
            os << "\tmul" << rnd << ".f64 %fd" << t + 3 << ", " << fd(a, 1) << ", " << fd(b, 0) << ";\n";
            os << "\tsub" << rnd << ".f64 " << fd(dst, 0) << ", %fd" << t << ", %fd" << t + 1 << ";\n";
            os << "\tadd" << rnd << ".f64 " << fd(dst, 1) << ", %fd" << t + 2 << ", %fd" << t + 3 << ";\n";
            return;
        }
        for (int i = 0; i < S; ++i) os << "\t" << op << rnd << ".f64 " << fd(dst, i) << ", " << fd(a, i) << ", " << fd(b, i) << ";\n";
    }
    // Complex z^n, n >= 4: Base.power_by_squaring (base/intfuncs.jl), unrolled for the known exponent
    int cpow(int x0, unsigned n) {
        int x = x0;
        auto sq = [&](int v) {
            const int r = new_val();
            binop("mul", r, v, v);
            return r;
        };
        int t = __builtin_ctz(n) + 1;
        n >>= t;
        while (--t > 0) x = sq(x);
        int y = x;
        while (n > 0) {
            t = __builtin_ctz(n) + 1;
            n >>= t;
            while (--t >= 0) x = sq(x);
            const int r = new_val();
            binop("mul", r, y, x);
            y = r;
        }
        return y;
    }
    void scale(int dst, int a, double f) {
        for (int i = 0; i < S; ++i) os << "\tmul" << rnd << ".f64 " << fd(dst, i) << ", " << fd(a, i) << ", " << dimm(f) << ";\n";
    }
    void neg(int dst, int a) {
        for (int i = 0; i < S; ++i) os << "\tneg.f64 " << fd(dst, i) << ", " << fd(a, i) << ";\n";
    }
    bool wide = false;  // leading dimensions of 4 GiB or more: 64-bit row offsets (three instructions instead of one)
    // %rd<a> = base + index * stride; the strides live in %rd2 / %rd4 (bytes, 64 bit) and %r13 / %r14 (low halves)
    void row_addr(std::ostringstream &o2, int a, const std::string &base, const std::string &stride, int64_t index) {
        if (wide)
            o2 << "\tmad.lo.u64 %rd" << a << ", " << stride << ", " << index << ", " << base << ";\n";
        else
            o2 << "\tmad.wide.u32 %rd" << a << ", " << (stride == "%rd2" ? "%r13" : (stride == "%rd4" ? "%r14" : "%r20")) << ", " << index << ", " << base << ";\n";
    }
    // value = load from  base + index * stride  (two f64 for S == 2)
    int load(const char *space, const std::string &base, const std::string &stride, int64_t index) {
        const int a = nrd++;
        row_addr(os, a, base, stride, index);
        const int r = new_val();
        if (S == 2)
            os << "\t" << space << ".v2.f64 {" << fd(r, 0) << ", " << fd(r, 1) << "}, [%rd" << a << "];\n";
        else
            os << "\t" << space << ".f64 " << fd(r, 0) << ", [%rd" << a << "];\n";
        return r;
    }

    // x^n, n >= 4: Julia >= 1.9 pow_body (base/math.jl), unrolled for the known exponent
    int pow_body(int x0, int n) {
        int x = x0;
        int y = new_val(), xnlo = new_val(), ynlo = new_val();
        for (int i = 0; i < S; ++i) {
            os << "\tmov.f64 " << fd(y, i) << ", 0d3FF0000000000000;\n";
            os << "\tmov.f64 " << fd(xnlo, i) << ", 0d0000000000000000;\n";
            os << "\tmov.f64 " << fd(ynlo, i) << ", 0d0000000000000000;\n";
        }
        auto fma = [&](int d, int a, int b, int c) {
            for (int i = 0; i < S; ++i)
                os << "\tfma.rn.f64 " << fd(d, i) << ", " << fd(a, i) << ", " << fd(b, i) << ", " << fd(c, i) << ";\n";
        };
        auto neg = [&](int d, int a) {
            for (int i = 0; i < S; ++i) os << "\tneg.f64 " << fd(d, i) << ", " << fd(a, i) << ";\n";
        };
        while (n > 1) {
            if (n & 1) {
                const int t = new_val(), err = new_val(), p = new_val(), np_ = new_val(), lo = new_val(), yl = new_val();
                binop("mul", t, x, ynlo);
                fma(err, y, xnlo, t);
                binop("mul", p, x, y);
                neg(np_, p);
                fma(lo, x, y, np_);
                binop("add", yl, lo, err);
                y = p;
                ynlo = yl;
            }
            const int t2 = new_val(), err = new_val(), p = new_val(), np_ = new_val(), lo = new_val(), xl = new_val();
            scale(t2, x, 2.0);
            binop("mul", err, t2, xnlo);
            binop("mul", p, x, x);
            neg(np_, p);
            fma(lo, x, x, np_);
            binop("add", xl, lo, err);
            x = p;
            xnlo = xl;
            n >>= 1;
        }
        const int t = new_val(), err = new_val(), r1 = new_val(), r2 = new_val(), res = new_val();
        binop("mul", t, x, ynlo);
        fma(err, y, xnlo, t);
        fma(r1, x, y, err);
        binop("mul", r2, x, y);
        for (int i = 0; i < S; ++i) {
            const int p1 = np++, p2 = np++;
            os << "\ttestp.finite.f64 %p" << p1 << ", " << fd(x, i) << ";\n";
            os << "\ttestp.finite.f64 %p" << p2 << ", " << fd(err, i) << ";\n";
            os << "\tand.pred %p" << p1 << ", %p" << p1 << ", %p" << p2 << ";\n";
            os << "\tselp.f64 " << fd(res, i) << ", " << fd(r1, i) << ", " << fd(r2, i) << ", %p" << p1 << ";\n";
        }
        return res;
    }

    void root_out(int r, int32_t root) {
        if (!acc) {
            const int a = nrd++;
            os << "\tmad.lo.u64 %rd" << a << ", %rd5, " << root << ", %rd6;\n";  // rootbase + root * ld_root_bytes
            if (cplx) {
                os << "\t@%p0 st.global.v2.f64 [%rd" << a << "], {" << fd(r, 0) << ", " << fd(r, 1) << "};\n";
            } else if (S == 2) {
                os << "\t@%p1 st.global.v2.f64 [%rd" << a << "], {" << fd(r, 0) << ", " << fd(r, 1) << "};\n";
                os << "\t@%p2 st.global.f64 [%rd" << a << "], " << fd(r, 0) << ";\n";  // odd tail: first sample only
            } else {
                os << "\t@%p0 st.global.f64 [%rd" << a << "], " << fd(r, 0) << ";\n";
            }
            return;
        }
        if ((int)sacc_pos.size() + (cplx ? 2 : 1) <= sacc_cap) {
            for (int c = 0; c < (cplx ? 2 : 1); ++c) {
                const int s = nfd++, t = nfd++;
                const int off = (int)sacc_pos.size() * sacc_stride;
                sacc_pos.push_back(cplx ? 2 * root + c : root);
                os << "\tselp.f64 %fd" << s << ", " << fd(r, c) << ", 0d0000000000000000, %p0;\n";
                os << "\tld.shared.f64 %fd" << t << ", [%r16+" << off << "];\n";
                os << "\tadd.rn.f64 %fd" << t << ", %fd" << t << ", %fd" << s << ";\n";
                os << "\tst.shared.f64 [%r16+" << off << "], %fd" << t << ";\n";
            }
            return;
        }
        if (cplx) {
            // re and im are summed separately: columns 2r and 2r + 1
            for (int c = 0; c < 2; ++c) {
                const int s = nfd++, t = nfd++;
                os << "\tselp.f64 %fd" << s << ", " << fd(r, c) << ", 0d0000000000000000, %p0;\n";
                if (persistent) {
                    os << "\tadd.rn.f64 %fd" << racc0 + 2 * root + c << ", %fd" << racc0 + 2 * root + c << ", %fd" << s << ";\n";
                    continue;
                }
                for (int m = 16; m >= 1; m >>= 1) {
                    os << "\tmov.b64 {%r8, %r9}, %fd" << s << ";\n";
                    os << "\tshfl.sync.bfly.b32 %r10, %r8, " << m << ", 31, 0xffffffff;\n";
                    os << "\tshfl.sync.bfly.b32 %r11, %r9, " << m << ", 31, 0xffffffff;\n";
                    os << "\tmov.b64 %fd" << t << ", {%r10, %r11};\n";
                    os << "\tadd.rn.f64 %fd" << s << ", %fd" << s << ", %fd" << t << ";\n";
                }
                const int64_t off = ((int64_t)root * 2 + c) * 8;
                os << "\t@%p3 ld.global.f64 %fd" << t << ", [%rd7+" << off << "];\n";
                os << "\t@%p3 add.rn.f64 %fd" << t << ", %fd" << t << ", %fd" << s << ";\n";
                os << "\t@%p3 st.global.f64 [%rd7+" << off << "], %fd" << t << ";\n";
            }
            return;
        }
        if (persistent) {
            // running sum of this thread: samples in order, masked
            const int s = nfd++, t = nfd++;
            os << "\tselp.f64 %fd" << s << ", " << fd(r, 0) << ", 0d0000000000000000, %p0;\n";
            if (S == 2) {
                os << "\tselp.f64 %fd" << t << ", " << fd(r, 1) << ", 0d0000000000000000, %p1;\n";
                os << "\tadd.rn.f64 %fd" << s << ", %fd" << s << ", %fd" << t << ";\n";
            }
            os << "\tadd.rn.f64 %fd" << racc0 + root << ", %fd" << racc0 + root << ", %fd" << s << ";\n";
            return;
        }
        // masked sum of the thread's samples, xor-tree over the warp, lane 0 adds into the warp's partial row
        const int s = nfd++, t = nfd++;
        os << "\tselp.f64 %fd" << s << ", " << fd(r, 0) << ", 0d0000000000000000, %p0;\n";
        if (S == 2) {
            os << "\tselp.f64 %fd" << t << ", " << fd(r, 1) << ", 0d0000000000000000, %p1;\n";
            os << "\tadd.rn.f64 %fd" << s << ", %fd" << s << ", %fd" << t << ";\n";
        }
        for (int m = 16; m >= 1; m >>= 1) {
            os << "\tmov.b64 {%r8, %r9}, %fd" << s << ";\n";
            os << "\tshfl.sync.bfly.b32 %r10, %r8, " << m << ", 31, 0xffffffff;\n";
            os << "\tshfl.sync.bfly.b32 %r11, %r9, " << m << ", 31, 0xffffffff;\n";
            os << "\tmov.b64 %fd" << t << ", {%r10, %r11};\n";
            os << "\tadd.rn.f64 %fd" << s << ", %fd" << s << ", %fd" << t << ";\n";
        }
        os << "\t@%p3 ld.global.f64 %fd" << t << ", [%rd7+" << (int64_t)root * 8 << "];\n";
        os << "\t@%p3 add.rn.f64 %fd" << t << ", %fd" << t << ", %fd" << s << ";\n";
        os << "\t@%p3 st.global.f64 [%rd7+" << (int64_t)root * 8 << "], %fd" << t << ";\n";
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// planning + PTX
// ---------------------------------------------------------------------------------------------------------------------
// Linear SSA form of the emitted function in FOLD order: a single-use node is evaluated at the point of use inside
// its parent's left fold (so at most one partial accumulator per nesting level is live), a multi-use node or root is
// evaluated at its first use and its value id remembered.  The arithmetic is exactly the emitter's (static.jl:13-46).
struct IrOp {
    uint8_t kind;   // 0 MUL(a,b)  1 ADD(a,b)  2 SCALE(a,f)  3 POW(a,n)  4 ROOT(a -> root position n)  5 NEG(a)
    int32_t a, b;   // operands: value id >= 0, or -(leaf+1) for leaf `leaf`
    int32_t n;
    double f;
};
enum { IR_MUL = 0, IR_ADD = 1, IR_SCALE = 2, IR_POW = 3, IR_ROOT = 4, IR_NEG = 5 };

// Order in which the root statements are expanded.  Statement ORDER is free (each statement keeps its own fold, so
// every value keeps its bits); what the order decides is how long shared values stay live, i.e. how many of them cross
// kernel boundaries.  The emitter's order walks `graphs` as given, and diagram front ends list roots by channel and
// time configuration, which puts roots that share sub-vertices far apart (Parquet vertex4 order 4: 1800 values live).
// Greedy: next comes the root whose cone retires the most already-computed shared values and opens the fewest new
// ones; ties by emitter order.  (Same graph: <= 190 values live, cross traffic 4x smaller.)
static thread_local double g_root_leaf_w = 0.0;
static thread_local long g_root_leaf_window = 4;
void jit_set_root_leaf_bias(double w, long window) {
    g_root_leaf_w = w;
    g_root_leaf_window = window;
}

static void order_roots(const Lowered &low, std::vector<int32_t> &order) {
    const auto &st = low.st;
    const auto &ops = low.ops;
    std::vector<int32_t> roots;
    for (int32_t v = 0; v < (int32_t)st.size(); ++v)
        if (st[(size_t)v].root >= 0) roots.push_back(v);
    order = roots;
    const char *env = getenv("FDG_JIT_ROOT_ORDER");
    if (roots.size() < 3 || (env && atoi(env) == 0)) return;
    const size_t n = st.size(), R = roots.size();
    // cones: inner statements reachable from each root (sorted statement indices)
    std::vector<std::vector<int32_t>> cone(R);
    std::vector<int32_t> mark(n, -1), work;
    size_t total = 0;
    for (size_t r = 0; r < R; ++r) {
        if (st[(size_t)roots[r]].op < 0) continue;  // root[r] = leafVal[k]
        work.assign(1, roots[r]);
        mark[(size_t)roots[r]] = (int32_t)r;
        while (!work.empty()) {
            const int32_t v = work.back();
            work.pop_back();
            cone[r].push_back(v);
            const Stmt &s = st[(size_t)v];
            for (int32_t i = 0; i < s.count; ++i) {
                const int32_t c = ops[(size_t)(s.first + i)].val;
                if (st[(size_t)c].op >= 0 && mark[(size_t)c] != (int32_t)r) {
                    mark[(size_t)c] = (int32_t)r;
                    work.push_back(c);
                }
            }
        }
        total += cone[r].size();
        if (total > 100000000u) return;  // the cones alone would take > 400 MB: keep emitter order
    }
    // Scoring every remaining root at every step is quadratic in the number of roots.  For big root sets only a short
    // list is scored each step: the roots with the most already-computed values in their cone (the ones that can retire
    // something), plus the next root in emitter order.
    const char *sl = getenv("FDG_JIT_ROOT_SHORTLIST");
    const bool shortlist = sl ? atoi(sl) != 0 : (double)total * (double)R > 2e9;
    std::vector<std::vector<int32_t>> roots_of;  // statement -> roots whose cone holds it (short-list mode only)
    std::vector<int32_t> warm(R, 0);             // computed values inside the cone of every root not taken yet
    if (shortlist) {
        roots_of.resize(n);
        for (size_t r = 0; r < R; ++r)
            for (const int32_t v : cone[r]) roots_of[(size_t)v].push_back((int32_t)r);
    }
    // readers of every statement (distinct statements)
    std::vector<std::vector<int32_t>> readers(n);
    for (int32_t v = 0; v < (int32_t)n; ++v) {
        const Stmt &s = st[(size_t)v];
        if (!s.live || s.op < 0) continue;
        for (int32_t i = 0; i < s.count; ++i) {
            const int32_t c = ops[(size_t)(s.first + i)].val;
            if (st[(size_t)c].op >= 0 && (readers[(size_t)c].empty() || readers[(size_t)c].back() != v)) readers[(size_t)c].push_back(v);
        }
    }
    std::vector<uint8_t> computed(n, 0), taken(R, 0);
    std::vector<int32_t> in_cone(n, -1);
    // Leaf bias: a root is also charged w for every leaf its new statements read that none of the roots taken in the last
    // `window` steps has read -- with a small w a tie-break between roots that retire equally many shared values, which
    // brings roots over the same leaves into the same kernel.  The effect on the planned traffic is a few per cent and not
    // monotone in w, so the caller tries a few (w, window) pairs and keeps the plan the model likes best (jit_set_root_leaf_bias;
    // FDG_JIT_ROOT_LEAF_WEIGHT / _WINDOW force one).
    double leaf_w = g_root_leaf_w;
    long leaf_window = g_root_leaf_window;
    if (const char *e = getenv("FDG_JIT_ROOT_LEAF_WEIGHT")) leaf_w = atof(e);
    if (const char *e = getenv("FDG_JIT_ROOT_LEAF_WINDOW")) leaf_window = atol(e);
    std::vector<long> leaf_step((size_t)low.L, -1000000);
    std::vector<int32_t> leaf_seen((size_t)low.L, -1);
    order.clear();
    std::vector<size_t> cand;
    size_t next_emitter = 0;
    for (size_t step = 0; step < R; ++step) {
        long best_score = LONG_MIN;
        size_t best = R;
        cand.clear();
        if (shortlist) {
            while (next_emitter < R && taken[next_emitter]) ++next_emitter;
            const size_t K = 32;
            for (size_t r = 0; r < R; ++r) {
                if (taken[r] || warm[r] == 0) continue;
                cand.push_back(r);
            }
            if (cand.size() > K) {
                std::partial_sort(cand.begin(), cand.begin() + (long)K, cand.end(),
                                  [&](size_t a, size_t b) { return warm[a] != warm[b] ? warm[a] > warm[b] : a < b; });
                cand.resize(K);
            }
            if (next_emitter < R && std::find(cand.begin(), cand.end(), next_emitter) == cand.end()) cand.push_back(next_emitter);
        } else {
            for (size_t r = 0; r < R; ++r)
                if (!taken[r]) cand.push_back(r);
        }
        for (const size_t r : cand) {
            for (const int32_t v : cone[r]) in_cone[(size_t)v] = (int32_t)r;
            long kills = 0, creates = 0;
            for (const int32_t v : cone[r]) {
                if (computed[(size_t)v]) {
                    // a computed value this cone reads: retired if every reader is computed or inside the cone
                    bool read_here = false, dies = true;
                    for (const int32_t u : readers[(size_t)v]) {
                        if (computed[(size_t)u]) continue;
                        if (in_cone[(size_t)u] == (int32_t)r) read_here = true;
                        else dies = false;
                    }
                    if (read_here && dies) ++kills;
                } else {
                    for (const int32_t u : readers[(size_t)v])
                        if (in_cone[(size_t)u] != (int32_t)r) {
                            ++creates;
                            break;
                        }
                }
            }
            long score = kills - creates;
            if (leaf_w > 0.0) {
                long cold = 0;
                for (const int32_t v : cone[r]) {
                    if (computed[(size_t)v]) continue;
                    const Stmt &sv = st[(size_t)v];
                    for (int32_t i = 0; i < sv.count; ++i) {
                        const int32_t c = ops[(size_t)(sv.first + i)].val;
                        if (st[(size_t)c].op >= 0) continue;
                        const int32_t lf = st[(size_t)c].leaf;
                        if (leaf_seen[(size_t)lf] == (int32_t)(step * R + r) % 2000000000) continue;
                        leaf_seen[(size_t)lf] = (int32_t)(step * R + r) % 2000000000;
                        if ((long)step - leaf_step[(size_t)lf] > leaf_window) ++cold;
                    }
                }
                score = (long)(1000.0 * (double)score - 1000.0 * leaf_w * (double)cold);
            } else {
                score *= 1000;
            }
            if (score > best_score) {
                best_score = score;
                best = r;
            }
        }
        taken[best] = 1;
        order.push_back(roots[best]);
        if (leaf_w > 0.0)
            for (const int32_t v : cone[best]) {
                if (computed[(size_t)v]) continue;
                const Stmt &sv = st[(size_t)v];
                for (int32_t i = 0; i < sv.count; ++i) {
                    const int32_t c = ops[(size_t)(sv.first + i)].val;
                    if (st[(size_t)c].op < 0) leaf_step[(size_t)st[(size_t)c].leaf] = (long)step;
                }
            }
        for (const int32_t v : cone[best]) {
            if (shortlist && !computed[(size_t)v])
                for (const int32_t r2 : roots_of[(size_t)v]) warm[(size_t)r2]++;
            computed[(size_t)v] = 1;
        }
    }
}

static void build_ir(const Lowered &low, std::vector<IrOp> &ir, const Lowered *merged = nullptr, int scope = 0) {
    const auto &st = low.st;
    const auto &ops = low.ops;
    std::vector<int32_t> val_of(st.size(), INT32_MIN);  // value id of a materialised statement once computed
    // scoped merging: value and position of the latest evaluation of every class of equal statements
    const bool scoped = merged && merged->canon.size() == st.size() && scope > 0;
    std::vector<int32_t> cls_val, cls_at;
    if (scoped) {
        cls_val.assign(st.size(), INT32_MIN);
        cls_at.assign(st.size(), 0);
    }
    auto cls = [&](int32_t v) { return merged->canon[(size_t)v]; };
    // `x * (-1.0)` is the sign flip of x for every finite and infinite double (round-to-nearest is symmetric), so it is
    // written as a negation, which sm_100a folds into the operand modifiers of the DMUL / DADD that reads it: the
    // multiply disappears, the bits stay (NaN sign/payload is hardware-specific either way).  FDG_JIT_NEGFOLD=0 keeps it.
    const char *nf = getenv("FDG_JIT_NEGFOLD");
    const bool negfold = !(nf && atoi(nf) == 0);
    auto emit = [&](uint8_t kind, int32_t a, int32_t b, int32_t n, double f) -> int32_t {
        if (kind == IR_SCALE && f == -1.0 && negfold) kind = IR_NEG;
        ir.push_back({kind, a, b, n, f});
        return (int32_t)ir.size() - 1;
    };
    struct Frame {
        int32_t v, i, acc;
        bool have_ret;
        int32_t ret;
    };
    std::vector<Frame> stack;
    std::vector<int32_t> root_order;
    order_roots(scoped ? *merged : low, root_order);  // statement indices are the same in both views
    for (const int32_t root_stmt : root_order) {
        if (st[(size_t)root_stmt].op < 0) {  // root[r] = leafVal[k]
            emit(IR_ROOT, -(st[(size_t)root_stmt].leaf + 1), 0, st[(size_t)root_stmt].root, 1.0);
            continue;
        }
        if (val_of[(size_t)root_stmt] != INT32_MIN) continue;
        stack.push_back({root_stmt, 0, 0, false, 0});
        while (!stack.empty()) {
            Frame &fr = stack.back();
            const Stmt &s = st[(size_t)fr.v];
            if (fr.i == s.count) {
                int32_t res = fr.acc;
                if (s.op == FDG_OP_POWER) {
                    res = emit(IR_POW, res, 0, s.pow_n, 1.0);
                    const double f = ops[(size_t)s.first].f;
                    if (f != 1.0) res = emit(IR_SCALE, res, 0, 0, f);
                }
                if (s.root >= 0) emit(IR_ROOT, res, 0, s.root, 1.0);
                if (s.root >= 0 || s.uses >= 2) val_of[(size_t)fr.v] = res;
                if (scoped && res >= 0) {
                    cls_val[(size_t)cls(fr.v)] = res;
                    cls_at[(size_t)cls(fr.v)] = (int32_t)ir.size();
                }
                stack.pop_back();
                if (!stack.empty()) {
                    stack.back().have_ret = true;
                    stack.back().ret = res;
                }
                continue;
            }
            const Operand &o = ops[(size_t)(s.first + fr.i)];
            int32_t t;
            if (fr.have_ret) {
                t = fr.ret;
                fr.have_ret = false;
            } else if (st[(size_t)o.val].op < 0) {
                t = -(st[(size_t)o.val].leaf + 1);
            } else if (val_of[(size_t)o.val] != INT32_MIN) {
                t = val_of[(size_t)o.val];
            } else if (scoped && cls_val[(size_t)cls(o.val)] != INT32_MIN && (int32_t)ir.size() - cls_at[(size_t)cls(o.val)] <= scope) {
                t = cls_val[(size_t)cls(o.val)];  // an equal statement was evaluated a moment ago: same bits, no work
                if (st[(size_t)o.val].uses >= 2) val_of[(size_t)o.val] = t;
            } else {
                const int32_t child = o.val;
                stack.push_back({child, 0, 0, false, 0});  // NB: invalidates fr
                continue;
            }
            if (s.op == FDG_OP_POWER) {
                fr.acc = t;
            } else if (s.op == FDG_OP_SUM) {
                if (o.f != 1.0) t = emit(IR_SCALE, t, 0, 0, o.f);
                fr.acc = fr.i == 0 ? t : emit(IR_ADD, fr.acc, t, 0, 1.0);
            } else {
                if (fr.i == 0) {
                    fr.acc = o.f != 1.0 ? emit(IR_SCALE, t, 0, 0, o.f) : t;
                } else {
                    fr.acc = emit(IR_MUL, fr.acc, t, 0, 1.0);
                    if (o.f != 1.0) fr.acc = emit(IR_SCALE, fr.acc, 0, 0, o.f);
                }
            }
            stack.back().i++;
        }
    }
}

static int plan_from_ir(const Lowered &low, const std::vector<IrOp> &ir, int spt, bool acc, int seg_ops, bool wide_strides, bool fma,
                        JitPlan &plan, std::string &err, const PipeOptions *pipe, bool bulk = false);

// prefix sums of the per-operation instruction estimate (pipeline form: in 1/16 instructions, stretched by the stage weights)
static void plan_costs(const Lowered &low, const std::vector<IrOp> &ir, int spt, bool acc, const PipeOptions *pipe, std::vector<int64_t> &cost) {
    const bool cplx = low.dtype == FDG_C128;
    const int W = cplx ? 2 : 1;
    const int S = cplx ? 1 : spt;
    const size_t nops = ir.size();
    cost.assign(nops + 1, 0);
    for (size_t i = 0; i < nops; ++i) {
        const IrOp &o = ir[i];
        int64_t w = 0;
        switch (o.kind) {
            case IR_MUL: w = cplx ? 6 : S; break;
            case IR_ADD:
            case IR_SCALE: w = cplx ? 2 : S; break;
            case IR_NEG: w = 0; break;
            case IR_POW: {
                int lg = 0;
                while ((1 << (lg + 1)) <= o.n) ++lg;
                w = o.n <= 3 ? (o.n - 1) * (cplx ? 6 : S) : (cplx ? 12 * lg : 14 * lg * S);
            } break;
            default: w = acc ? (pipe ? 4 * W : 28 * W) : 2; break;  // a root: warp reduction + partial-row update, or a store
        }
        if (pipe) {
            double sc = 16.0;
            if (!pipe->weight.empty() && pipe->prev_start.size() == pipe->weight.size() + 1) {
                const size_t k = (size_t)(std::upper_bound(pipe->prev_start.begin(), pipe->prev_start.end(), (int32_t)i) - pipe->prev_start.begin());
                if (k >= 1 && k <= pipe->weight.size()) sc *= std::max(0.25, std::min(4.0, pipe->weight[k - 1]));
            }
            w = (int64_t)((double)w * sc + 0.5);
        }
        cost[i + 1] = cost[i] + w;
    }
}

int jit_plan(const Lowered &low, int spt, bool acc, int seg_ops, bool wide_strides, bool fma, JitPlan &plan, std::string &err,
             const PipeOptions *pipe, const Lowered *merged, int scope, bool bulk) {
    std::vector<IrOp> ir;
    if (merged && scope <= 0) scope = (seg_ops > 0 ? seg_ops : 4000) * 2 / 3;
    if (const char *e = getenv("FDG_CSE_SCOPE")) scope = atoi(e);
    build_ir(low, ir, merged, scope);
    if (const char *dump = getenv("FDG_JIT_DUMP_IR")) {  // debugging aid: the fold-order IR as raw records
        if (FILE *fp = std::fopen(dump, "wb")) {
            for (const IrOp &o : ir) {
                const int32_t rec[4] = {(int32_t)o.kind, o.a, o.b, o.n};
                std::fwrite(rec, sizeof(rec), 1, fp);
            }
            std::fclose(fp);
        }
    }
    if (!pipe) return plan_from_ir(low, ir, spt, acc, seg_ops, wide_strides, fma, plan, err, nullptr, bulk);
    // Pipeline form: the cuts follow a cost estimate per operation, but what a stage really costs is only known once its
    // code is written (input rows are charged where they are first read in the stage, roots cost a warp reduction).  So the
    // plan is made, the stages are re-weighted with (cost of the written code) / (estimate), and the cuts are placed again.
    PipeOptions po = *pipe;
    int rc = FDG_OK;
    if (!po.measured.empty() && po.prev_start.size() == po.measured.size() + 1) {
        // weights from a profile: (share of the measured time) / (share of the estimate) of every stage of the previous plan
        PipeOptions plain = *pipe;
        plain.weight.clear();
        plain.prev_start.clear();
        plain.measured.clear();
        std::vector<int64_t> c;
        plan_costs(low, ir, spt, acc, &plain, c);
        double sum_m = 0, sum_c = (double)std::max<int64_t>(c.back(), 1);
        for (const double m : po.measured) sum_m += m;
        po.weight.assign(po.measured.size(), 1.0);
        for (size_t k = 0; k < po.measured.size(); ++k) {
            const double ck = (double)(c[(size_t)po.prev_start[k + 1]] - c[(size_t)po.prev_start[k]]);
            if (ck > 0 && sum_m > 0) po.weight[k] = (po.measured[k] / sum_m) / (ck / sum_c);
        }
    }
    JitPlan best;
    double best_load = 1e30;
    // time of all passes (each as slow as its slowest stage) relative to a perfect split of the work over all SMs
    auto imbalance = [](const JitPlan &pl) {
        double tot = 0, t = 0;
        for (const int64_t c : pl.stage_cost) tot += (double)c;
        for (int p0 = 0; p0 < (int)pl.stage_cost.size(); p0 += pl.stages_per_pass) {
            double worst = 0;
            for (int k = p0; k < std::min<int>(p0 + pl.stages_per_pass, (int)pl.stage_cost.size()); ++k)
                worst = std::max(worst, (double)pl.stage_cost[(size_t)k] / (double)pl.stage_blocks[(size_t)k]);
            t += worst;
        }
        return t / (tot / (double)pl.n_sm);
    };
    for (int iter = 0; iter < 6; ++iter) {
        JitPlan cand;
        rc = plan_from_ir(low, ir, spt, acc, seg_ops, wide_strides, fma, cand, err, &po);
        if (rc != FDG_OK) return rc;
        const double load = imbalance(cand);
        const bool better = load < best_load;
        if (better) best_load = load;
        plan = cand;
        if (better) best = std::move(cand);
        // measured weights (a profile of a previous plan) are used as given: the model must not pull them back
        if (iter == 5 || plan.seg.size() < 2 || best_load < 1.02 || !pipe->weight.empty() || !pipe->measured.empty()) break;
        // estimated cost of each stage as the cuts saw it (already weighted) vs. the cost of the code
        std::vector<double> w(plan.seg.size());
        double worst = 0;
        double sum_c = 0, sum_e = 0;
        for (size_t k = 0; k < w.size(); ++k) {
            sum_c += (double)plan.stage_cost[k];
            sum_e += (double)plan.stage_estimate[k];
        }
        for (size_t k = 0; k < w.size(); ++k) {
            const double r = ((double)plan.stage_cost[k] / sum_c) / std::max((double)plan.stage_estimate[k] / sum_e, 1e-9);
            worst = std::max(worst, std::fabs(r - 1.0));
            // the new weight of an operation = its old weight x r of the stage it was in
            w[k] = r;
        }
        if (getenv("FDG_PIPE_TRACE")) {
            std::fprintf(stderr, "pipeline plan iteration %d: worst %.3f, cost/estimate:", iter, worst);
            for (size_t k = 0; k < w.size(); ++k) std::fprintf(stderr, " %.2f", w[k]);
            std::fprintf(stderr, "\n");
        }
        if (worst < 0.02) break;
        std::vector<double> neww(w.size());
        for (size_t k = 0; k < w.size(); ++k) {
            // weight the stage had in this round (piecewise over the previous boundaries: take the one at the stage's middle)
            double old = 1.0;
            if (!po.weight.empty()) {
                const int32_t mid = (plan.stage_start[k] + plan.stage_start[k + 1]) / 2;
                const size_t j = (size_t)(std::upper_bound(po.prev_start.begin(), po.prev_start.end(), mid) - po.prev_start.begin());
                if (j >= 1 && j <= po.weight.size()) old = po.weight[j - 1];
            }
            neww[k] = std::max(0.25, std::min(4.0, old * w[k]));
        }
        po.prev_start = plan.stage_start;
        po.weight = neww;
    }
    plan = std::move(best);
    return rc;
}

static int plan_from_ir(const Lowered &low, const std::vector<IrOp> &ir, int spt, bool acc, int seg_ops, bool wide_strides, bool fma,
                        JitPlan &plan, std::string &err, const PipeOptions *pipe, bool bulk) {
    const bool cplx = low.dtype == FDG_C128;
    if (bulk && !cplx && spt != 1) {
        err = "the bulk form evaluates one sample per thread";
        return FDG_ERR_BAD_ARG;
    }
    if (pipe && !cplx && spt != 1) {
        err = "the pipeline form evaluates one sample per thread";
        return FDG_ERR_BAD_ARG;
    }
    if (cplx) spt = 2;  // two f64 registers per value: (re, im) of one sample
    const int W = cplx ? 2 : 1;                // doubles per sample
    const int samples_per_thread = cplx ? 1 : spt;
    const int esh = cplx ? 4 : 3;              // log2(bytes per sample element)
    plan = JitPlan();
    plan.spt = samples_per_thread;
    plan.acc = acc;
    plan.fma = fma;
    // `seg_ops` is a budget of machine instructions per kernel (estimated below): what bounds a kernel is the
    // instruction cache -- straight-line code of more than about 100 KB stalls on instruction fetch (measured,
    // DESIGN.md section 6) -- so a complex multiply counts six and a negation folded into its reader none
    if (seg_ops <= 0) seg_ops = 4000;
    bool ring_on = true;
    int ring_rows = 0;
    // bulk form: a row is fetched again when its next use is more than a third of a kernel away (see `reload_gap` below)
    // (ring form: only ComplexF64 gains -- Taylor-AD sigma order 4 120 -> 129 M samples/s at 600; Float64 does not)
    int64_t reload_gap = bulk ? seg_ops / 3 : (cplx ? 600 : 0);
    if (const char *rg = getenv("FDG_JIT_RELOAD_GAP")) reload_gap = atoll(rg);
    if (const char *rg = getenv("FDG_JIT_RING")) ring_on = atoi(rg) != 0;
    if (const char *rr = getenv("FDG_JIT_RING_ROWS")) ring_rows = atoi(rr);
    const size_t nops = ir.size();
    auto is_binary = [](const IrOp &o) { return o.kind == IR_MUL || o.kind == IR_ADD; };
    // ---- cut points: about seg_ops operations per kernel, each cut placed where the fewest values are live ---------
    std::vector<int32_t> last_use(nops, -1);
    for (size_t i = 0; i < nops; ++i) {
        const IrOp &o = ir[i];
        if (o.a >= 0) last_use[(size_t)o.a] = (int32_t)i;
        if (is_binary(o) && o.b >= 0) last_use[(size_t)o.b] = (int32_t)i;
    }
    std::vector<int32_t> seg_start{0};
    std::vector<int64_t> cost;  // prefix sums of the instruction estimate
    plan_costs(low, ir, spt, acc, pipe, cost);
    if (pipe) seg_ops *= 16;
    int pipe_stages_per_pass = 1, pipe_n_pass = 1;
    std::vector<int> pipe_stage_sms;
    {
        // live[p] = values defined before op p and read at or after p (what a cut in front of p sends through memory)
        std::vector<int32_t> live(nops + 2, 0);
        for (size_t v = 0; v < nops; ++v)
            if (last_use[v] > (int32_t)v) {
                live[v + 1] += 1;
                live[(size_t)last_use[v] + 1] -= 1;
            }
        for (size_t q = 1; q < live.size(); ++q) live[q] += live[q - 1];
        const char *nc = getenv("FDG_JIT_NARROW_CUTS");
        const bool narrow = !(nc && atoi(nc) == 0);
        size_t start = 0;
        // first position whose cost since `start` reaches c
        auto pos_at = [&](size_t from, int64_t c) {
            return (size_t)(std::lower_bound(cost.begin() + (long)from, cost.end(), cost[from] + c) - cost.begin());
        };
        if (pipe) {
            // Pipeline form: a pass has one stage per instruction-cache group and the stage's share of the pass is the
            // group's share of the SMs, so the cuts go where the cumulative cost reaches those shares, each one moved inside
            // a narrow window (+-3 % of a stage) to the position with the fewest live values.  The number of passes follows
            // from the code the groups can hold: `seg_ops` estimated instructions per stage.
            const int64_t total = std::max<int64_t>(cost[nops], 1);
            const int G = (int)pipe->groups.size();
            const int n_sm = pipe->n_sm();
            int n_pass = (int)std::max<int64_t>((total + (int64_t)seg_ops * G - 1) / ((int64_t)seg_ops * G), 1);
            if (const char *e = getenv("FDG_PIPE_PASSES")) n_pass = std::max(1, atoi(e));
            // a small program does not need every group to run its own stage: neighbouring groups then share one
            int gs = G;  // stages per pass
            if (n_pass == 1) gs = (int)std::max<int64_t>(1, std::min<int64_t>(G, (total + seg_ops / 4 - 1) / (seg_ops / 4)));
            if (const char *e = getenv("FDG_PIPE_STAGES")) gs = std::max(1, std::min(atoi(e), G));
            pipe_stages_per_pass = gs;
            pipe_n_pass = n_pass;
            // groups of one stage: stage j of a pass gets groups [j * G / gs, (j + 1) * G / gs)
            pipe_stage_sms.assign((size_t)gs, 0);
            for (int g = 0; g < G; ++g) pipe_stage_sms[(size_t)((int64_t)g * gs / G)] += pipe->groups[(size_t)g];
            int slack_pct = 3;
            if (const char *e = getenv("FDG_PIPE_SLACK")) slack_pct = std::max(0, std::min(25, atoi(e)));
            const int64_t slack = std::max<int64_t>(total / ((int64_t)gs * n_pass) * slack_pct / 100, 1);
            size_t prev = 0;
            for (int ps = 0; ps < n_pass; ++ps) {
                int64_t cum_sm = 0;
                for (int j = 0; j < gs; ++j) {
                    cum_sm += pipe_stage_sms[(size_t)j];
                    if (ps == n_pass - 1 && j == gs - 1) break;
                    const int64_t target = (int64_t)((double)total * ((double)ps + (double)cum_sm / (double)n_sm) / (double)n_pass);
                    size_t cut = std::min(nops - 1, std::max(prev + 1, pos_at(0, target)));
                    if (narrow) {
                        const size_t lo_w = std::max(prev + 1, pos_at(0, std::max<int64_t>(target - slack, 0)));
                        const size_t hi_w = std::min(nops - 1, pos_at(0, target + slack));
                        for (size_t q = lo_w; q <= hi_w; ++q)
                            if (live[q] < live[cut] || (live[q] == live[cut] && std::llabs(cost[q] - target) < std::llabs(cost[cut] - target))) cut = q;
                    }
                    if (cut <= prev || cut >= nops) {
                        // an empty stage cannot be expressed: give up on the pipeline form for this program
                        err = "program too small for the pipeline form";
                        return FDG_ERR_CAPACITY;
                    }
                    seg_start.push_back((int32_t)cut);
                    prev = cut;
                }
            }
            start = nops;  // the loop below has nothing left to do
        }
        while (start < nops && cost[nops] - cost[start] > (int64_t)seg_ops + (narrow ? seg_ops / 4 : 0)) {
            size_t cut = std::min(nops - 1, std::max(start + 1, pos_at(start, seg_ops)));
            if (narrow) {
                const size_t lo_w = std::max(start + 1, pos_at(start, (int64_t)seg_ops * 3 / 4));
                const size_t hi_w = std::min(nops - 1, pos_at(start, (int64_t)seg_ops * 5 / 4));
                for (size_t q = lo_w; q <= hi_w; ++q)
                    if (live[q] <= live[cut]) cut = q;
            }
            seg_start.push_back((int32_t)cut);
            start = cut;
        }
    }
    const int nseg = (int)seg_start.size();
    seg_start.push_back((int32_t)nops);
    std::vector<int32_t> seg_id(nops);
    for (int sg = 0; sg < nseg; ++sg)
        for (int32_t i = seg_start[(size_t)sg]; i < seg_start[(size_t)sg + 1]; ++i) seg_id[(size_t)i] = sg;
    auto seg_of = [&](int32_t id) { return seg_id[(size_t)id]; };
    // ---- values read in a later segment than the one defining them travel through the cross buffer; a row is reused
    //      once the last kernel reading it has run (kernels of one stream run in order)
    std::vector<int32_t> cross(nops, -1);
    std::vector<int32_t> bcross(nops, -1);  // pipeline form: row of the pass-boundary buffer instead (value read in a later pass)
    int32_t n_cross = 0, n_cross_values = 0, n_boundary = 0;
    {
        std::vector<int32_t> last_seg(nops, -1);
        for (size_t i = 0; i < nops; ++i) {
            const IrOp &o = ir[i];
            const int sg = seg_of((int32_t)i);
            if (o.a >= 0 && seg_of(o.a) != sg) last_seg[(size_t)o.a] = sg;
            if (is_binary(o) && o.b >= 0 && seg_of(o.b) != sg) last_seg[(size_t)o.b] = sg;
        }
        const int spp = pipe ? pipe_stages_per_pass : nseg + 1;  // stages per pass (classic: everything is one "pass")
        std::vector<std::vector<int32_t>> expire((size_t)nseg + 1);  // rows that become free after segment s
        std::vector<std::vector<int32_t>> bexpire((size_t)nseg / (size_t)spp + 2);  // boundary rows free after pass p
        std::vector<int32_t> free_rows, bfree_rows;
        for (int sg = 0; sg < nseg; ++sg) {
            if (pipe && sg % spp == 0) {
                // a new pass: the ring rows start afresh, boundary rows whose last reader was an earlier pass are free
                if (sg > 0)
                    for (const int32_t row : bexpire[(size_t)(sg / spp - 1)]) bfree_rows.push_back(row);
            }
            for (int32_t i = seg_start[(size_t)sg]; i < seg_start[(size_t)sg + 1]; ++i) {
                if (last_seg[(size_t)i] < 0) continue;
                ++n_cross_values;
                if (pipe && last_seg[(size_t)i] / spp != sg / spp) {
                    int32_t row;
                    if (!bfree_rows.empty()) {
                        row = bfree_rows.back();
                        bfree_rows.pop_back();
                    } else {
                        row = n_boundary++;
                    }
                    bcross[(size_t)i] = row;
                    bexpire[(size_t)(last_seg[(size_t)i] / spp)].push_back(row);
                    continue;
                }
                int32_t row;
                if (!free_rows.empty()) {
                    row = free_rows.back();
                    free_rows.pop_back();
                } else {
                    row = n_cross++;
                }
                cross[(size_t)i] = row;
                expire[(size_t)last_seg[(size_t)i]].push_back(row);
            }
            // rows whose last reader is THIS segment are free for values defined in later segments only
            for (const int32_t row : expire[(size_t)sg]) free_rows.push_back(row);
        }
    }
    plan.n_boundary = n_boundary;
    for (const IrOp &o : ir) {
        switch (o.kind) {
            case IR_MUL: plan.fp64_instr += cplx ? 6 : samples_per_thread; break;
            case IR_ADD:
            case IR_SCALE: plan.fp64_instr += cplx ? 2 : samples_per_thread; break;
            case IR_POW: plan.fp64_instr += (o.n <= 3 ? (o.n - 1) : 14) * (cplx ? 6 : samples_per_thread); break;
            default: break;
        }
    }
    plan.fp64_instr /= samples_per_thread;  // per sample
    plan.n_cross_values = n_cross_values;
    plan.n_cross = n_cross;
    plan.seg.resize((size_t)nseg);
    // a single accumulate kernel with few roots runs as a grid-stride loop: per-thread running sums in registers,
    // one warp reduction per root at the very end (instead of one per tile)
    plan.persistent = !pipe && !bulk && acc && nseg == 1 && low.R * W <= 32;
    plan.bulk = bulk;
    int bulk_smem_max = 0;
    plan.pipeline = pipe != nullptr;
    plan.stage_start.assign(seg_start.begin(), seg_start.end());
    plan.stage_estimate.clear();
    for (int sg = 0; sg < nseg; ++sg) plan.stage_estimate.push_back(cost[(size_t)seg_start[(size_t)sg + 1]] - cost[(size_t)seg_start[(size_t)sg]]);
    int ring_bytes_max = 0;
    for (int sg = 0; sg < nseg; ++sg) {
        Emitter e(low, spt, acc);
        e.cplx = cplx;
        e.wide = wide_strides;
        if (fma) e.rnd = "";
        e.persistent = plan.persistent;
        if (e.persistent) {
            e.racc0 = e.nfd;
            e.nfd += (int)low.R * W;
        }
        std::ostringstream &os = e.os;
        const size_t lo = (size_t)seg_start[(size_t)sg], hi = (size_t)seg_start[(size_t)sg + 1];
        std::vector<int32_t> reg_of(hi - lo, -1);          // register of a value defined in this segment
        std::vector<int32_t> leaf_reg((size_t)low.L, -1);  // register of a leaf loaded in this segment
        std::unordered_map<int32_t, int32_t> cross_reg;    // register of a cross value loaded in this segment
        // ---- input rows (leaves, cross values) in order of first use ------------------------------------------------
        std::vector<std::pair<int, int32_t>> in_rows;  // (0 leaf / 1 cross, row)
        // A row whose next use lies more than `reload_gap` operations after the previous one is fetched AGAIN instead of
        // being held in a register all that time: at the widest point of a kernel of the headline graph ~150 values are
        // live and 60-95 of them are leaves waiting for their next use, which is what made ptxas spill.  The second copy
        // comes out of L2, where the same SM put it microseconds ago (ncu: +1 % DRAM reads for +17 % rows fetched; spill
        // stores 3.5 KB -> 0.7 KB per sample; 199 -> 213 M samples/s).  0 = every row is fetched once per kernel.
        {
            std::vector<int64_t> seen_leaf((size_t)low.L, -1);  // position of the latest use
            std::unordered_map<int32_t, int64_t> seen_cross;
            auto note = [&](int32_t a, int64_t at) {
                if (a < 0) {
                    int64_t &last = seen_leaf[(size_t)(-a - 1)];
                    if (last < 0 || (reload_gap > 0 && at - last > reload_gap)) in_rows.emplace_back(0, -a - 1);
                    last = at;
                } else if ((size_t)a < lo) {
                    auto it = seen_cross.find(a);
                    if (it == seen_cross.end() || (reload_gap > 0 && at - it->second > reload_gap)) {
                        if (bcross[(size_t)a] >= 0) in_rows.emplace_back(2, bcross[(size_t)a]);
                        else in_rows.emplace_back(1, cross[(size_t)a]);
                    }
                    seen_cross[a] = at;
                }
            };
            for (size_t i = lo; i < hi; ++i) {
                note(ir[i].a, (int64_t)i);
                if (is_binary(ir[i])) note(ir[i].b, (int64_t)i);
            }
        }
        std::vector<int64_t> leaf_last((size_t)low.L, -1);
        std::unordered_map<int32_t, int64_t> cross_last;
        int64_t cur = 0;  // the operation being written
        // The ring: every thread streams its own element of each input row global -> shared with cp.async (LDGSTS),
        // `ring_rows` rows ahead of the row the arithmetic is reading, in the order the straight-line code first needs
        // them.  A copy in flight holds no register and no scoreboard entry, so the depth of the memory pipeline no
        // longer depends on what ptxas can hoist with 250 registers taken; `cp.async.wait_group` counts are known here.
        const int n_in = (int)in_rows.size();
        const int ES = cplx ? 16 : 8 * samples_per_thread;  // bytes per thread per row
        const int G = 4;
        int NR = ring_rows > 0 ? ring_rows : (ES == 8 ? 32 : 24);
        const int T = pipe ? pipe->threads : (bulk ? 256 : 128);  // (consumer) threads per block
        // static shared memory stops at 48 KB; the pipeline kernel asks for dynamic shared memory and may go deeper
        NR = std::max(G, std::min(NR, (pipe ? 131072 : 49152) / (T * ES)) / G * G);
        // bulk form: NG groups of BG rows; a group is what one mbarrier phase covers
        int BG = 16, NG = ES == 8 ? 4 : 2;  // 128 KB of rows in flight per SM; few, long groups: every wait costs the consumers code
        if (const char *x = getenv("FDG_JIT_BULK_GROUPS")) NG = std::max(2, std::min(16, atoi(x)));
        if (const char *x = getenv("FDG_JIT_BULK_GROUP_ROWS")) BG = std::max(1, std::min(64, atoi(x)));
        bool bulk_guard = true;
        int bulk_guard_mode = 1;
        if (const char *x = getenv("FDG_JIT_BULK_GUARD")) bulk_guard = atoi(x) != 0, bulk_guard_mode = atoi(x);
        int bulk_prefetch = 0;  // groups ahead of the ring that are prefetched into L2 (experiment)
        if (const char *x = getenv("FDG_JIT_BULK_PREFETCH")) bulk_prefetch = std::max(0, atoi(x));
        int bulk_psleep = 0;  // ns a producer warp sleeps after a failed poll of an empty-slot barrier (experiment)
        if (const char *x = getenv("FDG_JIT_BULK_PSLEEP")) bulk_psleep = std::max(0, atoi(x));
        int bulk_hint = 1000000;  // ns
        if (const char *x = getenv("FDG_JIT_BULK_HINT")) bulk_hint = std::max(0, atoi(x));
        if (bulk) {
            // deeper ring where the kernel's running sums leave room for it (experiment: FDG_JIT_BULK_GROUPS_MAX)
            if (const char *x = getenv("FDG_JIT_BULK_GROUPS_MAX")) {
                int n_roots_here = 0;
                for (size_t i = lo; i < hi; ++i) n_roots_here += ir[i].kind == IR_ROOT ? (cplx ? 2 : 1) : 0;
                const int room = 226 * 1024 - 256 - (acc ? n_roots_here * (T + 1) * 8 : 0);
                NG = std::max(NG, std::min(atoi(x), room / (BG * T * ES)));
            }
            NR = NG * BG;
        }
        const int ROWB = T * ES;  // bytes of one ring row (one input row of one tile)
        const bool ring = (ring_on || bulk) && !e.persistent && n_in > 0;
        const int sacc0 = 256 + (ring ? NR * T * ES : 0);  // pipeline / bulk form: where the running sums of the roots start
        const int b_groups = (n_in + BG - 1) / BG;                                  // bulk form: groups with rows in them ...
        const int b_groups_padded = (b_groups + 2 * NG - 1) / (2 * NG) * (2 * NG);  // ... padded so that slot and phase of a group do not depend on the tile
        int b_wait_id = 0;
        if ((pipe || bulk) && acc) {
            e.sacc_stride = (T + 1) * 8;
            e.sacc_cap = std::max(0, ((pipe ? 200 : 226) * 1024 - sacc0) / e.sacc_stride);
        }
        const int n_groups = (n_in + G - 1) / G;
        int next_in = 0;  // rows consumed so far
        auto ring_issue = [&](std::ostringstream &o2, int j) {  // copy of input row j into its slot
            const auto &row = in_rows[(size_t)j];
            const int a = e.nrd++;
            e.row_addr(o2, a, row.first == 2 ? "%rd9" : (row.first ? "%rd3" : "%rd1"), row.first == 2 ? "%rd11" : (row.first ? "%rd4" : "%rd2"), row.second);
            o2 << "\tcp.async." << (ES == 16 ? "cg" : "ca") << ".shared.global [%r12+" << (j % NR) * T * ES << "], [%rd" << a << "], " << ES << ";\n";
        };
        // bulk form: wait until the producer's copies of group g have landed (full barrier of its slot, phase known here)
        auto bulk_wait = [&](std::ostringstream &o2, int g) {
            const int id = b_wait_id++;
            // (the last operand is a suspend-time hint in ns: the warp sleeps in the barrier unit instead of polling)
            if (bulk_guard_mode == 2) {  // experiment: unguarded, but sleeping between polls
                o2 << "FDG_BW" << id << ":\n\tmbarrier.try_wait.parity.shared::cta.b64 %p7, [fdg_ring+" << 8 * (g % NG) << "], " << ((g / NG) & 1)
                   << ", " << bulk_hint << ";\n\t@%p7 bra FDG_BG" << id << ";\n\tnanosleep.u32 40;\n\tbra FDG_BW" << id << ";\nFDG_BG" << id << ":\n";
                return;
            }
            if (!bulk_guard) {
                o2 << "FDG_BW" << id << ":\n\tmbarrier.try_wait.parity.shared::cta.b64 %p7, [fdg_ring+" << 8 * (g % NG) << "], " << ((g / NG) & 1)
                   << ", " << bulk_hint << ";\n\t@!%p7 bra FDG_BW" << id << ";\n";
                return;
            }
            // guarded form: a wait that does not end within 4096 (long) polls traps instead of hanging the device
            o2 << "\tmov.u32 %r21, 0;\nFDG_BW" << id << ":\n"
               << "\tmbarrier.try_wait.parity.shared::cta.b64 %p7, [fdg_ring+" << 8 * (g % NG) << "], " << ((g / NG) & 1) << ", " << bulk_hint << ";\n"
               << "\t@%p7 bra FDG_BG" << id << ";\n\tadd.u32 %r21, %r21, 1;\n\tsetp.lt.u32 %p7, %r21, 4096;\n\t@%p7 bra FDG_BW" << id << ";\n"
               << "\ttrap;\nFDG_BG" << id << ":\n";
        };
        // ... and hand the slot back once this warp has read the group's rows (empty barrier: one arrival per consumer
        // warp; the warp-level barrier in front orders the other lanes' reads before lane 0's arrival)
        auto bulk_release = [&](std::ostringstream &o2, int g) {
            o2 << "\tbar.warp.sync 0xffffffff;\n\t@%p3 mbarrier.arrive.shared::cta.b64 %rd31, [fdg_ring+" << 8 * (NG + g % NG) << "];\n";
        };
        auto ring_load = [&](int kind_, int32_t row_) -> int {
            const int j = next_in++;
            (void)kind_;
            (void)row_;
            if (bulk) {
                const int g = j / BG;
                if (j % BG == 0) bulk_wait(os, g);
                const int r = e.new_val();
                if (ES == 16)
                    os << "\tld.shared.v2.f64 {" << e.fd(r, 0) << ", " << e.fd(r, 1) << "}, [%r12+" << (j % NR) * ROWB << "];\n";
                else
                    os << "\tld.shared.f64 " << e.fd(r, 0) << ", [%r12+" << (j % NR) * ROWB << "];\n";
                if (j % BG == BG - 1 || j == n_in - 1) bulk_release(os, g);
                return r;
            }
            const int g = j / G;
            if (j % G == 0) os << "\tcp.async.wait_group " << std::min(NR / G - 1, n_groups - 1 - g) << ";\n";
            const int r = e.new_val();
            if (ES == 16)
                os << "\tld.shared.v2.f64 {" << e.fd(r, 0) << ", " << e.fd(r, 1) << "}, [%r12+" << (j % NR) * T * ES << "];\n";
            else
                os << "\tld.shared.f64 " << e.fd(r, 0) << ", [%r12+" << (j % NR) * T * ES << "];\n";
            if (j % G == G - 1 || j == n_in - 1) {
                const int j0 = (g + NR / G) * G;
                if (j0 < n_in) {
                    for (int q = j0; q < std::min(n_in, j0 + G); ++q) ring_issue(os, q);
                    os << "\tcp.async.commit_group;\n";
                }
            }
            return r;
        };
        auto operand = [&](int32_t a) -> int {
            if (a < 0) {
                const int32_t k = -a - 1;
                if (leaf_reg[(size_t)k] < 0 || (reload_gap > 0 && cur - leaf_last[(size_t)k] > reload_gap)) {
                    if (leaf_reg[(size_t)k] < 0) plan.leaf_loads++;
                    else plan.refetch_loads++;
                    leaf_reg[(size_t)k] = ring ? ring_load(0, k) : e.load("ld.global.nc", "%rd1", "%rd2", k);
                }
                leaf_last[(size_t)k] = cur;
                return leaf_reg[(size_t)k];
            }
            if ((size_t)a >= lo) return reg_of[(size_t)a - lo];
            auto it = cross_reg.find(a);
            if (it != cross_reg.end() && !(reload_gap > 0 && cur - cross_last[a] > reload_gap)) {
                cross_last[a] = cur;
                return it->second;
            }
            const int r = ring ? ring_load(1, cross[(size_t)a])
                               : (bcross[(size_t)a] >= 0 ? e.load("ld.global", "%rd9", "%rd11", bcross[(size_t)a]) : e.load("ld.global", "%rd3", "%rd4", cross[(size_t)a]));
            if (it == cross_reg.end()) plan.cross_loads++;
            else plan.refetch_loads++;
            cross_reg[a] = r;
            cross_last[a] = cur;
            return r;
        };
        for (size_t i = lo; i < hi; ++i) {
            const IrOp &o = ir[i];
            cur = (int64_t)i;
            int r = -1;
            switch (o.kind) {
                case IR_MUL:
                case IR_ADD: {
                    const int a = operand(o.a), b = operand(o.b);
                    r = e.new_val();
                    e.binop(o.kind == IR_MUL ? "mul" : "add", r, a, b);
                } break;
                case IR_SCALE: {
                    const int a = operand(o.a);
                    r = e.new_val();
                    e.scale(r, a, o.f);
                } break;
                case IR_NEG: {
                    const int a = operand(o.a);
                    r = e.new_val();
                    e.neg(r, a);
                } break;
                case IR_POW: {
                    const int x = operand(o.a);
                    if (o.n == 2) {
                        r = e.new_val();
                        e.binop("mul", r, x, x);
                    } else if (o.n == 3) {
                        const int t = e.new_val();
                        e.binop("mul", t, x, x);
                        r = e.new_val();
                        e.binop("mul", r, t, x);
                    } else {
                        r = cplx ? e.cpow(x, (unsigned)o.n) : e.pow_body(x, o.n);
                    }
                } break;
                default: e.root_out(operand(o.a), o.n); break;
            }
            reg_of[i - lo] = r;
            if (r >= 0 && (cross[i] >= 0 || bcross[i] >= 0)) {
                plan.cross_stores++;
                const int a = e.nrd++;
                if (bcross[i] >= 0) e.row_addr(os, a, "%rd9", "%rd11", bcross[i]);
                else e.row_addr(os, a, "%rd3", "%rd4", cross[i]);
                if (spt == 2)
                    os << "\tst.global.v2.f64 [%rd" << a << "], {" << e.fd(r, 0) << ", " << e.fd(r, 1) << "};\n";
                else
                    os << "\tst.global.f64 [%rd" << a << "], " << e.fd(r, 0) << ";\n";
            }
        }
        if (bulk && ring)  // the padding groups: no rows, but every slot sees the same (even) number of phases per tile
            for (int g = b_groups; g < b_groups_padded; ++g) {
                bulk_wait(os, g);
                bulk_release(os, g);
            }
        const std::string body = os.str();
        const int spt_hdr = samples_per_thread;
        JitSegment &js = plan.seg[(size_t)sg];
        js.name = (pipe ? "fdg_stage" : "fdg_seg") + std::to_string(sg);
        js.n_stmts = (int)(hi - lo);
        std::ostringstream p;
        p << ".version 8.7\n.target sm_100a\n.address_size 64\n\n";
        if (pipe) {
            // ---- stage function of the pipeline kernel -------------------------------------------------------------------
            // One warp = one tile of 32 samples at a time; the warps of a stage take tiles round-robin (warp w of W: tiles
            // w, w + W, ...: a fixed assignment, so the accumulators are summed in a fixed order).  progress[tile] = number of
            // stages the tile has passed: stage k waits for k, publishes k + 1; stage 0 waits until the tile that used its
            // cross slot `window` tiles ago has passed every stage.
            // Shared memory: [0, 128) the launch arguments (written by the entry kernel), [128, 256) one word per warp,
            // then the input ring.  The stage function takes no parameters and re-reads what it needs per tile, so that
            // nothing but the tile index has to stay in a register across the straight-line code.
            const int n_sacc = (int)e.sacc_pos.size();
            ring_bytes_max = std::max(ring_bytes_max, sacc0 + n_sacc * e.sacc_stride);
            p << ".extern .shared .align 16 .b8 fdg_ring[];\n";
            if (n_sacc > 0) {
                p << ".const .align 4 .u32 fdg_rt" << sg << "[" << n_sacc << "] = {";
                for (int j = 0; j < n_sacc; ++j) p << (j ? ", " : "") << e.sacc_pos[(size_t)j];
                p << "};\n";
            }
            p << ".visible .func " << js.name << "()\n{\n";
            p << "\t.reg .f64 %fd<" << e.nfd + 2 << ">;\n\t.reg .b64 %rd<" << e.nrd + 1 + (ring ? NR : 0) << ">;\n\t.reg .pred %p<" << e.np + 1
              << ">;\n\t.reg .b32 %r<" << e.nr + 1 << ">;\n";
            const int wpb = T / 32;
            // %rd20 = tile of this warp (the only value that lives across tiles besides what ptxas keeps of the setup)
            p << "\tmov.u32 %r2, %tid.x;\n\tand.b32 %r3, %r2, 31;\n\tsetp.eq.u32 %p3, %r3, 0;\n"
              << "\tmov.u32 %r0, %ctaid.x;\n\tshr.u32 %r5, %r2, 5;\n"
              << "\tld.volatile.shared.u32 %r15, [fdg_ring+96];\n\tmad.lo.u32 %r6, %r15, " << wpb << ", %r5;\n"
              << "\tcvt.u64.u32 %rd20, %r6;\n"
              << "\tmov.u64 %rd21, %clock64;\n\tshl.b32 %r7, %r5, 3;\n\tmov.u32 %r15, fdg_ring;\n\tadd.u32 %r7, %r7, %r15;\n"
              << "\tst.volatile.shared.u64 [%r7+128], %rd21;\n";
            if (ring) p << "\tmov.u32 %r12, fdg_ring;\n\tmad.lo.u32 %r12, %r2, " << ES << ", %r12;\n\tadd.u32 %r12, %r12, 256;\n";
            if (n_sacc > 0) {
                // running sums of this stage's roots: one column per thread, zeroed here, reduced after the last tile
                p << "\tmov.u32 %r16, fdg_ring;\n\tmad.lo.u32 %r16, %r2, 8, %r16;\n\tadd.u32 %r16, %r16, " << sacc0 << ";\n"
                  << "\tmov.u32 %r17, 0;\n\tmov.u32 %r18, %r16;\n\tmov.f64 %fd" << e.nfd << ", 0d0000000000000000;\n"
                  << "FDG_ZERO:\n\tst.shared.f64 [%r18], %fd" << e.nfd << ";\n\tadd.u32 %r18, %r18, " << e.sacc_stride << ";\n"
                  << "\tadd.u32 %r17, %r17, 1;\n\tsetp.lt.u32 %p6, %r17, " << n_sacc << ";\n\t@%p6 bra FDG_ZERO;\n";
            }
            p << "FDG_TILE:\n\tld.volatile.shared.u64 %rd17, [fdg_ring+64];\n\tsetp.ge.u64 %p6, %rd20, %rd17;\n\t@%p6 bra FDG_DONE;\n"
              << "\tld.volatile.shared.u64 %rd16, [fdg_ring+56];\n\tld.volatile.shared.u64 %rd19, [fdg_ring+72];\n"
              << "\tmov.u64 %rd21, %clock64;\n";
            // ---- wait for the tile ----
            // The first stage of a pass waits until the tile that used its ring slot `window` tiles ago has passed every stage
            // of the pass (the tile itself is ready: the previous pass is a previous kernel); the others wait for the stage
            // before them.  Polling is a relaxed load; the fence after it orders everything that follows.
            const int spp = pipe_stages_per_pass;
            const bool first_of_pass = sg % spp == 0;
            if (first_of_pass) {
                p << "\tsetp.lt.u64 %p6, %rd20, %rd19;\n\t@%p6 bra FDG_GO;\n"
                  << "\tsub.u64 %rd22, %rd20, %rd19;\n\tshl.b64 %rd22, %rd22, 2;\n\tadd.u64 %rd22, %rd16, %rd22;\n";
            } else {
                p << "\tshl.b64 %rd22, %rd20, 2;\n\tadd.u64 %rd22, %rd16, %rd22;\n";
            }
            const int want = first_of_pass ? std::min(nseg, (sg / spp + 1) * spp) : sg;
            p << "\tmov.u64 %rd23, %globaltimer;\n"
              << "FDG_WAIT:\n\tld.relaxed.gpu.global.u32 %r15, [%rd22];\n\tsetp.ge.u32 %p7, %r15, " << want << ";\n\t@%p7 bra FDG_GO;\n"
              << "\tnanosleep.u32 64;\n\tmov.u64 %rd30, %globaltimer;\n\tsub.u64 %rd30, %rd30, %rd23;\n"
              << "\tsetp.gt.u64 %p7, %rd30, 4000000000;\n\t@!%p7 bra FDG_WAIT;\n"
              // a stalled pipeline ends instead of hanging: flag it and leave
              << "\tld.volatile.shared.u64 %rd30, [fdg_ring+80];\n\tmov.u32 %r15, 1;\n\tst.global.u32 [%rd30], %r15;\n\tbra FDG_DONE;\n"
              << "FDG_GO:\n\tfence.acq_rel.gpu;\n"
              << "\tmov.u64 %rd31, %clock64;\n\tsub.u64 %rd21, %rd31, %rd21;\n\tld.volatile.shared.u64 %rd30, [fdg_ring+80];\n"
              << "\t@%p3 red.global.add.u64 [%rd30+" << 24 + 16 * sg << "], %rd21;\n";
            // ---- per tile: sample index, validity, bases ----
            p << "\tcvt.u64.u32 %rd25, %r3;\n\tshl.b64 %rd0, %rd20, 5;\n\tadd.u64 %rd0, %rd0, %rd25;\n"
              << "\tld.volatile.shared.u64 %rd10, [fdg_ring+48];\n\tsetp.lt.s64 %p0, %rd0, %rd10;\n"
              << "\tselp.u64 %rd12, %rd0, 0, %p0;\n\tshl.b64 %rd12, %rd12, " << esh << ";\n"
              << "\tld.volatile.shared.u64 %rd1, [fdg_ring+0];\n\tadd.u64 %rd1, %rd1, %rd12;\n"
              << "\tld.volatile.shared.u64 %rd2, [fdg_ring+8];\n\tld.volatile.shared.u64 %rd4, [fdg_ring+24];\n"
              << "\tcvt.u32.u64 %r13, %rd2;\n\tcvt.u32.u64 %r14, %rd4;\n"
              << "\tld.volatile.shared.u64 %rd14, [fdg_ring+32];\n";
            if (n_boundary > 0)
                p << "\tshl.b64 %rd13, %rd0, " << esh << ";\n\tld.volatile.shared.u64 %rd9, [fdg_ring+104];\n\tadd.u64 %rd9, %rd9, %rd13;\n"
                  << "\tld.volatile.shared.u64 %rd11, [fdg_ring+112];\n\tcvt.u32.u64 %r20, %rd11;\n";
            if (n_cross > 0)
                p << "\trem.u64 %rd24, %rd20, %rd19;\n\tshl.b64 %rd13, %rd24, 5;\n\tadd.u64 %rd13, %rd13, %rd25;\n\tshl.b64 %rd13, %rd13, " << esh << ";\n"
                  << "\tld.volatile.shared.u64 %rd3, [fdg_ring+16];\n\tadd.u64 %rd3, %rd3, %rd13;\n";
            if (acc)
                // partial row of this warp: out + (block * warps per block + warp of the block) * nroots * 8
                p << "\tld.volatile.shared.u64 %rd5, [fdg_ring+40];\n\tmad.lo.u32 %r6, %r0, " << wpb << ", %r5;\n\tcvt.u64.u32 %rd15, %r6;\n"
                  << "\tmul.lo.u64 %rd15, %rd15, %rd5;\n\tshl.b64 %rd15, %rd15, 3;\n\tadd.u64 %rd7, %rd14, %rd15;\n";
            else
                p << "\tld.volatile.shared.u64 %rd5, [fdg_ring+40];\n\tshl.b64 %rd13, %rd0, " << esh << ";\n\tadd.u64 %rd6, %rd14, %rd13;\n";
            if (ring) {
                std::ostringstream pro;
                for (int j = 0; j < std::min(NR, n_in); ++j) {
                    ring_issue(pro, j);
                    if (j % G == G - 1 || j == std::min(NR, n_in) - 1) pro << "\tcp.async.commit_group;\n";
                }
                p << pro.str();
            }
            p << body;
            // ---- publish the tile, next tile of this warp ----
            p << "\tfence.acq_rel.gpu;\n\tbar.warp.sync 0xffffffff;\n"
              << "\tld.volatile.shared.u64 %rd16, [fdg_ring+56];\n\tshl.b64 %rd22, %rd20, 2;\n\tadd.u64 %rd22, %rd16, %rd22;\n\tmov.u32 %r15, " << sg + 1 << ";\n"
              << "\t@%p3 st.release.gpu.global.u32 [%rd22], %r15;\n"
              << "\tld.volatile.shared.u64 %rd18, [fdg_ring+88];\n\tadd.u64 %rd20, %rd20, %rd18;\n\tbra FDG_TILE;\n"
              << "FDG_DONE:\n";
            if (n_sacc > 0) {
                // lane l sums entries l, l + 32, ... over the 32 columns of its warp, in lane order, and writes them into the
                // warp's row of partial sums (each entry of the row has exactly one writer)
                p << "\tbar.warp.sync 0xffffffff;\n"
                  << "\tld.volatile.shared.u64 %rd14, [fdg_ring+32];\n\tld.volatile.shared.u64 %rd5, [fdg_ring+40];\n"
                  << "\tmad.lo.u32 %r6, %r0, " << wpb << ", %r5;\n\tcvt.u64.u32 %rd15, %r6;\n"
                  << "\tmul.lo.u64 %rd15, %rd15, %rd5;\n\tshl.b64 %rd15, %rd15, 3;\n\tadd.u64 %rd7, %rd14, %rd15;\n"
                  << "\tmov.u32 %r17, %r3;\n"
                  << "FDG_RED:\n\tsetp.ge.u32 %p6, %r17, " << n_sacc << ";\n\t@%p6 bra FDG_RED_END;\n"
                  << "\tmov.u32 %r18, fdg_ring;\n\tmad.lo.u32 %r18, %r17, " << e.sacc_stride << ", %r18;\n\tmad.lo.u32 %r18, %r5, 256, %r18;\n"
                  << "\tmov.f64 %fd" << e.nfd << ", 0d0000000000000000;\n\tmov.u32 %r19, 0;\n"
                  << "FDG_RED_IN:\n\tld.shared.f64 %fd" << e.nfd + 1 << ", [%r18+" << sacc0 << "];\n"
                  << "\tadd.rn.f64 %fd" << e.nfd << ", %fd" << e.nfd << ", %fd" << e.nfd + 1 << ";\n"
                  << "\tadd.u32 %r18, %r18, 8;\n\tadd.u32 %r19, %r19, 1;\n\tsetp.lt.u32 %p7, %r19, 32;\n\t@%p7 bra FDG_RED_IN;\n"
                  << "\tmov.u64 %rd22, fdg_rt" << sg << ";\n\tmul.wide.u32 %rd23, %r17, 4;\n\tadd.u64 %rd22, %rd22, %rd23;\n\tld.const.u32 %r19, [%rd22];\n"
                  << "\tmul.wide.u32 %rd23, %r19, 8;\n\tadd.u64 %rd23, %rd7, %rd23;\n\tst.global.f64 [%rd23], %fd" << e.nfd << ";\n"
                  << "\tadd.u32 %r17, %r17, 32;\n\tbra FDG_RED;\n"
                  << "FDG_RED_END:\n";
            }
            p
              // statistics: clocks this stage was alive / spent waiting, summed over its warps (lane 0 of each)
              << "\tmov.u64 %rd31, %clock64;\n\tld.volatile.shared.u64 %rd21, [%r7+128];\n\tsub.u64 %rd21, %rd31, %rd21;\n"
              << "\tld.volatile.shared.u64 %rd30, [fdg_ring+80];\n"
              << "\t@%p3 red.global.add.u64 [%rd30+" << 16 + 16 * sg << "], %rd21;\n"
              << "\tret;\n}\n";
            js.ptx = p.str();
            // issue-cycle estimate of one tile: an FP64 instruction holds the pipe two cycles, everything else one issue slot
            {
                int64_t n_fp64 = 0, n_other = 0;
                size_t pos = 0;
                while (pos < body.size()) {
                    const size_t eol = body.find('\n', pos);
                    const std::string line = body.substr(pos, eol == std::string::npos ? std::string::npos : eol - pos);
                    pos = eol == std::string::npos ? body.size() : eol + 1;
                    if (line.size() < 2 || line[0] != '\t') continue;
                    const bool f64 = line.find(".f64 ") != std::string::npos &&
                                     (line.compare(1, 3, "mul") == 0 || line.compare(1, 3, "add") == 0 || line.compare(1, 3, "sub") == 0 ||
                                      line.compare(1, 3, "fma") == 0);
                    if (f64) ++n_fp64;
                    else if (line.compare(1, 3, "neg") != 0) ++n_other;
                }
                plan.stage_cost.push_back(std::max<int64_t>(2 * n_fp64, n_fp64 + n_other) + 64);
            }
            continue;
        }
        if (bulk) {
            // ---- bulk form: persistent, warp-specialised -------------------------------------------------------------------
            // 384 threads, one block per SM.  Warps 0-7 (two warpgroups, 240 registers each after setmaxnreg) are the
            // consumers: tile after tile of 256 samples they run the straight-line code, reading every input row out of the
            // ring.  Warp 8 (its warpgroup shrinks to 24 registers) is the producer: it walks the row table of the kernel
            // and copies row after row of the tile global -> shared with cp.async.bulk, eight rows per mbarrier phase.
            // Shared memory: [0, 8 NG) full barriers, [8 NG, 16 NG) empty barriers, [256, ...) the ring, then the per-thread
            // running sums of the roots (accumulate mode).
            const int n_sacc = (int)e.sacc_pos.size();
            const int smem = sacc0 + n_sacc * e.sacc_stride;
            bulk_smem_max = std::max(bulk_smem_max, smem);
            const int CT = T;  // consumer threads
            p << ".extern .shared .align 128 .b8 fdg_ring[];\n";
            if (n_in > 0) {
                p << ".const .align 4 .u32 fdg_tab[" << n_in << "] = {";
                for (int j = 0; j < n_in; ++j) p << (j ? ", " : "") << (((uint32_t)in_rows[(size_t)j].first << 31) | (uint32_t)in_rows[(size_t)j].second);
                p << "};\n";
            }
            if (n_sacc > 0) {
                p << ".const .align 4 .u32 fdg_rt[" << n_sacc << "] = {";
                for (int j = 0; j < n_sacc; ++j) p << (j ? ", " : "") << e.sacc_pos[(size_t)j];
                p << "};\n";
            }
            p << ".visible .entry " << js.name << "(\n"
              << "\t.param .u64 p_leaf, .param .u64 p_ld_leaf, .param .u64 p_cross, .param .u64 p_ld_cross,\n"
              << "\t.param .u64 p_out, .param .u64 p_ld_root, .param .u64 p_batch, .param .u64 p_nroots)\n"
              << ".maxntid " << CT + 128 << ", 1, 1\n.minnctapersm 1\n{\n";
            p << "\t.reg .f64 %fd<" << e.nfd + 2 << ">;\n\t.reg .b64 %rd<" << e.nrd + 1 << ">;\n\t.reg .pred %p<" << e.np + 1
              << ">;\n\t.reg .b32 %r<" << e.nr + 1 << ">;\n";
            // ---- common prologue: barriers, role split ----
            p << "\tmov.u32 %r2, %tid.x;\n\tsetp.ne.u32 %p6, %r2, 0;\n\t@%p6 bra FDG_INITED;\n";
            for (int sl = 0; sl < NG; ++sl)
                p << "\tmbarrier.init.shared::cta.b64 [fdg_ring+" << 8 * sl << "], 4;\n"
                  << "\tmbarrier.init.shared::cta.b64 [fdg_ring+" << 8 * (NG + sl) << "], " << CT / 32 << ";\n";
            p << "\tfence.mbarrier_init.release.cluster;\nFDG_INITED:\n\tbar.sync 0;\n"
              << "\tsetp.ge.u32 %p6, %r2, " << CT << ";\n\t@%p6 bra FDG_PRODUCER;\n"
              << "\tsetmaxnreg.inc.sync.aligned.u32 " << (getenv("FDG_JIT_BULK_CREGS") ? atoi(getenv("FDG_JIT_BULK_CREGS")) : 240) << ";\n";
            // ---- consumers ----
            // %r2 tid, %r3 lane, %r5 warp, %p3 lane 0, %r4 tile (the one value carried from tile to tile), %rd10 batch
            p << "\tand.b32 %r3, %r2, 31;\n\tsetp.eq.u32 %p3, %r3, 0;\n\tshr.u32 %r5, %r2, 5;\n\tmov.u32 %r0, %ctaid.x;\n"
              << "\tmov.u32 %r12, fdg_ring;\n\tmad.lo.u32 %r12, %r2, " << ES << ", %r12;\n\tadd.u32 %r12, %r12, 256;\n";
            if (n_sacc > 0) {
                p << "\tmov.u32 %r16, fdg_ring;\n\tmad.lo.u32 %r16, %r2, 8, %r16;\n\tadd.u32 %r16, %r16, " << sacc0 << ";\n"
                  << "\tmov.u32 %r17, 0;\n\tmov.u32 %r18, %r16;\n\tmov.f64 %fd" << e.nfd << ", 0d0000000000000000;\n"
                  << "FDG_ZERO:\n\tst.shared.f64 [%r18], %fd" << e.nfd << ";\n\tadd.u32 %r18, %r18, " << e.sacc_stride << ";\n"
                  << "\tadd.u32 %r17, %r17, 1;\n\tsetp.lt.u32 %p6, %r17, " << n_sacc << ";\n\t@%p6 bra FDG_ZERO;\n";
            }
            p << "\tmov.u32 %r4, %r0;\n"
              << "FDG_TILE:\n"
              << "\tld.param.u64 %rd10, [p_batch];\n\tadd.u64 %rd8, %rd10, " << CT - 1 << ";\n\tshr.u64 %rd8, %rd8, " << (CT == 256 ? 8 : 7) << ";\n"  // tiles
              << "\tcvt.u64.u32 %rd9, %r4;\n\tsetp.ge.u64 %p6, %rd9, %rd8;\n\t@%p6 bra FDG_DONE;\n"
              << "\tshl.b64 %rd9, %rd9, " << (CT == 256 ? 8 : 7) << ";\n\tcvt.u64.u32 %rd8, %r2;\n\tadd.u64 %rd0, %rd9, %rd8;\n"
              << "\tsetp.lt.s64 %p0, %rd0, %rd10;\n\tshl.b64 %rd13, %rd0, " << esh << ";\n";
            if (n_cross > 0)
                p << "\tld.param.u64 %rd3, [p_cross];\n\tcvta.to.global.u64 %rd3, %rd3;\n\tadd.u64 %rd3, %rd3, %rd13;\n"
                  << "\tld.param.u64 %rd4, [p_ld_cross];\n\tshl.b64 %rd4, %rd4, " << esh << ";\n\tcvt.u32.u64 %r14, %rd4;\n";
            p << "\tld.param.u64 %rd14, [p_out];\n\tcvta.to.global.u64 %rd14, %rd14;\n";
            if (acc) {
                // partial row of this warp (roots beyond the shared-memory running sums go there tile by tile)
                p << "\tld.param.u64 %rd5, [p_nroots];\n\tmad.lo.u32 %r6, %r0, " << CT / 32 << ", %r5;\n\tcvt.u64.u32 %rd15, %r6;\n"
                  << "\tmul.lo.u64 %rd15, %rd15, %rd5;\n\tshl.b64 %rd15, %rd15, 3;\n\tadd.u64 %rd7, %rd14, %rd15;\n";
            } else {
                p << "\tld.param.u64 %rd5, [p_ld_root];\n\tshl.b64 %rd5, %rd5, " << esh << ";\n\tadd.u64 %rd6, %rd14, %rd13;\n";
            }
            p << body;
            p << "\tmov.u32 %r1, %nctaid.x;\n\tadd.u32 %r4, %r4, %r1;\n\tbra FDG_TILE;\nFDG_DONE:\n";
            if (n_sacc > 0) {
                // lane l adds up entries l, l + 32, ... over the 32 columns of its warp, in lane order, and adds the sum to the
                // warp's row of partial sums (every entry of that row has exactly one writer; the host zeroed it)
                p << "\tbar.warp.sync 0xffffffff;\n"
                  << "\tld.param.u64 %rd14, [p_out];\n\tcvta.to.global.u64 %rd14, %rd14;\n\tld.param.u64 %rd5, [p_nroots];\n"
                  << "\tmad.lo.u32 %r6, %r0, " << CT / 32 << ", %r5;\n\tcvt.u64.u32 %rd15, %r6;\n"
                  << "\tmul.lo.u64 %rd15, %rd15, %rd5;\n\tshl.b64 %rd15, %rd15, 3;\n\tadd.u64 %rd7, %rd14, %rd15;\n"
                  << "\tmov.u32 %r17, %r3;\n"
                  << "FDG_RED:\n\tsetp.ge.u32 %p6, %r17, " << n_sacc << ";\n\t@%p6 bra FDG_RED_END;\n"
                  << "\tmov.u32 %r18, fdg_ring;\n\tmad.lo.u32 %r18, %r17, " << e.sacc_stride << ", %r18;\n\tmad.lo.u32 %r18, %r5, 256, %r18;\n"
                  << "\tmov.f64 %fd" << e.nfd << ", 0d0000000000000000;\n\tmov.u32 %r19, 0;\n"
                  << "FDG_RED_IN:\n\tld.shared.f64 %fd" << e.nfd + 1 << ", [%r18+" << sacc0 << "];\n"
                  << "\tadd.rn.f64 %fd" << e.nfd << ", %fd" << e.nfd << ", %fd" << e.nfd + 1 << ";\n"
                  << "\tadd.u32 %r18, %r18, 8;\n\tadd.u32 %r19, %r19, 1;\n\tsetp.lt.u32 %p7, %r19, 32;\n\t@%p7 bra FDG_RED_IN;\n"
                  << "\tmov.u64 %rd22, fdg_rt;\n\tmul.wide.u32 %rd23, %r17, 4;\n\tadd.u64 %rd22, %rd22, %rd23;\n\tld.const.u32 %r19, [%rd22];\n"
                  << "\tmul.wide.u32 %rd23, %r19, 8;\n\tadd.u64 %rd23, %rd7, %rd23;\n"
                  << "\tld.global.f64 %fd" << e.nfd + 1 << ", [%rd23];\n\tadd.rn.f64 %fd" << e.nfd << ", %fd" << e.nfd + 1 << ", %fd" << e.nfd << ";\n"
                  << "\tst.global.f64 [%rd23], %fd" << e.nfd << ";\n"
                  << "\tadd.u32 %r17, %r17, 32;\n\tbra FDG_RED;\n"
                  << "FDG_RED_END:\n";
            }
            p << "\tret;\n";
            // ---- producer ----
            // All four warps of the producer's warpgroup copy: warp w takes rows w, w + 4, ... of every group, so four copies
            // are being set up at any time (one warp alone needs ~50 cycles per copy and cannot keep the ring full).  Every
            // value below is the same in all lanes of a warp (uniform datapath); one elected lane issues.  Warp 0 announces
            // the bytes of the group (expect_tx) -- copies of the other warps may complete before that: the transaction count
            // is signed and the phase cannot end before the arrival of all four warps.
            p << "FDG_PRODUCER:\n\tsetmaxnreg.dec.sync.aligned.u32 24;\n"
              << "\tshr.u32 %r5, %r2, 5;\n\tsub.u32 %r5, %r5, " << CT / 32 << ";\n\tsetp.eq.u32 %p3, %r5, 0;\n";
            if (n_in > 0) {
                const int tsh = CT == 256 ? 8 : 7;
                p << "\tld.param.u64 %rd10, [p_batch];\n\tadd.u64 %rd8, %rd10, " << CT - 1 << ";\n\tshr.u64 %rd8, %rd8, " << tsh << ";\n"  // %rd8 tiles
                  << "\tld.param.u64 %rd1, [p_leaf];\n\tcvta.to.global.u64 %rd1, %rd1;\n\tld.param.u64 %rd2, [p_ld_leaf];\n\tshl.b64 %rd2, %rd2, " << esh << ";\n"
                  << "\tld.param.u64 %rd3, [p_cross];\n\tcvta.to.global.u64 %rd3, %rd3;\n\tld.param.u64 %rd4, [p_ld_cross];\n\tshl.b64 %rd4, %rd4, " << esh << ";\n"
                  << "\tmov.u32 %r4, %ctaid.x;\n\tmov.u32 %r1, %nctaid.x;\n\tmov.u32 %r7, fdg_ring;\n"
                  << "FDG_PTILE:\n\tcvt.u64.u32 %rd9, %r4;\n\tsetp.ge.u64 %p6, %rd9, %rd8;\n\t@%p6 bra FDG_PEND;\n"
                  << "\tshl.b64 %rd9, %rd9, " << tsh << ";\n"                                                       // first sample of the tile
                  << "\tsub.u64 %rd11, %rd10, %rd9;\n\tmin.u64 %rd11, %rd11, " << CT << ";\n\tcvt.u32.u64 %r8, %rd11;\n"  // valid samples
                  << "\tshl.b32 %r8, %r8, " << esh << ";\n\tadd.u32 %r8, %r8, 15;\n\tand.b32 %r8, %r8, 0xfffffff0;\n"      // bytes per row, whole 16-byte units
                  << "\tshl.b64 %rd9, %rd9, " << esh << ";\n"                                                       // byte offset of the tile within a row
                  << "\tmov.u32 %r9, 0;\n\tmov.u32 %r10, 0;\n\tmov.u32 %r11, 1;\n"                               // group, its slot, parity to wait for
                  << "FDG_PGROUP:\n"
                  << "\tshl.b32 %r15, %r10, 3;\n\tadd.u32 %r15, %r15, %r7;\n"                                      // full barrier of the slot (+ 8 NG: empty)
                  << "\tmul.lo.u32 %r17, %r9, " << BG << ";\n"                                                       // first row of the group
                  << "\tsetp.ge.u32 %p5, %r17, " << n_in << ";\n"                                                    // a padding group: nothing to copy
                  << "\tsub.u32 %r18, " << n_in << ", %r17;\n\tmin.u32 %r18, %r18, " << BG << ";\n\tselp.u32 %r18, 0, %r18, %p5;\n"  // rows in the group
                  << "\tmul.lo.u32 %r19, %r18, %r8;\n"                                                               // bytes of the group
                  << "\tmov.u32 %r21, 0;\n"
                  << "FDG_PWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 %p7, [%r15+" << 8 * NG << "], %r11, " << bulk_hint << ";\n\t@%p7 bra FDG_PGO;\n"
                  << (bulk_psleep > 0 ? "\tnanosleep.u32 " + std::to_string(bulk_psleep) + ";\n" : std::string())  // experiment: idle producers poll less often
                  << "\tadd.u32 %r21, %r21, 1;\n\tsetp.lt.u32 %p7, %r21, " << (bulk_psleep > 0 ? 1u << 22 : 4096u) << ";\n\t@%p7 bra FDG_PWAIT;\n\ttrap;\n"
                  << "FDG_PGO:\n"
                  // every producer warp arrives on the full barrier of every group, rows or not (count 4): a phase cannot end
                  // before all four have passed their wait for it, so the parity a warp polls is never more than one phase old
                  << "\telect.sync %r22|%p4, 0xffffffff;\n\tand.pred %p5, %p4, %p3;\n\tnot.pred %p6, %p3;\n\tand.pred %p6, %p6, %p4;\n"
                  << "\t@%p5 mbarrier.arrive.expect_tx.shared::cta.b64 %rd31, [%r15], %r19;\n"
                  << "\t@%p6 mbarrier.arrive.shared::cta.b64 %rd31, [%r15];\n"
                  << "\tmov.u32 %r6, %r5;\n"                                                                         // row of the group this warp copies next
                  << "FDG_PROW:\n\tsetp.ge.u32 %p6, %r6, %r18;\n\t@%p6 bra FDG_PNEXT;\n"
                  << "\tadd.u32 %r20, %r17, %r6;\n\tmov.u64 %rd12, fdg_tab;\n\tmul.wide.u32 %rd13, %r20, 4;\n\tadd.u64 %rd12, %rd12, %rd13;\n\tld.const.u32 %r20, [%rd12];\n"
                  << "\tand.b32 %r22, %r20, 0x7fffffff;\n\tcvt.u64.u32 %rd13, %r22;\n\tsetp.ge.u32 %p6, %r20, 0x80000000;\n"
                  << "\tselp.u64 %rd14, %rd4, %rd2, %p6;\n\tselp.u64 %rd15, %rd3, %rd1, %p6;\n"
                  << "\tmad.lo.u64 %rd15, %rd13, %rd14, %rd15;\n\tadd.u64 %rd15, %rd15, %rd9;\n"
                  << "\tmad.lo.u32 %r23, %r10, " << BG << ", %r6;\n\tmad.lo.u32 %r23, %r23, " << ROWB << ", %r7;\n"
                  << "\t@%p4 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%r23+256], [%rd15], %r8, [%r15];\n";
                if (bulk_prefetch > 0)
                    // ... and the row `bulk_prefetch` groups further down the table is asked into L2 now: DRAM keeps working
                    // through the arithmetic-heavy stretches of the code, when the ring is full
                    p << "\tadd.u32 %r20, %r17, %r6;\n\tadd.u32 %r20, %r20, " << bulk_prefetch * BG << ";\n\tsetp.ge.u32 %p6, %r20, " << n_in << ";\n\t@%p6 bra FDG_PNOPF;\n"
                      << "\tmov.u64 %rd12, fdg_tab;\n\tmul.wide.u32 %rd13, %r20, 4;\n\tadd.u64 %rd12, %rd12, %rd13;\n\tld.const.u32 %r20, [%rd12];\n"
                      << "\tand.b32 %r22, %r20, 0x7fffffff;\n\tcvt.u64.u32 %rd13, %r22;\n\tsetp.ge.u32 %p6, %r20, 0x80000000;\n"
                      << "\tselp.u64 %rd14, %rd4, %rd2, %p6;\n\tselp.u64 %rd15, %rd3, %rd1, %p6;\n"
                      << "\tmad.lo.u64 %rd15, %rd13, %rd14, %rd15;\n\tadd.u64 %rd15, %rd15, %rd9;\n"
                      << "\t@%p4 cp.async.bulk.prefetch.L2.global [%rd15], %r8;\nFDG_PNOPF:\n";
                p << "\tadd.u32 %r6, %r6, 4;\n\tbra FDG_PROW;\n"
                  << "FDG_PNEXT:\n"
                  << "\tadd.u32 %r10, %r10, 1;\n\tsetp.eq.u32 %p6, %r10, " << NG << ";\n\t@%p6 xor.b32 %r11, %r11, 1;\n\t@%p6 mov.u32 %r10, 0;\n"  // next slot; the parity flips every NG groups
                  << "\tadd.u32 %r9, %r9, 1;\n\tsetp.lt.u32 %p6, %r9, " << b_groups_padded << ";\n\t@%p6 bra FDG_PGROUP;\n"
                  << "\tadd.u32 %r4, %r4, %r1;\n\tbra FDG_PTILE;\n";
            }
            p << "FDG_PEND:\n\tret;\n}\n";
            js.ptx = p.str();
            continue;
        }
        p << ".visible .entry " << js.name << "(\n"
          << "\t.param .u64 p_leaf, .param .u64 p_ld_leaf, .param .u64 p_cross, .param .u64 p_ld_cross,\n"
          << "\t.param .u64 p_out, .param .u64 p_ld_root, .param .u64 p_batch, .param .u64 p_nroots)\n"
          << ".maxntid 128, 1, 1\n";
        if (const char *mr = getenv("FDG_JIT_MAXNREG")) p << ".maxnreg " << atoi(mr) << "\n";
        p << "{\n";
        if (ring) p << "\t.shared .align 16 .b8 fdg_ring[" << NR * T * ES << "];\n";
        p << "\t.reg .f64 %fd<" << e.nfd + 2 << ">;\n\t.reg .b64 %rd<" << e.nrd + 1 + (ring ? NR : 0) << ">;\n\t.reg .pred %p<" << e.np + 1
          << ">;\n\t.reg .b32 %r<" << e.nr + 1 << ">;\n";
        // %rd0 = first sample of the thread, %rd1 = leaf base, %rd2 = ld_leaf bytes, %rd3 = cross base, %rd4 = ld_cross bytes,
        // %rd5 = ld_root bytes, %rd6 = root base (eval), %rd7 = partial row of the warp (accumulate)
        p << "\tmov.u32 %r0, %ctaid.x;\n\tmov.u32 %r1, %ntid.x;\n\tmov.u32 %r2, %tid.x;\n"
          << "\tmul.wide.u32 %rd8, %r0, %r1;\n\tcvt.u64.u32 %rd9, %r2;\n\tadd.u64 %rd8, %rd8, %rd9;\n"  // global thread id
          << "\tmul.lo.u64 %rd0, %rd8, " << spt_hdr << ";\n"
          << "\tld.param.u64 %rd10, [p_batch];\n"
          << "\tld.param.u64 %rd2, [p_ld_leaf];\n\tshl.b64 %rd2, %rd2, " << esh << ";\n"
          << "\tld.param.u64 %rd4, [p_ld_cross];\n\tshl.b64 %rd4, %rd4, " << esh << ";\n"
          << "\tcvt.u32.u64 %r13, %rd2;\n\tcvt.u32.u64 %r14, %rd4;\n"
          << "\tld.param.u64 %rd14, [p_out];\n\tcvta.to.global.u64 %rd14, %rd14;\n";
        if (acc) {
            p << "\tshr.u64 %rd15, %rd8, 5;\n\tld.param.u64 %rd5, [p_nroots];\n\tmul.lo.u64 %rd15, %rd15, %rd5;\n"
              << "\tshl.b64 %rd15, %rd15, 3;\n\tadd.u64 %rd7, %rd14, %rd15;\n"
              << "\tand.b32 %r3, %r2, 31;\n\tsetp.eq.u32 %p3, %r3, 0;\n";
        } else {
            p << "\tld.param.u64 %rd5, [p_ld_root];\n\tshl.b64 %rd5, %rd5, " << esh << ";\n";
        }
        if (e.persistent) {
            for (int r = 0; r < (int)low.R * W; ++r) p << "\tmov.f64 %fd" << e.racc0 + r << ", 0d0000000000000000;\n";
            p << "\tmov.u32 %r4, %nctaid.x;\n\tmul.wide.u32 %rd9, %r4, %r1;\n\tmul.lo.u64 %rd9, %rd9, " << spt_hdr << ";\n";  // grid stride
            p << "FDG_LOOP:\n";
        }
        // per tile: validity predicates and the bases that depend on the sample index
        p << "\tsetp.lt.s64 %p0, %rd0, %rd10;\n";  // first sample valid
        if (spt_hdr == 2)
            p << "\tadd.u64 %rd11, %rd0, 1;\n\tsetp.lt.s64 %p1, %rd11, %rd10;\n"  // both samples valid
              << "\tnot.pred %p4, %p1;\n\tand.pred %p2, %p0, %p4;\n";             // only the first one
        p << "\tselp.u64 %rd12, %rd0, 0, %p0;\n"  // inactive threads read sample 0: every address stays in bounds
          << "\tshl.b64 %rd12, %rd12, " << esh << ";\n"
          << "\tld.param.u64 %rd1, [p_leaf];\n\tcvta.to.global.u64 %rd1, %rd1;\n\tadd.u64 %rd1, %rd1, %rd12;\n"
          << "\tshl.b64 %rd13, %rd0, " << esh << ";\n";
        if (n_cross > 0) p << "\tld.param.u64 %rd3, [p_cross];\n\tcvta.to.global.u64 %rd3, %rd3;\n\tadd.u64 %rd3, %rd3, %rd13;\n";
        if (!acc) p << "\tadd.u64 %rd6, %rd14, %rd13;\n";
        if (ring) {
            p << "\tmov.u32 %r12, fdg_ring;\n\tmad.lo.u32 %r12, %r2, " << ES << ", %r12;\n";  // this thread's column of the ring
            std::ostringstream pro;
            for (int j = 0; j < std::min(NR, n_in); ++j) {
                ring_issue(pro, j);
                if (j % G == G - 1 || j == std::min(NR, n_in) - 1) pro << "\tcp.async.commit_group;\n";
            }
            p << pro.str();
        }
        p << body;
        if (e.persistent) {
            p << "\tadd.u64 %rd0, %rd0, %rd9;\n\tsetp.lt.s64 %p5, %rd0, %rd10;\n\t@%p5 bra FDG_LOOP;\n";
            for (int r = 0; r < (int)low.R * W; ++r) {
                const int sreg = e.racc0 + r, t = e.nfd;
                for (int m = 16; m >= 1; m >>= 1) {
                    p << "\tmov.b64 {%r8, %r9}, %fd" << sreg << ";\n"
                      << "\tshfl.sync.bfly.b32 %r10, %r8, " << m << ", 31, 0xffffffff;\n"
                      << "\tshfl.sync.bfly.b32 %r11, %r9, " << m << ", 31, 0xffffffff;\n"
                      << "\tmov.b64 %fd" << t << ", {%r10, %r11};\n"
                      << "\tadd.rn.f64 %fd" << sreg << ", %fd" << sreg << ", %fd" << t << ";\n";
                }
                p << "\t@%p3 st.global.f64 [%rd7+" << r * 8 << "], %fd" << sreg << ";\n";
            }
        }
        p << "\tret;\n}\n";
        js.ptx = p.str();
    }
    plan.bulk_smem = bulk_smem_max;
    if (pipe) {
        const int S = nseg, spp = pipe_stages_per_pass;
        plan.n_sm = pipe->n_sm();
        plan.n_pass = (S + spp - 1) / spp;
        plan.stages_per_pass = spp;
        plan.ring_bytes = ring_bytes_max;
        plan.stage_blocks.assign((size_t)S, 0);
        for (int k = 0; k < S; ++k) plan.stage_blocks[(size_t)k] = pipe_stage_sms[(size_t)(k % spp)];
        // ---- the entry kernel of each pass: the launch arguments go to shared memory, then SM -> stage ------------------
        // Which stage a block runs is looked up by the SM it landed on (p_smtab[smid] = stage of the pass, index of the block
        // within the stage, blocks of the stage): SMs that share an instruction cache run the same stage, and only the host
        // knows which SMs those are (fdg_capi.cu).
        plan.dispatch_ptx.clear();
        for (int ps = 0; ps < plan.n_pass; ++ps) {
            const int k0 = ps * spp, k1 = std::min(S, k0 + spp);
            std::ostringstream d;
            d << ".version 8.7\n.target sm_100a\n.address_size 64\n\n.extern .shared .align 16 .b8 fdg_ring[];\n";
            for (int k = k0; k < k1; ++k) d << ".extern .func fdg_stage" << k << "();\n";
            d << ".visible .entry fdg_pipe" << ps << "(\n"
              << "\t.param .u64 p_leaf, .param .u64 p_ld_leaf, .param .u64 p_cross, .param .u64 p_ld_cross,\n"
              << "\t.param .u64 p_out, .param .u64 p_ld_root, .param .u64 p_batch, .param .u64 p_nroots,\n"
              << "\t.param .u64 p_progress, .param .u64 p_ntiles, .param .u64 p_window, .param .u64 p_stats, .param .u64 p_smtab,\n"
              << "\t.param .u64 p_boundary, .param .u64 p_ld_boundary)\n"
              << ".maxntid " << pipe->threads << ", 1, 1\n{\n"
              << "\t.reg .b64 %rd<20>;\n\t.reg .b32 %r<8>;\n\t.reg .pred %p<2>;\n"
              << "\tld.param.u64 %rd0, [p_leaf];\n\tcvta.to.global.u64 %rd0, %rd0;\n\tld.param.u64 %rd1, [p_ld_leaf];\n\tshl.b64 %rd1, %rd1, " << esh << ";\n"
              << "\tld.param.u64 %rd2, [p_cross];\n\tcvta.to.global.u64 %rd2, %rd2;\n\tld.param.u64 %rd3, [p_ld_cross];\n\tshl.b64 %rd3, %rd3, " << esh << ";\n"
              << "\tld.param.u64 %rd4, [p_out];\n\tcvta.to.global.u64 %rd4, %rd4;\n";
            if (acc) d << "\tld.param.u64 %rd5, [p_nroots];\n";
            else d << "\tld.param.u64 %rd5, [p_ld_root];\n\tshl.b64 %rd5, %rd5, " << esh << ";\n";
            d << "\tld.param.u64 %rd6, [p_batch];\n"
              << "\tld.param.u64 %rd8, [p_progress];\n\tcvta.to.global.u64 %rd8, %rd8;\n\tld.param.u64 %rd9, [p_ntiles];\n"
              << "\tld.param.u64 %rd10, [p_window];\n\tld.param.u64 %rd11, [p_stats];\n\tcvta.to.global.u64 %rd11, %rd11;\n"
              << "\tld.param.u64 %rd15, [p_boundary];\n\tcvta.to.global.u64 %rd15, %rd15;\n\tld.param.u64 %rd16, [p_ld_boundary];\n\tshl.b64 %rd16, %rd16, " << esh << ";\n"
              << "\tmov.u32 %r1, %tid.x;\n\tsetp.eq.u32 %p1, %r1, 0;\n"
              << "\tld.param.u64 %rd13, [p_smtab];\n\tcvta.to.global.u64 %rd13, %rd13;\n\tmov.u32 %r0, %smid;\n"
              << "\tmul.wide.u32 %rd14, %r0, 16;\n\tadd.u64 %rd13, %rd13, %rd14;\n"
              << "\tld.global.v4.u32 {%r3, %r2, %r4, %r5}, [%rd13];\n"  // stage of the pass, index within the stage, blocks of the stage
              << "\tmul.lo.u32 %r4, %r4, " << pipe->threads / 32 << ";\n\tcvt.u64.u32 %rd12, %r4;\n";  // warps of the stage
            const int off[13] = {0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 104, 112};
            const int src[13] = {0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 15, 16};
            for (int q = 0; q < 13; ++q) d << "\t@%p1 st.volatile.shared.u64 [fdg_ring+" << off[q] << "], %rd" << src[q] << ";\n";
            d << "\t@%p1 st.volatile.shared.u64 [fdg_ring+88], %rd12;\n\t@%p1 st.volatile.shared.u32 [fdg_ring+96], %r2;\n"
              << "\tbar.sync 0;\n";
            for (int k = k0; k < k1; ++k) d << "\tsetp.eq.u32 %p0, %r3, " << k - k0 << ";\n\t@%p0 bra FDG_S" << k << ";\n";
            // an SM the table does not know: flag the launch as stalled and leave (the other blocks give up after their time limit)
            d << "\tmov.u32 %r6, 1;\n\t@%p1 st.global.u32 [%rd11], %r6;\n\tret;\n";
            for (int k = k0; k < k1; ++k) d << "FDG_S" << k << ":\n\tcall.uni fdg_stage" << k << ", ();\n\tret;\n";
            d << "}\n";
            plan.dispatch_ptx.push_back(d.str());
        }
    }
    (void)err;
    return FDG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX -> cubin (sm_100a), segments in parallel
// ---------------------------------------------------------------------------------------------------------------------
static int compile_one(JitSegment &js, bool fma, std::string &err, bool relocatable = false) {
    nvPTXCompilerHandle h = nullptr;
    nvPTXCompileResult rc = nvPTXCompilerCreate(&h, js.ptx.size(), js.ptx.c_str());
    if (rc != NVPTXCOMPILE_SUCCESS) {
        err = "nvPTXCompilerCreate failed (" + std::to_string((int)rc) + ")";
        return FDG_ERR_UNSUPPORTED;
    }
    const char *opts[] = {"--gpu-name=sm_100a", "--opt-level=3", fma ? "--fmad=true" : "--fmad=false", "--verbose", "--compile-only"};
    rc = nvPTXCompilerCompile(h, relocatable ? 5 : 4, opts);
    size_t n = 0;
    if (rc != NVPTXCOMPILE_SUCCESS) {
        nvPTXCompilerGetErrorLogSize(h, &n);
        std::string log(n + 1, '\0');
        if (n) nvPTXCompilerGetErrorLog(h, &log[0]);
        err = "ptxas failed for " + js.name + ": " + log.c_str();
        nvPTXCompilerDestroy(&h);
        return FDG_ERR_UNSUPPORTED;
    }
    nvPTXCompilerGetCompiledProgramSize(h, &n);
    js.cubin.resize(n);
    nvPTXCompilerGetCompiledProgram(h, js.cubin.data());
    if (const char *dir = getenv("FDG_JIT_DUMP_CUBIN")) {  // debugging aid: cuobjdump -sass / nvdisasm -plr on the kernels
        const std::string path = std::string(dir) + "/" + js.name + ".cubin";
        if (FILE *fp = std::fopen(path.c_str(), "wb")) {
            std::fwrite(js.cubin.data(), 1, js.cubin.size(), fp);
            std::fclose(fp);
        }
    }
    size_t ln = 0;
    nvPTXCompilerGetInfoLogSize(h, &ln);
    if (ln) {
        js.info.assign(ln + 1, '\0');
        nvPTXCompilerGetInfoLog(h, &js.info[0]);
        js.info.resize(std::strlen(js.info.c_str()));
    }
    nvPTXCompilerDestroy(&h);
    return FDG_OK;
}

double jit_spill_bytes(const JitPlan &plan) {
    double spill = 0;
    for (const JitSegment &sg : plan.seg) {
        const char *p = sg.info.c_str();
        while ((p = std::strstr(p, "bytes spill")) != nullptr) {  // "... N bytes spill stores, M bytes spill loads"
            const char *q = p;
            while (q > sg.info.c_str() && (q[-1] == ' ')) --q;
            while (q > sg.info.c_str() && q[-1] >= '0' && q[-1] <= '9') --q;
            spill += std::atof(q);
            p += 11;
        }
    }
    return spill;
}

double jit_model_ns(const JitPlan &plan, int bytes_per_element) {
    // Calibrated on B200 measurements of seven workloads with and without merged sub-expressions (profiles/r02_experiments/
    // cse_ab.log): the kernels move their planned bytes at up to ~5.0 TB/s and issue FP64 instructions at up to ~13.8 T/s
    // (75 % of the measured DMUL + DADD rate); the two overlap imperfectly (a smooth maximum, exponent 3).  A spilled
    // register (known once the kernels are assembled) costs about six instructions per 8 bytes stored or reloaded, and
    // half of the spilled bytes end up as memory traffic: local memory is cached, but 38 000 threads' worth of it competes
    // with the streamed rows for L2 (measured on the headline graph: 1.8 KB of DRAM writes per sample beyond the plan
    // at 5.7 KB of spill stores, profiles/r02_bulk_segments_summary.txt; scope experiments in r02_experiments/bulk_form.log).
    const double spill = jit_spill_bytes(plan);
    // (rows fetched a second time within a kernel come out of L2 -- measured: +1 % of DRAM reads for +17 % of rows -- and
    // are charged a sixth of a DRAM row)
    const double bytes = ((double)(plan.leaf_loads + plan.cross_loads + plan.cross_stores) + (double)plan.refetch_loads / 6.0) * bytes_per_element + 0.5 * spill;
    const double t_mem = bytes / 5000.0;                                         // bytes / (GB/s) = ns
    const double t_fp = ((double)plan.fp64_instr + spill / 8.0 * 6.0) / 13800.0;  // instructions / (G lane-instructions/s) = ns
    return std::cbrt(t_mem * t_mem * t_mem + t_fp * t_fp * t_fp) + 0.01 * (double)plan.seg.size();
}

int jit_assemble(const std::string &ptx, int opt_level, std::vector<char> &cubin, std::string &err) {
    nvPTXCompilerHandle h = nullptr;
    if (nvPTXCompilerCreate(&h, ptx.size(), ptx.c_str()) != NVPTXCOMPILE_SUCCESS) {
        err = "nvPTXCompilerCreate failed";
        return FDG_ERR_UNSUPPORTED;
    }
    const std::string ol = "--opt-level=" + std::to_string(opt_level);
    const char *opts[] = {"--gpu-name=sm_100a", ol.c_str()};
    size_t n = 0;
    if (nvPTXCompilerCompile(h, 2, opts) != NVPTXCOMPILE_SUCCESS) {
        nvPTXCompilerGetErrorLogSize(h, &n);
        std::string log(n + 1, '\0');
        if (n) nvPTXCompilerGetErrorLog(h, &log[0]);
        err = std::string("ptxas failed: ") + log.c_str();
        nvPTXCompilerDestroy(&h);
        return FDG_ERR_UNSUPPORTED;
    }
    nvPTXCompilerGetCompiledProgramSize(h, &n);
    cubin.resize(n);
    nvPTXCompilerGetCompiledProgram(h, cubin.data());
    nvPTXCompilerDestroy(&h);
    return FDG_OK;
}

// bytes of machine code of the kernel in a cubin: the size of its `.text.<name>` section (ELF64, little endian)
static size_t text_bytes(const std::vector<char> &cubin) {
    auto rd = [&](size_t off, int bytes) -> uint64_t {
        uint64_t v = 0;
        if (off + (size_t)bytes > cubin.size()) return 0;
        std::memcpy(&v, cubin.data() + off, (size_t)bytes);
        return v;
    };
    if (cubin.size() < 64 || std::memcmp(cubin.data(), "\177ELF", 4) != 0 || cubin[4] != 2) return cubin.size();
    const uint64_t shoff = rd(0x28, 8);
    const uint64_t shentsize = rd(0x3A, 2), shnum = rd(0x3C, 2), shstrndx = rd(0x3E, 2);
    if (shentsize < 64 || shstrndx >= shnum) return cubin.size();
    const uint64_t stroff = rd(shoff + shstrndx * shentsize + 0x18, 8);
    size_t best = 0;
    for (uint64_t i = 0; i < shnum; ++i) {
        const uint64_t sh = shoff + i * shentsize;
        const uint64_t name = stroff + rd(sh, 4);
        if (name + 6 <= cubin.size() && std::memcmp(cubin.data() + name, ".text.", 6) == 0) best = std::max<size_t>(best, (size_t)rd(sh + 0x20, 8));
    }
    return best ? best : cubin.size();
}

int jit_compile(JitPlan &plan, std::string &err) {
    const int n = (int)plan.seg.size();
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = std::max(1, std::min<int>(n, hw ? (int)hw : 4));
    std::atomic<int> next{0};
    std::vector<std::string> errs((size_t)n);
    std::vector<int> rcs((size_t)n, FDG_OK);
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            rcs[(size_t)i] = compile_one(plan.seg[(size_t)i], plan.fma, errs[(size_t)i], plan.pipeline);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    for (int i = 0; i < n; ++i)
        if (rcs[(size_t)i] != FDG_OK) {
            err = errs[(size_t)i];
            return rcs[(size_t)i];
        }
    plan.max_code_bytes = 0;
    for (auto &sg : plan.seg) plan.max_code_bytes = std::max<int64_t>(plan.max_code_bytes, (int64_t)text_bytes(sg.cubin));
    if (plan.pipeline) {
        // the stage functions were assembled as relocatable objects; the entry kernel of a pass joins the stages of that pass
        // (device link, no GPU needed)
        plan.linked.clear();
        for (int ps = 0; ps < plan.n_pass; ++ps) {
            JitSegment entry;
            entry.name = "fdg_pipe" + std::to_string(ps);
            entry.ptx = plan.dispatch_ptx[(size_t)ps];
            int rc = compile_one(entry, plan.fma, err, true);
            if (rc != FDG_OK) return rc;
            nvJitLinkHandle lh = nullptr;
            const char *lopts[] = {"-arch=sm_100a"};
            if (nvJitLinkCreate(&lh, 1, lopts) != NVJITLINK_SUCCESS) {
                err = "nvJitLinkCreate failed";
                return FDG_ERR_UNSUPPORTED;
            }
            bool ok = nvJitLinkAddData(lh, NVJITLINK_INPUT_CUBIN, entry.cubin.data(), entry.cubin.size(), entry.name.c_str()) == NVJITLINK_SUCCESS;
            const int k0 = ps * plan.stages_per_pass, k1 = std::min<int>((int)plan.seg.size(), k0 + plan.stages_per_pass);
            for (int k = k0; k < k1; ++k) {
                auto &sg = plan.seg[(size_t)k];
                ok = ok && nvJitLinkAddData(lh, NVJITLINK_INPUT_CUBIN, sg.cubin.data(), sg.cubin.size(), sg.name.c_str()) == NVJITLINK_SUCCESS;
            }
            ok = ok && nvJitLinkComplete(lh) == NVJITLINK_SUCCESS;
            if (!ok) {
                size_t ln = 0;
                nvJitLinkGetErrorLogSize(lh, &ln);
                std::string log(ln + 1, '\0');
                if (ln) nvJitLinkGetErrorLog(lh, &log[0]);
                err = std::string("device link of the pipeline kernel failed: ") + log.c_str();
                nvJitLinkDestroy(&lh);
                return FDG_ERR_UNSUPPORTED;
            }
            size_t nbytes = 0;
            nvJitLinkGetLinkedCubinSize(lh, &nbytes);
            plan.linked.emplace_back(nbytes);
            nvJitLinkGetLinkedCubin(lh, plan.linked.back().data());
            nvJitLinkDestroy(&lh);
            if (const char *dir = getenv("FDG_JIT_DUMP_CUBIN")) {
                const std::string path = std::string(dir) + "/" + entry.name + ".cubin";
                if (FILE *fp = std::fopen(path.c_str(), "wb")) {
                    std::fwrite(plan.linked.back().data(), 1, plan.linked.back().size(), fp);
                    std::fclose(fp);
                }
            }
        }
    }
    return FDG_OK;
}

}  // namespace fdg
