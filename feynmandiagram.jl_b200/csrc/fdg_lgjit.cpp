// fdg_lgjit.cpp -- the leaf generator specialised per graph (example/benchmark.jl:44-81,113-127).
//
// The table-driven kernel of fdg_capi.cu (fdg_leafgen2_kernel) interprets the leaf metadata: two thirds of the
// instructions it issues decode tables.  Here the same arithmetic is written out as straight-line PTX for the graph at
// hand -- loop-basis coefficients, time indices and row numbers are immediates, the sample's (K, tau) live in registers
// -- and assembled for sm_100a like the evaluator's kernels.  Covered: order-0 Green's functions, order-0 interactions
// and constant leaves (everything a bare Parquet / GV graph has); counter-term leaves (derivative orders >= 1) stay with
// the table-driven kernel, which is then run on those leaves only.
//
// One thread = one sample.  The leaves are walked momentum by momentum (loop-basis vector); the code of ~40 momenta
// fills one kernel (the instruction cache bounds a kernel), each kernel reads the (K, tau) rows it
// needs once and writes its rows of the leaf matrix.
//
// Arithmetic: the operations of the table-driven kernel in the same order -- dot products as fused multiply-adds over
// the non-zero coefficients, exp(x <= 0) from the 32-entry table of 2^(j/32) and a degree-6 polynomial
// (tools/check_exp_table.py), 1 / (1 + e) from the hardware estimate and two Newton steps -- except that results below
// 2^-1022 are flushed to zero (the table-driven kernel returns the denormal).
#include "fdg_jit.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <sstream>

namespace fdg {
namespace {

std::string imm(double f) {
    uint64_t u;
    std::memcpy(&u, &f, 8);
    char buf[32];
    std::snprintf(buf, sizeof(buf), "0d%016llX", (unsigned long long)u);
    return buf;
}

struct Gen {
    std::ostringstream os;
    int nfd = 0;
    int lines = 0;
    int fd() { return nfd++; }
    std::string F(int r) const { return "%fd" + std::to_string(r); }
    void op(const std::string &s) {
        os << "\t" << s << ";\n";
        ++lines;
    }
    // exp(x), x <= 0; result flushed to zero below 2^-1022
    int exp_neg(int x) {
        const int t = fd(), nd = fd(), r1 = fd(), r = fd(), tj = fd(), p = fd(), res = fd(), out = fd();
        op("fma.rn.f64 " + F(t) + ", " + F(x) + ", " + imm(46.16624130844683) + ", " + imm(6755399441055744.0));
        op("mov.b64 {%r4, %r5}, " + F(t));
        op("add.rn.f64 " + F(nd) + ", " + F(t) + ", " + imm(-6755399441055744.0));
        op("fma.rn.f64 " + F(r1) + ", " + F(nd) + ", " + imm(-0.021660849364707246) + ", " + F(x));
        op("fma.rn.f64 " + F(r) + ", " + F(nd) + ", " + imm(-2.7791044496520866e-11) + ", " + F(r1));
        int q = fd();
        op("fma.rn.f64 " + F(q) + ", " + F(r) + ", " + imm(1.0 / 720.0) + ", " + imm(1.0 / 120.0));
        const double c[4] = {1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
        for (int i = 0; i < 4; ++i) {
            const int q2 = fd();
            op("fma.rn.f64 " + F(q2) + ", " + F(q) + ", " + F(r) + ", " + imm(c[i]));
            q = q2;
        }
        const int qr = fd();
        op("mul.rn.f64 " + F(qr) + ", " + F(q) + ", " + F(r));
        op("shl.b32 %r6, %r4, 3");
        op("and.b32 %r6, %r6, 248");
        op("add.u32 %r6, %r6, %r3");
        op("ld.shared.f64 " + F(tj) + ", [%r6]");
        op("fma.rn.f64 " + F(p) + ", " + F(tj) + ", " + F(qr) + ", " + F(tj));
        // 2^m onto the exponent: p in [1, 2), m >= -1022 wherever the result is kept
        op("shr.s32 %r7, %r4, 5");
        op("shl.b32 %r7, %r7, 20");
        op("mov.b64 {%r8, %r9}, " + F(p));
        op("add.s32 %r9, %r9, %r7");
        op("mov.b64 " + F(res) + ", {%r8, %r9}");
        op("setp.lt.f64 %p2, " + F(x) + ", " + imm(-708.0));
        op("selp.f64 " + F(out) + ", " + imm(0.0) + ", " + F(res) + ", %p2");
        return out;
    }
    // 1 / d, d in [1, 2]
    int rcp(int d) {
        const int y0 = fd(), e0 = fd(), y1 = fd(), e1 = fd(), y2 = fd(), nd = fd();
        op("rcp.approx.ftz.f64 " + F(y0) + ", " + F(d));
        op("neg.f64 " + F(nd) + ", " + F(d));
        op("fma.rn.f64 " + F(e0) + ", " + F(nd) + ", " + F(y0) + ", " + imm(1.0));
        op("fma.rn.f64 " + F(y1) + ", " + F(y0) + ", " + F(e0) + ", " + F(y0));
        op("fma.rn.f64 " + F(e1) + ", " + F(nd) + ", " + F(y1) + ", " + imm(1.0));
        op("fma.rn.f64 " + F(y2) + ", " + F(y1) + ", " + F(e1) + ", " + F(y1));
        return y2;
    }
};

}  // namespace

int lgjit_build(const std::vector<LgJitBasis> &bases, int n_loops, int dim, int n_tau, double kF2, double beta, double lambda, bool wide,
                int budget, std::vector<JitSegment> &out, std::string &err) {
    out.clear();
    if (dim != 2 && dim != 3) {
        err = "dim must be 2 or 3";
        return FDG_ERR_UNSUPPORTED;
    }
    int maxnreg = 0;  // registers per thread (0: ptxas decides -- 150-170, three blocks per SM)
    if (const char *e = getenv("FDG_LG_JIT_MAXNREG")) maxnreg = atoi(e);
    if (budget <= 0) budget = 4400;  // lines of PTX per kernel: ~100 KB of machine code, inside the instruction cache
    // 2^(j/32), correctly rounded
    static const char *tab[32] = {
        "3FF0000000000000", "3FF059B0D3158574", "3FF0B5586CF9890F", "3FF11301D0125B51", "3FF172B83C7D517B", "3FF1D4873168B9AA",
        "3FF2387A6E756238", "3FF29E9DF51FDEE1", "3FF306FE0A31B715", "3FF371A7373AA9CB", "3FF3DEA64C123422", "3FF44E086061892D",
        "3FF4BFDAD5362A27", "3FF5342B569D4F82", "3FF5AB07DD485429", "3FF6247EB03A5585", "3FF6A09E667F3BCD", "3FF71F75E8EC5F74",
        "3FF7A11473EB0187", "3FF82589994CCE13", "3FF8ACE5422AA0DB", "3FF93737B0CDC5E5", "3FF9C49182A3F090", "3FFA5503B23E255D",
        "3FFAE89F995AD3AD", "3FFB7F76F2FB5E47", "3FFC199BDD85529C", "3FFCB720DCEF9069", "3FFD5818DCFBA487", "3FFDFC97337B9B5F",
        "3FFEA4AFA2A490DA", "3FFF50765B6E4540"};
    size_t ib = 0;
    while (ib < bases.size()) {
        Gen g;
        std::set<int> k_rows, t_rows;  // rows of K (loop * dim + c) and of T this kernel reads
        // registers of the variables are named by row: %fk<row>, %ft<row>
        auto K = [&](int loop, int c) { return "%fk" + std::to_string(loop * dim + c); };
        auto T = [&](int t) { return "%ft" + std::to_string(t); };
        int n_store = 0;
        auto store = [&](int row, const std::string &val) {
            if (wide) g.op("mad.lo.u64 %rd8, %rd7, " + std::to_string(row) + ", %rd6");
            else g.op("mad.wide.u32 %rd8, %r2, " + std::to_string(row) + ", %rd6");
            g.op("st.global.f64 [%rd8], " + val);
            ++n_store;
        };
        const size_t first = ib;
        for (; ib < bases.size() && (ib == first || g.lines < budget); ++ib) {
            const LgJitBasis &B = bases[ib];
            if (B.g0.empty() && B.w0_out.empty() && B.one_out.empty()) continue;
            for (const int row : B.one_out) store(row, imm(1.0));
            if (B.g0.empty() && B.w0_out.empty()) continue;
            // |K . basis|^2
            int kq[3] = {-1, -1, -1};
            for (int c = 0; c < dim; ++c) {
                for (int n = 0; n < B.nnz; ++n) {
                    k_rows.insert(B.idx[n] * dim + c);
                    const int r = g.fd();
                    if (n == 0) g.op("mul.rn.f64 " + g.F(r) + ", " + K(B.idx[n], c) + ", " + imm(B.coef[n]));
                    else g.op("fma.rn.f64 " + g.F(r) + ", " + K(B.idx[n], c) + ", " + imm(B.coef[n]) + ", " + g.F(kq[c]));
                    kq[c] = r;
                }
                if (B.nnz == 0) {
                    kq[c] = g.fd();
                    g.op("mov.f64 " + g.F(kq[c]) + ", " + imm(0.0));
                }
            }
            int q2 = g.fd();
            g.op("mul.rn.f64 " + g.F(q2) + ", " + g.F(kq[0]) + ", " + g.F(kq[0]));
            for (int c = 1; c < dim; ++c) {
                const int r = g.fd();
                g.op("fma.rn.f64 " + g.F(r) + ", " + g.F(kq[c]) + ", " + g.F(kq[c]) + ", " + g.F(q2));
                q2 = r;
            }
            if (!B.w0_out.empty()) {
                // 8 pi / invK with invK = 1 / (q^2 + lambda)
                const int sm = g.fd(), wv = g.fd();
                g.op("add.rn.f64 " + g.F(sm) + ", " + g.F(q2) + ", " + imm(lambda));
                g.op("mul.rn.f64 " + g.F(wv) + ", " + g.F(sm) + ", " + imm(25.132741228718345));
                for (const int row : B.w0_out) store(row, g.F(wv));
            }
            if (B.g0.empty()) continue;
            // green(tau, w, beta) = s exp(-|w| x) / (1 + exp(-|w| beta)), x in (0, beta], s = sign(tau); st = tau sign(w),
            // x = st if st > 0 else st + beta
            const int w = g.fd(), aw = g.fd(), naw = g.fd(), ab = g.fd(), den = g.fd(), sw = g.fd();
            g.op("add.rn.f64 " + g.F(w) + ", " + g.F(q2) + ", " + imm(-kF2));
            g.op("abs.f64 " + g.F(aw) + ", " + g.F(w));
            g.op("neg.f64 " + g.F(naw) + ", " + g.F(aw));
            g.op("mul.rn.f64 " + g.F(ab) + ", " + g.F(naw) + ", " + imm(beta));
            const int ebeta = g.exp_neg(ab);
            g.op("add.rn.f64 " + g.F(den) + ", " + g.F(ebeta) + ", " + imm(1.0));
            const int inv = g.rcp(den);
            g.op("setp.gt.f64 %p3, " + g.F(w) + ", " + imm(0.0));
            g.op("selp.f64 " + g.F(sw) + ", " + imm(1.0) + ", " + imm(-1.0) + ", %p3");
            for (const LgJitBasis::G0 &lf : B.g0) {
                t_rows.insert(lf.tau_in);
                t_rows.insert(lf.tau_out);
                const int tau0 = g.fd(), tau = g.fd(), st = g.fd(), xb = g.fd(), x = g.fd(), arg = g.fd(), v = g.fd(), sv = g.fd();
                g.op("sub.rn.f64 " + g.F(tau0) + ", " + T(lf.tau_out) + ", " + T(lf.tau_in));
                g.op("setp.eq.f64 %p4, " + g.F(tau0) + ", " + imm(0.0));
                g.op("selp.f64 " + g.F(tau) + ", " + imm(-1e-10) + ", " + g.F(tau0) + ", %p4");
                g.op("mul.rn.f64 " + g.F(st) + ", " + g.F(tau) + ", " + g.F(sw));
                g.op("add.rn.f64 " + g.F(xb) + ", " + g.F(st) + ", " + imm(beta));
                g.op("setp.gt.f64 %p5, " + g.F(st) + ", " + imm(0.0));
                g.op("selp.f64 " + g.F(x) + ", " + g.F(st) + ", " + g.F(xb) + ", %p5");
                g.op("mul.rn.f64 " + g.F(arg) + ", " + g.F(naw) + ", " + g.F(x));
                const int e = g.exp_neg(arg);
                g.op("mul.rn.f64 " + g.F(v) + ", " + g.F(e) + ", " + g.F(inv));
                g.op("copysign.f64 " + g.F(sv) + ", " + g.F(tau) + ", " + g.F(v));
                store(lf.out, g.F(sv));
            }
        }
        if (n_store == 0) continue;
        JitSegment js;
        js.name = "fdg_lg" + std::to_string(out.size());
        js.n_stmts = g.lines;
        std::ostringstream p;
        p << ".version 8.7\n.target sm_100a\n.address_size 64\n\n.const .align 8 .b64 fdg_lg_tab[32] = {";
        for (int j = 0; j < 32; ++j) p << (j ? ", " : "") << "0x" << tab[j];
        p << "};\n.visible .entry " << js.name
          << "(\n\t.param .u64 p_K, .param .u64 p_T, .param .u64 p_ld_var, .param .u64 p_batch, .param .u64 p_leaf, .param .u64 p_ld_leaf)\n"
          << ".maxntid 128, 1, 1\n" << (maxnreg > 0 ? ".maxnreg " + std::to_string(maxnreg) + "\n" : std::string()) << "{\n\t.shared .align 8 .b8 lgtab[256];\n"
          << "\t.reg .f64 %fd<" << g.nfd + 2 << ">;\n\t.reg .f64 %fk<" << n_loops * dim + 1 << ">;\n\t.reg .f64 %ft<" << n_tau + 1 << ">;\n"
          << "\t.reg .b64 %rd<12>;\n\t.reg .b32 %r<12>;\n\t.reg .pred %p<8>;\n"
          // the table of 2^(j/32) into shared memory
          << "\tmov.u32 %r0, %tid.x;\n\tmov.u32 %r1, %ctaid.x;\n\tmov.u32 %r3, lgtab;\n\tsetp.ge.u32 %p0, %r0, 32;\n\t@%p0 bra FDG_LG_TAB;\n"
          << "\tmul.wide.u32 %rd0, %r0, 8;\n\tmov.u64 %rd1, fdg_lg_tab;\n\tadd.u64 %rd1, %rd1, %rd0;\n\tld.const.f64 %fd" << g.nfd << ", [%rd1];\n"
          << "\tshl.b32 %r4, %r0, 3;\n\tadd.u32 %r4, %r4, %r3;\n\tst.shared.f64 [%r4], %fd" << g.nfd << ";\n"
          << "FDG_LG_TAB:\n\tbar.sync 0;\n"
          // this thread's sample; threads past the end of the batch are done
          << "\tmul.wide.u32 %rd2, %r1, 128;\n\tcvt.u64.u32 %rd3, %r0;\n\tadd.u64 %rd2, %rd2, %rd3;\n"
          << "\tld.param.u64 %rd3, [p_batch];\n\tsetp.ge.u64 %p1, %rd2, %rd3;\n\t@%p1 bra FDG_LG_DONE;\n"
          << "\tshl.b64 %rd2, %rd2, 3;\n"
          << "\tld.param.u64 %rd4, [p_ld_var];\n\tshl.b64 %rd4, %rd4, 3;\n"
          << "\tld.param.u64 %rd5, [p_K];\n\tcvta.to.global.u64 %rd5, %rd5;\n\tadd.u64 %rd5, %rd5, %rd2;\n";
        for (const int row : k_rows)
            p << "\tmad.lo.u64 %rd8, %rd4, " << row << ", %rd5;\n\tld.global.nc.f64 %fk" << row << ", [%rd8];\n";
        p << "\tld.param.u64 %rd5, [p_T];\n\tcvta.to.global.u64 %rd5, %rd5;\n\tadd.u64 %rd5, %rd5, %rd2;\n";
        for (const int row : t_rows)
            p << "\tmad.lo.u64 %rd8, %rd4, " << row << ", %rd5;\n\tld.global.nc.f64 %ft" << row << ", [%rd8];\n";
        p << "\tld.param.u64 %rd6, [p_leaf];\n\tcvta.to.global.u64 %rd6, %rd6;\n\tadd.u64 %rd6, %rd6, %rd2;\n"
          << "\tld.param.u64 %rd7, [p_ld_leaf];\n\tshl.b64 %rd7, %rd7, 3;\n\tcvt.u32.u64 %r2, %rd7;\n"
          << g.os.str() << "FDG_LG_DONE:\n\tret;\n}\n";
        js.ptx = p.str();
        out.push_back(std::move(js));
    }
    return FDG_OK;
}

}  // namespace fdg
