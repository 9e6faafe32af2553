// fdg_lower.cpp -- lowering of the reference's Graph DAG to the VM program (host only, no CUDA).
//
// Stage A reproduces the reference emitter's statement order and leaf numbering
//   (src/backend/static.jl:98-133: post-order DFS over `graphs`, children in stored order,
//    first visit of an id wins, leaf k = k-th distinct leaf id, `root[r] = g` right after the
//    statement of a node whose id is in `root`, r = first position of that id).
// Stage B decides which values live where: single-use inner nodes are folded straight into their
//   parent's accumulator (acc_d registers), multi-use nodes and roots are materialised in the
//   shared-memory slot file, leaves are staged into slots by asynchronous loads.
// Stage C allocates slots (Belady eviction, spills to global scratch when a program does not fit),
//   hoists leaf loads ahead of their first use (software prefetch) and packs packets.
//
// None of this changes any arithmetic: every fold keeps the stored operand order, so the packet
// program computes exactly the values of the emitted function (bit for bit, see fdg_isa.h).
#include "fdg_lower.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <unordered_map>

#include "fdg_isa.h"

namespace fdg {
namespace {

// symbolic (pre-allocation) VM operation
struct Sym {
    uint8_t op;  // FDG_OP_*
    uint8_t k;   // operand count (TERM / MOV / MUL / ADD)
    uint8_t push, first;
    int32_t vals[FDG_TERM_MAX];  // values read; vals[0] is also the value defined by ST
    int32_t arg;                 // POW exponent / ROOT position
    double f;
};

inline int reads_count(const Sym &s) {
    switch (s.op) {
        case FDG_OP_TERM:
        case FDG_OP_MOV:
        case FDG_OP_MUL:
        case FDG_OP_ADD: return s.k;
        case FDG_OP_MULF:
        case FDG_OP_XADDF:
        case FDG_OP_XMULF: return 1;
        default: return 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Stage A: emitter order
// ------------------------------------------------------------------------------------------------
int build_statements(const fdg_graph_desc &g, std::vector<Stmt> &st, std::vector<Operand> &ops, Lowered &out,
                     std::string &err) {
    const int64_t n = g.n_nodes;
    if (n < 0 || g.n_edges < 0 || g.n_graphs < 0 || g.n_roots < 0) {
        err = "negative size in graph description";
        return FDG_ERR_BAD_ARG;
    }
    if (n > 0 && (!g.node_id || !g.node_op || !g.node_pow || !g.child_ptr)) {
        err = "null node array";
        return FDG_ERR_BAD_ARG;
    }
    if (g.n_edges > 0 && (!g.child_node || !g.child_factor)) {
        err = "null edge array";
        return FDG_ERR_BAD_ARG;
    }
    if ((g.n_graphs > 0 && !g.graphs) || (g.n_roots > 0 && !g.root_id)) {
        err = "null graphs/root array";
        return FDG_ERR_BAD_ARG;
    }
    if (n > 0 && (g.child_ptr[0] != 0 || g.child_ptr[n] != g.n_edges)) {
        err = "child_ptr does not span [0, n_edges]";
        return FDG_ERR_BAD_GRAPH;
    }
    for (int64_t i = 0; i < n; ++i) {
        const int64_t a = g.child_ptr[i], b = g.child_ptr[i + 1];
        if (b < a) {
            err = "child_ptr not monotone";
            return FDG_ERR_BAD_GRAPH;
        }
        const int32_t op = g.node_op[i];
        if (op < FDG_OP_UNITARY || op > FDG_OP_POWER) {
            // static.jl:6-11: "Static representation ... with operator X not yet implemented"
            err = "unknown operator code " + std::to_string(op) + " at node " + std::to_string(i);
            return FDG_ERR_BAD_GRAPH;
        }
        const int64_t nc = b - a;
        if (op == FDG_OP_UNITARY && nc != 0) {
            err = "Unitary node with subgraphs";
            return FDG_ERR_BAD_GRAPH;
        }
        if (op == FDG_OP_POWER && nc != 0) {
            if (nc != 1) {
                err = "Power node must have exactly one subgraph";
                return FDG_ERR_BAD_GRAPH;
            }
            if (g.node_pow[i] < 2) {
                err = "Power{N} with N < 2 is not supported (N=" + std::to_string(g.node_pow[i]) + ")";
                return FDG_ERR_BAD_GRAPH;
            }
            if (g.node_pow[i] >= (1 << 30)) {
                err = "Power exponent too large";
                return FDG_ERR_UNSUPPORTED;
            }
        }
        for (int64_t e = a; e < b; ++e) {
            const int32_t c = g.child_node[e];
            if (c < 0 || c >= n) {
                err = "child index out of range";
                return FDG_ERR_BAD_GRAPH;
            }
        }
    }
    for (int64_t i = 0; i < g.n_graphs; ++i)
        if (g.graphs[i] < 0 || g.graphs[i] >= n) {
            err = "graphs[] index out of range";
            return FDG_ERR_BAD_GRAPH;
        }

    // `findfirst(x -> x == g_id, root)`: the first position of an id is the one written (static.jl:112)
    std::unordered_map<int64_t, int32_t> root_pos;
    root_pos.reserve((size_t)g.n_roots * 2 + 1);
    for (int64_t r = 0; r < g.n_roots; ++r) root_pos.emplace(g.root_id[r], (int32_t)r);
    out.R = g.n_roots;
    out.root_set.assign((size_t)g.n_roots, 0);
    out.last_root = -1;

    std::unordered_map<int64_t, int32_t> val_of_id;
    val_of_id.reserve((size_t)n * 2 + 1);
    std::vector<uint8_t> color((size_t)n, 0);  // 0 new, 1 on the DFS stack, 2 done
    struct Fr {
        int32_t node;
        int64_t e;
    };
    std::vector<Fr> stack;
    for (int64_t gi = 0; gi < g.n_graphs; ++gi) {
        if (color[g.graphs[gi]] == 2) continue;  // the whole object sub-DAG was walked already
        stack.push_back({g.graphs[gi], g.child_ptr[g.graphs[gi]]});
        color[g.graphs[gi]] = 1;
        while (!stack.empty()) {
            Fr &fr = stack.back();
            const int32_t u = fr.node;
            if (fr.e < g.child_ptr[u + 1]) {
                const int32_t c = g.child_node[fr.e++];
                if (color[c] == 1) {
                    err = "graph has a cycle through node " + std::to_string(c);
                    return FDG_ERR_BAD_GRAPH;
                }
                if (color[c] == 0) {
                    color[c] = 1;
                    stack.push_back({c, g.child_ptr[c]});
                }
                continue;
            }
            // post-visit of object u
            stack.pop_back();
            color[u] = 2;
            const int64_t uid = g.node_id[u];
            {
                auto seen = val_of_id.find(uid);
                if (seen != val_of_id.end()) {  // `g_id in inds_visited... && continue`
                    // static.jl:116,122 keep one visited list for leaves and one for inner nodes: an id carried by both kinds
                    // of object would be emitted twice there.  That graph is ambiguous: reject it.
                    const bool was_leaf = st[(size_t)seen->second].op < 0, is_leaf = g.child_ptr[u] == g.child_ptr[u + 1];
                    if (was_leaf != is_leaf) {
                        err = "node id " + std::to_string(uid) + " is carried by a leaf and by an inner node";
                        return FDG_ERR_BAD_GRAPH;
                    }
                    continue;
                }
            }
            Stmt s;
            const int64_t a = g.child_ptr[u], b = g.child_ptr[u + 1];
            if (a == b) {  // leaf: isempty(subgraphs(g)), whatever the operator tag (static.jl:115)
                s.op = -1;
                s.leaf = (int32_t)out.L++;
                out.leaf_node.push_back(u);
            } else {
                s.op = (int8_t)g.node_op[u];
                s.pow_n = g.node_pow[u];
                s.first = (int64_t)ops.size();
                s.count = (int32_t)(b - a);
                for (int64_t e = a; e < b; ++e) {
                    auto it = val_of_id.find(g.node_id[g.child_node[e]]);
                    if (it == val_of_id.end()) {
                        err = "internal: child statement missing";
                        return FDG_ERR_BAD_GRAPH;
                    }
                    ops.push_back({it->second, g.child_factor[e]});
                }
                out.N++;
            }
            auto rp = root_pos.find(uid);
            if (rp != root_pos.end()) {
                s.root = rp->second;
                out.root_set[(size_t)rp->second] = 1;
                out.last_root = rp->second;
            }
            val_of_id.emplace(uid, (int32_t)st.size());
            st.push_back(s);
        }
    }
    if (out.L >= FDG_MAX_LEAVES) {
        err = "too many leaves for the packet encoding";
        return FDG_ERR_UNSUPPORTED;
    }
    if (out.R >= (1 << 30)) {
        err = "too many roots for the packet encoding";
        return FDG_ERR_UNSUPPORTED;
    }
    return FDG_OK;
}

// liveness and use counts (operands always precede their users)
void mark_live(std::vector<Stmt> &st, const std::vector<Operand> &ops) {
    for (Stmt &s : st) {
        s.live = false;
        s.uses = 0;
    }
    for (int64_t v = (int64_t)st.size() - 1; v >= 0; --v) {
        Stmt &s = st[(size_t)v];
        if (s.root >= 0) s.live = true;
        if (!s.live || s.op < 0) continue;
        for (int32_t i = 0; i < s.count; ++i) {
            Stmt &c = st[(size_t)ops[(size_t)(s.first + i)].val];
            c.live = true;
            c.uses++;
        }
    }
}

// Common-subexpression elimination on the emitted function (the hash-based analogue of the reference's
// optimize!(level=1) / remove_duplicated_nodes!, src/computational_graph/optimize.jl:345-390): two statements with the
// same operator and the same (operand, factor) sequence compute the same bits, so later copies read the first one.
// Operand ORDER is part of the key -- the fold order defines the rounding -- so nothing is re-associated.
int64_t eliminate_common_subexpressions(std::vector<Stmt> &st, std::vector<Operand> &ops, int64_t min_cost, std::vector<int32_t> &canon) {
    canon.assign(st.size(), 0);
    std::vector<int64_t> cost(st.size(), 0);  // operations needed to recompute the statement from leaves
    std::unordered_map<uint64_t, std::vector<int32_t>> buckets;
    buckets.reserve(st.size() * 2 + 1);
    int64_t removed = 0;
    for (int32_t v = 0; v < (int32_t)st.size(); ++v) {
        Stmt &s = st[(size_t)v];
        canon[(size_t)v] = v;
        if (s.op < 0) continue;
        uint64_t h = 1469598103934665603ull ^ (uint64_t)(s.op * 131 + s.pow_n);
        for (int32_t i = 0; i < s.count; ++i) {
            Operand &o = ops[(size_t)(s.first + i)];
            o.val = canon[(size_t)o.val];
            uint64_t fb;
            std::memcpy(&fb, &o.f, 8);
            h = (h ^ (uint64_t)(uint32_t)o.val) * 1099511628211ull;
            h = (h ^ fb) * 1099511628211ull;
        }
        int64_t c = s.op == FDG_OP_POWER ? s.pow_n : s.count;
        for (int32_t i = 0; i < s.count; ++i) c += cost[(size_t)ops[(size_t)(s.first + i)].val] + (ops[(size_t)(s.first + i)].f != 1.0);
        cost[(size_t)v] = std::min<int64_t>(c, 1 << 30);
        // a cheap copy is cheaper to recompute where it is needed than to keep alive (registers, slots, cross buffer)
        if (!s.live || cost[(size_t)v] < min_cost) continue;
        auto &b = buckets[h];
        int32_t found = -1;
        for (int32_t u : b) {
            const Stmt &t = st[(size_t)u];
            if (t.op != s.op || t.pow_n != s.pow_n || t.count != s.count) continue;
            bool same = true;
            for (int32_t i = 0; i < s.count && same; ++i) {
                const Operand &x = ops[(size_t)(s.first + i)], &y = ops[(size_t)(t.first + i)];
                same = x.val == y.val && std::memcmp(&x.f, &y.f, 8) == 0;
            }
            if (same) {
                found = u;
                break;
            }
        }
        if (found >= 0 && s.root < 0) {
            canon[(size_t)v] = found;  // every later reader uses the first copy; this statement becomes dead
            ++removed;
        } else if (found < 0) {
            b.push_back(v);
        }
    }
    return removed;
}

// ------------------------------------------------------------------------------------------------
// Stage B: symbolic code generation (stack machine: A on top of R1..R3)
// ------------------------------------------------------------------------------------------------
struct CodeGen {
    const std::vector<Stmt> &st;
    const std::vector<Operand> &ops;
    const std::vector<uint8_t> &remat;  // multi-use statements that are recomputed at every use instead of kept
    std::vector<uint8_t> computed;      // materialised statements whose value already sits in the slot file
    std::vector<Sym> code;
    int32_t n_values;  // statements + temporaries
    int32_t max_depth = 0;
    bool pending_push = false;
    // eager: before the fold of a unit starts, every materialised value its (inlined) sub-tree reads is computed as a
    // unit of its own (statement order of the emitted function, restricted to what the unit needs).  Folds then only
    // read available slots, so sums of products become TERM blocks and nesting stays shallow; the price is longer
    // live ranges.  lazy: materialised values are computed in the middle of the fold that uses them first.
    bool eager = true;
    std::vector<int32_t> mark;  // scratch for collect()
    int32_t mark_gen = 0;

    CodeGen(const std::vector<Stmt> &s, const std::vector<Operand> &o, const std::vector<uint8_t> &r)
        : st(s), ops(o), remat(r), computed(s.size(), 0), n_values((int32_t)s.size()), mark(s.size(), 0) {}

    // materialised, not yet computed values read by the fold of `v` (through inlined children), in first-use order
    void collect(int32_t v, std::vector<int32_t> &deps) {
        ++mark_gen;
        std::vector<std::pair<int32_t, int32_t>> stk;  // (statement, next operand)
        stk.push_back({v, 0});
        while (!stk.empty()) {
            auto &top = stk.back();
            const Stmt &s = st[(size_t)top.first];
            if (top.second == s.count) {
                stk.pop_back();
                continue;
            }
            const int32_t c = ops[(size_t)(s.first + top.second++)].val;
            if (is_leaf(c) || mark[(size_t)c] == mark_gen) continue;
            mark[(size_t)c] = mark_gen;
            if (materialised(c)) {
                if (!computed[(size_t)c]) deps.push_back(c);
            } else {
                stk.push_back({c, 0});
            }
        }
    }

    // computes `unit` (and, in eager mode, first everything materialised it depends on) -- iterative
    void gen_unit(int32_t unit) {
        if (!eager) {
            gen_fold(unit);
            return;
        }
        std::vector<std::pair<int32_t, bool>> work;
        work.push_back({unit, false});
        std::vector<int32_t> deps;
        while (!work.empty()) {
            const int32_t v = work.back().first;
            if (computed[(size_t)v]) {
                work.pop_back();
                continue;
            }
            if (!work.back().second) {
                work.back().second = true;
                deps.clear();
                collect(v, deps);
                for (auto it = deps.rbegin(); it != deps.rend(); ++it) work.push_back({*it, false});
            } else {
                work.pop_back();
                gen_fold(v);
            }
        }
    }

    bool is_leaf(int32_t v) const { return st[(size_t)v].op < 0; }
    // kept in the slot file once computed: roots and multi-use nodes (unless rematerialised)
    bool materialised(int32_t v) const {
        const Stmt &s = st[(size_t)v];
        return s.op >= 0 && (s.root >= 0 || (s.uses >= 2 && !remat[(size_t)v]));
    }
    // readable from the slot file right now
    bool available(int32_t v) const { return is_leaf(v) || (materialised(v) && computed[(size_t)v]); }

    Sym &emit(uint8_t op, int32_t val = -1, int32_t arg = 0, double f = 1.0) {
        Sym s{};
        s.op = op;
        s.k = val >= 0 ? 1 : 0;
        s.vals[0] = val;
        s.arg = arg;
        s.f = f;
        if (op == FDG_OP_MOV) {  // a MOV always starts a fold: it is the packet that pushes
            s.push = pending_push;
            pending_push = false;
        }
        code.push_back(s);
        return code.back();
    }
    void emit_term(bool first, const int32_t *vals, int k, double f) {
        Sym s{};
        s.op = FDG_OP_TERM;
        s.k = (uint8_t)k;
        s.first = first;
        for (int i = 0; i < k; ++i) s.vals[i] = vals[i];
        s.f = f;
        if (first) {
            s.push = pending_push;
            pending_push = false;
        }
        code.push_back(s);
    }

    // a Sum operand that is an unmaterialised product of slot values with unit factors: one TERM packet
    bool termable(int32_t v) const {
        const Stmt &c = st[(size_t)v];
        if (c.op != FDG_OP_PROD || materialised(v) || c.count > FDG_TERM_MAX) return false;
        for (int32_t i = 0; i < c.count; ++i) {
            const Operand &o = ops[(size_t)(c.first + i)];
            if (o.f != 1.0 || !available(o.val)) return false;
        }
        return true;
    }

    struct Frame {
        int32_t v;
        int32_t i;
        int8_t mode;  // pending combine once the child returns: 0 none, 1 first operand, 2 stack, 3 parked in a slot
        double f;
        int32_t tmp;
    };

    // computes statement `unit` into A with an empty stack; materialised nodes met on the way (lazy mode) are computed
    // at their first use, stored, and read from the slot file afterwards
    void gen_fold(int32_t unit) {
        std::vector<Frame> stack;
        int depth = 0;  // stack registers holding partial folds of enclosing nodes
        stack.push_back({unit, 0, 0, 1.0, -1});
        while (!stack.empty()) {
            Frame &fr = stack.back();
            const Stmt &s = st[(size_t)fr.v];
            max_depth = std::max<int32_t>(max_depth, (int32_t)stack.size());
            const bool sum = s.op == FDG_OP_SUM;
            if (fr.mode != 0) {
                // a child computed inline has just returned in A: fold it into this node
                if (fr.mode == 1) {
                    if (fr.f != 1.0 && s.op != FDG_OP_POWER) emit(FDG_OP_SCALE, -1, 0, fr.f);
                } else if (fr.mode == 2) {
                    emit(sum ? FDG_OP_RADDF : FDG_OP_RMULF, -1, 0, fr.f);
                    --depth;
                } else {
                    emit(sum ? FDG_OP_XADDF : FDG_OP_XMULF, fr.tmp, 0, fr.f);
                }
                fr.mode = 0;
                fr.i++;
                continue;
            }
            if (fr.i == s.count) {
                if (s.op == FDG_OP_POWER) {
                    emit(FDG_OP_POW, -1, s.pow_n, 1.0);
                    const double f = ops[(size_t)s.first].f;
                    if (f != 1.0) emit(FDG_OP_SCALE, -1, 0, f);
                }
                if (materialised(fr.v)) {
                    if (s.root >= 0) emit(FDG_OP_ROOT, -1, s.root);
                    if (s.uses >= 1) emit(FDG_OP_ST, fr.v);
                    computed[(size_t)fr.v] = 1;
                }
                stack.pop_back();
                continue;
            }
            const Operand &o = ops[(size_t)(s.first + fr.i)];
            const bool first = fr.i == 0;
            if (available(o.val)) {
                // operand comes from the slot file
                if (s.op == FDG_OP_POWER) {
                    emit(FDG_OP_MOV, o.val);
                } else if (first) {
                    if (o.f == 1.0)
                        emit(FDG_OP_MOV, o.val);
                    else
                        emit_term(true, &o.val, 1, o.f);  // A = v * f
                } else if (sum) {
                    if (o.f == 1.0)
                        emit(FDG_OP_ADD, o.val);
                    else
                        emit_term(false, &o.val, 1, o.f);  // A = A + v * f
                } else {
                    if (o.f == 1.0)
                        emit(FDG_OP_MUL, o.val);
                    else
                        emit(FDG_OP_MULF, o.val, 0, o.f);
                }
                fr.i++;
                continue;
            }
            if ((sum || first) && s.op != FDG_OP_POWER && termable(o.val)) {
                // A (+)= (v1 * v2 * ... * vk) * f as one term record, no stack traffic.  As the first operand of a
                // Prod this is A = (v1 * ... * vk) * f, the same left fold the MOV/MUL/SCALE sequence would compute.
                const Stmt &c = st[(size_t)o.val];
                int32_t vals[FDG_TERM_MAX];
                for (int32_t q = 0; q < c.count; ++q) vals[q] = ops[(size_t)(c.first + q)].val;
                emit_term(first, vals, c.count, o.f);
                fr.i++;
                continue;
            }
            // an inner node that is not in the slot file: evaluate it right here
            fr.f = o.f;
            const int32_t child = o.val;
            if (first) {
                fr.mode = 1;  // A is free: the child's fold simply starts in A
            } else if (depth < FDG_STACK_REGS) {
                fr.mode = 2;  // the child's first packet pushes this node's partial fold
                pending_push = true;
                ++depth;
            } else {
                fr.mode = 3;  // stack full: park the partial fold in a slot
                fr.tmp = n_values++;
                emit(FDG_OP_ST, fr.tmp);
            }
            stack.push_back({child, 0, 0, 1.0, -1});  // NB: invalidates fr
        }
    }

    void run() {
        // roots in the order the emitted function assigns them; everything else on demand
        for (int32_t v = 0; v < (int32_t)st.size(); ++v) {
            const Stmt &s = st[(size_t)v];
            if (s.root < 0) continue;
            if (s.op < 0) {  // a leaf that is itself a root: root[r] = leafVal[k]
                emit(FDG_OP_MOV, v);
                emit(FDG_OP_ROOT, -1, s.root);
                continue;
            }
            if (!computed[(size_t)v]) gen_unit(v);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Stage C: slots, prefetch, packets
// ------------------------------------------------------------------------------------------------
struct Op2 {  // post-allocation, non-LDL operation
    uint8_t kind;  // 0 VM op, 1 SPILL, 2 FILL
    Sym s;
    int32_t slots[FDG_TERM_MAX];  // slots read (kind 0) / slots[0] = slot written by ST, SPILL source, FILL target
    int32_t scratch;              // SPILL / FILL
};
struct Ldl {
    int32_t slot, leaf;
    int64_t anchor;    // index into ops2 before which the load originally sits
    int64_t earliest;  // smallest legal anchor (after the last access of the slot's previous content)
};

struct Allocator {
    const std::vector<Stmt> &st;
    const std::vector<Sym> &code;
    int32_t n_values, max_slots;
    std::vector<Op2> ops2;
    std::vector<Ldl> ldls;
    std::vector<std::vector<int64_t>> use_pos;  // per value: positions in `code` that read it
    std::vector<int32_t> use_ptr, val_slot, val_scratch;
    std::vector<int32_t> slot_val;
    std::vector<int64_t> slot_last;  // index in ops2 of the last op touching the slot, -1 if none
    std::vector<uint8_t> slot_pinned;
    std::vector<int32_t> free_slots;
    int32_t n_slots = 0, n_scratch = 0;
    int64_t leaf_loads = 0;

    Allocator(const std::vector<Stmt> &s, const std::vector<Sym> &c, int32_t nv, int32_t ms)
        : st(s), code(c), n_values(nv), max_slots(ms) {}

    int64_t next_use(int32_t v) const {
        const auto &u = use_pos[(size_t)v];
        const int32_t p = use_ptr[(size_t)v];
        return p < (int32_t)u.size() ? u[(size_t)p] : INT64_MAX;
    }
    bool is_leaf(int32_t v) const { return v < (int32_t)st.size() && st[(size_t)v].op < 0; }

    int32_t alloc_slot() {
        if (!free_slots.empty()) {
            // take the free slot that has been idle the longest: most room to hoist a load into it
            size_t best = 0;
            for (size_t i = 1; i < free_slots.size(); ++i)
                if (slot_last[(size_t)free_slots[i]] < slot_last[(size_t)free_slots[best]]) best = i;
            const int32_t s = free_slots[best];
            free_slots[best] = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        if (n_slots < max_slots) {
            slot_val.push_back(-1);
            slot_last.push_back(-1);
            slot_pinned.push_back(0);
            return n_slots++;
        }
        // evict the resident value whose next use is farthest away (Belady); never an operand of this packet
        int32_t victim = -1;
        int64_t far = -1;
        for (int32_t s = 0; s < n_slots; ++s) {
            const int32_t v = slot_val[(size_t)s];
            if (v < 0 || slot_pinned[(size_t)s]) continue;
            const int64_t nu = next_use(v);
            if (nu > far) {
                far = nu;
                victim = s;
            }
        }
        const int32_t v = slot_val[(size_t)victim];
        if (!is_leaf(v) && next_use(v) != INT64_MAX && val_scratch[(size_t)v] < 0) {
            val_scratch[(size_t)v] = n_scratch++;
            Op2 o{};
            o.kind = 1;
            o.slots[0] = victim;
            o.scratch = val_scratch[(size_t)v];
            slot_last[(size_t)victim] = (int64_t)ops2.size();
            ops2.push_back(o);
        }
        val_slot[(size_t)v] = -1;
        slot_val[(size_t)victim] = -1;
        return victim;
    }

    void release_if_dead(int32_t v) {
        if (next_use(v) != INT64_MAX) return;
        const int32_t s = val_slot[(size_t)v];
        if (s < 0) return;
        val_slot[(size_t)v] = -1;
        slot_val[(size_t)s] = -1;
        free_slots.push_back(s);
    }

    void run() {
        use_pos.assign((size_t)n_values, {});
        for (size_t p = 0; p < code.size(); ++p) {
            const int k = reads_count(code[p]);
            for (int q = 0; q < k; ++q) {
                auto &u = use_pos[(size_t)code[p].vals[q]];
                if (u.empty() || u.back() != (int64_t)p) u.push_back((int64_t)p);
            }
        }
        use_ptr.assign((size_t)n_values, 0);
        val_slot.assign((size_t)n_values, -1);
        val_scratch.assign((size_t)n_values, -1);
        for (size_t p = 0; p < code.size(); ++p) {
            const Sym &s = code[p];
            Op2 o{};
            o.kind = 0;
            o.s = s;
            const int k = reads_count(s);
            if (k > 0) {
                // bring every operand in, pinning the ones already placed for this packet
                for (int q = 0; q < k; ++q) {
                    const int32_t v = s.vals[q];
                    if (val_slot[(size_t)v] >= 0) slot_pinned[(size_t)val_slot[(size_t)v]] = 1;
                }
                for (int q = 0; q < k; ++q) {
                    const int32_t v = s.vals[q];
                    if (val_slot[(size_t)v] < 0) {
                        const int32_t sl = alloc_slot();
                        if (is_leaf(v)) {
                            ldls.push_back({sl, st[(size_t)v].leaf, (int64_t)ops2.size(), slot_last[(size_t)sl] + 1});
                            leaf_loads++;
                        } else {
                            Op2 f{};
                            f.kind = 2;
                            f.slots[0] = sl;
                            f.scratch = val_scratch[(size_t)v];
                            slot_last[(size_t)sl] = (int64_t)ops2.size();
                            ops2.push_back(f);
                        }
                        val_slot[(size_t)v] = sl;
                        slot_val[(size_t)sl] = v;
                        slot_pinned[(size_t)sl] = 1;
                    }
                    o.slots[q] = val_slot[(size_t)v];
                }
                for (int q = 0; q < k; ++q) {
                    const int32_t v = s.vals[q];
                    slot_pinned[(size_t)o.slots[q]] = 0;
                    slot_last[(size_t)o.slots[q]] = (int64_t)ops2.size();
                    if (next_use(v) == (int64_t)p) use_ptr[(size_t)v]++;
                }
                ops2.push_back(o);
                for (int q = 0; q < k; ++q) release_if_dead(s.vals[q]);
            } else if (s.op == FDG_OP_ST) {
                const int32_t v = s.vals[0];
                if (use_pos[(size_t)v].empty()) continue;  // never read: nothing to keep
                const int32_t sl = alloc_slot();
                val_slot[(size_t)v] = sl;
                slot_val[(size_t)sl] = v;
                o.slots[0] = sl;
                slot_last[(size_t)sl] = (int64_t)ops2.size();
                ops2.push_back(o);
            } else {
                ops2.push_back(o);
            }
        }
    }
};

struct Packer {
    const Allocator &al;
    std::vector<uint32_t> &words;
    int32_t dist;
    int64_t committed = 0, completed = 0;  // cp.async groups
    std::vector<int64_t> slot_group;       // group that fills the slot, -1 when filled synchronously
    uint32_t wait_field = 0;               // header wait field of the packet being assembled

    Packer(const Allocator &a, std::vector<uint32_t> &w, int32_t d) : al(a), words(w), dist(d) {}

    void packet(uint32_t w0, uint32_t w1 = 0, uint32_t w2 = 0, uint32_t w3 = 0) {
        words.push_back(w0);
        words.push_back(w1);
        words.push_back(w2);
        words.push_back(w3);
    }
    int64_t n_packets() const { return (int64_t)words.size() / 4; }
    // the packet about to be emitted reads `slot`: make sure its cp.async group has landed
    void need(int32_t slot) {
        const int64_t g = slot_group[(size_t)slot];
        if (g < 0 || g < completed) return;
        int64_t pend = committed - (g + 1);  // newer groups may stay in flight
        if (pend > FDG_MAX_WAIT) pend = FDG_MAX_WAIT;
        if (wait_field == 0 || (uint32_t)(pend + 1) < wait_field) wait_field = (uint32_t)(pend + 1);
        completed = std::max(completed, committed - pend);
    }
    uint32_t take_wait() {
        const uint32_t w = wait_field;
        wait_field = 0;
        return w;
    }
    static void split(double f, uint32_t &lo, uint32_t &hi) {
        uint64_t u;
        std::memcpy(&u, &f, 8);
        lo = (uint32_t)u;
        hi = (uint32_t)(u >> 32);
    }

    void pad_to_chunk() {
        while (n_packets() % FDG_CHUNK) packet(FDG_HDR(FDG_OP_NOP, 0, 0, 0, 0, 0));
    }
    int64_t room() const { return FDG_CHUNK - n_packets() % FDG_CHUNK; }  // packets left in the current chunk

    void run() {
        const auto &ops2 = al.ops2;
        const int64_t M = (int64_t)ops2.size();
        slot_group.assign((size_t)std::max(al.n_slots, 1), -1);
        // maximal runs of TERM ops that one block could cover: run_start[j] = first op of the run containing j
        std::vector<int64_t> run_start((size_t)M + 1);
        for (int64_t j = 0; j < M; ++j) {
            const Op2 &o = ops2[(size_t)j];
            const bool cont = j > 0 && o.kind == 0 && o.s.op == FDG_OP_TERM && !o.s.first && ops2[(size_t)(j - 1)].kind == 0 &&
                              ops2[(size_t)(j - 1)].s.op == FDG_OP_TERM && ops2[(size_t)(j - 1)].s.k == o.s.k;
            run_start[(size_t)j] = cont ? run_start[(size_t)(j - 1)] : j;
        }
        run_start[(size_t)M] = M;
        // bucket the loads by their hoisted anchor; a load that would land inside a run of terms moves to the start
        // of the run when that is legal (so the run stays one block), anchors are quantised so neighbours share packets
        std::vector<std::vector<int32_t>> at((size_t)M + 1);
        for (size_t i = 0; i < al.ldls.size(); ++i) {
            const Ldl &l = al.ldls[i];
            int64_t t = l.anchor - dist;
            if (dist > 0) t &= ~(int64_t)3;
            if (t < l.earliest) t = l.earliest;
            if (t > l.anchor) t = l.anchor;
            if (run_start[(size_t)t] >= l.earliest) t = run_start[(size_t)t];
            at[(size_t)t].push_back((int32_t)i);
        }
        int64_t j = 0;
        while (j <= M) {
            // loads anchored before op j, up to seven per packet pair, one cp.async group each
            const auto &lst = at[(size_t)j];
            for (size_t k = 0; k < lst.size();) {
                size_t n = std::min<size_t>(FDG_LDL_MAX, lst.size() - k);
                if (n > 3 && room() < 2) n = 3;
                uint32_t w[FDG_LDL_MAX] = {0, 0, 0, 0, 0, 0, 0};
                for (size_t q = 0; q < n; ++q) {
                    const Ldl &l = al.ldls[(size_t)lst[k + q]];
                    w[q] = FDG_LDL_WORD(l.slot, l.leaf);
                    slot_group[(size_t)l.slot] = committed;
                }
                packet(FDG_HDR(FDG_OP_LDL, n, 0, 0, 0, 0), w[0], w[1], w[2]);
                if (n > 3) packet(w[3], w[4], w[5], w[6]);
                committed++;
                k += n;
            }
            if (j == M) break;
            const Op2 &o = ops2[(size_t)j];
            if (o.kind == 1) {  // SPILL
                need(o.slots[0]);
                packet(FDG_HDR(FDG_OP_SPILL, 0, take_wait(), 0, 0, 0), (uint32_t)o.slots[0], (uint32_t)o.scratch);
                ++j;
                continue;
            }
            if (o.kind == 2) {  // FILL
                slot_group[(size_t)o.slots[0]] = -1;
                packet(FDG_HDR(FDG_OP_FILL, 0, 0, 0, 0, 0), (uint32_t)o.slots[0], (uint32_t)o.scratch);
                ++j;
                continue;
            }
            const Sym &s = o.s;
            uint32_t lo = 0, hi = 0;
            split(s.f, lo, hi);
            switch (s.op) {
                case FDG_OP_MOV:
                case FDG_OP_MUL:
                case FDG_OP_ADD: {
                    // pack a run of the same fold: MOV a [MUL b [MUL c]] / MUL a b c / ADD a b c
                    const uint8_t follow = s.op == FDG_OP_ADD ? FDG_OP_ADD : FDG_OP_MUL;
                    uint32_t w[3] = {(uint32_t)o.slots[0], 0, 0};
                    int n = 1;
                    while (n < 3 && j + n < M && at[(size_t)(j + n)].empty()) {
                        const Op2 &o2 = ops2[(size_t)(j + n)];
                        if (o2.kind != 0 || o2.s.op != follow) break;
                        w[n] = (uint32_t)o2.slots[0];
                        ++n;
                    }
                    for (int q = 0; q < n; ++q) need((int32_t)w[q]);
                    packet(FDG_HDR(s.op, n, take_wait(), s.push, 0, 0), w[0], w[1], w[2]);
                    j += n;
                    continue;
                }
                case FDG_OP_TERM: {
                    // one block: this term and the following terms of the same fold with the same operand count,
                    // as far as no load is anchored in between and the records fit the current chunk
                    const int k = s.k;
                    const int rec = k > 4 ? 2 : 1;
                    if (room() < 1 + rec) pad_to_chunk();
                    int64_t n = 1;
                    while (j + n < M && run_start[(size_t)(j + n)] == run_start[(size_t)j] && at[(size_t)(j + n)].empty() &&
                           1 + (n + 1) * rec <= room())
                        ++n;
                    for (int64_t t = 0; t < n; ++t)
                        for (int q = 0; q < k; ++q) need(ops2[(size_t)(j + t)].slots[q]);
                    packet(FDG_HDR(FDG_OP_TERM, k, take_wait(), s.push, s.first, 0), (uint32_t)n);
                    for (int64_t t = 0; t < n; ++t) {
                        const Op2 &ot = ops2[(size_t)(j + t)];
                        uint32_t e[6] = {0, 0, 0, 0, 0, 0};
                        for (int q = 0; q < k; ++q) e[q >> 1] |= (uint32_t)ot.slots[q] << (16 * (q & 1));
                        uint32_t flo, fhi;
                        split(ot.s.f, flo, fhi);
                        packet(e[0], e[1], flo, fhi);
                        if (rec == 2) packet(e[2], e[3], e[4], e[5]);
                    }
                    j += n;
                    continue;
                }
                case FDG_OP_MULF:
                case FDG_OP_XADDF:
                case FDG_OP_XMULF:
                    need(o.slots[0]);
                    packet(FDG_HDR(s.op, 1, take_wait(), 0, 0, 0), (uint32_t)o.slots[0], lo, hi);
                    break;
                case FDG_OP_SCALE:
                case FDG_OP_RADDF:
                case FDG_OP_RMULF:
                    packet(FDG_HDR(s.op, 0, 0, 0, 0, 0), 0, lo, hi);
                    break;
                case FDG_OP_POW:
                case FDG_OP_ROOT:
                    packet(FDG_HDR(s.op, 0, 0, 0, 0, 0), (uint32_t)s.arg);
                    break;
                case FDG_OP_ST:
                    slot_group[(size_t)o.slots[0]] = -1;
                    packet(FDG_HDR(s.op, 0, 0, 0, 0, 0), (uint32_t)o.slots[0]);
                    break;
                default:
                    break;
            }
            ++j;
        }
        packet(FDG_HDR(FDG_OP_END, 0, 0, 0, 0, 0));
    }
};

}  // namespace

int lower(const fdg_graph_desc &g, const fdg_options &opt, Lowered &out, std::string &err) {
    if (opt.dtype != FDG_F64 && opt.dtype != FDG_C128) {
        err = "unsupported dtype (only FDG_F64 and FDG_C128)";  // static.jl:151 "Unsupported type"
        return FDG_ERR_UNSUPPORTED;
    }
    out = Lowered();
    out.dtype = opt.dtype;
    std::vector<Stmt> &st = out.st;
    std::vector<Operand> &ops = out.ops;
    int rc = build_statements(g, st, ops, out, err);
    if (rc != FDG_OK) return rc;
    mark_live(st, ops);

    // algorithmic operation counts of the reference function (count_operation, tree_properties.jl:165-185,
    // plus one multiply per factor != 1 and N-1 multiplies per Power{N}); independent of how the VM evaluates it
    for (const Stmt &s : st) {
        if (!s.live || s.op < 0) continue;
        out.n_operands += s.count;
        for (int32_t i = 0; i < s.count; ++i)
            if (ops[(size_t)(s.first + i)].f != 1.0) out.muls_vf++;
        if (s.op == FDG_OP_SUM) out.adds_vv += s.count - 1;
        if (s.op == FDG_OP_PROD) out.muls_vv += s.count - 1;
        if (s.op == FDG_OP_POWER) out.pow_muls += s.pow_n - 1;
    }

    if (opt.cse != 0) {
        int64_t min_cost = 2;  // everything that saves at least one operation
        if (const char *e = getenv("FDG_CSE_MIN_COST")) min_cost = atoll(e);
        out.cse_removed = eliminate_common_subexpressions(st, ops, min_cost, out.canon);
        mark_live(st, ops);
    }

    int32_t max_slots = opt.max_slots > 0 ? opt.max_slots : 48;
    if (max_slots < FDG_TERM_MAX + 1) max_slots = FDG_TERM_MAX + 1;  // every operand of a TERM must be resident
    if (max_slots > FDG_MAX_SLOTS) max_slots = FDG_MAX_SLOTS;

    // cost of recomputing a statement from leaves (capped): cheap multi-use values that the allocator had to
    // spill to global scratch are rematerialised instead (recomputing gives the same bits)
    const int32_t kCheap = 16;
    std::vector<int32_t> cost(st.size(), 0);
    for (size_t v = 0; v < st.size(); ++v) {
        const Stmt &s = st[v];
        if (s.op < 0) {
            cost[v] = 1;
            continue;
        }
        int64_t c = 0;
        for (int32_t i = 0; i < s.count; ++i) c += cost[(size_t)ops[(size_t)(s.first + i)].val];
        cost[v] = (int32_t)std::min<int64_t>(c, 1 << 20);
    }
    std::vector<uint8_t> remat(st.size(), 0);
    std::vector<Sym> code;
    int32_t n_values = 0;
    Allocator *al_ptr = nullptr;
    std::unique_ptr<Allocator> al_owner;
    for (int iter = 0; iter < 4; ++iter) {
        CodeGen cg(st, ops, remat);
        cg.eager = opt.schedule != 1;
        cg.run();
        out.max_depth = cg.max_depth;
        code.swap(cg.code);
        n_values = cg.n_values;
        al_owner.reset(new Allocator(st, code, n_values, max_slots));
        al_owner->run();
        al_ptr = al_owner.get();
        if (al_ptr->n_scratch == 0 || iter == 3) break;
        bool changed = false;
        for (size_t v = 0; v < st.size(); ++v) {
            if (al_ptr->val_scratch[v] >= 0 && !remat[v] && st[v].root < 0 && cost[v] <= kCheap) {
                remat[v] = 1;
                changed = true;
            }
        }
        if (!changed) break;
    }
    Allocator &al = *al_ptr;
    if (al.n_scratch >= (1 << 30)) {
        err = "program needs too many scratch values";
        return FDG_ERR_CAPACITY;
    }
    out.n_slots = std::max(al.n_slots, 1);
    out.n_scratch = al.n_scratch;
    out.leaf_loads = al.leaf_loads;

    int32_t dist = opt.prefetch == 0 ? 24 : (opt.prefetch < 0 ? 0 : opt.prefetch);
    Packer pk(al, out.words, dist);
    pk.run();
    return FDG_OK;
}

}  // namespace fdg
