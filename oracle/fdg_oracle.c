/*
 * fdg_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm for the graph-evaluation path, used only as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product (feynmandiagram.jl_b200/, libfdgraph.so) never links, imports or calls this file.
 *
 * What is restated, and from where (paths relative to the reference checkout):
 *   - statement order, leaf numbering, root assignment of the emitted function
 *         src/backend/static.jl:98-133  (to_julia_str), :155-197 (to_Cstr),
 *         src/backend/compiler_python.jl:9-52 (to_python_str)
 *     post-order DFS over `graphs` with children in stored order (AbstractTrees.PostOrderDFS with
 *     children(g) = subgraphs(g), src/computational_graph/tree_properties.jl:20-22), first visit of
 *     an id wins, leaf k = k-th distinct leaf id, `root[r] = g` right after the node's statement.
 *   - the arithmetic of one statement, `to_static`                      src/backend/static.jl:13-46
 *         Sum   -> (g1 * f1 + g2 + g3 * f3)   left fold, `* f` omitted iff f == 1
 *         Prod  -> (g1 * f1 * g2 * g3 * f3)   left fold, `* f` omitted iff f == 1
 *         Power -> ((g)^N * f)                 literal powers 2, 3 are x*x, x*x*x (Base.literal_pow)
 *   - the interpreter `eval!` / `apply`                 src/computational_graph/eval.jl:1-3, :15-39
 *         Sum: sum(w_i * f_i), Prod: prod(w_i * f_i)  (always multiplies by f_i, different rounding
 *         from the emitter for Prod), Power: w^N * f
 *   Julia never contracts a*b+c into an FMA: this file must be compiled with -ffp-contract=off.
 *
 * Parity status: pinned against the reference's own known-answer tests: tests/test_oracle_kat.py holds
 * test/compiler.jl:4-15 (4.5), test/computational_graph.jl:874-887 (26, 27, 702), the derivative known answers
 * of test/computational_graph.jl:930-1071 (through the restated graph-level AD, oracle/frontend/ad.py, and again
 * through the restated Taylor expansion) and test/taylor.jl:42-56, :96-112; tests/test_frontends.py holds the
 * diagram counts and filters of test/front_end.jl:186-219, :398-443, :446-598, :600-825, :221-310.  Power{N>=4}, ComplexF64 and random-leaf values through the compiled path have no pinned
 * numbers in the reference ("parity unpinned" for those sub-cases; covered by emitter-vs-interpreter-
 * vs-exact-rational self-consistency).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t op; /* -1 leaf, 1 Sum, 2 Prod, 3 Power */
    int32_t pow_n;
    int64_t first;
    int32_t count;
    int32_t root; /* root position assigned after this statement, -1 none */
    int32_t leaf; /* leaf index */
} ostmt;

typedef struct {
    int64_t n_stmt, n_operand, L, R;
    ostmt *stmt;
    int32_t *opnd_val;
    double *opnd_f;
    int32_t *leaf_node;
    int32_t last_root;
} oprog;

/* tiny open-addressing map int64 -> int32 */
typedef struct {
    int64_t *k;
    int32_t *v;
    uint8_t *used;
    int64_t cap;
} omap;
static void omap_init(omap *m, int64_t n) {
    int64_t c = 16;
    while (c < 2 * n + 2) c <<= 1;
    m->cap = c;
    m->k = (int64_t *)calloc((size_t)c, sizeof(int64_t));
    m->v = (int32_t *)calloc((size_t)c, sizeof(int32_t));
    m->used = (uint8_t *)calloc((size_t)c, 1);
}
static void omap_free(omap *m) {
    free(m->k);
    free(m->v);
    free(m->used);
}
static int64_t omap_slot(const omap *m, int64_t key) {
    uint64_t h = (uint64_t)key * 0x9E3779B97F4A7C15ull;
    int64_t i = (int64_t)(h >> 20) & (m->cap - 1);
    while (m->used[i] && m->k[i] != key) i = (i + 1) & (m->cap - 1);
    return i;
}
static int omap_get(const omap *m, int64_t key, int32_t *out) {
    int64_t i = omap_slot(m, key);
    if (!m->used[i]) return 0;
    *out = m->v[i];
    return 1;
}
static void omap_put_first(omap *m, int64_t key, int32_t val) { /* keeps the first value of a key */
    int64_t i = omap_slot(m, key);
    if (m->used[i]) return;
    m->used[i] = 1;
    m->k[i] = key;
    m->v[i] = val;
}

void oracle_free(oprog *p) {
    if (!p) return;
    free(p->stmt);
    free(p->opnd_val);
    free(p->opnd_f);
    free(p->leaf_node);
    free(p);
}

/* Restates to_julia_str's traversal (static.jl:98-133).  Returns NULL on malformed input. */
oprog *oracle_lower(int64_t n_nodes, const int64_t *node_id, const int32_t *node_op, const int32_t *node_pow,
                    const int64_t *child_ptr, const int32_t *child_node, const double *child_factor,
                    int64_t n_graphs, const int32_t *graphs, int64_t n_roots, const int64_t *root_id) {
    oprog *p = (oprog *)calloc(1, sizeof(oprog));
    int64_t n_edges = n_nodes ? child_ptr[n_nodes] : 0;
    p->stmt = (ostmt *)calloc((size_t)n_nodes + 1, sizeof(ostmt));
    p->opnd_val = (int32_t *)calloc((size_t)n_edges + 1, sizeof(int32_t));
    p->opnd_f = (double *)calloc((size_t)n_edges + 1, sizeof(double));
    p->leaf_node = (int32_t *)calloc((size_t)n_nodes + 1, sizeof(int32_t));
    p->R = n_roots;
    p->last_root = -1;
    omap rootpos, visited;
    omap_init(&rootpos, n_roots);
    omap_init(&visited, n_nodes);
    for (int64_t r = 0; r < n_roots; ++r) omap_put_first(&rootpos, root_id[r], (int32_t)r); /* findfirst */
    /* explicit-stack post-order DFS; `done` prunes re-walks of an object (a re-walk emits nothing) */
    uint8_t *done = (uint8_t *)calloc((size_t)n_nodes + 1, 1);
    int32_t *stk_node = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n_nodes + 1));
    int64_t *stk_edge = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_nodes + 1));
    int bad = 0;
    for (int64_t gi = 0; gi < n_graphs && !bad; ++gi) {
        int64_t sp = 0;
        if (done[graphs[gi]]) continue;
        stk_node[0] = graphs[gi];
        stk_edge[0] = child_ptr[graphs[gi]];
        done[graphs[gi]] = 1;
        sp = 1;
        while (sp > 0 && !bad) {
            int32_t u = stk_node[sp - 1];
            if (stk_edge[sp - 1] < child_ptr[u + 1]) {
                int32_t c = child_node[stk_edge[sp - 1]++];
                if (c < 0 || c >= n_nodes) {
                    bad = 1;
                    break;
                }
                if (!done[c]) {
                    if (sp >= n_nodes) {
                        bad = 1;
                        break;
                    }
                    done[c] = 1;
                    stk_node[sp] = c;
                    stk_edge[sp] = child_ptr[c];
                    ++sp;
                }
                continue;
            }
            --sp;
            int32_t dummy;
            if (omap_get(&visited, node_id[u], &dummy)) continue; /* `g_id in inds_visited && continue` */
            ostmt *s = &p->stmt[p->n_stmt];
            int64_t a = child_ptr[u], b = child_ptr[u + 1];
            s->root = -1;
            if (a == b) { /* isempty(subgraphs(g)) */
                s->op = -1;
                s->leaf = (int32_t)p->L;
                p->leaf_node[p->L++] = u;
            } else {
                s->op = node_op[u];
                s->pow_n = node_pow[u];
                s->first = p->n_operand;
                s->count = (int32_t)(b - a);
                if (s->op < 1 || s->op > 3 || (s->op == 3 && (s->count != 1 || s->pow_n < 2))) bad = 1;
                for (int64_t e = a; e < b && !bad; ++e) {
                    int32_t v;
                    if (!omap_get(&visited, node_id[child_node[e]], &v)) {
                        bad = 1; /* a cycle: the child is still on the stack */
                        break;
                    }
                    p->opnd_val[p->n_operand] = v;
                    p->opnd_f[p->n_operand] = child_factor[e];
                    p->n_operand++;
                }
            }
            int32_t r;
            if (omap_get(&rootpos, node_id[u], &r)) {
                s->root = r;
                p->last_root = r;
            }
            omap_put_first(&visited, node_id[u], (int32_t)p->n_stmt);
            p->n_stmt++;
        }
    }
    free(done);
    free(stk_node);
    free(stk_edge);
    omap_free(&rootpos);
    omap_free(&visited);
    if (bad) {
        oracle_free(p);
        return NULL;
    }
    return p;
}

int64_t oracle_num_leaves(const oprog *p) { return p->L; }
int64_t oracle_num_stmts(const oprog *p) { return p->n_stmt; }
int32_t oracle_last_root(const oprog *p) { return p->last_root; }
void oracle_leafmap(const oprog *p, int32_t *out) { memcpy(out, p->leaf_node, sizeof(int32_t) * (size_t)p->L); }
/* statement table export (for the C-text emitter in oracle/emit_c.py) */
void oracle_stmt(const oprog *p, int64_t i, int32_t *op, int32_t *pow_n, int64_t *first, int32_t *count, int32_t *root,
                 int32_t *leaf) {
    const ostmt *s = &p->stmt[i];
    *op = s->op, *pow_n = s->pow_n, *first = s->first, *count = s->count, *root = s->root, *leaf = s->leaf;
}
void oracle_operand(const oprog *p, int64_t e, int32_t *val, double *f) { *val = p->opnd_val[e], *f = p->opnd_f[e]; }

/* ---- Float64 powers ------------------------------------------------------------------------------ */
/* Julia >= 1.9 base/math.jl pow_body (x::Float64, n::Integer), n >= 0 branch. muladd == fma on FMA hardware. */
static double pow_body(double x, int64_t n) {
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
    if (n == 3) return x * x * x;
    while (n > 1) {
        if (n & 1) {
            double err = fma(y, xnlo, x * ynlo);
            double pr = x * y;
            ynlo = fma(x, y, -pr) + err;
            y = pr;
        }
        double err = x * 2 * xnlo;
        double pr = x * x;
        xnlo = fma(x, x, -pr) + err;
        x = pr;
        n >>= 1;
    }
    double err = fma(y, xnlo, x * ynlo);
    return (isfinite(x) && isfinite(err)) ? fma(x, y, err) : x * y;
}
static double pow_literal(double x, int32_t n) { /* Base.literal_pow: x^2 -> x*x, x^3 -> x*x*x */
    if (n == 2) return x * x;
    if (n == 3) return x * x * x;
    return pow_body(x, n);
}

typedef struct {
    double re, im;
} cplx;
static inline cplx cmul(cplx a, cplx b) { /* Julia *(z::Complex, w::Complex), base/complex.jl */
    cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
    return r;
}
static inline cplx cscale(cplx a, double f) { /* *(z::Complex, x::Real) */
    cplx r = {a.re * f, a.im * f};
    return r;
}
static inline cplx cadd(cplx a, cplx b) {
    cplx r = {a.re + b.re, a.im + b.im};
    return r;
}
static cplx cpow_int(cplx x, int32_t p) { /* literal_pow for 2, 3; Base.power_by_squaring otherwise */
    if (p == 2) return cmul(x, x);
    if (p == 3) return cmul(cmul(x, x), x);
    int t = __builtin_ctz((unsigned)p) + 1;
    p >>= t;
    while (--t > 0) x = cmul(x, x);
    cplx y = x;
    while (p > 0) {
        t = __builtin_ctz((unsigned)p) + 1;
        p >>= t;
        while (--t >= 0) x = cmul(x, x);
        y = cmul(y, x);
    }
    return y;
}

/* ---- one sample ---------------------------------------------------------------------------------------
 * mode 0: emitter semantics (static.jl:13-46); mode 1: interpreter semantics (eval.jl:1-3).
 * leaf[k * ls] is leaf k, root[r * rs] is root r (strides in elements). `g` is scratch of n_stmt values. */
static void eval_one_f64(const oprog *p, int mode, const double *leaf, int64_t ls, double *root, int64_t rs, double *g) {
    for (int64_t i = 0; i < p->n_stmt; ++i) {
        const ostmt *s = &p->stmt[i];
        double v;
        if (s->op < 0) {
            v = leaf[(int64_t)s->leaf * ls];
        } else {
            const int32_t *cv = p->opnd_val + s->first;
            const double *cf = p->opnd_f + s->first;
            if (s->op == 3) {
                if (mode == 0) {
                    v = pow_literal(g[cv[0]], s->pow_n);
                    if (cf[0] != 1.0) v = v * cf[0];
                } else {
                    v = pow_body(g[cv[0]], s->pow_n) * cf[0];
                }
            } else if (mode == 0) {
                v = g[cv[0]];
                if (cf[0] != 1.0) v = v * cf[0];
                if (s->op == 1) {
                    for (int32_t k = 1; k < s->count; ++k) {
                        double t = g[cv[k]];
                        if (cf[k] != 1.0) t = t * cf[k];
                        v = v + t;
                    }
                } else {
                    for (int32_t k = 1; k < s->count; ++k) {
                        v = v * g[cv[k]];
                        if (cf[k] != 1.0) v = v * cf[k];
                    }
                }
            } else {
                v = g[cv[0]] * cf[0];
                if (s->op == 1)
                    for (int32_t k = 1; k < s->count; ++k) v = v + g[cv[k]] * cf[k];
                else
                    for (int32_t k = 1; k < s->count; ++k) v = v * (g[cv[k]] * cf[k]);
            }
        }
        g[i] = v;
        if (s->root >= 0) root[(int64_t)s->root * rs] = v;
    }
}

static void eval_one_c128(const oprog *p, int mode, const cplx *leaf, int64_t ls, cplx *root, int64_t rs, cplx *g) {
    for (int64_t i = 0; i < p->n_stmt; ++i) {
        const ostmt *s = &p->stmt[i];
        cplx v;
        if (s->op < 0) {
            v = leaf[(int64_t)s->leaf * ls];
        } else {
            const int32_t *cv = p->opnd_val + s->first;
            const double *cf = p->opnd_f + s->first;
            if (s->op == 3) {
                v = cpow_int(g[cv[0]], s->pow_n);
                if (mode == 1 || cf[0] != 1.0) v = cscale(v, cf[0]);
            } else if (mode == 0) {
                v = g[cv[0]];
                if (cf[0] != 1.0) v = cscale(v, cf[0]);
                if (s->op == 1) {
                    for (int32_t k = 1; k < s->count; ++k) {
                        cplx t = g[cv[k]];
                        if (cf[k] != 1.0) t = cscale(t, cf[k]);
                        v = cadd(v, t);
                    }
                } else {
                    for (int32_t k = 1; k < s->count; ++k) {
                        v = cmul(v, g[cv[k]]);
                        if (cf[k] != 1.0) v = cscale(v, cf[k]);
                    }
                }
            } else {
                v = cscale(g[cv[0]], cf[0]);
                if (s->op == 1)
                    for (int32_t k = 1; k < s->count; ++k) v = cadd(v, cscale(g[cv[k]], cf[k]));
                else
                    for (int32_t k = 1; k < s->count; ++k) v = cmul(v, cscale(g[cv[k]], cf[k]));
            }
        }
        g[i] = v;
        if (s->root >= 0) root[(int64_t)s->root * rs] = v;
    }
}

/* ---- a batch of samples ----------------------------------------------------------------------------------
 * layout 0: batch-major   leaf[k*ld_leaf + b], root[r*ld_root + b]   (the GPU library's layout)
 * layout 1: sample-major  leaf[b*ld_leaf + k], root[b*ld_root + r]   (one contiguous leafVal per call, the
 *           reference's natural one-sample-at-a-time layout; used for the timed CPU baseline)
 * dtype 0 = Float64, 1 = ComplexF64 (interleaved).  nthreads <= 0: all OpenMP threads.
 * Returns the number of threads used. */
int oracle_eval_batch(const oprog *p, int dtype, int mode, int layout, const void *leaf, int64_t ld_leaf, void *root,
                      int64_t ld_root, int64_t batch, int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#else
    (void)nthreads;
#endif
    const size_t es = dtype ? sizeof(cplx) : sizeof(double);
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        void *g = malloc(es * (size_t)(p->n_stmt + 1));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t b = 0; b < batch; ++b) {
            if (dtype == 0) {
                const double *lf = (const double *)leaf;
                double *rt = (double *)root;
                if (layout == 0)
                    eval_one_f64(p, mode, lf + b, ld_leaf, rt + b, ld_root, (double *)g);
                else
                    eval_one_f64(p, mode, lf + b * ld_leaf, 1, rt + b * ld_root, 1, (double *)g);
            } else {
                const cplx *lf = (const cplx *)leaf;
                cplx *rt = (cplx *)root;
                if (layout == 0)
                    eval_one_c128(p, mode, lf + b, ld_leaf, rt + b, ld_root, (cplx *)g);
                else
                    eval_one_c128(p, mode, lf + b * ld_leaf, 1, rt + b * ld_root, 1, (cplx *)g);
            }
        }
        free(g);
    }
    return used;
}

/* Driver for a function emitted by the restated to_Cstr (oracle/emit_c.py) and compiled on its own:
 * calls fn(root_b, leafVal_b) once per sample, sample-major layout, OpenMP over samples. */
typedef void (*emitted_fn)(double *root, double *leafVal);
int oracle_run_emitted(emitted_fn fn, double *leaf, int64_t ld_leaf, double *root, int64_t ld_root, int64_t batch,
                       int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int64_t b = 0; b < batch; ++b) fn(root + b * ld_root, leaf + b * ld_leaf);
    return used;
}

/* same for `complex double` (ComplexF64) functions: ld_* count complex elements */
typedef void (*emitted_fn_c128)(cplx *root, cplx *leafVal);
int oracle_run_emitted_c128(emitted_fn_c128 fn, cplx *leaf, int64_t ld_leaf, cplx *root, int64_t ld_root, int64_t batch,
                            int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int64_t b = 0; b < batch; ++b) fn(root + b * ld_root, leaf + b * ld_leaf);
    return used;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
