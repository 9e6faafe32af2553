/*
 * fdgraph.h -- C ABI of libfdgraph.so, the B200 (sm_100a) back end for FeynmanDiagram.jl's
 * computational-graph evaluator.
 *
 * This is the drop-in boundary for the ONE hot path this project accelerates:
 *
 *   reference                                                    replaced by
 *   ---------------------------------------------------------   ---------------------------------
 *   Compilers.compile(graphs; root)   src/backend/static.jl:221-227      fdg_compile + fdg_leafmap
 *   to_julia_str / to_Cstr / to_python_str (emit + leaf numbering)
 *        static.jl:98-133, :155-197, compiler_python.jl:9-52             fdg_compile (lowering)
 *   generated eval_graph!(root, leafVal)   static.jl:100,117,123,127,131 fdg_eval (batch of samples)
 *   batched torch form  eval_graph(leafVal[B,L]) -> root[B,R]
 *        compiler_python.jl:23,28,43-49                                  fdg_eval / fdg_eval_host
 *   (no counterpart: MC accumulation over samples, multi-GPU)            fdg_eval_accumulate,
 *                                                                        fdg_comm_* / fdg_allreduce
 *
 * Conventions
 *   - plain C types only; every function returns an int status (FDG_OK == 0) and never throws.
 *     fdg_last_error() returns a thread-local message for the last non-zero status.
 *   - the caller owns every data buffer and the CUDA stream; a handle owns the lowered program
 *     and its device copy (one per device, uploaded lazily on first evaluation).
 *   - a handle may be used from several threads: device-pointer calls are serialised while they are ISSUED (work buffers
 *     are kept per stream, so calls on different streams overlap on the device); the host-buffer calls run one at a time.
 *   - node ids: the reference emitter keeps separate visited lists for leaves and inner nodes (static.jl:116,122), so an id
 *     carried by a leaf object AND by an inner object would be emitted twice there.  Such a graph is ambiguous and is
 *     rejected here with FDG_ERR_BAD_GRAPH.
 *   - data layout is batch-major ("one column per leaf / root", the layout of a Julia B x L
 *     matrix and of the reference's torch emitter): leaf value l of sample b lives at
 *     leaf[l * ld_leaf + b], root r of sample b at root[r * ld_root + b].  Element type is
 *     double (FDG_F64) or interleaved (re, im) doubles (FDG_C128, Julia ComplexF64).
 *   - leaf numbering is the reference's: leaf k (0-based here, k+1 in Julia) is the k-th distinct
 *     leaf (by Graph id) met by the post-order DFS over `graphs` (static.jl:106-120).
 *   - arithmetic is the reference's: n-ary Sum / Prod are left folds in stored subgraph order,
 *     a `* factor` exists only where the factor != 1, no FMA contraction.  Results are
 *     bit-identical to the emitted Julia / C function for Sum, Prod and Power{2,3}.
 */
#ifndef FDGRAPH_H
#define FDGRAPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDG_ABI_VERSION 1

/* status codes */
enum {
    FDG_OK = 0,
    FDG_ERR_BAD_ARG = 1,      /* null pointer, negative size, misaligned buffer ...          */
    FDG_ERR_BAD_GRAPH = 2,    /* cycle, child index out of range, unknown operator, Power N<2,
                                 Power with != 1 subgraph (static.jl:6-11 `error(...)`)      */
    FDG_ERR_UNSUPPORTED = 3,  /* dtype / option not supported (static.jl:151)                 */
    FDG_ERR_CUDA = 4,         /* CUDA runtime error (message has cudaGetErrorString)          */
    FDG_ERR_NCCL = 5,         /* NCCL missing or failing                                      */
    FDG_ERR_NO_DEVICE = 6,    /* no CUDA device: there is NO CPU fallback                     */
    FDG_ERR_CAPACITY = 7      /* program does not fit the on-chip slot file even with spills  */
};

/* node operators: abstractgraph.jl:3-12.  A node with no subgraphs is a LEAF for the back end
 * whatever its operator tag (static.jl:115). */
enum { FDG_OP_UNITARY = 0, FDG_OP_SUM = 1, FDG_OP_PROD = 2, FDG_OP_POWER = 3 };

/* back ends.  VM: the packet interpreter kernel (any program, Float64 and ComplexF64).  JIT: the emitted
 * function written as PTX and assembled for sm_100a at first use (Float64 and ComplexF64).  AUTO picks JIT
 * and falls back to the VM if the kernels cannot be built; both compute bit-identical results. */
enum { FDG_BACKEND_AUTO = 0, FDG_BACKEND_VM = 1, FDG_BACKEND_JIT = 2 };

/* element types of leafVal / root (julia_to_C_typestr, static.jl:135-153) */
enum { FDG_F64 = 0, FDG_C128 = 1 };

/* The DAG as the Julia/Python side flattens it: one entry per node OBJECT, in any order.
 * Replaces walking Graph{F,W} / FeynmanGraph{F,W} (graph.jl:28-40, feynmangraph.jl:72-82)
 * through id/operator/subgraphs/subgraph_factors. */
typedef struct fdg_graph_desc {
    int64_t n_nodes;
    int64_t n_edges;
    const int64_t *node_id;      /* [n_nodes] Graph.id -- the emitter dedupes by id, first visit wins */
    const int32_t *node_op;      /* [n_nodes] FDG_OP_*                                        */
    const int32_t *node_pow;     /* [n_nodes] N of Power{N} (ignored otherwise)               */
    const int64_t *child_ptr;    /* [n_nodes+1] CSR offsets into child_node / child_factor    */
    const int32_t *child_node;   /* [n_edges] node index of each subgraph, stored order       */
    const double *child_factor;  /* [n_edges] subgraph_factors                                */
    int64_t n_graphs;
    const int32_t *graphs;       /* [n_graphs] node index of each element of `graphs`         */
    int64_t n_roots;
    const int64_t *root_id;      /* [n_roots] the `root` kwarg: ids, default id.(graphs)      */
} fdg_graph_desc;

typedef struct fdg_options {
    int32_t dtype;        /* FDG_F64 (default) or FDG_C128                                     */
    int32_t max_slots;    /* cap on on-chip value slots per sample (0 = choose automatically)  */
    int32_t prefetch;     /* leaf prefetch distance in packets (0 = default, <0 = demand only) */
    int32_t schedule;     /* 0 = eager: multi-use values are computed as statements of their own
                             before the fold that reads them (emitter order); 1 = lazy: inside it */
    int32_t backend;      /* FDG_BACKEND_AUTO (0), FDG_BACKEND_VM (1), FDG_BACKEND_JIT (2)            */
    int32_t jit_segment;  /* machine instructions per specialised kernel, estimated (0 = 4000)  */
    int32_t cse;          /* common sub-expressions evaluated once (hash-based analogue of optimize!(level=1),
                             optimize.jl:345-390; operand order is part of the key, so every value keeps its bits):
                             0 (default) = automatic: the specialised back end plans the program with and without
                             merging and keeps the plan whose modelled time is lower -- merging saves operations but
                             makes more values cross kernel boundaries, which the memory-bound order-4 graphs cannot
                             afford; 1 = always; -1 = never                      */
    int32_t fma;          /* 0 (default): every multiply and add is rounded on its own -- the bits of the
                             emitted Julia / C function.  1 (opt-in, specialised kernels only): a multiply
                             may be fused into the add that reads it (DFMA).  One rounding fewer per fused
                             pair: results differ from the reference in the last bits (|error| stays below
                             the reference's own bound, see tests) but are NOT bit-identical.                */
} fdg_options;

typedef struct fdg_program *fdg_handle;

/* counters describing a lowered program (fdg_stats) */
typedef struct fdg_stats_t {
    int64_t n_leaves;       /* L: distinct leaves = columns of leafVal                         */
    int64_t n_inner;        /* N: distinct inner nodes (statements of the emitted function)    */
    int64_t n_roots;        /* R: columns of root                                              */
    int64_t n_operands;     /* operand references (edges after dedupe)                         */
    int64_t n_packets;      /* 16-byte VM packets                                              */
    int64_t n_slots;        /* on-chip slots per sample the program uses                       */
    int64_t n_scratch;      /* spilled values per sample kept in global scratch                */
    int64_t leaf_loads;     /* leaf fetches per sample issued by the program (>= L)            */
    int64_t flops_add;      /* real additions per sample (complex: x2)                         */
    int64_t flops_mul;      /* real multiplications per sample                                 */
    int64_t bytes_in;       /* algorithmic input bytes per sample  = sizeof(W) * L             */
    int64_t bytes_out;      /* algorithmic output bytes per sample = sizeof(W) * R (eval mode) */
    int64_t max_depth;      /* accumulator nesting depth of the program                        */
    int64_t cse_removed;    /* statements removed as copies of an earlier statement            */
    int64_t reserved[2];
} fdg_stats_t;

int fdg_abi_version(void);
const char *fdg_last_error(void);

/* --- compile: host-only, needs no GPU ---------------------------------------------------------- */
int fdg_compile(const fdg_graph_desc *graph, const fdg_options *opts /* may be NULL */, fdg_handle *out);
int fdg_destroy(fdg_handle h);
/* A flattened graph as a file, so that graphs built by a Julia session elsewhere can be evaluated here.
 * Little-endian: 8 bytes "FDGRAPH\1", int64 n_nodes, n_edges, n_graphs, n_roots, then the arrays of fdg_graph_desc in
 * declaration order (node_id, node_op, node_pow, child_ptr[n_nodes+1], child_node, child_factor, graphs, root_id),
 * each with its own element type and no padding. */
int fdg_graph_write(const fdg_graph_desc *graph, const char *path);
int fdg_compile_file(const char *path, const fdg_options *opts /* may be NULL */, fdg_handle *out);
int fdg_stats(fdg_handle h, fdg_stats_t *out);
/* leafmap: node index (into the desc arrays) of leaf k, k = 0..L-1 -- same numbering as the
 * reference's `leafmap::Dict{Int,Graph}` (static.jl:118-119), 0-based. */
int fdg_leafmap(fdg_handle h, int32_t *leaf_node /* [L] */);
/* position in `root` (0-based) written last -- eval_graph! returns that value (static.jl:127,131);
 * -1 if no root statement exists. */
int fdg_last_root(fdg_handle h, int32_t *out);
/* the lowered VM program (for inspection / tests): n_packets * 4 uint32 words */
int fdg_program_words(fdg_handle h, const uint32_t **words, int64_t *n_words);

/* --- evaluate: device pointers, asynchronous on `stream` (a cudaStream_t, may be NULL) ---------- */
/* root[r*ld_root + b] = graph r at sample b, b < batch.  Columns of `root` whose id was never met
 * are left untouched, like the reference. */
int fdg_eval(fdg_handle h, const void *leaf, int64_t ld_leaf, void *root, int64_t ld_root,
             int64_t batch, void *stream);
/* acc[r] += sum_b root_r(b): deterministic fixed-order reduction on device (no atomics).
 * acc has R doubles (2R for FDG_C128) in device memory. */
int fdg_eval_accumulate(fdg_handle h, const void *leaf, int64_t ld_leaf, int64_t batch, double *acc,
                        void *stream);
/* host-buffer convenience (the reference-facing call: leafVal / root are ordinary host arrays):
 * chunked H2D -> fdg_eval -> D2H on two streams straight from / into the caller's arrays; synchronous.  The copies run at
 * full PCIe speed when the arrays are page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory); pageable memory
 * works and is slower.  One host call at a time per handle: callers on other threads wait their turn. */
int fdg_eval_host(fdg_handle h, const void *leaf_host, int64_t ld_leaf, void *root_host, int64_t ld_root,
                  int64_t batch);
/* choose launch shape: threads per block, samples per thread (1, 2 or 4), blocks per SM (0 = auto). */
int fdg_set_launch(fdg_handle h, int32_t threads, int32_t samples_per_thread, int32_t blocks_per_sm);
/* number of kernel launches issued by this handle so far (bench's gpu_launches) */
int fdg_launch_count(fdg_handle h, int64_t *out);
/* specialised back end: generate + assemble the kernels now (host only, no GPU needed) instead of at first
 * use.  accumulate: 0 = fdg_eval variant, 1 = fdg_eval_accumulate variant.  Reports the number of kernels,
 * the rows of the cross-segment buffer and the total cubin size. */
int fdg_jit_prepare(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int32_t *n_kernels,
                    int32_t *n_cross, int64_t *cubin_bytes);
/* counters of a prepared variant, out[0..n_out): kernels, rows of the cross buffer, values crossing a kernel
 * boundary, then per sample over all kernels: leaf loads, cross loads, cross stores, operations; [7] = 1 if the
 * variant is a single grid-stride kernel; [8] = bytes of machine code of the largest kernel; [9] = 1 if the plan was made
 * from the program with common sub-expressions merged; [10] = FP64 instructions per sample the kernels execute;
 * [11] = modelled time of one sample, ps; [12] = 1 for the bulk form (persistent warp-specialised kernels fed by
 * cp.async.bulk), [13] = its dynamic shared memory per block, [14] = rows fetched a second time within a kernel
 * (served by L2, not counted in the loads above).  (leaf loads + cross loads + cross stores) x sizeof(W) is
 * the traffic the plan asks of the memory system per sample -- the figure DESIGN.md compares with ncu's dram bytes.
 * samples_per_thread = 0 asks for the variant the last launch of this handle ran (fdg_jit_ptx likewise). */
int fdg_jit_info(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int64_t *out, int32_t n_out);
/* Pipeline form of the specialised back end: ONE cooperative kernel, one block per SM; the blocks
 * of stage k run only the code of segment k (resident in that SM's instruction cache) and tiles of 32 samples flow from
 * stage to stage through L2.  fdg_pipeline_prepare builds it for a device with n_sm SMs (host only, no GPU needed) and
 * answers a query: what = 0: stages, cross rows, cross values, then per sample leaf loads, cross loads, cross stores,
 * operations, bytes of machine code of the largest stage, bytes of the linked kernel, bytes of the input ring;
 * what = 1: SMs given to each stage; what = 2: estimated issue cycles of one tile in each stage. */
int fdg_pipeline_prepare(fdg_handle h, int32_t accumulate, int32_t n_sm, int32_t what, int64_t *out, int32_t n_out);
/* after a pipeline launch on `stream` (synchronises it): out[0] = 1 if a stage gave up waiting (the results are then
 * invalid), out[1 + 2k], out[2 + 2k] = clocks stage k spent computing / waiting, summed over its warps. */
int fdg_pipeline_stats(fdg_handle h, void *stream, int64_t *out, int32_t n_out);
/* PTX text of kernel `index` of a prepared variant (for inspection / tests) */
int fdg_jit_ptx(fdg_handle h, int32_t samples_per_thread, int32_t accumulate, int32_t index, const char **ptx,
                const char **ptxas_log);

/* measurement aid: launches a kernel of 8 independent FP64 chains per thread on `stream` (8 blocks of 256 threads per SM),
 * `iters` steps of either DMUL + DADD (fma = 0: what the bit-exact kernels issue) or DFMA + DFMA (fma = 1) per chain;
 * *ops_per_launch = FP64 instructions x lanes executed.  Timed by the caller with events: the roofline's FP64 denominator. */
int fdg_probe_fp64(int32_t fma, int64_t iters, void *stream, double *sink_device, int64_t *ops_per_launch);

/* --- leaf values computed on the device from the Monte-Carlo variables ------------------------
 * What the integrand of the reference's driver does between `compile` and `eval_graph!` for every sample
 * (example/benchmark.jl:44-81, with the per-leaf metadata of `leafstates`, src/frontend/frontends.jl:175-232):
 *     kq   = K[:, 1:n_loops] * loop_basis[:, loop_index[l]]                       (FrontEnds.update / loop, pool.jl:69-78)
 *     type 1 (BareGreenId):        leaf = green(T[tau_out] - T[tau_in], dot(kq, kq) - kF^2, beta)   (benchmark.jl:113-127)
 *     type 2 (BareInteractionId):  invK = 1 / (dot(kq, kq) + lambda);  leaf = 8pi / invK * (lambda * invK)^order
 *     type 0:                      leaf = 1.0
 * Green's-function derivative orders 1..5 (green_derive, benchmark.jl:93-111: (-1)^n / n! d^n/dw^n green) are taken by
 * the reference from Lehmann.Spectral.kernelFermiT_dw^n, a dependency that is not vendored: the function is
 * unambiguous and is evaluated here in closed form (checked against 60-digit differentiation), its bits in the
 * reference are unpinned; order > 5 is FDG_ERR_UNSUPPORTED ("not implemented!", benchmark.jl:107).  exp() is the
 * device's (<= 1 ulp): leaf values agree with a CPU evaluation of the same formulas to ~1e-15 relative, not bit for
 * bit; the graph evaluation on top of them stays bit-exact. */
typedef struct fdg_leafgen_desc {
    int64_t n_leaves;
    const int32_t *leaf_type;   /* [n_leaves] 0, 1 or 2 (FrontEnds.index(typeof(properties)), diagram_id.jl:342-354)   */
    const int32_t *leaf_order;  /* [n_leaves * 2] derivative orders: (Green's function, interaction)                 */
    const int32_t *tau_in;      /* [n_leaves] 0-based index into T of extT[1]                                        */
    const int32_t *tau_out;     /* [n_leaves] 0-based index into T of extT[2]                                        */
    const int32_t *loop_index;  /* [n_leaves] 0-based index into loop_basis                                          */
    int64_t n_basis;
    int64_t n_loops;            /* <= 8 */
    int64_t dim;                /* 2 or 3 */
    int64_t n_tau;
    const double *loop_basis;   /* [n_basis * n_loops], basis vector i at loop_basis[i * n_loops ...]                */
    double kF, beta, lambda;    /* example/benchmark.jl:11-18 */
} fdg_leafgen_desc;
typedef struct fdg_leafgen *fdg_leafgen_t;
int fdg_leafgen_create(const fdg_leafgen_desc *desc, fdg_leafgen_t *out);
int fdg_leafgen_destroy(fdg_leafgen_t g);
/* The order-0 propagators, order-0 interactions and constant leaves are filled by kernels written for the graph at hand
 * (straight-line PTX, loop-basis coefficients / time indices / rows as immediates), the counter-term leaves by a
 * table-driven kernel.  This builds and assembles the specialised kernels now (host only, no GPU needed) and reports
 * out[0..n_out): kernels, bytes of machine code, instructions written, leaves covered, bytes of the largest kernel;
 * `ptx` (optional) receives the text of kernel `index`.  wide: the variant for ld_leaf * 8 >= 4 GiB. */
int fdg_leafgen_jit_prepare(fdg_leafgen_t g, int32_t wide, int32_t index, int64_t *out, int32_t n_out, const char **ptx);
/* device pointers, batch-major like everything else: component c of loop momentum j of sample b at
 * K[(j * dim + c) * ld_var + b], time t at T[t * ld_var + b]; writes leaf[l * ld_leaf + b]. */
int fdg_leafgen_fill(fdg_leafgen_t g, const double *K, const double *T, int64_t ld_var, int64_t batch, double *leaf,
                     int64_t ld_leaf, void *stream);
/* generate + evaluate without ever materialising more than a sub-batch of the leaf matrix:
 * acc[r] += sum over the batch of graph r (device pointers; acc has R doubles). */
int fdg_eval_generated_accumulate(fdg_handle h, fdg_leafgen_t g, const double *K, const double *T, int64_t ld_var,
                                  int64_t batch, double *acc, void *stream);
/* the same from HOST arrays of (K, T) -- (dim * n_loops + n_tau) doubles per sample cross the bus instead of L --
 * into a host array of R sums; synchronous. */
int fdg_eval_generated_host(fdg_handle h, fdg_leafgen_t g, const double *K_host, const double *T_host, int64_t ld_var,
                            int64_t batch, double *acc_host);

/* --- multi-GPU: one process per GPU; the only exchange is the sum of the per-root accumulators ---- */
typedef struct fdg_comm *fdg_comm_t;
int fdg_comm_unique_id(void *id128 /* 128 bytes out */);
int fdg_comm_init(fdg_comm_t *out, int32_t nranks, int32_t rank, const void *id128);
int fdg_comm_destroy(fdg_comm_t c);
/* in-place sum over ranks of n doubles in device memory (ncclAllReduce, ncclSum) */
int fdg_allreduce(fdg_comm_t c, double *acc, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FDGRAPH_H */
