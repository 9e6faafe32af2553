/* The Julia shim (feynmandiagram.jl_b200/julia/FDGraphB200.jl) cannot be run here: there is no julia binary in the image.
 * What CAN be checked is everything the shim assumes about the C side.  A Julia `struct` of Int64 / Int32 / Float64 / Ptr
 * fields has the C layout of the same fields (natural alignment, declaration order), so the byte offsets written below are
 * the `fieldoffset`s of the shim's GraphDesc, Options, LeafGenDesc and of the Int64 vector it passes as fdg_stats_t; they
 * are asserted against include/fdgraph.h at compile time.  main() then makes the calls of FDGraphB200.compile() in the
 * shim's order and argument types on a four-node graph (host only: no GPU needed) and prints what the shim reads back. */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "fdgraph.h"

/* GraphDesc (FDGraphB200.jl:22-35) */
_Static_assert(sizeof(fdg_graph_desc) == 96, "GraphDesc");
_Static_assert(offsetof(fdg_graph_desc, n_nodes) == 0 && offsetof(fdg_graph_desc, n_edges) == 8, "GraphDesc");
_Static_assert(offsetof(fdg_graph_desc, node_id) == 16 && offsetof(fdg_graph_desc, node_op) == 24 && offsetof(fdg_graph_desc, node_pow) == 32, "GraphDesc");
_Static_assert(offsetof(fdg_graph_desc, child_ptr) == 40 && offsetof(fdg_graph_desc, child_node) == 48 && offsetof(fdg_graph_desc, child_factor) == 56, "GraphDesc");
_Static_assert(offsetof(fdg_graph_desc, n_graphs) == 64 && offsetof(fdg_graph_desc, graphs) == 72, "GraphDesc");
_Static_assert(offsetof(fdg_graph_desc, n_roots) == 80 && offsetof(fdg_graph_desc, root_id) == 88, "GraphDesc");
/* Options (FDGraphB200.jl:37-46): eight Int32 */
_Static_assert(sizeof(fdg_options) == 32, "Options");
_Static_assert(offsetof(fdg_options, dtype) == 0 && offsetof(fdg_options, max_slots) == 4 && offsetof(fdg_options, prefetch) == 8, "Options");
_Static_assert(offsetof(fdg_options, schedule) == 12 && offsetof(fdg_options, backend) == 16 && offsetof(fdg_options, jit_segment) == 20, "Options");
_Static_assert(offsetof(fdg_options, cse) == 24 && offsetof(fdg_options, fma) == 28, "Options");
/* fdg_stats_t is read as zeros(Int64, 16): stats[1] = n_leaves, stats[3] = n_roots (FDGraphB200.jl:116-118) */
_Static_assert(sizeof(fdg_stats_t) == 16 * 8, "stats");
_Static_assert(offsetof(fdg_stats_t, n_leaves) == 0 && offsetof(fdg_stats_t, n_roots) == 16, "stats");
/* LeafGenDesc (FDGraphB200.jl:165-180) */
_Static_assert(sizeof(fdg_leafgen_desc) == 112, "LeafGenDesc");
_Static_assert(offsetof(fdg_leafgen_desc, n_leaves) == 0 && offsetof(fdg_leafgen_desc, leaf_type) == 8 && offsetof(fdg_leafgen_desc, leaf_order) == 16, "LeafGenDesc");
_Static_assert(offsetof(fdg_leafgen_desc, tau_in) == 24 && offsetof(fdg_leafgen_desc, tau_out) == 32 && offsetof(fdg_leafgen_desc, loop_index) == 40, "LeafGenDesc");
_Static_assert(offsetof(fdg_leafgen_desc, n_basis) == 48 && offsetof(fdg_leafgen_desc, n_loops) == 56 && offsetof(fdg_leafgen_desc, dim) == 64, "LeafGenDesc");
_Static_assert(offsetof(fdg_leafgen_desc, n_tau) == 72 && offsetof(fdg_leafgen_desc, loop_basis) == 80, "LeafGenDesc");
_Static_assert(offsetof(fdg_leafgen_desc, kF) == 88 && offsetof(fdg_leafgen_desc, beta) == 96 && offsetof(fdg_leafgen_desc, lambda) == 104, "LeafGenDesc");

int main(void) {
    /* g = (v1 + v2) * 1.5, test/compiler.jl:4-15: leaves v1, v2 (ids 1, 2), Sum (id 3), unary Prod with factor 1.5 (id 4) */
    int64_t node_id[4] = {1, 2, 3, 4};
    int32_t node_op[4] = {FDG_OP_SUM, FDG_OP_SUM, FDG_OP_SUM, FDG_OP_PROD}, node_pow[4] = {0, 0, 0, 0};
    int64_t child_ptr[5] = {0, 0, 0, 2, 3};
    int32_t child_node[3] = {0, 1, 2};
    double child_factor[3] = {1.0, 1.0, 1.5};
    int32_t graphs[1] = {3};
    int64_t root_id[1] = {4};
    fdg_graph_desc d = {4, 3, node_id, node_op, node_pow, child_ptr, child_node, child_factor, 1, graphs, 1, root_id};
    fdg_options o;
    memset(&o, 0, sizeof(o)); /* Options(dtype <: Complex ? 1 : 0, 0, 0, 0, 0, 0, 0, 0) */
    fdg_handle h = NULL;
    if (fdg_compile(&d, &o, &h) != FDG_OK) {
        printf("fdg_compile: %s\n", fdg_last_error());
        return 1;
    }
    int64_t stats[16];
    memset(stats, 0, sizeof(stats));
    if (fdg_stats(h, (fdg_stats_t *)stats) != FDG_OK) return 2;
    int32_t leaf_node[2] = {-1, -1}, last = -7;
    if (fdg_leafmap(h, leaf_node) != FDG_OK || fdg_last_root(h, &last) != FDG_OK) return 3;
    printf("L=%lld R=%lld leafmap=%d,%d last_root=%d abi=%d\n", (long long)stats[0], (long long)stats[2], leaf_node[0], leaf_node[1], last,
           fdg_abi_version());
    if (fdg_destroy(h) != FDG_OK) return 4;
    /* leafgen(leafstat, loopbasis; ...): two leaves of one momentum */
    int32_t ltype[2] = {1, 2}, lorder[4] = {0, 0, 0, 0}, tin[2] = {0, 0}, tout[2] = {1, 0}, lidx[2] = {0, 0};
    double basis[2] = {1.0, -1.0};
    fdg_leafgen_desc g = {2, ltype, lorder, tin, tout, lidx, 1, 2, 3, 2, basis, 1.919, 3.0, 1.2};
    fdg_leafgen_t lg = NULL;
    if (fdg_leafgen_create(&g, &lg) != FDG_OK) {
        printf("fdg_leafgen_create: %s\n", fdg_last_error());
        return 5;
    }
    if (fdg_leafgen_destroy(lg) != FDG_OK) return 6;
    printf("ok\n");
    return 0;
}
