// fdg_lower.h -- host-side lowering: flattened Graph DAG -> emitter-order statements -> VM packets.
#ifndef FDG_LOWER_H
#define FDG_LOWER_H
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fdgraph.h"

namespace fdg {

struct Operand {
    int32_t val;  // statement index of the operand
    double f;     // subgraph factor
};

struct Stmt {          // one value of the emitted function, in emitter order (static.jl:106-129)
    int8_t op = -1;    // -1 leaf, else FDG_OP_SUM / PROD / POWER
    int32_t pow_n = 0;
    int64_t first = 0;  // operands[first .. first+count)
    int32_t count = 0;
    int32_t root = -1;  // root position assigned right after this statement
    int32_t uses = 0;   // operand references from live statements
    bool live = false;
    int32_t leaf = -1;  // leaf index k for leaves
};

struct Lowered {
    int32_t dtype = FDG_F64;
    // emitter-order view (static.jl:98-133)
    int64_t L = 0, N = 0, R = 0;
    std::vector<int32_t> leaf_node;  // leaf k -> desc node index   (the reference's leafmap)
    std::vector<uint8_t> root_set;   // root position r is assigned by the program
    int32_t last_root = -1;          // root position assigned last (eval_graph!'s return value)
    std::vector<Stmt> st;      // the emitted function, statement by statement
    std::vector<Operand> ops;  // operands of all statements
    // VM program
    std::vector<uint32_t> words;  // 4 per packet
    int32_t n_slots = 0;
    int32_t n_scratch = 0;
    int32_t max_depth = 0;
    // counters
    int64_t n_operands = 0, leaf_loads = 0;
    // operation counts per sample: value*value, value*real-factor, value+value, multiplies inside Power
    int64_t muls_vv = 0, muls_vf = 0, adds_vv = 0, pow_muls = 0;
    int64_t cse_removed = 0;  // statements that turned out to be copies of an earlier one
    std::vector<int32_t> canon;  // with merging: statement -> the earlier statement it is a copy of (itself otherwise)
};

// returns FDG_OK or an FDG_ERR_* code with `err` filled
int lower(const fdg_graph_desc &g, const fdg_options &opt, Lowered &out, std::string &err);

}  // namespace fdg
#endif
