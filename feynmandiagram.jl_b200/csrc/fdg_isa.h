// fdg_isa.h -- the packet ISA of the graph VM: what fdg_lower.cpp emits and fdg_vm.cu executes.
//
// A lowered program is a sequence of 16-byte packets {w0, w1, w2, w3} executed in order by every
// thread for its own sample(s).  The machine state per sample is
//   * NREG = 4 accumulator registers acc0..acc3 (the partial folds of the nodes currently being
//     evaluated: acc_d belongs to nesting depth d),
//   * a slot file v[0 .. n_slots) in shared memory (staged leaves, values of multi-use nodes,
//     spilled accumulators), and
//   * optional per-sample scratch in global memory for values that do not fit the slot file.
//
// w0 = opcode | n << 8 | arg << 10
//   opcode < FDG_FIRST_REG_OP : control / memory packets
//   opcode >= FDG_FIRST_REG_OP: (base - FDG_FIRST_REG_OP') * 4 + d, register packets on acc_d
//
// Every arithmetic packet is one step of a LEFT FOLD in the reference's order
// (src/backend/static.jl:13-46): Sum is ((g1*f1) + g2*f2) + ..., Prod is ((g1*f1)*g2)*f2 ...;
// a `* f` that the emitter omits (f == 1) is either absent here or a multiplication by exactly
// 1.0, which is the identity on IEEE doubles, so results are bit-identical.
#ifndef FDG_ISA_H
#define FDG_ISA_H
#include <stdint.h>

#define FDG_NREG 4
#define FDG_MAX_WAIT 7  /* cp.async.wait_group immediate range used by the VM */

enum {
    // ---- control / memory ----
    FDG_OP_END = 0,    // end of program
    FDG_OP_LDL = 1,    // n leaf loads, async: w[i] = slot | leaf << 12     (cp.async, one commit group)
    FDG_OP_WAIT = 2,   // cp.async.wait_group arg   (arg <= FDG_MAX_WAIT)
    FDG_OP_SPILL = 3,  // scratch[arg] = v[w1]
    FDG_OP_FILL = 4,   // v[w1] = scratch[arg]      (synchronous)
    FDG_FIRST_REG_OP = 8,
};

// register packets: opcode = FDG_FIRST_REG_OP + base * 4 + d
enum {
    FDG_R_MOV = 0,    // acc_d = v[w1]; then *= v[w2], *= v[w3] for n = 2, 3   (start of a fold)
    FDG_R_MUL = 1,    // acc_d = ((acc_d * v[w1]) * v[w2]) * v[w3]              (n = 1..3)
    FDG_R_ADD = 2,    // acc_d = ((acc_d + v[w1]) + v[w2]) + v[w3]
    FDG_R_MOVF = 3,   // acc_d = v[w1] * f                                      f = (w2, w3) as double
    FDG_R_MULF = 4,   // acc_d = (acc_d * v[w1]) * f
    FDG_R_ADDF = 5,   // acc_d = acc_d + (v[w1] * f)
    FDG_R_SCALE = 6,  // acc_d = acc_d * f
    FDG_R_RADDF = 7,  // acc_{d-1} = acc_{d-1} + (acc_d * f)                    (d >= 1)
    FDG_R_RMULF = 8,  // acc_{d-1} = (acc_{d-1} * acc_d) * f
    FDG_R_XADDF = 9,  // acc_d = v[w1] + (acc_d * f)      (parent partial spilled to slot w1)
    FDG_R_XMULF = 10, // acc_d = (v[w1] * acc_d) * f
    FDG_R_POW = 11,   // acc_d = acc_d ^ arg              (arg >= 2; x*x, x*x*x, Julia >= 1.8 pow for >= 4)
    FDG_R_ST = 12,    // v[arg] = acc_d
    FDG_R_ROOT = 13,  // root[arg] = acc_d  (eval)   or   racc[arg] += acc_d  (accumulate)
    FDG_R_NBASE = 14,
};

#define FDG_NUM_OPCODES (FDG_FIRST_REG_OP + FDG_R_NBASE * FDG_NREG)

#define FDG_HDR(op, n, arg) ((uint32_t)(op) | ((uint32_t)(n) << 8) | ((uint32_t)(arg) << 10))
#define FDG_HDR_OP(w) ((w) & 0xffu)
#define FDG_HDR_N(w) (((w) >> 8) & 3u)
#define FDG_HDR_ARG(w) ((w) >> 10)
#define FDG_REGOP(base, d) (FDG_FIRST_REG_OP + (base) * FDG_NREG + (d))

#define FDG_LDL_SLOT_BITS 12
#define FDG_LDL_WORD(slot, leaf) ((uint32_t)(slot) | ((uint32_t)(leaf) << FDG_LDL_SLOT_BITS))
#define FDG_MAX_SLOTS (1 << FDG_LDL_SLOT_BITS)
#define FDG_MAX_LEAVES (1 << (32 - FDG_LDL_SLOT_BITS))
#define FDG_MAX_ARG (1 << 22)

#endif
