// fdg_isa.h -- the packet ISA of the graph VM: what fdg_lower.cpp emits and fdg_vm.cuh executes.
//
// A lowered program is a sequence of 16-byte packets {w0, w1, w2, w3} executed in order by every thread for
// its own sample(s).  The machine state per sample is
//   * an accumulator stack of four registers  A (top), R1, R2, R3 : A holds the fold being computed, R1..R3 the
//     partial folds of the enclosing nodes (PUSH shifts A->R1->R2->R3, the pop-combine packets shift back),
//   * a slot file v[0 .. n_slots) in shared memory (staged leaves, values of multi-use nodes, overflow of the
//     accumulator stack), and
//   * optional per-sample scratch in global memory for values that do not fit the slot file.
//
// w0 (header) = op | k << 6 | wait << 10 | push << 13 | first << 14 | slot0 << 20
//   wait  : 0 = nothing, w > 0 = cp.async.wait_group(w - 1) before the packet executes
//   push  : shift the accumulator stack before A is overwritten (MOV, TERM with `first`)
//
// Every arithmetic packet is one or more steps of a LEFT FOLD in the reference's order
// (src/backend/static.jl:13-46): Sum is ((g1*f1) + g2*f2) + ..., Prod is ((g1*f1)*g2)*f2 ...; a `* f` the
// emitter omits (f == 1) is either absent here or a multiplication by exactly 1.0, which is the identity on
// IEEE doubles, so results are bit-identical.
#ifndef FDG_ISA_H
#define FDG_ISA_H
#include <stdint.h>

#define FDG_STACK_REGS 3     /* R1..R3 below the accumulator A */
#define FDG_MAX_WAIT 6       /* largest cp.async.wait_group immediate used (header field is wait+1 <= 7) */
#define FDG_TERM_MAX 11      /* operands of one TERM record */
#define FDG_LDL_MAX 7        /* loads of one LDL packet pair */
#define FDG_CHUNK 32         /* packets per program chunk (one per lane of the fetching warp) */

enum {
    FDG_OP_END = 0,    // end of program
    FDG_OP_NOP = 1,    // padding (keeps a TERM and its extension packet inside one chunk)
    FDG_OP_LDL = 2,    // k <= 7 leaf loads, async (cp.async, one commit group): slot | leaf << 12 in w1..w3 and, for
                       //   k > 3, in the four words of the NEXT packet
    FDG_OP_SPILL = 3,  // scratch[w2] = v[w1]
    FDG_OP_FILL = 4,   // v[w1] = scratch[w2]                      (synchronous)
    FDG_OP_TERM = 5,   // a BLOCK of w1 terms with k operands each, run by a dispatch-free inner loop:
                       //     for each record: t = v[s0] * v[s1] * ... * v[s(k-1)];  A = A + t * f
                       //   (`first`: the first record sets A = t * f instead).  Records follow the header packet:
                       //     {s0 | s1 << 16, s2 | s3 << 16, f_lo, f_hi} and, for k > 4, {s4 | s5 << 16, ..., s10}.
                       //   header and records never straddle a chunk boundary.
    FDG_OP_MOV = 6,    // A = v[w1]; then *= v[w2], *= v[w3] for k = 2, 3                        (start of a fold)
    FDG_OP_MUL = 7,    // A = ((A * v[w1]) * v[w2]) * v[w3]                                     (k = 1..3)
    FDG_OP_ADD = 8,    // A = ((A + v[w1]) + v[w2]) + v[w3]
    FDG_OP_MULF = 9,   // A = (A * v[w1]) * f
    FDG_OP_SCALE = 10, // A = A * f
    FDG_OP_RADDF = 11, // A = R1 + (A * f);   pop                (Sum parent continues its fold)
    FDG_OP_RMULF = 12, // A = (R1 * A) * f;   pop                (Prod parent continues its fold)
    FDG_OP_XADDF = 13, // A = v[w1] + (A * f)                    (parent partial was parked in slot w1)
    FDG_OP_XMULF = 14, // A = (v[w1] * A) * f
    FDG_OP_POW = 15,   // A = A ^ w1       (w1 >= 2; x*x, x*x*x, Julia >= 1.9 pow_body / power_by_squaring for >= 4)
    FDG_OP_ST = 16,    // v[w1] = A
    FDG_OP_ROOT = 17,  // root[w1] = A  (eval)   or   racc[w1] += A  (accumulate)
    FDG_NUM_OPCODES = 18,
};

#define FDG_HDR(op, k, wait, push, first, slot0) \
    ((uint32_t)(op) | ((uint32_t)(k) << 6) | ((uint32_t)(wait) << 10) | ((uint32_t)(push) << 13) | \
     ((uint32_t)(first) << 14) | ((uint32_t)(slot0) << 20))
#define FDG_HDR_OP(w) ((w) & 63u)
#define FDG_HDR_K(w) (((w) >> 6) & 15u)
#define FDG_HDR_WAIT(w) (((w) >> 10) & 7u)
#define FDG_HDR_PUSH(w) (((w) >> 13) & 1u)
#define FDG_HDR_FIRST(w) (((w) >> 14) & 1u)
#define FDG_HDR_SLOT0(w) ((w) >> 20)

#define FDG_LDL_SLOT_BITS 12
#define FDG_LDL_WORD(slot, leaf) ((uint32_t)(slot) | ((uint32_t)(leaf) << FDG_LDL_SLOT_BITS))
#define FDG_MAX_SLOTS (1 << FDG_LDL_SLOT_BITS)
#define FDG_MAX_LEAVES (1 << (32 - FDG_LDL_SLOT_BITS))

#endif
