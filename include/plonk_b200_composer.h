/* plonk_b200_composer.h - C ABI of the host-side circuit front end of libplonk_b200.
 *
 * The reference builds circuits with its turbo `Composer` (src/composer.rs) and a library of
 * gadgets; `benches/plonk.rs` (the circuit BASELINE.json's metric is quoted on) is written against
 * it.  That code is CPU-side, O(gates) and stays in Rust when libplonk_b200 is linked under the
 * reference; this ABI exists so that C, C++ and Python callers of the device
 * prover (plonk_b200.h: pb200_prover_new / pb200_prove) can describe the same circuits - in
 * particular the reference's own BenchCircuit - without a Rust toolchain.  No GPU is involved.
 *
 * Reference interface each entry point mirrors:
 *   pb200_composer_new                       Composer::initialized            src/composer.rs:177-190
 *   pb200_composer_append_witness            Composer::append_witness         src/composer.rs:243-260
 *   pb200_composer_append_gate               append_gate / append_custom_gate src/composer.rs:262-276
 *   pb200_composer_append_evaluated_output   append_evaluated_output          src/composer.rs:298-352
 *   pb200_composer_gate_add                  gate_add / gate_mul              src/composer.rs:402-417
 *   pb200_composer_append_constant / _public append_constant / append_public  src/composer.rs:342-370
 *   pb200_composer_assert_equal(_constant)   assert_equal(_constant)          src/composer.rs:373-400
 *   pb200_composer_component_boolean         component_boolean                src/composer/bits.rs:36-47
 *   pb200_composer_component_decomposition   component_decomposition::<N>     src/composer/bits.rs:60-98
 *   pb200_composer_component_range_bits      component_range_bits::<BITS>     src/composer/range.rs:27-40
 *   pb200_composer_component_range           component_range::<BIT_PAIRS>     src/composer/range.rs:52-57
 *   pb200_composer_append_logic              append_logic_and / _xor          src/composer/logic.rs:44-236
 *   pb200_composer_component_truncate        component_truncate::<N>          src/composer/truncate.rs:46-63
 *   pb200_composer_component_select*         component_select / _one / _zero  src/composer/select.rs:20-92
 *   pb200_composer_append_point              append_point / append_constant_point / append_public_point
 *                                                                             src/composer/point.rs:40-121
 *   pb200_composer_assert_equal_point        assert_equal_point               src/composer/point.rs:124-127
 *   pb200_composer_assert_equal_public_point assert_equal_public_point        src/composer/point.rs:134-158
 *   pb200_composer_assert_torsion_free_point assert_torsion_free_point        src/composer/point.rs:171-221
 *   pb200_composer_point_op                  component_add_point / _sub_point / _neg_point
 *                                                                             src/composer/point.rs:224-264
 *   pb200_composer_component_select_identity component_select_identity        src/composer/point.rs:322-343
 *   pb200_composer_component_select_point    component_select_point           src/composer/point.rs:387-397
 *   pb200_composer_component_mul_point       component_mul_point              src/composer/point.rs:361-378
 *   pb200_composer_component_mul_generator   component_mul_generator          src/composer/fixed_base.rs:47-226
 *   pb200_composer_bench_circuit             BenchCircuit<DEGREE>::circuit    benches/plonk.rs:12-82
 *   pb200_composer_export                    what Compiler::preprocess / Prover::prove read from the
 *                                            Composer                         src/compiler.rs:132-170
 *
 * Layout: a BlsScalar is 4 x u64 little-endian limbs in Montgomery form (as in plonk_b200.h); a
 * Witness is its u32 index; a witness point is two indices (x, y); a JubJub affine point is u then
 * v (8 x u64).  Every function returns PB200_OK or a negative status; pb200_last_error() (plonk_b200.h)
 * holds the message.  A composer handle is not thread-safe; distinct handles are independent.
 */
#ifndef PLONK_B200_COMPOSER_H
#define PLONK_B200_COMPOSER_H

#include <stddef.h>
#include <stdint.h>

#include "plonk_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Error variants of src/error.rs this path can produce, beyond those of plonk_b200.h */
#define PB200_ERR_JUBJUB_POINT (-7)     /* Error::JubJubPointNotTorsionFree */
#define PB200_ERR_JUBJUB_GENERATOR (-8) /* Error::JubJubGeneratorNotPrimeOrder */
#define PB200_ERR_JUBJUB_SCALAR (-9)    /* Error::JubJubScalarMalformed */

enum pb200_point_op { PB200_POINT_ADD = 0, PB200_POINT_SUB = 1, PB200_POINT_NEG = 2 };

typedef struct pb200_composer pb200_composer_t;

int pb200_composer_new(pb200_composer_t** out);
/* Witness-only mode, for the composer that Prover::prove runs per proof (src/compiler/prover.rs:425): the
 * circuit is re-run for its witness table and public inputs only - the gate layout is the one the prover
 * was compiled from - so gates are validated and counted but not stored; pb200_composer_export then
 * refuses `selectors` / `wires`.  Switch it on right after pb200_composer_new. */
int pb200_composer_set_witness_only(pb200_composer_t* c, int on);
void pb200_composer_free(pb200_composer_t* c);
size_t pb200_composer_constraints(const pb200_composer_t* c);
size_t pb200_composer_witnesses(const pb200_composer_t* c);
size_t pb200_composer_public_inputs(const pb200_composer_t* c);
int pb200_composer_witness_value(const pb200_composer_t* c, uint32_t w, uint64_t* out);

int pb200_composer_append_witness(pb200_composer_t* c, const uint64_t* value, uint32_t* out_w);
/* selectors: 11 scalars in the order q_m, q_l, q_r, q_o, q_f, q_c, q_arith, q_range, q_logic,
 * q_fixed_group_add, q_variable_group_add; wires: a, b, c, d; pi: NULL or the public input.
 * custom = 0 forces q_arith = 1 (append_gate), custom = 1 takes the selectors as given. */
int pb200_composer_append_gate(pb200_composer_t* c, const uint64_t* selectors, const uint32_t* wires, const uint64_t* pi, int custom);
/* *solved = 0 when q_o = 0: the gate is appended, no witness is allocated */
int pb200_composer_append_evaluated_output(pb200_composer_t* c, const uint64_t* selectors, const uint32_t* wires, const uint64_t* pi,
                                           uint32_t* out_w, int* solved);
/* gate_add and gate_mul are the same operation: q_o := -1, c := the evaluated polynomial */
int pb200_composer_gate_add(pb200_composer_t* c, const uint64_t* selectors, const uint32_t* wires, const uint64_t* pi, uint32_t* out_w);
int pb200_composer_append_constant(pb200_composer_t* c, const uint64_t* value, uint32_t* out_w);
int pb200_composer_append_public(pb200_composer_t* c, const uint64_t* value, uint32_t* out_w);
int pb200_composer_assert_equal(pb200_composer_t* c, uint32_t a, uint32_t b);
int pb200_composer_assert_equal_constant(pb200_composer_t* c, uint32_t a, const uint64_t* constant, const uint64_t* pi);

int pb200_composer_component_boolean(pb200_composer_t* c, uint32_t a);
int pb200_composer_component_decomposition(pb200_composer_t* c, uint32_t scalar, uint32_t n_bits, uint32_t* out_bits /* n_bits */);
int pb200_composer_component_range_bits(pb200_composer_t* c, uint32_t w, uint32_t bits);
int pb200_composer_component_range(pb200_composer_t* c, uint32_t w, uint32_t bit_pairs);
int pb200_composer_append_logic(pb200_composer_t* c, uint32_t a, uint32_t b, uint32_t bit_pairs, int is_xor, uint32_t* out_w);
int pb200_composer_component_truncate(pb200_composer_t* c, uint32_t w, uint32_t n_bits, uint32_t* out_w);
int pb200_composer_component_select(pb200_composer_t* c, uint32_t bit, uint32_t a, uint32_t b, uint32_t* out_w);
int pb200_composer_component_select_one(pb200_composer_t* c, uint32_t bit, uint32_t value, uint32_t* out_w);
int pb200_composer_component_select_zero(pb200_composer_t* c, uint32_t bit, uint32_t value, uint32_t* out_w);

/* kind: 0 = append_point, 1 = append_constant_point, 2 = append_public_point */
int pb200_composer_append_point(pb200_composer_t* c, const uint64_t* uv, int kind, uint32_t* out_xy);
int pb200_composer_assert_equal_point(pb200_composer_t* c, const uint32_t* a_xy, const uint32_t* b_xy);
int pb200_composer_assert_equal_public_point(pb200_composer_t* c, const uint32_t* p_xy, const uint64_t* uv);
int pb200_composer_assert_torsion_free_point(pb200_composer_t* c, const uint32_t* p_xy);
int pb200_composer_point_op(pb200_composer_t* c, int op, const uint32_t* a_xy, const uint32_t* b_xy, uint32_t* out_xy);
int pb200_composer_component_select_identity(pb200_composer_t* c, uint32_t bit, const uint32_t* a_xy, uint32_t* out_xy);
int pb200_composer_component_select_point(pb200_composer_t* c, uint32_t bit, const uint32_t* a_xy, const uint32_t* b_xy, uint32_t* out_xy);
int pb200_composer_component_mul_point(pb200_composer_t* c, uint32_t jubjub, const uint32_t* p_xy, uint32_t* out_xy);
/* generator_uv = NULL selects dusk_jubjub::GENERATOR */
int pb200_composer_component_mul_generator(pb200_composer_t* c, uint32_t jubjub, const uint64_t* generator_uv, uint32_t* out_xy);

/* dusk_jubjub::GENERATOR and scalar multiplication (scalar: canonical, 4 x u64 little-endian) */
int pb200_jubjub_generator(uint64_t* out_uv);
int pb200_jubjub_mul(const uint64_t* point_uv, const uint64_t* scalar, uint64_t* out_uv);

/* Appends BenchCircuit<degree>::circuit with its Default values (benches/plonk.rs:22-31). */
int pb200_composer_bench_circuit(pb200_composer_t* c, size_t degree);

/* Flat arrays in the layout pb200_prover_new / pb200_prove take (any pointer may be NULL):
 *   selectors 11 x constraints scalars (column-major), wires 4 x constraints u32 (column-major),
 *   witnesses n_witnesses scalars, pi_idx / pi_vals public-input gate indices (ascending) and values. */
int pb200_composer_export(const pb200_composer_t* c, uint64_t* selectors, uint32_t* wires, uint64_t* witnesses, uint64_t* pi_idx,
                          uint64_t* pi_vals);

#ifdef __cplusplus
}
#endif
#endif
