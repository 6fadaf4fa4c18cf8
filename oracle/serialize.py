"""Byte formats of the reference's serialized keys, restated for the tests of the product's loaders
(pb200_commit_key_from_raw_var_bytes, pb200_prover_from_bytes).  TEST INFRASTRUCTURE - never imported from
plonk_b200/.

  CommitKey::to_raw_var_bytes         src/commitment_scheme/kzg10/key.rs:215-229
  PublicParameters::to_raw_var_bytes  src/commitment_scheme/kzg10/srs.rs:114-119
  Prover::to_bytes                    src/compiler/prover.rs:211-263
  ProverKey::to_var_bytes             src/proof_system/widget.rs:347-445
  VerifierKey::to_bytes               src/proof_system/widget.rs:84-111
  Evaluations::to_var_bytes           src/fft/evaluations.rs:52-61
  EvaluationDomain::to_bytes          src/fft/domain.rs:59-80
  Polynomial::to_var_bytes            src/fft/polynomial.rs:141-149

G1Affine::to_raw_bytes lives in dusk-bls12_381 0.14 (not under the reference checkout): x then y as little-endian
u64 Montgomery limbs and one infinity byte, 97 bytes (its published source; the reference has no golden bytes for
it, so this layout is NOT pinned by a reference vector - parity unpinned for the raw point codec)."""
from __future__ import annotations

from typing import Sequence

from . import pyref as R

# file order of the 15 polynomials / commitments (widget.rs:360-440, :93-108)
FILE_ORDER = ["q_m", "q_l", "q_r", "q_o", "q_f", "q_c", "q_arith", "q_logic", "q_range", "q_fixed_group_add",
              "q_variable_group_add", "s_sigma_1", "s_sigma_2", "s_sigma_3", "s_sigma_4"]
G1_RAW_SIZE = 97
OPENING_KEY_BYTES = 48 + 96 + 96


def g1_to_raw_bytes_97(p) -> bytes:
    if p is None:  # G1Affine::identity(): x = 0, y = 1 (Montgomery one), infinity = 1
        return bytes(48) + ((1 << 384) % R.P_MOD).to_bytes(48, "little") + b"\x01"
    return R.g1_to_raw_bytes(p) + b"\x00"


def commit_key_to_raw_var_bytes(points: Sequence) -> bytes:
    return len(points).to_bytes(8, "little") + b"".join(g1_to_raw_bytes_97(p) for p in points)


def domain_to_bytes(size: int) -> bytes:
    d = R.EvaluationDomain(size)
    out = d.size.to_bytes(8, "little") + (d.size.bit_length() - 1).to_bytes(4, "little")
    for v in (d.size % R.R_MOD, d.size_inv, d.group_gen, d.group_gen_inv, pow(R.GENERATOR, -1, R.R_MOD)):
        out += R.fr_to_bytes(v)
    return out


def evaluations_to_var_bytes(evals: Sequence[int]) -> bytes:
    return domain_to_bytes(len(evals)) + b"".join(R.fr_to_bytes(v) for v in evals)


def prover_key_to_var_bytes(pd: "R.ProverData") -> bytes:
    n8 = 8 * pd.size
    eval_size = n8 * 32 + 172
    out = pd.size.to_bytes(8, "little") + eval_size.to_bytes(8, "little")
    for k in FILE_ORDER:
        coeffs = pd.polys[k]
        out += len(coeffs).to_bytes(8, "little") + b"".join(R.fr_to_bytes(c) for c in coeffs)
        out += evaluations_to_var_bytes(pd.evals_8n[k])
    out += evaluations_to_var_bytes(pd.evals_8n["linear"]) + evaluations_to_var_bytes(pd.evals_8n["v_h"])
    return out


def verifier_key_to_bytes(pd: "R.ProverData") -> bytes:
    out = pd.size.to_bytes(8, "little") + b"".join(R.g1_compress(pd.comms[k]) for k in FILE_ORDER)
    return out + bytes(20 * 48 + 8 - len(out))


def prover_to_bytes(pd: "R.ProverData") -> bytes:
    pk, ck, vk = prover_key_to_var_bytes(pd), commit_key_to_raw_var_bytes(pd.commit_key), verifier_key_to_bytes(pd)
    head = b"".join(v.to_bytes(8, "big") for v in (len(pd.label), len(pk), len(ck), len(vk), pd.size, pd.constraints))
    return head + pd.label + pk + ck + vk
