"""plonk_b200 - B200-native backend for the dusk-plonk prover hot path (NTT + G1 MSM).

Host-side mirror of the reference's crate-private seam:

    EvaluationDomain.{fft, ifft, coset_fft, coset_ifft}   reference src/fft/domain.rs:166-232
    CommitKey.commit                                      reference src/commitment_scheme/kzg10/key.rs:376-388

Everything computes on the GPU through the C ABI in include/plonk_b200.h; there is no CPU path."""
from ._lib import Pb200Error, lib  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .kzg import CommitKey, Commitment, PolynomialDegreeTooLarge  # noqa: F401
from .prover import CircuitUnsatisfied, Prover  # noqa: F401
