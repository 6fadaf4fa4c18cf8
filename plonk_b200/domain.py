"""EvaluationDomain mirror (reference src/fft/domain.rs:35-232) over the CUDA NTT.

Vectors cross this boundary as ``bytes``/buffers of 32-byte Fr elements in the reference's
in-memory layout (4 x u64 little-endian limbs, Montgomery form)."""
from __future__ import annotations

import ctypes

from ._lib import check, lib

TWO_ADACITY = 32
FR_BYTES = 32


class InvalidEvalDomainSize(ValueError):
    """Error::InvalidEvalDomainSize (reference src/fft/domain.rs:132-137)."""


class EvaluationDomain:
    def __init__(self, num_coeffs: int):
        size = 1 if num_coeffs <= 1 else 1 << (num_coeffs - 1).bit_length()
        log = size.bit_length() - 1
        if log >= TWO_ADACITY:
            raise InvalidEvalDomainSize(f"log_size_of_group {log} >= {TWO_ADACITY}")
        self.size = size
        self.log_size_of_group = log

    def _run(self, data: bytes, inverse: int, coset: int) -> bytes:
        assert len(data) % FR_BYTES == 0
        in_len = len(data) // FR_BYTES
        out = ctypes.create_string_buffer(self.size * FR_BYTES)
        src = (ctypes.c_char * len(data)).from_buffer_copy(data) if data else None
        check(lib().pb200_ntt(src, in_len, out, self.log_size_of_group, inverse, coset, 1, in_len, self.size))
        return out.raw

    def fft(self, coeffs: bytes) -> bytes:
        return self._run(coeffs, 0, 0)

    def ifft(self, evals: bytes) -> bytes:
        return self._run(evals, 1, 0)

    def coset_fft(self, coeffs: bytes) -> bytes:
        return self._run(coeffs, 0, 1)

    def coset_ifft(self, evals: bytes) -> bytes:
        return self._run(evals, 1, 1)

    def batch(self, vectors, inverse: int, coset: int):
        """Several transforms in one launch set (the reference issues 4 iffts / 5 coset_ffts
        concurrently from rayon workers, src/compiler/prover.rs:163-185, quotient_poly.rs:139-157)."""
        if not vectors:
            return []
        in_len = max(len(v) for v in vectors) // FR_BYTES
        stride = max(in_len, 1)
        buf = bytearray(stride * FR_BYTES * len(vectors))
        for i, v in enumerate(vectors):
            buf[i * stride * FR_BYTES : i * stride * FR_BYTES + len(v)] = v
        out = ctypes.create_string_buffer(self.size * FR_BYTES * len(vectors))
        src = (ctypes.c_char * len(buf)).from_buffer(buf)
        check(lib().pb200_ntt(src, in_len, out, self.log_size_of_group, inverse, coset, len(vectors), stride, self.size))
        raw = out.raw
        step = self.size * FR_BYTES
        return [raw[i * step : (i + 1) * step] for i in range(len(vectors))]
