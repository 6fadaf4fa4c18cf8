"""ctypes wrapper around oracle/_build/libcref.so (the multi-threaded C++ restatement).

TEST INFRASTRUCTURE / CPU BASELINE ONLY - never imported from plonk_b200/."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence

from . import pyref as R

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        # libgomp reads OMP_NUM_THREADS when it initialises; torchrun exports OMP_NUM_THREADS=1, under
        # which even explicit num_threads() clauses were measured to run slower than one thread.
        os.environ["OMP_NUM_THREADS"] = str(threads())
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", _HERE])
        L = ctypes.CDLL(_SO)
        c = ctypes
        L.cref_ntt.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p, c.c_uint32, c.c_int, c.c_int, c.c_int]
        L.cref_msm.argtypes = [c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_int]
        L.cref_srs_from_secret.argtypes = [c.c_size_t, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int]
        L.cref_prover_new.restype = c.c_void_p
        L.cref_prover_new.argtypes = [c.c_void_p, c.c_size_t, c.c_size_t, c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_int]
        L.cref_prover_free.argtypes = [c.c_void_p]
        L.cref_prover_commitments.argtypes = [c.c_void_p, c.c_void_p]
        L.cref_prove.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_void_p]
        _lib = L
    return _lib


def threads() -> int:
    """Host threads for the CPU baseline: every core this process may run on.  (Not
    omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its children, which would silently
    make the reference arm single-threaded; every parallel region in cref.cpp takes an explicit
    num_threads.)"""
    env = os.environ.get("PB200_CPU_THREADS")
    if env:
        return int(env)
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _cgroup_cpu_limit() -> Optional[int]:
    """CPU bandwidth quota of this container (cgroup v2 cpu.max or v1 cfs quota), in whole CPUs."""
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            return max(1, -(-int(quota) // int(period)))
    except (OSError, ValueError):
        pass
    try:
        quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if quota > 0 and period > 0:
            return max(1, -(-quota // period))
    except (OSError, ValueError):
        pass
    return None


_best_threads: Optional[int] = None


PORT_MAX_THREADS = 32


def best_threads() -> int:
    """Thread count of the CPU baseline - ONE fixed policy, so that the number does not swing with a
    calibration run (round 1 picked the fastest of T, T/2, T/4 on a short sample, which chose 16, 48 or 96
    threads on different boxes): min(32, the CPUs this process may run on, the container's CPU quota).
    32 is where the restatement stops scaling - its MSM is window-parallel (20 windows at 2^16 points) and
    the prover runs 4- to 5-way outer concurrency like the reference's rayon::join sites - and more threads
    were measured slower (0.27 proofs/s at 64 threads against 0.66 at 16 on a 128-CPU box).
    PB200_CPU_THREADS overrides."""
    global _best_threads
    if _best_threads is None:
        if os.environ.get("PB200_CPU_THREADS"):
            _best_threads = threads()
        else:
            top = min(PORT_MAX_THREADS, threads())
            limit = _cgroup_cpu_limit()
            _best_threads = max(1, min(top, limit) if limit else top)
    return _best_threads


def thread_policy() -> str:
    return (f"{best_threads()} threads = min({PORT_MAX_THREADS}, the {threads()} CPUs usable by this process, the cgroup CPU quota): "
            "fixed policy (the restatement stops scaling there), PB200_CPU_THREADS overrides")


def _default_threads(work_items: int) -> int:
    """Threads for a stand-alone call: small inputs stay narrow, large ones use the calibrated width."""
    if work_items < (1 << 13):
        return max(1, min(4, threads()))
    return best_threads()


def ntt(data: bytes, log_n: int, inverse: int, coset: int, nthreads: Optional[int] = None) -> bytes:
    out = ctypes.create_string_buffer(32 << log_n)
    lib().cref_ntt(data, len(data) // 32, out, log_n, inverse, coset, nthreads or _default_threads(1 << log_n))
    return out.raw


def msm(bases_raw: bytes, scalars: bytes, nthreads: Optional[int] = None) -> bytes:
    n = min(len(bases_raw) // 96, len(scalars) // 32)
    out = ctypes.create_string_buffer(96)
    lib().cref_msm(bases_raw, scalars, n, out, nthreads or _default_threads(n))
    return out.raw


def srs_from_secret(n: int, x: int, g_scalar: int, nthreads: Optional[int] = None) -> bytes:
    out = ctypes.create_string_buffer(96 * n)
    lib().cref_srs_from_secret(n, R.fr_to_mont_bytes(x), R.fr_to_mont_bytes(g_scalar), out, nthreads or _default_threads(n))
    return out.raw


class CircuitArrays:
    """Flat description of a composed circuit, shared by the C++ oracle and the GPU prover:
    11 selector columns (Montgomery Fr), 4 wire columns (u32 witness indices), witnesses,
    sorted public-input positions and values."""

    def __init__(self, comp: R.Composer):
        n = len(comp.constraints)
        self.constraints = n
        self.selectors = b"".join(R.fr_vec_to_mont_bytes([g.sel[k] for g in comp.constraints]) for k in R.SELECTORS)
        cols = [[g.a for g in comp.constraints], [g.b for g in comp.constraints], [g.c for g in comp.constraints], [g.d for g in comp.constraints]]
        self.wires = b"".join(int(w).to_bytes(4, "little") for col in cols for w in col)
        self.n_witnesses = len(comp.witnesses)
        self.witnesses = R.fr_vec_to_mont_bytes(comp.witnesses)
        idx = comp.public_input_indexes()
        self.pi_idx = b"".join(i.to_bytes(8, "little") for i in idx)
        self.pi_vals = R.fr_vec_to_mont_bytes(comp.public_inputs_vec())
        self.n_pi = len(idx)


def draw_blinders(rng: R.StdRng) -> bytes:
    """The 14 BlsScalar::random draws of one Prover::prove call, in RNG order (prover.rs:154-161,
    503, 553-555)."""
    return R.fr_vec_to_mont_bytes([R.fr_random(rng) for _ in range(14)])


class CrefProver:
    def __init__(self, label: bytes, arrays: CircuitArrays, srs_raw: bytes, nthreads: Optional[int] = None):
        self.arrays = arrays
        if nthreads is None:
            # small circuits gain nothing from a wide OpenMP team (and lose a lot on an oversubscribed host)
            nthreads = max(1, min(best_threads(), arrays.constraints // 512))
        self._h = lib().cref_prover_new(label, len(label), arrays.constraints, arrays.selectors, arrays.wires,
                                        arrays.n_witnesses, srs_raw, len(srs_raw) // 96, nthreads)
        if not self._h:
            raise ValueError("cref_prover_new failed (SRS too small?)")

    def commitments(self) -> List[bytes]:
        out = ctypes.create_string_buffer(48 * 15)
        lib().cref_prover_commitments(self._h, out)
        return [out.raw[48 * i : 48 * (i + 1)] for i in range(15)]

    def prove(self, blinders: bytes, arrays: Optional[CircuitArrays] = None) -> bytes:
        a = arrays or self.arrays
        out = ctypes.create_string_buffer(1008)
        rc = lib().cref_prove(self._h, a.witnesses, a.pi_idx, a.pi_vals, a.n_pi, blinders, out)
        if rc != 0:
            raise ValueError(f"cref_prove failed: {rc}")
        return out.raw

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cref_prover_free(self._h)
            self._h = None
