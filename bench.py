#!/usr/bin/env python
"""bench.py - proofs/sec at 2^16 gates (BASELINE.json metric) for the B200-native prover backend.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA prover
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

A "step" is one pass of the hot path over one batch of synthetic input: `--inflight` independent
2^16-gate proofs issued concurrently (one host thread + CUDA stream each) on every rank.  Ranks
hold independent proofs (replicas, no data-path collective), so scaling is
weak.  `value` = proofs/s with the witness tables already resident in HBM; `e2e` = the same through
the host-pointer C ABI call (pb200_prove), i.e. including the pinned-host -> device copy of every
proof's witnesses and the device -> host read of the proof.  The roofline block describes the
dominant kernel (MSM bucket accumulation), timed live with CUDA events on its launching stream.
"""
import argparse
import math
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
LOG_GATES = 16
N_GATES = (1 << LOG_GATES) - 6          # fills the 2^16 domain, commit key trimmed to 2^16 + 7 points
SRS_POINTS = (1 << LOG_GATES) + 7
SRS_X, SRS_G = 0x1234567, 0x7654321     # synthetic "toxic waste" (seeded; a real SRS comes from a ceremony)
LABEL = b"bench-2^16"
# per G1 mixed addition (XYZZ madd = 8M + 2S in Fp) the SASS of k_msm_accumulate issues 276 IMAD.WIDE per product,
# 210 per squaring (symmetric partial products once) and 420 for the fused pair R(Q - X3) - Y1 PPP (two products,
# one Montgomery reduction): 6 products + 2 squarings + 1 fused pair
IMAD_PER_ADD = 6 * 276 + 2 * 210 + 420
FP_PRODUCTS_PER_ADD = IMAD_PER_ADD / 276.0  # in units of one carry-chained Fp product
IMAD_PER_BUTTERFLY = 137


def mont(v: int) -> bytes:
    return ((v << 256) % R_MOD).to_bytes(32, "little")


def blinders_for(i: int) -> bytes:
    import random

    rng = random.Random(0xB200_0000 + i)
    return b"".join(mont(rng.randrange(R_MOD)) for _ in range(14))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                    if val.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def result(self):
        busy = sorted(self.samples)[len(self.samples) // 2:] if self.samples else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def finish_dist(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def build_workload(circuit: str):
    """Returns (arrays, description).  "bench" is the reference's own benchmark circuit
    (benches/plonk.rs BenchCircuit<2^16>: 64129 gates of every gadget family), built by the native
    composer of libplonk_b200; "synthetic" is a random arithmetic-gate circuit filling the domain."""
    if circuit == "bench":
        from plonk_b200.gadgets import bench_circuit

        arrays = bench_circuit(1 << LOG_GATES).arrays()
        what = f"reference BenchCircuit<2^16> (benches/plonk.rs: {arrays.constraints} gates of all gadget families"
    else:
        from plonk_b200.composer import synthetic_circuit

        arrays = synthetic_circuit(N_GATES, seed=16).arrays()
        what = f"2^16-gate synthetic arithmetic circuit ({arrays.constraints} constraints"
    return arrays, what + ", domain 2^16, quotient domain 2^19)"


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from plonk_b200 import Prover
    from plonk_b200._lib import check, lib

    L = lib()
    check(L.pb200_init(local))
    arrays, workload = build_workload(args.circuit)
    srs_raw = ctypes.create_string_buffer(SRS_POINTS * 96)
    check(L.pb200_srs_setup_from_secret(mont(SRS_X), mont(SRS_G), SRS_POINTS, srs_raw))
    prover = Prover(LABEL, arrays.constraints, arrays.selectors, arrays.wires, arrays.n_witnesses, srs_raw.raw)
    inflight = args.inflight
    n_wit = arrays.n_witnesses
    # pinned host staging buffers (one per in-flight slot) and device-resident copies
    host_wit = [torch.empty(n_wit * 32, dtype=torch.uint8).pin_memory() for _ in range(inflight)]
    for t in host_wit:
        t.copy_(torch.frombuffer(bytearray(arrays.witnesses), dtype=torch.uint8))
    dev_wit = [t.cuda() for t in host_wit]
    proofs = [ctypes.create_string_buffer(1008) for _ in range(inflight)]
    pool = ThreadPoolExecutor(inflight)
    pi_idx, pi_vals, n_pi = arrays.pi_idx, arrays.pi_vals, arrays.n_pi

    def one(slot, step, resident):
        bl = blinders_for(step * inflight + slot)
        if resident:
            check(L.pb200_prove_dev(prover._h, dev_wit[slot].data_ptr(), pi_idx, pi_vals, n_pi, bl, proofs[slot], None))
        else:
            check(L.pb200_prove(prover._h, host_wit[slot].data_ptr(), n_wit, pi_idx, pi_vals, n_pi, bl, proofs[slot]))

    def run_steps(k, resident, base=0):
        # k steps = k proofs on each of the `inflight` slots; the slots run back to back without a
        # barrier between steps (independent proofs: nothing to wait for)
        def worker(slot):
            for s in range(k):
                one(slot, base + s, resident)

        list(pool.map(worker, range(inflight)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, resident):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(k, resident, base=1000)
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    run_steps(args.warmup, True)
    run_steps(max(1, args.warmup // 2), False)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.pb200_launch_count()
    ms_res = timed(args.steps, True)
    launches = L.pb200_launch_count() - launches0
    ms_e2e = timed(args.steps, False)
    # Dominant kernel (MSM bucket accumulation), timed with CUDA events on its launching stream while
    # proofs run one at a time, so the event pairs bracket the kernel alone (with several proofs in
    # flight the kernels of different streams overlap and a per-kernel duration is not meaningful).
    check(L.pb200_profile_enable(1))
    barrier()
    t_single0 = time.time()
    for s in range(3):
        one(0, 5000 + s, True)
    torch.cuda.synchronize()
    single_ms = (time.time() - t_single0) * 1e3 / 3
    acc_ms, acc_adds, acc_launches, acc_points = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    check(L.pb200_profile_read(ctypes.byref(acc_ms), ctypes.byref(acc_adds), ctypes.byref(acc_launches), ctypes.byref(acc_points)))
    check(L.pb200_profile_enable(0))
    sampler.stop_flag = True
    sampler.join()

    total_proofs = args.steps * inflight * world
    value = total_proofs / (ms_res * 1e-3)
    e2e_value = total_proofs / (ms_e2e * 1e-3)
    if rank != 0:
        finish_dist(world)
        return
    # dominant kernel: MSM bucket accumulation
    hbm_peak, peak_src = measured_peaks()
    imad = ctypes.c_double()
    check(L.pb200_imad_peak(ctypes.byref(imad)))
    acc_s = acc_ms.value * 1e-3
    adds_per_s = acc_adds.value / acc_s if acc_s else 0.0
    algo_bytes = 128.0 * acc_points.value  # 96 B affine base + 32 B scalar per MSM point
    roof = {
        "kernel": "k_msm_accumulate", "bound": "hbm",
        "achieved": algo_bytes / acc_s / 1e9 if acc_s else 0.0, "peak": hbm_peak, "unit": "GB/s",
        "frac": (algo_bytes / acc_s / 1e9 / hbm_peak) if acc_s else 0.0,
        "traffic": TRAFFIC_PER_LAUNCH, "peak_source": peak_src,
        "launches": acc_launches.value, "avg_launch_ms": acc_ms.value / max(1, acc_launches.value),
        "share_of_step": (acc_ms.value / 3) / single_ms if single_ms else None,
        "measured": "3 proofs issued one at a time after the timed region (exclusive kernel durations)",
        "single_stream_ms_per_proof": single_ms,
        "alu": {"unit": "G1 adds/s", "achieved": adds_per_s, "peak": imad.value / IMAD_PER_ADD,
                "frac": adds_per_s / (imad.value / IMAD_PER_ADD), "imad_wide_per_s_measured": imad.value,
                "carry_chain_ceiling": {"fp_products_per_s": 30.3e9, "adds_per_s": 30.3e9 / FP_PRODUCTS_PER_ADD,
                                        "frac": adds_per_s * FP_PRODUCTS_PER_ADD / 30.3e9,
                                        "source": "tools/mulbench on this pool's B200: IMAD.WIDE.U32.X (carry in/out) issues at half the "
                                                  "rate of carry-free IMAD.WIDE; 9.0 Fp-product equivalents per mixed addition"},
                "note": "both hot kernels are IMAD-pipe bound at 256/381-bit precision; HBM fraction is reported because the contract asks for it"},
    }
    ntt = ntt_microbench(L, torch, imad.value)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(arrays, 1)
    line = {
        "metric": "proofs/sec @ 2^16 gates", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (Fr 256-bit / Fp 384-bit Montgomery)", "data": "synthetic",
        "config": {"workload": workload + "; Proof bytes == CPU restatement of the reference (tests/test_gpu_prover.py)",
                   "proofs_per_step_per_gpu": inflight, "parallelism": f"replicas x{world}, no collective",
                   "l2": "per-proof working set ~0.7 GB (prover key 240 MB + MSM tables 100 MB + scratch) > 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": inflight * (n_wit * 32 + n_pi * 32 + 14 * 32),
                "d2h_bytes_per_step": inflight * (11 * 96 + 15 * 32)},
        "gpu_launches": int(launches), "clocks": sampler.result(), "roofline": roof, "ntt": ntt,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    finish_dist(world)


# dram__bytes_read.sum + dram__bytes_write.sum of one k_msm_accumulate launch (batch of 4 MSMs over 2^16+7
# points: 364.9 MB + 22.9 MB) from the committed capture profiles/ncu_r01_accumulate.ncu-rep.  The
# algorithmic 128 B/point would be 33.6 MB: the x11.5 is the deliberate table of window multiples
# (16 x 96 B gathered per point), not a re-read to fix.
TRAFFIC_PER_LAUNCH = 387.8e6


def ntt_microbench(L, torch, imad_peak):
    from plonk_b200._lib import check

    s = torch.cuda.Stream()
    out = {}
    with torch.cuda.stream(s):
        for log_n, batch in ((16, 4), (19, 5)):
            n = 1 << log_n
            x = torch.randint(0, 2**62, (batch * n, 4), dtype=torch.int64, device="cuda")
            y = torch.empty_like(x)
            f = lambda: check(L.pb200_ntt_dev(x.data_ptr(), n, y.data_ptr(), log_n, 0, 1, batch, n, n, s.cuda_stream))
            for _ in range(3):
                f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.synchronize()
            e0.record(s)
            for _ in range(10):
                f()
            e1.record(s)
            s.synchronize()
            ms = e0.elapsed_time(e1) / 10
            bf = batch * (n // 2) * log_n / (ms * 1e-3)
            out[f"coset_ntt_2^{log_n}_batch{batch}"] = {"ms": ms, "butterflies_per_s": bf, "algorithmic_GBps": 64 * n * batch / ms / 1e6,
                                                         "alu_frac": bf / (imad_peak / IMAD_PER_BUTTERFLY)}
    return out


def cpu_baseline(arrays, n_proofs):
    """Restated reference (oracle/cref.cpp, all host threads) on a bounded sample of the same workload."""
    from oracle import cref

    threads = cref.best_threads()
    srs = cref.srs_from_secret(SRS_POINTS, SRS_X, SRS_G, threads)
    ca = arrays
    prover = cref.CrefProver(LABEL, ca, srs, threads)
    t0 = time.time()
    for i in range(n_proofs):
        prover.prove(blinders_for(i), ca)
    dt = time.time() - t0
    # the two kernels on their own, same sizes as the GPU microbenchmarks
    import random

    rng = random.Random(19)
    n19 = 1 << (LOG_GATES + 3)
    vec = b"".join(mont(rng.randrange(R_MOD)) for _ in range(1 << 12)) * (n19 >> 12)
    t0 = time.time()
    for _ in range(2):
        cref.ntt(vec, LOG_GATES + 3, 0, 1, threads)
    ntt_s = (time.time() - t0) / 2
    scalars = vec[: SRS_POINTS * 32]
    t0 = time.time()
    cref.msm(srs, scalars, threads)
    msm_s = time.time() - t0
    window = int(math.log(SRS_POINTS)) + 2  # msm_variable_base's window rule
    return {"value": n_proofs / dt, "unit": "proofs/s", "cores": threads, "kind": "port",
            "sample": f"{n_proofs} proof(s) of the same 2^16-gate circuit, C++/OpenMP restatement of the reference prover "
                      f"(the Rust crate cannot be built here: no cargo/rustc); "
                      f"thread count = fastest of T, T/2, T/4 for the {cref.threads()} usable host threads",
            "coset_ntt_2^19": {"ms": ntt_s * 1e3, "butterflies_per_s": (n19 // 2) * (LOG_GATES + 3) / ntt_s},
            "msm_2^16": {"ms": msm_s * 1e3, "points_per_s": SRS_POINTS / msm_s,
                         "bucket_adds_per_s": SRS_POINTS * math.ceil(255 / window) / msm_s, "window_bits": window}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import cref

    if args.circuit == "bench":
        # the reference arm runs none of the product's code: the circuit comes from the oracle's own composer
        # (byte-identical to the native one, tests/test_gadgets.py)
        from oracle import gadgets as oracle_gadgets

        comp = oracle_gadgets.GadgetComposer.initialized()
        oracle_gadgets.bench_circuit(comp, 1 << LOG_GATES)
        arrays = cref.CircuitArrays(comp)
        workload = (f"reference BenchCircuit<2^16> (benches/plonk.rs: {arrays.constraints} gates of all gadget families, "
                    "domain 2^16, quotient domain 2^19)")
    else:
        arrays, workload = build_workload(args.circuit)

    threads = cref.best_threads()
    srs = cref.srs_from_secret(SRS_POINTS, SRS_X, SRS_G, threads)
    prover = cref.CrefProver(LABEL, arrays, srs, threads)
    for i in range(args.warmup):
        prover.prove(blinders_for(i), arrays)
    t0 = time.time()
    for i in range(args.steps):
        prover.prove(blinders_for(1000 + i), arrays)
    dt = time.time() - t0
    value = args.steps / dt
    sample = (f"each step = 1 proof of the 2^16-gate workload on {threads} host threads (fastest of T, T/2, T/4 for the "
              f"{cref.threads()} usable ones; C++/OpenMP restatement; the Rust reference cannot be built here)")
    print(json.dumps({
        "impl": "reference", "metric": "proofs/sec @ 2^16 gates", "value": value, "unit": "proofs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (CPU)", "data": "synthetic",
        "config": {"workload": workload, "proofs_per_step": 1},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("PB200_INFLIGHT", "8")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--circuit", default="bench", choices=["bench", "synthetic"],
                    help="bench = the reference's BenchCircuit<2^16> (default); synthetic = random arithmetic gates")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 20
        args.warmup = max(3, args.warmup if args.warmup is not None else 3)
        run_ours(args)


if __name__ == "__main__":
    main()
