#!/bin/bash
# first GPU run of the batched-affine bucket accumulation (PB200_MSM_AFFINE=1): MSM + prover parity, then bench A/B
O=gpurun_out/${1:-affine}
mkdir -p $O
Q="--no-cpu-baseline --no-msm-sweep --no-proof20"
PB200_MSM_AFFINE=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "msm or srs" > $O/pytest_msm_affine.log 2>&1; rc=$?; echo "msm_affine rc=$rc" > $O/status.txt; if [ $rc -ne 0 ]; then PB200_MSM_AFFINE=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "known_discrete_log" > $O/sanitizer.log 2>&1; tail -40 $O/sanitizer.log; fi
PB200_MSM_AFFINE=1 timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_gadget_circuits.py -m gpu -x -q -k "not 2_18 and not 2_20 and not cpp_mirror" > $O/pytest_prover_affine.log 2>&1; echo "prover_affine rc=$?" >> $O/status.txt
PB200_MSM_AFFINE=1 timeout 300 python bench.py $Q > $O/bench_affine.json 2> $O/bench_affine.err; echo "bench_affine rc=$?" >> $O/status.txt
PB200_MSM_AFFINE=1 timeout 300 python bench.py --inflight 16 $Q > $O/bench_affine16.json 2> $O/bench_affine16.err; echo "bench_affine16 rc=$?" >> $O/status.txt
PB200_MSM_AFFINE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_affine.csv python bench.py --steps 1 --warmup 3 --inflight 4 $Q > $O/ncu_bench.log 2>&1; echo "ncu rc=$?" >> $O/status.txt
cat $O/status.txt; tail -n 12 $O/pytest_msm_affine.log; tail -n 5 $O/pytest_prover_affine.log; tail -n 3 $O/bench_affine.err
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    x=d.get("extra",{})
    print(round(d["value"],1), round(d["e2e"]["value"],1), "single", round(d["roofline"]["single_stream_ms_per_proof"],2), "acc_avg_ms", round(d["roofline"]["avg_launch_ms"],3), "synth", x.get("e2e_with_synthesis",{}).get("value"))
except Exception as e: print("ERR", e)
PY
done
