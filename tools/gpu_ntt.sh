#!/bin/bash
# GPU check of an NTT kernel change: NTT parity (all directions, all sizes), prover parity, bench with and without TMA staging
O=gpurun_out/${1:-ntt}
mkdir -p $O
Q="--no-cpu-baseline --no-msm-sweep --no-proof20"
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ntt" > $O/pytest_ntt.log 2>&1; echo "ntt rc=$?" > $O/status.txt
PB200_NTT_TMA=0 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ntt" > $O/pytest_ntt_notma.log 2>&1; echo "ntt_notma rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_gadget_circuits.py tests/test_gpu_level1_dropin.py -m gpu -x -q -k "not 2_18 and not 2_20 and not cpp_mirror" > $O/pytest_prover.log 2>&1; echo "prover rc=$?" >> $O/status.txt
timeout 300 python bench.py $Q > $O/bench_tma.json 2> $O/bench_tma.err; echo "bench_tma rc=$?" >> $O/status.txt
PB200_NTT_TMA=0 timeout 300 python bench.py $Q > $O/bench_notma.json 2> $O/bench_notma.err; echo "bench_notma rc=$?" >> $O/status.txt
cat $O/status.txt; tail -n 6 $O/pytest_ntt.log; tail -n 4 $O/pytest_ntt_notma.log; tail -n 4 $O/pytest_prover.log; tail -n 3 $O/bench_tma.err
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["e2e"]["value"],1), "single", round(d["roofline"]["single_stream_ms_per_proof"],2), {k:(round(v["ms"],4), round(v["butterflies_per_s"]/1e9,1)) for k,v in d["ntt"].items()})
except Exception as e: print("ERR", e)
PY
done
