#!/bin/bash
O=gpurun_out/${1:-ntt2}
mkdir -p $O
Q="--no-cpu-baseline --no-msm-sweep --no-proof20"
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ntt" > $O/pytest_ntt.log 2>&1; echo "ntt rc=$?" > $O/status.txt
timeout 900 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "not 2_18 and not 2_20 and not cpp_mirror" > $O/pytest_prover.log 2>&1; echo "prover rc=$?" >> $O/status.txt
timeout 300 python bench.py $Q > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/status.txt
PB200_NTT_RADIX8=0 timeout 300 python bench.py $Q > $O/bench_radix2.json 2> $O/bench_radix2.err
PB200_NTT_TMA=0 timeout 300 python bench.py $Q > $O/bench_notma.json 2> $O/bench_notma.err
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ntt_pass" --launch-skip 40 -c 10 -f -o $O/ncu_r02_ntt python bench.py $Q --steps 1 --warmup 3 --inflight 2 > $O/ncu_ntt.log 2>&1
python tools/ncu_export.py $O/ncu_r02_ntt.ncu-rep --json $O/ncu_r02_ntt.json > /dev/null 2>> $O/status.txt; rm -f $O/ncu_r02_ntt.ncu-rep
cat $O/status.txt; tail -n 3 $O/pytest_ntt.log; tail -n 3 $O/pytest_prover.log
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["e2e"]["value"],1), "single", round(d["roofline"]["single_stream_ms_per_proof"],2), {k:(round(v["ms"],4), round(v["butterflies_per_s"]/1e9,1)) for k,v in d["ntt"].items()})
except Exception as e: print("ERR", e)
PY
done
