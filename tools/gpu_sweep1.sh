#!/bin/bash
# round-2 measurement pass: full new bench line, ncu launch list, proofs-in-flight sweep, Lagrange-key window sweep
O=gpurun_out/${1:-sweep1}
mkdir -p $O
timeout 900 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench_full rc=$?" > $O/status.txt
Q="--no-cpu-baseline --no-msm-sweep --no-proof20"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_inflight8.csv python bench.py --steps 1 --warmup 3 $Q > $O/ncu_bench.log 2>&1; echo "ncu rc=$?" >> $O/status.txt
for n in 4 6 12 16; do timeout 300 python bench.py --inflight $n --steps 12 $Q > $O/bench_inflight$n.json 2> $O/bench_inflight$n.err; done
for c in 10 12 13 14; do PB200_LAG_C=$c timeout 300 python bench.py $Q > $O/bench_lagc$c.json 2> $O/bench_lagc$c.err; done
cat $O/status.txt; tail -n 5 $O/bench_full.err
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    x=d.get("extra",{})
    print(round(d["value"],1), round(d["e2e"]["value"],1), "single", round(d["roofline"]["single_stream_ms_per_proof"],2), "synth", x.get("e2e_with_synthesis",{}).get("value"))
except Exception as e: print("ERR", e)
PY
done
cut -c1-6000 $O/bench_full.json
