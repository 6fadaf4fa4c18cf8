#!/bin/bash
# Round-end verification under gpurun: the full -m gpu suite, smoke(), the default bench line, the reference arm,
# and the ncu launch list of the default command (outputs in gpurun_out/$1, copied to profiles/ afterwards).
O=gpurun_out/${1:-final_r02}
mkdir -p $O
( time timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/status.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench_ref rc=$?" >> $O/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_default.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-msm-sweep --no-proof20 > $O/ncu_bench.log 2>&1; echo "ncu rc=$?" >> $O/status.txt
cat $O/status.txt; tail -n 16 $O/pytest_gpu.log; tail -n 2 $O/smoke.log; cut -c1-9000 $O/bench_default.json; echo; cut -c1-900 $O/bench_reference.json; wc -l $O/launches_default.csv
