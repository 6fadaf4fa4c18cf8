#!/bin/bash
# quick GPU check: parity tests without the slow ones, then the full default bench line (outputs in gpurun_out/$1)
O=gpurun_out/${1:-quick}
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q -k "not 2_20 and not legacy and not 2_18 and not 2_22" ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/status.txt
timeout 900 python bench.py ${BENCH_ARGS} > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?" >> $O/status.txt
cat $O/status.txt; tail -n 5 $O/pytest_gpu.log; tail -n 5 $O/bench_full.err; cut -c1-7000 $O/bench_full.json
