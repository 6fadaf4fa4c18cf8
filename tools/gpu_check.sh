#!/bin/bash
# Full GPU check under gpurun: the -m gpu suite, smoke(), the default bench line (outputs in gpurun_out/$1).
O=gpurun_out/${1:-check}
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 600 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/status.txt
cat $O/status.txt; tail -n 6 $O/pytest_gpu.log; tail -n 2 $O/smoke.log; cut -c1-1800 $O/bench_default.json; tail -n 3 $O/bench_default.err
