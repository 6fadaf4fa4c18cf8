#!/bin/bash
# First GPU run of the two pending patches (run tools/build_patched.sh on the CPU first), under gpurun:
#   1. the group-element inverse NTT alone (pb200_g1_lagrange_key, in the default library already)
#   2. the patched library with each switch: the prover parity tests must stay green (Proof bytes unchanged)
#   3. bench.py: default library, patched library without switches, with each, with both
O=gpurun_out/ab_patches
mkdir -p $O
rm -f $O/status.txt
PB200_TEST_LAGRANGE=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k lagrange > $O/lagrange_key.log 2>&1
echo "lagrange_key rc=$?" >> $O/status.txt
export PB200_LIB=$PWD/plonk_b200/libplonk_b200_patched.so
K="not 2_18 and not 2_20 and not cpp_mirror"
timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_gpu_gadget_circuits.py -m gpu -x -q -k "$K" > $O/parity_patched_off.log 2>&1
echo "parity_patched_off rc=$?" >> $O/status.txt
PB200_QUOT4N=1 timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_gpu_gadget_circuits.py -m gpu -x -q -k "$K" > $O/parity_quot4n.log 2>&1
echo "parity_quot4n rc=$?" >> $O/status.txt
PB200_LAGRANGE=1 timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_gpu_gadget_circuits.py -m gpu -x -q -k "$K" > $O/parity_lagrange.log 2>&1
echo "parity_lagrange rc=$?" >> $O/status.txt
PB200_QUOT4N=1 PB200_LAGRANGE=1 timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_gpu_gadget_circuits.py -m gpu -x -q -k "$K" > $O/parity_both.log 2>&1
echo "parity_both rc=$?" >> $O/status.txt
( unset PB200_LIB; timeout 300 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err )
timeout 300 python bench.py --no-cpu-baseline > $O/bench_patched_off.json 2> $O/bench_patched_off.err
PB200_QUOT4N=1 timeout 300 python bench.py --no-cpu-baseline > $O/bench_quot4n.json 2> $O/bench_quot4n.err
PB200_LAGRANGE=1 timeout 300 python bench.py --no-cpu-baseline > $O/bench_lagrange.json 2> $O/bench_lagrange.err
PB200_QUOT4N=1 PB200_LAGRANGE=1 timeout 300 python bench.py --no-cpu-baseline > $O/bench_both.json 2> $O/bench_both.err
cat $O/status.txt
for f in $O/parity_*.log $O/lagrange_key.log; do echo $f; tail -n 3 $f; done
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], d["roofline"].get("alu",{}).get("achieved"))
except Exception as e: print("ERR", e)
PY
done
