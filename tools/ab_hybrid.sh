#!/bin/bash
# A/B of the two-pipe products under gpurun.  Variants (all from the same sources):
#   imad: ./build.sh                                   -> plonk_b200/libplonk_b200.so (the default build)
#   hyb : PB200_OUT=plonk_b200/libplonk_b200_hyb.so PB200_OBJ=build/obj_hyb ./build.sh -DPB_FP_HYBRID=1
#   frh : PB200_OUT=plonk_b200/libplonk_b200_frh.so PB200_OBJ=build/obj_frh ./build.sh -DPB_FP_HYBRID=1 -DPB_FR_HYBRID=1
# (when profiles/ab_hybrid_r01 was recorded the two-pipe Fp product was still the default build and the
# all-IMAD one was the variant; the file names below follow the current defaults)
O=gpurun_out/ab_hybrid
mkdir -p $O
rm -f $O/status.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 120 tools/mulbench/mulbench > $O/mulbench.log 2>&1
echo "mulbench rc=$?" >> $O/status.txt
export PB200_LIB=$PWD/plonk_b200/libplonk_b200_hyb.so
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "montgomery or product_forms or msm or compress or setup" > $O/parity_kernels_hyb.log 2>&1
echo "parity_kernels_hyb rc=$?" >> $O/status.txt
export PB200_LIB=$PWD/plonk_b200/libplonk_b200_frh.so
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "montgomery or product_forms or ntt_matches" > $O/parity_kernels_frh.log 2>&1
echo "parity_kernels_frh rc=$?" >> $O/status.txt
timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "golden_digest or matches_cpu_oracle and not 2_1 and not 2_20" > $O/parity_prover_frh.log 2>&1
echo "parity_prover_frh rc=$?" >> $O/status.txt
for v in frh hyb imad; do
  if [ $v = imad ]; then unset PB200_LIB; else export PB200_LIB=$PWD/plonk_b200/libplonk_b200_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
  echo "bench_$v rc=$?" >> $O/status.txt
done
for v in frh imad; do
  if [ $v = imad ]; then unset PB200_LIB; else export PB200_LIB=$PWD/plonk_b200/libplonk_b200_$v.so; fi
  timeout 200 python bench.py --no-cpu-baseline --inflight 1 --steps 10 > $O/bench_${v}_inflight1.json 2> $O/bench_${v}_inflight1.err
done
cat $O/status.txt; tail -12 $O/mulbench.log; tail -2 $O/parity_kernels_hyb.log $O/parity_kernels_frh.log $O/parity_prover_frh.log
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["e2e"]["value"], d["roofline"].get("alu",{}).get("achieved"), d.get("ntt"))
except Exception as e: print("ERR", e)
PY
done
