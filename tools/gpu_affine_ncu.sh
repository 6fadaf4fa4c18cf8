O=gpurun_out/affine_ncu; mkdir -p $O
Q="--no-cpu-baseline --no-msm-sweep --no-proof20 --steps 1 --warmup 3 --inflight 2"
PB200_MSM_AFFINE=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_msm_affine_round" --launch-skip 60 -c 8 -f -o $O/ncu_r02_affine python bench.py $Q > $O/ncu.log 2>&1
python tools/ncu_export.py $O/ncu_r02_affine.ncu-rep --json $O/ncu_r02_affine.json > /dev/null
ncu -i $O/ncu_r02_affine.ncu-rep --page details --csv > $O/ncu_r02_affine.details.csv 2>/dev/null
rm -f $O/ncu_r02_affine.ncu-rep
for n in 16 24; do PB200_MSM_AFFINE=1 timeout 300 python bench.py --inflight $n --no-cpu-baseline --no-msm-sweep --no-proof20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('affine inflight', $n, d['value'], d['e2e']['value'])"; done
timeout 300 python bench.py --inflight 16 --no-cpu-baseline --no-msm-sweep --no-proof20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('xyzz inflight 16', d['value'], d['e2e']['value'])"
