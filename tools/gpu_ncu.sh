#!/bin/bash
# ncu --set full captures of the round's main kernels (one GPU; each kernel replayed ~40 times) + the launch list
# of the default bench command.  Reports land in gpurun_out/$1; export them with tools/ncu_export.py.
O=gpurun_out/${1:-ncu}
mkdir -p $O
Q="--no-cpu-baseline --no-msm-sweep --no-proof20 --steps 1 --warmup 3 --inflight 2"
cap() {  # name, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 -c $4 -f -o $O/$1 python bench.py $Q > $O/$1.log 2>&1
  echo "$1 rc=$?" >> $O/status.txt
  # text exports on the box (gpurun brings back at most 64 MiB: the binary reports stay behind)
  python tools/ncu_export.py $O/$1.ncu-rep --json $O/$1.json > /dev/null 2>> $O/status.txt
  rm -f $O/$1.ncu-rep
}
rm -f $O/status.txt
cap ncu_r02_accumulate "k_msm_accumulate" 12 4
cap ncu_r02_quotient "k_quotient_4n" 3 1
cap ncu_r02_ntt "k_ntt_pass" 40 10
cap ncu_r02_groups "k_msm_groups|k_msm_group_classes|k_msm_final" 12 3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_default.csv python bench.py --no-cpu-baseline --no-msm-sweep --no-proof20 --steps 1 --warmup 3 > $O/launches.log 2>&1; echo "launches rc=$?" >> $O/status.txt
cat $O/status.txt; ls -la $O
