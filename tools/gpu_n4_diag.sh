#!/bin/bash
# 4-GPU diagnostics: where do proofs/s go when N grows? (host wake-up latency vs nvidia-smi polling)
O=gpurun_out/n4c; mkdir -p $O
uptime; cat /proc/loadavg
run() {  # name, port, env...
  name=$1; port=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --msm-sizes 16,20 > $O/$name.json 2> $O/$name.err
  python - "$O/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    h=d["extra"]["host"]
    print(sys.argv[2], round(d["value"],1), round(d["e2e"]["value"],1), round(d["extra"]["e2e_with_synthesis"]["value"],1), [(r["log_n"], round(r["ms"],2)) for r in d["extra"]["msm_sweep"]["sizes"]], "cpu_s", round(h["cpu_s_value"],2), round(h["cpu_s_e2e"],2), "thr", h["after_synthesis"]["nr_throttled"]-h["before"]["nr_throttled"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
}
run default 29531 PB200_X=0
run nosampler 29532 PB200_NO_SAMPLER=1
run spin 29533 PB200_SPIN=1 PB200_NO_SAMPLER=1
cat /proc/loadavg
