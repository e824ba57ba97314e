#!/bin/bash
# e2e host-wait modes on the 8-GPU box (32 host cores): spin / yield / block, C2, 8 ranks
o=gpurun_out/$1; mkdir -p $o
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
for m in yield block spin; do
  ECB_BENCH_SYNC=$m $R --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > $o/c2_n8_$m.json 2> $o/c2_n8_$m.err
  python - <<PY
import json
try:
    d=json.loads(open("$o/c2_n8_$m.json").read().strip().splitlines()[-1]); e=d["e2e"]
    print("$m", "step %.3f"%d["ms_per_step"], "e2e_ms %.3f"%e["ms_per_step"], "frac_of_host_ceiling", e.get("frac_of_host_ceiling"), e["pipeline"][:20])
except Exception as ex: print("$m ERR", ex)
PY
done
