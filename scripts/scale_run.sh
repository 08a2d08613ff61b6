#!/bin/bash
# N-GPU run of the default bench (config 4, topic-sharded, all-gather inside the step):
#   gpurun --gpus N --timeout 900 -- 'bash scripts/scale_run.sh N'
N=${1:-2}
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/scale_n$N.json
python - <<PY
import json
d=json.load(open("gpurun_out/scale_n$N.json"))
print("N=$N value=%.4e ms/step=%.4f compute=%.4f gather=%.4f e2e=%.3e parity=%s exchange=%s"%(d["value"],d["ms_per_step"],d["compute_ms_per_step"],d["gather_ms_per_step"],d["e2e"]["value"],d["parity_spot_check"],d.get("exchange_ms")))
PY
