#!/bin/bash
# N-GPU sanity of the exchange path with tight limits:  gpurun --gpus 4 --timeout 400 -- 'bash scripts/scale_check4.sh 4'
N=${1:-4}
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -q -x --timeout=140 -k "sharded_nccl and $N" 2>&1 | tail -2
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/err_n$N.txt | grep '^{' | tail -1 > gpurun_out/scale_n$N.json
python - <<PY || tail -5 gpurun_out/err_n$N.txt
import json
d=json.load(open("gpurun_out/scale_n$N.json"))
print("N=$N value=%.4e ms/step=%.4f kernels=%.4f e2e=%.3e parity=%s exchange=%s"%(d["value"],d["ms_per_step"],d["compute_ms_per_step"],d["e2e"]["value"],d["parity_spot_check"],{k:(round(v,4) if isinstance(v,float) else v) for k,v in (d.get("exchange_ms") or {}).items() if k!="note"}))
PY
