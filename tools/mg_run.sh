#!/bin/bash
# usage: tools/mg_run.sh MODE N [N ...]   -- bench.py at N GPUs (same launch line as the driver) + the per-stage profile at the last N
mode=$1; shift
for N in "$@"; do
  BELLA_MG_MODE=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg_${mode}_n$N.json 2> gpurun_out/mg_${mode}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/mg_${mode}_n$N.json").read().strip().splitlines()[-1])
    print("$mode N=$N", round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "M nnz/s", "e2e", round(d["e2e"]["ms_per_step"],2), d["roofline"].get("rank0_phase_ms"), d.get("parity"))
except Exception as e:
    print("$mode N=$N failed", e); print(open("gpurun_out/mg_${mode}_n$N.err").read()[-1500:])
PY
done
BELLA_MG_PROFILE=1 BELLA_MG_MODE=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29599 tools/mgprof.py 2>&1 | grep "^[0-9]" 
