#!/bin/bash
# N=1,2,4,8 scaling run of bench.py (same launch lines as the driver)
python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/s1.err | tail -1 > gpurun_out/scale_n1.json
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/s$N.err
done
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/scale_n{N}.json").read().strip().splitlines()[-1])
        print(N, round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "M nnz/s", "e2e", round(d["e2e"]["ms_per_step"],2), d["roofline"].get("rank0_phase_ms", d["roofline"].get("phase_ms")))
    except Exception as e:
        print(N, "failed", e)
PY
