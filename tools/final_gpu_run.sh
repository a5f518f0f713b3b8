#!/bin/bash
# The round's closing measurements in one gpurun call (one B200):  gpurun --timeout 600 -- 'bash tools/final_gpu_run.sh'
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_final_tests.txt
timeout 240 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; tail -2 gpurun_out/r02_final_bench.err; cut -c1-400 gpurun_out/r02_final_bench.json
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 80 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_rp1|k_rp2|k_bucket|k_scatter|k_group_fold|k_compact" -s 51 -c 20 -o gpurun_out/r02_final python bench.py --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1
timeout 100 python tools/phase_prof.py 50000 gpurun_out/phase_r02_final.json 2>&1 | tail -16
timeout 100 python bench.py --config 1 --steps 10 --warmup 3 --no-cpu 2>/dev/null > gpurun_out/r02_config1.json; cut -c1-300 gpurun_out/r02_config1.json
timeout 150 python tools/xdrop_vs_logan.py gpurun_out/xdrop_vs_logan_r02.json 2>&1 | tail -2
