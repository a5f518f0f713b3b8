#!/bin/bash
# Everything that was written after round 1's GPU time was spent, in ONE gpurun call (about 4 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/first_gpu_run.sh'
# 1. the never-run GPU tests (non-strict xfail: read the X / x marks), 2. every X-drop shape + the reference's LOGAN on the
# bench batch, 3. the k-mer row against the reference's SplitCount, 4. ncu of the two thread kernels and the k-mer pipeline.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_xdrop_shim_gpu.py tests/test_zz_unvalidated_gpu.py -m gpu -q -rxX 2>&1 | tail -25 | tee gpurun_out/first_unvalidated_tests.txt
BELLA_RUN_LOGAN=1 timeout 60 python -m pytest tests/test_xdrop_gpu.py -m gpu -q -k logan -s 2>&1 | tail -5 | tee gpurun_out/first_logan_test.txt
timeout 90 python tools/xdrop_prof.py --logan 2>&1 | tail -20 | tee gpurun_out/first_xdrop_shapes.txt
timeout 90 python tools/kmers_bench.py 20000 10000 gpurun_out/kmers_bench.json 2>&1 | tail -3
timeout 60 ncu --set full --import-source on --clock-control none -k regex:k_xdrop_thread -c 1 -o gpurun_out/xdrop_thread64 python tools/xdrop_prof.py --prof 1 64 2>&1 | tail -2
timeout 60 ncu --set full --import-source on --clock-control none -k regex:k_xdrop_thread_packed -c 1 -o gpurun_out/xdrop_packed64 python tools/xdrop_prof.py --prof 3 64 2>&1 | tail -2
timeout 60 ncu --set full --import-source on --clock-control none -k regex:k_xdrop_thread_two -c 1 -o gpurun_out/xdrop_two64 python tools/xdrop_prof.py --prof 5 64 2>&1 | tail -2
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/kmers_launches.csv python tools/kmers_bench.py 4000 10000 gpurun_out/kmers_ncu.json 2>&1 | tail -2
