#!/bin/bash
# one gpurun call = several measurements (each under its own timeout; logs under gpurun_out/r2/)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
STAGE_TIMEOUT=200 timeout 500 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_native.log 2>&1
B200MOE_MX_NATIVE=0 STAGE_TIMEOUT=200 timeout 300 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_dequant.log 2>&1
STAGE_TIMEOUT=200 timeout 300 python tools/gpu_bringup.py bw > $O/${TAG}_bw_fp8.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -n 8 $O/${TAG}_pytest.log
grep -h "M=256\|M=16:\|M=1:" $O/${TAG}_bw4_native.log $O/${TAG}_bw4_dequant.log | cut -c1-420
tail -n 12 $O/${TAG}_bw_fp8.log | cut -c1-300
cat $O/${TAG}_bench_n1.json | cut -c1-400
