#!/bin/bash
# one gpurun call = several measurements (each under its own timeout; logs under gpurun_out/r2/)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c5}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "prefill or mla_ or large_batch" > $O/${TAG}_pytest.log 2>&1
STAGE_TIMEOUT=200 timeout 500 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_native.log 2>&1
timeout 300 python tools/prefill_bench.py fp8 8192 > $O/${TAG}_prefill_fp8.json 2> $O/${TAG}_prefill_fp8.err
timeout 300 python tools/prefill_bench.py bf16 8192 > $O/${TAG}_prefill_bf16.json 2> $O/${TAG}_prefill_bf16.err
timeout 300 python tools/prefill_bench.py fp8 2048 > $O/${TAG}_prefill_fp8_2048.json 2> $O/${TAG}_prefill_fp8_2048.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moe_gemm_kernel -s 4 -c 2 -f -o $O/${TAG}_prof_prefill python tools/prefill_bench.py fp8 8192 > $O/${TAG}_ncu_prefill.log 2>&1
STAGE_TIMEOUT=400 timeout 600 ncu --set full --clock-control none --import-source on -k regex:moe_fused_kernel -s 20 -c 1 -f -o $O/${TAG}_prof_moe_native python tools/gpu_bringup.py --child bw4 > $O/${TAG}_ncu_moe.log 2>&1
tail -n 6 $O/${TAG}_pytest.log
grep -h "M=256\|M=16:" $O/${TAG}_bw4_native.log | cut -c1-560
cat $O/${TAG}_prefill_*.json
tail -n 3 $O/${TAG}_ncu_prefill.log $O/${TAG}_ncu_moe.log
