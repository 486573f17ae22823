#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c7}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
STAGE_TIMEOUT=200 timeout 500 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_native.log 2>&1
timeout 300 python tools/prefill_bench.py fp8 8192 > $O/${TAG}_prefill_fp8.json 2> $O/${TAG}_prefill_fp8.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moe_gemm_kernel -s 4 -c 2 -f -o $O/${TAG}_prof_prefill python tools/prefill_bench.py fp8 8192 > $O/${TAG}_ncu_prefill.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 600 python tools/mla_vs_ref.py > $O/${TAG}_mla_vs_ref.jsonl 2> $O/${TAG}_mla_vs_ref.err
tail -n 6 $O/${TAG}_pytest.log
grep -h "M=256\|M=64:" $O/${TAG}_bw4_native.log | cut -c1-640
cat $O/${TAG}_prefill_*.json
cut -c1-200 $O/${TAG}_bench_n1.json
cat $O/${TAG}_mla_vs_ref.jsonl; tail -n 3 $O/${TAG}_mla_vs_ref.err
