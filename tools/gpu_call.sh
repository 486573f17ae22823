#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c9}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "prefill or large_batch or fp8_block or fp8_golden or baseline_width or relu2 or coarse or mla or 16bit or swiglu" > $O/${TAG}_pytest.log 2>&1
timeout 300 python tools/prefill_bench.py fp8 8192 > $O/${TAG}_prefill_fp8.json 2> $O/${TAG}_prefill_fp8.err
timeout 300 python tools/prefill_bench.py bf16 8192 > $O/${TAG}_prefill_bf16.json 2> $O/${TAG}_prefill_bf16.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moe_gemm_kernel -s 4 -c 2 -f -o $O/${TAG}_prof_prefill python tools/prefill_bench.py fp8 8192 > $O/${TAG}_ncu_prefill.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"b200|moe_gemm|route_sort|gather_rows|combine" -s 0 -c 40 --csv --log-file $O/${TAG}_launches_prefill.csv python tools/prefill_bench.py fp8 8192 > $O/${TAG}_ncu_prefill2.log 2>&1
timeout 600 python tools/mla_vs_ref.py > $O/${TAG}_mla_vs_ref.jsonl 2> $O/${TAG}_mla_vs_ref.err
tail -n 4 $O/${TAG}_pytest.log
cat $O/${TAG}_prefill_*.json
cat $O/${TAG}_mla_vs_ref.jsonl | cut -c1-400
