#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c12}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 300 python tools/e8m0_probe.py > $O/${TAG}_e8m0_probe.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "ue8m0 or w4_prefill or w4_pass_loop or large_batch or prefill_8192 or fp8" > $O/${TAG}_pytest.log 2>&1
for f in fp8 fp8e8m0 bf16 mxfp4; do
  timeout 300 python tools/prefill_bench.py $f 8192 > $O/${TAG}_prefill_$f.json 2> $O/${TAG}_prefill_$f.err
done
B200MOE_W4_PREFILL_MIN=0 timeout 300 python tools/prefill_bench.py mxfp4 8192 > $O/${TAG}_prefill_mxfp4_passloop.json 2> $O/${TAG}_prefill_mxfp4_passloop.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moe_gemm_kernel -s 4 -c 2 -f -o $O/${TAG}_prof_prefill_e8m0 python tools/prefill_bench.py fp8e8m0 8192 > $O/${TAG}_ncu_prefill.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${TAG}_launches_prefill_e8m0.csv python tools/prefill_bench.py fp8e8m0 8192 > /dev/null 2>&1
cat $O/${TAG}_e8m0_probe.log | tail -8
tail -n 6 $O/${TAG}_pytest.log
cat $O/${TAG}_prefill_*.json
