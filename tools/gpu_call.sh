#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c16}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
for ep in 1 2 8; do
  BW4_EP=$ep STAGE_TIMEOUT=200 B200MOE_ALIGN_G1=60 timeout 300 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_ep${ep}_g1.log 2>&1
done
for f in fp8 fp8e8m0 bf16 mxfp4; do
  timeout 300 python tools/prefill_bench.py $f 8192 > $O/${TAG}_prefill_$f.json 2> $O/${TAG}_prefill_$f.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_prefill_e8m0.csv -k regex:"route_|gather_rows|moe_gemm|combine_kernel" -c 70 python tools/prefill_bench.py fp8e8m0 8192 > /dev/null 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
# launch list of the default bench command (kernel shares) and one full capture of the dominant kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file $O/${TAG}_launches_bench_default.csv python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moe_fused_kernel -s 200 -c 1 -f -o $O/${TAG}_prof_moe_fused python bench.py --steps 2 --warmup 3 > $O/${TAG}_ncu_moe.log 2>&1
tail -n 5 $O/${TAG}_pytest.log
grep "^M=256" $O/${TAG}_bw4_ep*.log | cut -c1-330
cat $O/${TAG}_prefill_*.json
for f in n1; do echo "== $f"; cut -c1-250 $O/${TAG}_bench_$f.json; tail -n 2 $O/${TAG}_bench_$f.err; done
