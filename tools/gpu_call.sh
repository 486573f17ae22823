#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c14}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
for ep in 2 8; do
  BW4_EP=$ep STAGE_TIMEOUT=200 timeout 300 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_ep$ep.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "ue8m0 or large_batch or prefill_8192 or w4_prefill" > $O/${TAG}_pytest.log 2>&1
timeout 600 python bench.py --workload dsv3-fp8 --ep-shard-of 8 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_tpshard.json 2> $O/${TAG}_bench_dsv3_tpshard.err
BENCH_EP_ONLY=1 timeout 600 python bench.py --workload dsv3-fp8 --ep-shard-of 8 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_epshard.json 2> $O/${TAG}_bench_dsv3_epshard.err
for f in fp8 fp8e8m0; do
  timeout 300 python tools/prefill_bench.py $f 8192 > $O/${TAG}_prefill_$f.json 2> $O/${TAG}_prefill_$f.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_prefill_e8m0.csv -k regex:"route_sort|gather_rows|moe_gemm|combine_kernel" -c 40 python tools/prefill_bench.py fp8e8m0 8192 > /dev/null 2>&1
grep "^M=256" $O/${TAG}_bw4_ep*.log | cut -c1-700
tail -n 4 $O/${TAG}_pytest.log
for f in tpshard epshard; do echo "== $f"; python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_dsv3_$f.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["breakdown_us_per_layer_eager"], d["roofline"]["avg_launch_ms"], d["config"]["parallelism"])
PY
tail -n 2 $O/${TAG}_bench_dsv3_$f.err; done
cat $O/${TAG}_prefill_*.json
