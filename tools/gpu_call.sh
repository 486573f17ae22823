#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c17}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
for ep in 1 2 4 8; do
  BW4_EP=$ep STAGE_TIMEOUT=200 timeout 300 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_ep${ep}.log 2>&1
done
STAGE_TIMEOUT=200 timeout 300 python tools/gpu_bringup.py bw > $O/${TAG}_bw_fp8.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 600 python bench.py --workload dsv3-fp8 --ep-shard-of 8 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_shard.json 2> $O/${TAG}_bench_dsv3_shard.err
timeout 300 python __graft_entry__.py > $O/${TAG}_smoke.log 2>&1
tail -n 5 $O/${TAG}_pytest.log
grep "^M=" $O/${TAG}_bw4_ep*.log | cut -c1-120
grep "^M=" $O/${TAG}_bw_fp8.log | cut -c1-200
for f in n1 dsv3_shard; do echo "== $f"; cut -c1-250 $O/${TAG}_bench_$f.json; tail -n 2 $O/${TAG}_bench_$f.err; done
tail -n 3 $O/${TAG}_smoke.log
