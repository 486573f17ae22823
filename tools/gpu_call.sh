#!/bin/bash
# one gpurun call = several measurements (each under its own timeout; logs under gpurun_out/r2/)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c4}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
STAGE_TIMEOUT=200 timeout 500 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_native.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
timeout 600 python bench.py --workload dsv3-fp8 --ep-shard-of 8 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_shard.json 2> $O/${TAG}_bench_dsv3_shard.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload dsv3-fp8 --layers 6 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_n2_l6.json 2> $O/${TAG}_bench_dsv3_n2_l6.err
tail -n 6 $O/${TAG}_pytest.log
grep -h "M=256\|M=16:\|M=1:" $O/${TAG}_bw4_native.log | cut -c1-460
for f in n1 n2 dsv3_shard dsv3_n2_l6; do echo "== $f"; cut -c1-200 $O/${TAG}_bench_$f.json; tail -n 2 $O/${TAG}_bench_$f.err; done
