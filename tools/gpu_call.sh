#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c20}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err
echo "rc=$?" >> $O/${TAG}_bench_n8.err
for f in n8; do echo "== $f"; cut -c1-300 $O/${TAG}_bench_$f.json; tail -n 4 $O/${TAG}_bench_$f.err; done
