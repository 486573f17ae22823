#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c13}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err
echo "rc=$?" >> $O/${TAG}_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 10 --warmup 3 --no-sub > $O/${TAG}_bench_n4.json 2> $O/${TAG}_bench_n4.err
echo "rc=$?" >> $O/${TAG}_bench_n4.err
for f in n8 n4; do echo "== $f"; cut -c1-1500 $O/${TAG}_bench_$f.json; tail -n 4 $O/${TAG}_bench_$f.err; done
