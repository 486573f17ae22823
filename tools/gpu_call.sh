#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c22}
cat /sys/fs/cgroup/cpu.max > $O/${TAG}_cpu.log 2>&1; nproc >> $O/${TAG}_cpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 > $O/${TAG}_bench_n4.json 2> $O/${TAG}_bench_n4.err
echo "rc=$?" >> $O/${TAG}_bench_n4.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > $O/${TAG}_bench_ref_n4.json 2> $O/${TAG}_bench_ref_n4.err
echo "rc=$?" >> $O/${TAG}_bench_ref_n4.err
cat $O/${TAG}_cpu.log
for f in n4 ref_n4; do echo "== $f"; cut -c1-300 $O/${TAG}_bench_$f.json; tail -n 3 $O/${TAG}_bench_$f.err; done
