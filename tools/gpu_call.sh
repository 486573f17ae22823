#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c18}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
echo "rc=$?" >> $O/${TAG}_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --workload dsv3-fp8 --layers 6 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_n2_l6.json 2> $O/${TAG}_bench_dsv3_n2_l6.err
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "ep_ or two_gpus or multiproc" > $O/${TAG}_pytest_ep.log 2>&1
for f in n2 dsv3_n2_l6; do echo "== $f"; cut -c1-250 $O/${TAG}_bench_$f.json; tail -n 3 $O/${TAG}_bench_$f.err; done
tail -n 3 $O/${TAG}_pytest_ep.log
