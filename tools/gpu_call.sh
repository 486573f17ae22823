#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c11}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
timeout 300 python tools/prefill_bench.py fp8 8192 > $O/${TAG}_prefill_fp8.json 2> $O/${TAG}_prefill_fp8.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:moe_gemm_kernel -s 4 -c 2 -f -o $O/${TAG}_prof_prefill python tools/prefill_bench.py fp8 8192 > $O/${TAG}_ncu_prefill.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
timeout 600 python bench.py --workload dsv3-fp8 --ep-shard-of 8 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_shard.json 2> $O/${TAG}_bench_dsv3_shard.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload dsv3-fp8 --layers 6 --steps 20 --warmup 3 > $O/${TAG}_bench_dsv3_n2_l6.json 2> $O/${TAG}_bench_dsv3_n2_l6.err
tail -n 6 $O/${TAG}_pytest.log
cat $O/${TAG}_prefill_*.json
for f in n1 n2 dsv3_shard dsv3_n2_l6; do echo "== $f"; cut -c1-200 $O/${TAG}_bench_$f.json; tail -n 2 $O/${TAG}_bench_$f.err; done
