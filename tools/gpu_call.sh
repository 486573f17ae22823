#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c6}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
STAGE_TIMEOUT=200 timeout 500 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_native.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
timeout 300 python tools/prefill_bench.py fp8 8192 > $O/${TAG}_prefill_fp8.json 2> $O/${TAG}_prefill_fp8.err
timeout 300 python tools/prefill_bench.py bf16 8192 > $O/${TAG}_prefill_bf16.json 2> $O/${TAG}_prefill_bf16.err
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/${TAG}_bench_n2.json 2> $O/${TAG}_bench_n2.err
tail -n 6 $O/${TAG}_pytest.log
grep -h "M=256\|M=16:" $O/${TAG}_bw4_native.log | cut -c1-640
cat $O/${TAG}_prefill_*.json
for f in n1 n2; do echo "== $f"; cut -c1-200 $O/${TAG}_bench_$f.json; tail -n 2 $O/${TAG}_bench_$f.err; done
