#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c21}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref_default.json 2> $O/${TAG}_bench_ref_default.err
timeout 600 python bench.py --workload mixtral-bf16 --steps 20 --warmup 3 > $O/${TAG}_bench_mixtral_bf16.json 2> $O/${TAG}_bench_mixtral_bf16.err
timeout 600 python bench.py --impl reference --workload mixtral-bf16 --steps 2 --warmup 1 > $O/${TAG}_bench_ref_mixtral_bf16.json 2> $O/${TAG}_bench_ref_mixtral_bf16.err
tail -n 4 $O/${TAG}_pytest.log
tail -n 1 $O/${TAG}_smoke.log
for f in n1 ref_default mixtral_bf16 ref_mixtral_bf16; do echo "== $f"; cut -c1-400 $O/${TAG}_bench_$f.json; tail -n 2 $O/${TAG}_bench_$f.err; done
