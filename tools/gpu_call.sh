#!/bin/bash
# one GPU box call: full GPU test suite, smoke, default bench line (edit per call; outputs land in gpurun_out/r2/)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-final}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -n 4 $O/${TAG}_pytest.log
tail -n 1 $O/${TAG}_smoke.log
cut -c1-300 $O/${TAG}_bench_n1.json; tail -n 2 $O/${TAG}_bench_n1.err
