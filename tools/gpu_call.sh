#!/bin/bash
# One GPU call of the validation form used through round 2 (gpurun -- 'bash tools/gpu_call.sh cNN'): the full GPU suite, the
# default bench line and the opt-in variants' parity + timing.  Results land in gpurun_out/r2/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c28}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 120 python tools/variant_check.py $O/${TAG}_combine_check.jsonl B200MOE_COMBINE 1 2 > $O/${TAG}_combine_check.log 2>&1
timeout 120 python tools/variant_check.py $O/${TAG}_pair_check.jsonl B200MOE_GEMM_PAIR 0 1 > $O/${TAG}_pair_check.log 2>&1
tail -n 4 $O/${TAG}_pytest.log
tail -n 1 $O/${TAG}_bench_n1.json | cut -c1-400
tail -n 6 $O/${TAG}_combine_check.log
