#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c23}
timeout 400 python -m pytest tests -m gpu -q --timeout 120 -x -k "ue8m0" > $O/${TAG}_pytest.log 2>&1
B200MOE_E8M0_CLUSTER=1 timeout 120 python tools/prefill_bench.py fp8e8m0 8192 > $O/${TAG}_prefill_e8m0_cluster.json 2> $O/${TAG}_prefill_e8m0_cluster.err
timeout 120 python tools/prefill_bench.py fp8e8m0 8192 > $O/${TAG}_prefill_e8m0.json 2> $O/${TAG}_prefill_e8m0.err
tail -n 15 $O/${TAG}_pytest.log
cat $O/${TAG}_prefill_e8m0_cluster.json $O/${TAG}_prefill_e8m0.json
tail -n 3 $O/${TAG}_prefill_e8m0_cluster.err
