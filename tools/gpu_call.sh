#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c19}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "ep_shard_shapes or bench_shape" > $O/${TAG}_pytest.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1
tail -n 12 $O/${TAG}_pytest.log
tail -n 2 $O/${TAG}_smoke.log
