#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c25}
timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x -k "large_batch or w4_prefill or prefill_8192 or multi_cta" > $O/${TAG}_pytest.log 2>&1
timeout 100 python tools/prefill_bench.py bf16 8192 > $O/${TAG}_prefill_bf16_pair.json 2> $O/${TAG}_prefill_bf16_pair.err
B200MOE_GEMM_PAIR=0 timeout 100 python tools/prefill_bench.py bf16 8192 > $O/${TAG}_prefill_bf16_nopair.json 2> $O/${TAG}_prefill_bf16_nopair.err
timeout 100 python tools/prefill_bench.py mxfp4 8192 > $O/${TAG}_prefill_mxfp4_pair.json 2> $O/${TAG}_prefill_mxfp4_pair.err
tail -n 12 $O/${TAG}_pytest.log
cat $O/${TAG}_prefill_bf16_pair.json $O/${TAG}_prefill_bf16_nopair.json $O/${TAG}_prefill_mxfp4_pair.json
tail -n 2 $O/${TAG}_prefill_bf16_pair.err
