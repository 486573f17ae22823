#!/bin/bash
# one gpurun call = several measurements (each under its own timeout; logs under gpurun_out/r2/)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2
mkdir -p $O
TAG=${1:-c1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.log 2>&1
( timeout 60 tools/tma_probe tma 16384; timeout 60 tools/tma_probe tma 32768
  for v in "1 8" "8 1" "0 8" "1 1" "8 8" "0 1"; do timeout 60 tools/tma_probe cp $v; done ) > $O/${TAG}_tma_probe.log 2>&1
TX=16384
if grep -q "expect_tx 32768: barrier completed = 1" $O/${TAG}_tma_probe.log; then TX=32768; fi
echo "chosen MX tx bytes: $TX" >> $O/${TAG}_tma_probe.log
export B200MOE_MX_TX=$TX
STAGE_TIMEOUT=200 timeout 300 python tools/gpu_bringup.py bw4 > $O/${TAG}_bw4_dequant.log 2>&1
B200MOE_MX_NATIVE=1 STAGE_TIMEOUT=200 timeout 500 python tools/gpu_bringup.py mx bw4 > $O/${TAG}_bw4_native.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $O/${TAG}_pytest.log 2>&1
B200MOE_MX_NATIVE=1 B200MOE_TEST_MX_NATIVE=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "mxfp4 or w4_pass or fp16_classes" > $O/${TAG}_pytest_native.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
B200MOE_MX_NATIVE=1 timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1_native.json 2> $O/${TAG}_bench_n1_native.err
tail -5 $O/${TAG}_tma_probe.log $O/${TAG}_pytest.log $O/${TAG}_pytest_native.log
grep -h "M=256\|M=16:" $O/${TAG}_bw4_dequant.log $O/${TAG}_bw4_native.log | cut -c1-400
cat $O/${TAG}_bench_n1.json | cut -c1-600
