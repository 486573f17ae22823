"""bench.py's bookkeeping that needs no GPU: the algorithmic bytes per expert behind `roofline.achieved` against SURVEY.md 8(d)'s
figures, the one-workload-for-every-N rule, the config block (no model keys missing, workload named) and the committed ncu
traffic file the `roofline.traffic` field is read from."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_per_expert_match_survey(bench):
    # DSV3-FP8 44.051 MB, DSV3-NVFP4 24.77 MB, Mixtral-bf16 352.3 MB, Mixtral-int4-g32 99.09 MB, Qwen3-MXFP4 10.03 MB
    want = {"dsv3-fp8": 44.051e6, "dsv3-nvfp4": 24.77e6, "mixtral-bf16": 352.3e6, "mixtral-int4": 99.09e6, "qwen3-mxfp4": 10.03e6}
    for name, b in want.items():
        got = bench.bytes_per_expert(bench.WORKLOADS[name])
        assert abs(got - b) / b < 2e-3, (name, got, b)
    # decode bytes per token of the metric's configuration: k * bytes/expert * layers = 20.44 GB
    w = bench.WORKLOADS["dsv3-fp8"]
    assert abs(w["k"] * bench.bytes_per_expert(w) * w["layers"] - 20.44e9) / 20.44e9 < 2e-3
    # config 5 at batch 256 touches (nearly) all 128 experts: 1.283 GB per layer, 120.6 GB per step
    w = bench.WORKLOADS["qwen3-mxfp4"]
    assert abs(w["E"] * bench.bytes_per_expert(w) - 1.2835e9) / 1.2835e9 < 2e-3
    assert abs(w["E"] * bench.bytes_per_expert(w) * w["layers"] - 120.6e9) / 120.6e9 < 2e-3


def test_one_workload_for_every_n_and_config_block(bench):
    names = {bench.default_workload(n) for n in (1, 2, 4, 8)}
    assert names == {"qwen3-mxfp4"}                      # the driver's 1/2/4/8 runs form a curve
    for n in (1, 2, 8):
        cfg = bench._config("qwen3-mxfp4", dict(bench.WORKLOADS["qwen3-mxfp4"]), n)
        assert cfg["workload"] == "qwen3-mxfp4" and cfg["batch"] == 256 and cfg["experts"] == 128 and cfg["top_k"] == 8
        assert "l2" in cfg                                # says why no L2 flush is needed between timed iterations
        json.dumps(cfg)
    assert bench.METRIC.startswith("decode tok/s DeepSeek-V3 FP8")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "DeepSeek-V3" in json.dumps(base)              # the metric's configuration is the one BASELINE.json names


def test_committed_ncu_traffic_is_consistent_with_the_algorithmic_bytes(bench):
    t = bench.ncu_traffic("qwen3-mxfp4", 1)
    assert t is not None
    algo = bench.WORKLOADS["qwen3-mxfp4"]["E"] * bench.bytes_per_expert(bench.WORKLOADS["qwen3-mxfp4"])
    assert 1.0 <= t / algo < 1.15                          # DRAM traffic of one launch: no wasted re-reads (1.05 x)
    tp = bench.ncu_tensor_pipe()
    assert isinstance(tp, dict) and any("prefill" in k_ for k_ in tp)
