"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/b200moe.h
declares, the lk_moe surface matches the reference call site, the LVLLM_* scheduler predicates, and that
the product path fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200moe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200(?:moe)?_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from lvllm_b200 import build, _lib
    path = build.build_lib()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200moe.h but not exported"
    # python prototypes cover exactly the declared symbols
    assert sorted(_lib.PROTOTYPES) == declared
    assert _lib.lib().b200moe_version().startswith(b"b200moe")


def test_config_struct_matches_header():
    from lvllm_b200 import _lib
    src = open(os.path.join(ROOT, "include", "b200moe.h")).read()
    body = src[src.index("typedef struct b200moe_config {"):src.index("} b200moe_config;")]
    fields = re.findall(r"(?:int32_t|float)\s+([a-zA-Z_]+);", body)
    assert fields == [f for f, _ in _lib.B200Config._fields_]
    assert ctypes.sizeof(_lib.B200Config) == 4 * len(fields)


def test_lk_moe_surface_matches_reference_call_site():
    import inspect

    import lk_moe
    names = ["MOE_BF16", "MOE_FP16", "MOE_FP8", "MOE_FP8_FP16", "MOE_WNA16", "MOE_WNA16_FP16", "MOE_NVFP4",
             "MOE_NVFP4_FP16", "MOE_MXFP4", "MOE_MXFP4_FP16"]
    for n in names:
        assert hasattr(lk_moe, n)
    cfg = lk_moe.MOEConfigV2()
    # every attribute the reference sets (routed_experts.py:1490-1511) exists
    for f in ["num_processes", "process_id", "gpu_id", "has_gate_proj", "expert_num", "top_k", "hidden_size",
              "intermediate_size", "max_batch_size", "max_num_seqs", "stride", "group_min_len", "group_max_len",
              "groupN", "groupK", "activation_type", "swiglu_alpha", "swiglu_limit", "use_gpu_prefill"]:
        assert hasattr(cfg, f)
    base = lk_moe.MOE_BF16
    assert list(inspect.signature(base.cpu_decode).parameters)[1:] == [
        "stream_ptr", "num_tokens", "top_k", "hidden_ptr", "topk_ids_ptr", "topk_weights_ptr", "out_f32_ptr"]
    assert list(inspect.signature(base.cpu_prefill).parameters)[1:] == [
        "num_tokens", "top_k", "ids_host_ptr", "weights_host_ptr", "hidden_host_ptr", "out_f32_host_ptr"]
    assert list(inspect.signature(base.gpu_prefill).parameters)[1:] == [
        "hidden_ptr", "out_ptr", "topk_ids_ptr", "topk_weights_ptr", "num_tokens", "top_k", "stream_ptr"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    import lk_moe
    from lvllm_b200 import ops
    from lvllm_b200._lib import B200Error
    cfg = lk_moe.MOEConfigV2()
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = 4, 2, 256, 128
    w = torch.zeros(4, 256, 256, dtype=torch.bfloat16)
    with pytest.raises(B200Error):
        lk_moe.MOE_BF16(cfg, w.data_ptr(), w.data_ptr(), 0, 0, 0, 0)
    with pytest.raises(ValueError):
        ops.fused_topk(torch.randn(4, 8), 2, True)


def test_lvllm_scheduler_predicates(monkeypatch):
    from lvllm_b200 import envs
    monkeypatch.delenv("LVLLM_MOE_NUMA_ENABLED", raising=False)
    assert envs.is_lk_moe_gpu_resident_layer("model.layers.3.mlp.experts")  # feature off -> everything resident
    monkeypatch.setenv("LVLLM_MOE_NUMA_ENABLED", "1")
    monkeypatch.setenv("LVLLM_GPU_RESIDENT_MOE_LAYERS", "0-1, 33-34,x,7")
    assert envs.parse_layer_list("0-1, 33-34,x,7") == {0, 1, 33, 34, 7}
    assert envs.is_lk_moe_gpu_resident_layer("model.layers.1.mlp.experts")
    assert not envs.is_lk_moe_gpu_resident_layer("model.layers.2.mlp.experts")
    assert envs.is_lk_moe_gpu_resident_layer("mtp.layers.2.mlp.experts".replace("mtp.layers.2", "mtp.0"))
    assert envs.is_lk_moe_cpu_layer("model.layers.2.mlp.experts")
    assert envs.select_entry_point("model.layers.2.mlp.experts", 4, capturing=True) == "cpu_decode"
    assert envs.select_entry_point("model.layers.2.mlp.experts", 4, capturing=False) == "cpu_prefill"
    monkeypatch.setenv("LVLLM_GPU_PREFILL_MIN_BATCH_SIZE", "512")
    assert envs.is_lk_moe_gpu_prefill_layer("model.layers.2.mlp.experts")
    assert envs.select_entry_point("model.layers.2.mlp.experts", 1024, capturing=False) == "gpu_prefill"
    assert envs.select_entry_point("model.layers.2.mlp.experts", 16, capturing=False) == "cpu_prefill"
    assert envs.select_entry_point("model.layers.1.mlp.experts", 16, capturing=False) == "resident"
    old = envs.disable_lk_moe_gpu_prefill()
    assert old == 512 and not envs.is_lk_moe_use_gpu_prefill()
    envs.enable_lk_moe_gpu_prefill(old)
    assert envs.get_gpu_prefill_min_batch_size() == 512
    envs._overrides.clear()
    assert envs.get_gpu_prefetch_window() == 3
    assert envs.cuda_graph_sizes(24) == [1, 2, 4, 8, 16, 24]


def test_ep_a2a_layout_is_host_side_and_aligned():
    """b200_ep_a2a_layout is pure host arithmetic (no device needed): regions are 256-byte aligned, ordered and sized
    for X [M][H] 16-bit, IDS / W [M][k] 32-bit, Y [M][H] f32 (include/b200moe.h)."""
    from lvllm_b200 import _lib
    lib = _lib.lib()
    for M, H, k in [(256, 4096, 8), (8, 7168, 8), (3, 1000, 5)]:
        offs = [ctypes.c_int64() for _ in range(4)]
        total = lib.b200_ep_a2a_layout(M, H, k, *[ctypes.byref(o) for o in offs])
        x, ids, w, y = (o.value for o in offs)
        assert x == 0 and all(v % 256 == 0 for v in (ids, w, y, total))
        assert ids - x >= M * H * 2 and w - ids >= M * k * 4 and y - w >= M * k * 4 and total - y >= M * H * 4
    assert lib.b200_ep_a2a_layout(0, 4096, 8, None, None, None, None) == 0


def test_lvllm_predicates_match_reference_table(golden, monkeypatch):
    """lvllm_b200.envs against a table produced by the reference's own vllm/envs.py predicates (:2292-2410) under six
    environment configurations x eight layer names (tests/golden/make_golden.py)."""
    from lvllm_b200 import envs
    keys = ["LVLLM_MOE_NUMA_ENABLED", "LVLLM_GPU_RESIDENT_MOE_LAYERS", "LVLLM_GPU_PREFILL_MIN_BATCH_SIZE",
            "LVLLM_GPU_PREFETCH_WINDOW", "LVLLM_ENABLE_MOE_LAYERWISE_LOAD"]
    for row in golden["lvllm_env_predicates"]:
        envs._overrides.clear()
        for k in keys:
            monkeypatch.delenv(k, raising=False)
        for k, v in row["env"].items():
            monkeypatch.setenv(k, v)
        assert envs.is_lk_moe_feature_enabled() == row["feature"]
        assert envs.is_lk_moe_use_gpu_prefill() == row["use_gpu_prefill"]
        assert envs.get_gpu_prefill_min_batch_size() == row["min_batch"]
        assert envs.get_gpu_prefetch_window() == row["window"]
        for L in row["layers"]:
            nm = L["name"]
            assert envs.is_lk_moe_mtp_layer(nm) == L["mtp"], (row["env"], nm)
            assert envs.is_lk_moe_gpu_resident_layer(nm) == L["resident"], (row["env"], nm)
            assert envs.is_lk_moe_gpu_prefill_layer(nm) == L["gpu_prefill"], (row["env"], nm)
            assert envs.is_lk_moe_cpu_layer(nm) == L["cpu"], (row["env"], nm)


def test_library_sass_is_tcgen05_tma_and_no_legacy_mma():
    """The built library's hot kernels issue 5th-generation tensor-core MMAs (UTCHMMA / UTCQMMA), tensor-map TMA (UTMALDG)
    and bulk async copies (UBLKCP), read accumulators from TMEM (LDTM), and contain no warp-level mma.sync (HMMA / IMMA):
    counted from `cuobjdump -sass` by tools/sass_evidence.py (the committed copy is profiles/r02_sass_evidence.txt)."""
    import shutil
    import subprocess
    import sys
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not available")
    from lvllm_b200 import build
    build.build_lib()
    env = dict(os.environ, PATH=os.environ.get("PATH", "") + ":/usr/local/cuda/bin")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_evidence.py")], stdout=subprocess.PIPE, text=True,
                         check=True, env=env).stdout
    per = {}
    for ln in out.splitlines():
        if ln.startswith("#") or "{" not in ln:
            continue
        name, counts = ln.split("{", 1)
        per[name.strip()] = eval("{" + counts)   # noqa: S307 — our own tool's dict repr
    tot = {}
    for c in per.values():
        for k_, v in c.items():
            tot[k_] = tot.get(k_, 0) + v
    assert tot.get("HMMA", 0) + tot.get("IMMA", 0) + tot.get("QMMA", 0) == 0
    assert tot.get("UTCHMMA", 0) >= 100 and tot.get("UTCQMMA", 0) >= 50 and tot.get("UTMALDG", 0) >= 3 and tot.get("LDTM", 0) >= 100
    hot = [n for n in per if any(s in n for s in ("moe_fused_kernel", "moe_gemm_kernel", "router_gemm_topk_kernel",
                                                  "mla_decode_tc_kernel", "gqa_decode_tc_kernel"))]
    assert len(hot) >= 40
    for n in hot:
        assert per[n].get("UTCHMMA", 0) + per[n].get("UTCQMMA", 0) > 0, f"{n} issues no tcgen05 MMA"
