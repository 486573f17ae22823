"""lvllm_b200/plan.py against the sizes SURVEY.md 8 states for the BASELINE models (row a8 / 8d)."""
import pytest

from lvllm_b200 import plan as P


def test_expert_bytes_match_survey():
    for fmt, H, I, mb in (("fp8", 7168, 2048, 44.051), ("nvfp4", 7168, 2048, 24.77), ("bf16", 4096, 14336, 352.3),
                          ("wna16", 4096, 14336, 99.09), ("mxfp4", 4096, 1536, 10.03)):
        assert abs(P.expert_bytes(fmt, H, I) / 1e6 - mb) / mb < 2e-3, fmt
    with pytest.raises(ValueError):
        P.expert_bytes("gguf", 128, 128)


def test_baseline_deployments():
    # DeepSeek-V3 FP8: 654 GB of routed experts, 32 per GPU and 81.8 GB per GPU at EP8; needs >= 8 GPUs
    p8 = P.plan_experts("fp8", 256, 7168, 2048, 58, ep_size=8)
    assert p8.experts_per_rank == 32 and abs(p8.bytes_per_layer_per_rank / 1e9 - 1.41) < 0.01
    assert abs(p8.bytes_per_rank / 1e9 - 81.8) < 0.2 and abs(p8.bytes_total / 1e9 - 654) < 1 and p8.fits
    assert not P.plan_experts("fp8", 256, 7168, 2048, 58, ep_size=4, budget_frac=0.85).fits      # 163.5 GB on 180 GB
    assert P.min_ep_size("fp8", 256, 7168, 2048, 58) == 8
    # with the shared expert riding in the routed launch: 33 local experts
    assert P.plan_experts("fp8", 256, 7168, 2048, 58, ep_size=8, n_shared_experts=1).experts_per_rank == 33
    # Qwen3-235B MXFP4: 1.28 GB per layer, ~120.6 GB: the largest BASELINE configuration that fits ONE B200
    q = P.plan_experts("mxfp4", 128, 4096, 1536, 94)
    assert abs(q.bytes_per_layer_per_rank / 1e9 - 1.2835) < 0.002 and abs(q.bytes_per_rank / 1e9 - 120.6) < 0.2 and q.fits
    assert P.min_ep_size("mxfp4", 128, 4096, 1536, 94) == 1
    # Mixtral bf16 (90 GB) fits one GPU; TP2 halves the per-rank intermediate size
    m = P.plan_experts("bf16", 8, 4096, 14336, 32)
    assert abs(m.bytes_per_rank / 1e9 - 90.2) < 0.3 and m.fits
    m2 = P.plan_experts("bf16", 8, 4096, 14336, 32, tp_size=2)
    assert m2.intermediate_per_rank == 7168 and abs(m2.bytes_per_rank * 2 - m.bytes_per_rank) < 1e6
    # uneven expert counts: the fullest rank of the linear map
    assert P.plan_experts("bf16", 10, 1024, 1024, 1, ep_size=4).experts_per_rank == 3
    with pytest.raises(ValueError):
        P.plan_experts("bf16", 8, 4096, 14336, 32, tp_size=3)
    with pytest.raises(ValueError):
        P.plan_experts("bf16", 8, 4096, 1000, 1)
