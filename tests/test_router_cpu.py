"""lvllm_b200/router.py: the select_experts TEMPLATE — routing -> capture of the logical ids -> index
dtype — with the device operators replaced by CPU stand-ins built on the oracle, so that the control flow is checked without a
GPU (the operators themselves are parity-tested on the GPU: test_topk_*, test_grouped_topk_*, test_router_fused_vs_oracle)."""
import pytest
import torch

from lvllm_b200 import ops, router as R
from oracle import moe_oracle as O


@pytest.fixture()
def cpu_ops(monkeypatch):
    calls = []

    def fused_topk(g, topk, renorm, scoring="softmax", bias=None, rsf=1.0, return_token_expert_indices=False):
        calls.append("fused_topk")
        w, ids = O.topk_gating(g.float(), topk, renorm, scoring, bias)
        return w * rsf, ids

    def grouped_topk(g, topk, renorm, ng, tg, scoring="sigmoid", rsf=1.0, bias=None):
        calls.append("grouped_topk")
        return O.grouped_topk(g.float(), bias, ng, tg, topk, renorm, rsf)

    def g2l(ids, emap):
        calls.append("g2l")
        return O.global_to_local_expert_ids(ids, emap)

    def router_topk(h, wg, topk, renorm, scoring="softmax", bias=None, rsf=1.0, ng=0, tg=0, emap=None, return_logits=False,
                    n_shared=0, shared_base=-1, shared_w=1.0):
        calls.append("router_topk")
        logits = h.float() @ wg.float().t()
        if ng > 0:
            w, ids = O.grouped_topk(logits, bias, ng, tg, topk, renorm, rsf)
        else:
            w, ids = O.topk_gating(logits, topk, renorm, scoring, bias)
            w = w * rsf
        loc = O.global_to_local_expert_ids(ids, emap) if emap is not None else None
        if n_shared:
            M = ids.shape[0]
            sh = torch.arange(n_shared, dtype=torch.int32).expand(M, n_shared)
            ids = torch.cat([ids, wg.shape[0] + sh], 1)
            w = torch.cat([w, torch.full((M, n_shared), shared_w)], 1)
            loc = torch.cat([loc if loc is not None else ids[:, :topk], shared_base + sh], 1)
        return w, ids, loc

    monkeypatch.setattr(ops, "fused_topk", fused_topk)
    monkeypatch.setattr(ops, "grouped_topk", grouped_topk)
    monkeypatch.setattr(ops, "global_to_local_expert_ids", g2l)
    monkeypatch.setattr(ops, "router_topk", router_topk)
    return calls


def test_template_logits_form(cpu_ops):
    g = torch.Generator().manual_seed(0)
    M, E, k = 6, 16, 4
    logits = torch.randn(M, E, generator=g)
    hid = torch.randn(M, 32, generator=g).bfloat16()
    seen = []
    _, emap = O.determine_expert_map(2, 1, E)
    r = R.Router(k, E, renormalize=True, expert_map=emap, capture_fn=lambda ids: seen.append(ids.clone()))
    w, ids = r.select_experts(hid, logits, topk_indices_dtype=torch.int64)
    w_ref, ids_ref = O.topk_gating(logits, k, True, "softmax", None)
    assert cpu_ops == ["fused_topk", "g2l"]
    assert ids.dtype == torch.int64 and torch.equal(ids.int(), ids_ref) and torch.allclose(w, w_ref)
    assert len(seen) == 1 and seen[0].dtype == torch.int32 and torch.equal(seen[0], ids_ref)     # captured before the dtype step
    assert torch.equal(r.last_local_ids, O.global_to_local_expert_ids(ids_ref, emap))
    # DeepSeek-style grouped routing with bias and scaling factor
    cpu_ops.clear()
    bias = torch.randn(E, generator=g)
    r2 = R.Router(k, E, True, "sigmoid", num_expert_group=4, topk_group=2, routed_scaling_factor=2.5, e_score_correction_bias=bias)
    w2, ids2 = r2.select_experts(hid, logits)
    w2r, ids2r = O.grouped_topk(logits, bias, 4, 2, k, True, 2.5)
    assert cpu_ops == ["grouped_topk"] and ids2.dtype == torch.int32 and torch.equal(ids2, ids2r) and torch.allclose(w2, w2r)
    assert r2.last_local_ids is None


def test_template_fused_form_and_shared_columns(cpu_ops):
    g = torch.Generator().manual_seed(1)
    M, E, k, H = 5, 8, 2, 64
    hid = torch.randn(M, H, generator=g).bfloat16()
    wg = torch.randn(E, H, generator=g).bfloat16()
    _, emap = O.determine_expert_map(2, 0, E)
    r = R.Router(k, E, expert_map=emap, num_fused_shared_experts=1, shared_local_base=4, shared_weight=0.5)
    w, ids = r.select_experts(hid, gate_weight=wg)
    assert cpu_ops == ["router_topk"] and w.shape == (M, k + 1) and ids.shape == (M, k + 1)
    assert bool((ids[:, -1] == E).all()) and bool((w[:, -1] == 0.5).all()) and bool((r.last_local_ids[:, -1] == 4).all())


def test_router_argument_checks(cpu_ops):
    with pytest.raises(ValueError):
        R.Router(0, 8)
    with pytest.raises(ValueError):
        R.Router(9, 8)
    with pytest.raises(ValueError):
        R.Router(2, 8, num_expert_group=4)                       # topk_group missing
    with pytest.raises(ValueError):
        R.Router(2, 10, num_expert_group=4, topk_group=2)        # experts do not divide into the groups
    with pytest.raises(ValueError):
        R.Router(2, 8, scoring_func="tanh")
    with pytest.raises(NotImplementedError):
        R.Router(2, 8, eplb_state=object())
    with pytest.raises(ValueError):
        R.Router(2, 8, num_fused_shared_experts=1)               # no local base for the shared expert
    r = R.Router(2, 8)
    with pytest.raises(ValueError):
        r.select_experts(torch.zeros(1, 4))                      # neither logits nor gate weight
    with pytest.raises(ValueError):
        r.select_experts(torch.zeros(1, 4), torch.zeros(1, 7))   # wrong expert count
    with pytest.raises(ValueError):
        R.Router(2, 8, scoring_func="softmax", num_expert_group=4, topk_group=2).select_experts(
            torch.zeros(1, 4), gate_weight=torch.zeros(8, 4))    # fused grouped routing is sigmoid-only
