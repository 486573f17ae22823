"""Golden fixture for lvllm_b200/loader.py: the reference's OWN expert weight-loader narrowing, called unbound.

Run in the build container only (needs /root/reference):
    cd /tmp && PYTHONPATH=/root/reference python /root/repo/tests/golden/make_golden_loader.py

RoutedExperts._load_w13 / _load_w2 (reference vllm/model_executor/layers/fused_moe/routed_experts.py:528-606) are called on an
attribute bag that carries what they read from `self` (moe_config.is_act_and_mul, moe_config.moe_parallel_config.tp_size and
the two static helpers), for weights and for block-scale tensors, tp_size 1 / 2 / 4, gated and non-gated experts.  Inputs and
the resulting per-expert parameter slices go to tests/golden/golden_loader.pt.  Nothing here is imported by the product.
"""
import os
import sys
import types

import torch

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_loader.pt")


def main():
    sys.path.insert(0, "/root/reference")
    from vllm.model_executor.layers.fused_moe.routed_experts import RoutedExperts

    def fake_self(tp_size, gated):
        ns = types.SimpleNamespace()
        ns.moe_config = types.SimpleNamespace(is_act_and_mul=gated,
                                              moe_parallel_config=types.SimpleNamespace(tp_size=tp_size))
        ns._get_hidden_dim = RoutedExperts._get_hidden_dim                      # static helper
        ns._narrow_expert_data_for_padding = RoutedExperts._narrow_expert_data_for_padding
        return ns

    gen = torch.Generator().manual_seed(0)
    cases = []
    for (name, I, H, row_div, col_div, dtype, gated) in [
        ("weight bf16", 64, 32, 1, 1, torch.bfloat16, True),          # [I, H] / [H, I]
        ("fp8 block scales", 512, 256, 128, 128, torch.float32, True),  # [I/128, H/128] / [H/128, I/128]
        ("packed 4-bit weight", 64, 32, 1, 2, torch.uint8, True),     # [I, H/2] / [H, I/2]
        ("group-32 scales", 128, 64, 1, 32, torch.uint8, True),        # [I, H/32] / [H, I/32]
        ("non-gated weight", 48, 32, 1, 1, torch.bfloat16, False),
    ]:
        def rnd(r, c):
            if dtype == torch.uint8:
                return torch.randint(0, 256, (r, c), generator=gen, dtype=torch.uint8)
            return torch.randn(r, c, generator=gen).to(dtype)
        gate = rnd(I // row_div, H // col_div)
        up = rnd(I // row_div, H // col_div)
        down = rnd(H // row_div, I // col_div)
        outs = []
        for tp_size in (1, 2, 4):
            for tp_rank in range(tp_size):
                fs = fake_self(tp_size, gated)
                ipp = I // tp_size
                w13 = torch.zeros((2 if gated else 1) * ipp // row_div, H // col_div, dtype=dtype)
                w2 = torch.zeros(H // row_div, ipp // col_div, dtype=dtype)
                # non-gated experts (is_act_and_mul False) load their single projection as shard "w1" (routed_experts.py:543-546)
                RoutedExperts._load_w13(fs, expert_data=w13, shard_dim=0, shard_id="w1", loaded_weight=gate, tp_rank=tp_rank)
                if gated:
                    RoutedExperts._load_w13(fs, expert_data=w13, shard_dim=0, shard_id="w3", loaded_weight=up, tp_rank=tp_rank)
                RoutedExperts._load_w2(fs, expert_data=w2, shard_dim=1, loaded_weight=down, tp_rank=tp_rank)
                outs.append(dict(tp_size=tp_size, tp_rank=tp_rank, w13=w13.clone(), w2=w2.clone()))
        cases.append(dict(name=name, gated=gated, gate=gate, up=up, down=down, outs=outs))
    # ---- fused 3-D expert tensors (Llama-4 / Qwen3-VL-MoE style `experts.gate_up_proj`, `experts.down_proj`): the
    # reference orients them with _orient_fused_weight and splits gate / up with chunk(2, dim=1) (load_weights :988-1001)
    fused = []
    E, I, H = 3, 64, 32
    for orient in ("out_in", "in_out"):
        gu = torch.randn(E, 2 * I, H, generator=gen).bfloat16()
        dn = torch.randn(E, H, I, generator=gen).bfloat16()
        if orient == "in_out":
            gu, dn = gu.transpose(1, 2).contiguous(), dn.transpose(1, 2).contiguous()
        o13 = RoutedExperts._orient_fused_weight(gu, "w1", H)
        o2 = RoutedExperts._orient_fused_weight(dn, "w2", H)
        gate, up = o13.chunk(2, dim=1)
        outs = []
        for tp_size in (1, 2):
            for tp_rank in range(tp_size):
                fs = fake_self(tp_size, True)
                ipp = I // tp_size
                w13 = torch.zeros(E, 2 * ipp, H, dtype=torch.bfloat16)
                w2 = torch.zeros(E, H, ipp, dtype=torch.bfloat16)
                for e in range(E):
                    RoutedExperts._load_w13(fs, expert_data=w13[e], shard_dim=0, shard_id="w1", loaded_weight=gate[e], tp_rank=tp_rank)
                    RoutedExperts._load_w13(fs, expert_data=w13[e], shard_dim=0, shard_id="w3", loaded_weight=up[e], tp_rank=tp_rank)
                    RoutedExperts._load_w2(fs, expert_data=w2[e], shard_dim=1, loaded_weight=o2[e], tp_rank=tp_rank)
                outs.append(dict(tp_size=tp_size, tp_rank=tp_rank, w13=w13.clone(), w2=w2.clone()))
        fused.append(dict(orient=orient, hidden=H, gate_up=gu, down=dn, outs=outs))
    torch.save(dict(cases=cases, fused=fused), OUT)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
