"""HBM sizing of the routed experts for a B200 deployment (180 GB per GPU): how many bytes of expert weights each rank holds
under expert parallelism (the reference's linear expert map, expert_map_manager.py:65-90) and tensor parallelism (w13 rows /
w2 columns divided by tp, routed_experts.py:536-612), per weight format.  Bytes per weight follow SURVEY.md 8(d) / the
checkpoint layouts of row a8; `b200moe_device_bytes` reports the same figure for a constructed layer (plus tile padding).
Host-side arithmetic only."""
from __future__ import annotations

from dataclasses import dataclass

HBM_BYTES_B200 = 180 * 10 ** 9

# bytes per weight element: payload + scales
_PAYLOAD = {"bf16": 2.0, "fp16": 2.0, "fp8": 1.0, "wna16": 0.5, "int4": 0.5, "nvfp4": 0.5, "mxfp4": 0.5}
_GROUP_SCALE = {"wna16": 2.0 / 32, "int4": 2.0 / 32, "nvfp4": 1.0 / 16, "mxfp4": 1.0 / 32}


def expert_bytes(fmt: str, hidden_size: int, intermediate_size: int, gated: bool = True) -> float:
    """Bytes of ONE expert (w13 + w2 + scales) at `intermediate_size` (per rank when TP shards it)."""
    if fmt not in _PAYLOAD:
        raise ValueError(f"unknown weight format {fmt!r}")
    n = (3 if gated else 2) * hidden_size * intermediate_size
    b = n * _PAYLOAD[fmt]
    if fmt == "fp8":      # one f32 per 128 x 128 block of w13 and of w2
        rows13 = (2 if gated else 1) * -(-intermediate_size // 128)
        b += 4 * (rows13 * -(-hidden_size // 128) + -(-hidden_size // 128) * -(-intermediate_size // 128))
    else:
        b += n * _GROUP_SCALE.get(fmt, 0.0)
    return b


@dataclass(frozen=True)
class ExpertPlan:
    experts_per_rank: int          # the fullest rank (linear map: the first E % ep ranks hold one more)
    intermediate_per_rank: int
    bytes_per_expert: float
    bytes_per_layer_per_rank: float
    bytes_per_rank: float
    bytes_total: float
    fits: bool                     # under `hbm_bytes * budget_frac` on the fullest rank
    headroom_bytes: float


def plan_experts(fmt: str, num_experts: int, hidden_size: int, intermediate_size: int, num_moe_layers: int, ep_size: int = 1,
                 tp_size: int = 1, gated: bool = True, n_shared_experts: int = 0, hbm_bytes: float = HBM_BYTES_B200,
                 budget_frac: float = 0.85) -> ExpertPlan:
    """Expert-weight bytes per rank for `ep_size` x `tp_size` ranks.  Shared experts (replicated on every EP rank, the
    always-on columns of the fused router) count once per rank.  `budget_frac` leaves room for the KV cache, the dense
    layers and the workspaces (activation tiles, fp32 partials: tens of MB per device)."""
    if ep_size < 1 or tp_size < 1 or intermediate_size % tp_size:
        raise ValueError("ep_size / tp_size must be >= 1 and tp_size must divide the intermediate size")
    ipp = intermediate_size // tp_size
    if ipp % 128 or hidden_size % 128:
        raise ValueError("hidden_size and the per-rank intermediate size must be multiples of 128 (b200moe_create)")
    per = expert_bytes(fmt, hidden_size, ipp, gated)
    local = -(-num_experts // ep_size) + n_shared_experts
    layer = local * per
    rank = layer * num_moe_layers
    total = (num_experts + n_shared_experts) * expert_bytes(fmt, hidden_size, intermediate_size, gated) * num_moe_layers
    budget = hbm_bytes * budget_frac
    return ExpertPlan(local, ipp, per, layer, rank, total, rank <= budget, budget - rank)


def min_ep_size(fmt: str, num_experts: int, hidden_size: int, intermediate_size: int, num_moe_layers: int, **kw) -> int:
    """Smallest power-of-two EP size (1 .. 64) whose fullest rank fits."""
    ep = 1
    while ep <= 64:
        if plan_experts(fmt, num_experts, hidden_size, intermediate_size, num_moe_layers, ep_size=ep, **kw).fits:
            return ep
        ep *= 2
    raise ValueError("does not fit 64 GPUs")
