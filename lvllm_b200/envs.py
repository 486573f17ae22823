"""The LVLLM_* "hybrid scheduler" flags and per-layer predicates, re-stated for an all-HBM deployment.

Same names and semantics as reference vllm/envs.py:1929-1944 (flag getters) and :2292-2410 (predicates);
the dispatch they drive is reference .../fused_moe/runner/moe_runner.py:602-654.  With experts in HBM the
three lk_moe entry points all run the same device path, so the classification only decides which ENTRY
POINT the unmodified caller uses, not where the arithmetic happens.
"""
from __future__ import annotations

import os

_overrides: dict[str, object] = {}


def _get(name: str, default: str) -> str:
    if name in _overrides:
        return str(_overrides[name])
    return os.getenv(name, default)


def is_lk_moe_feature_enabled() -> bool:  # envs.py:2292
    return bool(int(_get("LVLLM_MOE_NUMA_ENABLED", "0")))


def is_numa_interleave_enabled() -> bool:  # envs.py:2295 — meaningless once experts live in HBM; kept for parity
    return bool(int(_get("LVLLM_ENABLE_NUMA_INTERLEAVE", "1")))


def get_gpu_prefill_min_batch_size() -> int:  # envs.py:2318
    return int(_get("LVLLM_GPU_PREFILL_MIN_BATCH_SIZE", "0"))


def is_lk_moe_use_gpu_prefill() -> bool:  # envs.py:2298
    return get_gpu_prefill_min_batch_size() > 0


def disable_lk_moe_gpu_prefill() -> int:  # envs.py:2301
    v = get_gpu_prefill_min_batch_size()
    _overrides["LVLLM_GPU_PREFILL_MIN_BATCH_SIZE"] = 0
    return v


def enable_lk_moe_gpu_prefill(value: int) -> int:  # envs.py:2306
    _overrides["LVLLM_GPU_PREFILL_MIN_BATCH_SIZE"] = int(value)
    return value


def get_gpu_prefetch_window() -> int:  # envs.py:2323 (getter default 3, :1942-1944)
    return int(_get("LVLLM_GPU_PREFETCH_WINDOW", "3"))


def enabled_layerwise_load() -> bool:  # envs.py:2409
    return bool(int(_get("LVLLM_ENABLE_MOE_LAYERWISE_LOAD", "0")))


def extract_layer_index(layer_name: str) -> int:  # envs.py:2327-2359 (num_attn_module == 1 form)
    ints = []
    for part in layer_name.split("."):
        try:
            ints.append(int(part))
        except ValueError:
            continue
    if len(ints) != 1:
        raise AssertionError(f"layer name {layer_name} should only contain one integer")
    return ints[0]


def parse_layer_list(spec: str | None) -> set[int]:
    """"0-1,33-34,40" -> {0,1,33,34,40}; malformed parts are skipped (envs.py:2383-2405)."""
    out: set[int] = set()
    if not spec:
        return out
    for part in spec.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            try:
                a, b = map(int, part.split("-"))
            except ValueError:
                continue
            if a <= b:
                out.update(range(a, b + 1))
        else:
            try:
                out.add(int(part))
            except ValueError:
                continue
    return out


def is_lk_moe_mtp_layer(layer_name: str) -> bool:  # envs.py:2361
    return layer_name.startswith("mtp.")


def is_lk_moe_gpu_resident_layer(layer_name: str) -> bool:  # envs.py:2371-2407
    if not is_lk_moe_feature_enabled():
        return True
    if is_lk_moe_mtp_layer(layer_name):
        return True
    layer_id = extract_layer_index(layer_name)   # before the spec check, like the reference: malformed names raise either way
    spec = _get("LVLLM_GPU_RESIDENT_MOE_LAYERS", "")
    if not spec:
        return False
    return layer_id in parse_layer_list(spec)


def is_lk_moe_gpu_prefill_layer(layer_name: str) -> bool:  # envs.py:2364
    return (is_lk_moe_use_gpu_prefill() and not is_lk_moe_gpu_resident_layer(layer_name)
            and not is_lk_moe_mtp_layer(layer_name))


def is_lk_moe_cpu_layer(layer_name: str) -> bool:  # envs.py:2367
    return (is_lk_moe_feature_enabled() and not is_lk_moe_gpu_resident_layer(layer_name)
            and not is_lk_moe_gpu_prefill_layer(layer_name) and not is_lk_moe_mtp_layer(layer_name))


def select_entry_point(layer_name: str, num_tokens: int, capturing: bool, cudagraph_mode_none: bool = True) -> str:
    """The 4-way branch of reference moe_runner.py:602-654 + routed_experts.py:1344-1357.
    Returns 'resident' | 'cpu_decode' | 'gpu_prefill' | 'cpu_prefill'."""
    if is_lk_moe_gpu_resident_layer(layer_name):
        return "resident"
    if capturing:
        return "cpu_decode"
    if (is_lk_moe_gpu_prefill_layer(layer_name) and cudagraph_mode_none
            and num_tokens >= get_gpu_prefill_min_batch_size()):
        return "gpu_prefill"
    return "cpu_prefill"


def cuda_graph_sizes(max_num_seqs: int) -> list[int]:
    """Decode sizes the reference captures (routed_experts.py:1829)."""
    return [1, 2, 4] + list(range(8, max_num_seqs + 1, 8))
