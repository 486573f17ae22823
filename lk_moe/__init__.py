"""`import lk_moe` drop-in: re-exports the B200-native implementation (lvllm_b200.lk_moe).

Lvllm imports this name when LVLLM_MOE_NUMA_ENABLED=1 (reference routed_experts.py:37-41)."""
from lvllm_b200.lk_moe import *  # noqa: F401,F403
from lvllm_b200.lk_moe import __all__  # noqa: F401
