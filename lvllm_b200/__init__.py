"""lvllm_b200 — B200-native (sm_100a) MoE expert path, routing, permutation and decode attention behind
Lvllm's `lk_moe` / FusedMoE call-site API.  Host side only binds the C ABI of libb200moe.so."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
