"""The caller side of the three `lk_moe` entry points, for users of this repository without
Lvllm around it: `ExpertsRunner.forward` is what `MoERunner._apply_quant_method` + `RoutedExperts._cpu_decode / _cpu_prefill /
_gpu_prefill` do in the reference (vllm/model_executor/layers/fused_moe/runner/moe_runner.py:577-664,
routed_experts.py:1824-1899):

  * under CUDA-graph capture -> `cpu_decode` on the current stream into the static fp32 buffer `[max_num_seqs, H]` shared by
    all layers of the process (`RoutedExperts.output_gpu`, :1827-1838), then the cast to the activation dtype;
  * eager and `num_tokens >= LVLLM_GPU_PREFILL_MIN_BATCH_SIZE` on a gpu-prefill layer -> `gpu_prefill` (activation dtype out);
  * otherwise -> `cpu_prefill` through HOST buffers after a stream synchronise (the reference's eager small-batch path).

With the experts in HBM all three run the same device kernels; the branch only reproduces which entry point an unmodified
Lvllm would use (lvllm_b200.envs.select_entry_point).  Layers Lvllm classifies as GPU-resident are not routed to lk_moe at
all in the reference; here they take `cpu_decode` under capture and `gpu_prefill` (no host traffic) otherwise.  Host-side
glue only.
"""
from __future__ import annotations

from typing import Callable

import torch

from . import envs


class ExpertsRunner:
    #: process-wide static decode output, one per (device, hidden size), like the reference's class attribute
    _decode_out: dict = {}

    def __init__(self, layer_name: str, moe, top_k: int, hidden_size: int, max_num_seqs: int,
                 num_speculative_tokens: int = 0, check_nan_in_output: bool = False,
                 is_capturing: Callable[[], bool] | None = None, stream_ptr: Callable[[], int] | None = None,
                 synchronize: Callable[[], None] | None = None):
        self.layer_name, self.moe = layer_name, moe
        self.top_k, self.hidden_size = int(top_k), int(hidden_size)
        # routed_experts.py:1814-1822: speculative decoding multiplies the decode batch
        self.max_num_seqs = int(max_num_seqs) * (1 + max(0, int(num_speculative_tokens)))
        self.check_nan_in_output = bool(check_nan_in_output)
        self._is_capturing = is_capturing or torch.cuda.is_current_stream_capturing
        self._stream_ptr = stream_ptr or (lambda: torch.cuda.current_stream().cuda_stream)
        self._synchronize = synchronize or (lambda: torch.cuda.current_stream().synchronize())

    # ---- buffers ----------------------------------------------------------------------------------------------------
    def decode_buffer(self, device: torch.device) -> torch.Tensor:
        key = (str(device), self.hidden_size)
        buf = ExpertsRunner._decode_out.get(key)
        if buf is None or buf.shape[0] < self.max_num_seqs:
            if buf is not None and self._is_capturing():
                raise RuntimeError("the static decode buffer cannot grow under CUDA-graph capture: construct the runners "
                                   "(or call decode_buffer) before capturing")
            buf = torch.zeros(self.max_num_seqs, self.hidden_size, dtype=torch.float32, device=device)
            ExpertsRunner._decode_out[key] = buf
        return buf

    # ---- the three entry points, with the reference's buffer handling --------------------------------------------------
    def _cpu_decode(self, hidden_states, topk_weights, topk_ids):
        M = hidden_states.size(0)
        if M > self.max_num_seqs:
            raise ValueError(f"decode batch {M} exceeds max_num_seqs x (1 + speculative tokens) = {self.max_num_seqs}")
        out = self.decode_buffer(hidden_states.device)
        self.moe.cpu_decode(self._stream_ptr(), M, self.top_k, hidden_states.data_ptr(), topk_ids.data_ptr(),
                            topk_weights.data_ptr(), out.data_ptr())
        y = out[:M]
        if self.check_nan_in_output:
            torch.nan_to_num(y, nan=0.0, out=y)
        return y.to(hidden_states.dtype)

    def _cpu_prefill(self, hidden_states, topk_weights, topk_ids):
        ids_cpu = topk_ids.to(dtype=torch.int32, device="cpu", non_blocking=True)
        w_cpu = topk_weights.to(dtype=torch.float32, device="cpu", non_blocking=True)
        h_cpu = hidden_states.to(device="cpu", non_blocking=True)
        out_cpu = torch.empty(hidden_states.shape, dtype=torch.float32, device="cpu")
        self._synchronize()
        self.moe.cpu_prefill(hidden_states.size(0), ids_cpu.size(1), ids_cpu.data_ptr(), w_cpu.data_ptr(), h_cpu.data_ptr(),
                             out_cpu.data_ptr())
        y = out_cpu.to(hidden_states.device, dtype=hidden_states.dtype, non_blocking=True)
        if self.check_nan_in_output:
            torch.nan_to_num(y, nan=0.0, out=y)
        return y

    def _gpu_prefill(self, hidden_states, topk_weights, topk_ids):
        out = torch.empty_like(hidden_states)
        self.moe.gpu_prefill(hidden_states.data_ptr(), out.data_ptr(), topk_ids.data_ptr(), topk_weights.data_ptr(),
                             hidden_states.size(0), topk_ids.size(1), self._stream_ptr())
        if self.check_nan_in_output:
            bad = torch.isnan(out) | torch.isinf(out)
            if bad.any():
                out.masked_fill_(bad, 0.0)
        return out

    # ---- the dispatch (moe_runner.py:602-654, routed_experts.py:1344-1357) ------------------------------------------------
    def entry_point(self, num_tokens: int, cudagraph_mode_none: bool = True) -> str:
        return envs.select_entry_point(self.layer_name, num_tokens, self._is_capturing(), cudagraph_mode_none)

    def forward(self, hidden_states: torch.Tensor, topk_weights: torch.Tensor, topk_ids: torch.Tensor,
                cudagraph_mode_none: bool = True) -> torch.Tensor:
        """hidden [M, H] (bf16 / fp16), topk_weights f32 [M, k], topk_ids i32 [M, k] (EP-local, < 0 = skip): the routed
        experts' weighted output [M, H] in the activation dtype."""
        if topk_ids.dtype != torch.int32 or topk_weights.dtype != torch.float32:
            raise ValueError("lk_moe takes int32 expert ids and float32 routing weights (routed_experts.py:1860)")
        if not (hidden_states.is_contiguous() and topk_ids.is_contiguous() and topk_weights.is_contiguous()):
            raise ValueError("the entry points take raw pointers: tensors must be contiguous")
        ep = self.entry_point(hidden_states.size(0), cudagraph_mode_none)
        if ep == "resident":
            # not an lk_moe layer in the reference (vLLM's own GPU experts run it); here: the capturable entry point under
            # capture, the host-traffic-free one otherwise
            ep = "cpu_decode" if (self._is_capturing() and hidden_states.size(0) <= self.max_num_seqs) else "gpu_prefill"
        if ep == "cpu_decode":
            return self._cpu_decode(hidden_states, topk_weights, topk_ids)
        if ep == "cpu_prefill":
            return self._cpu_prefill(hidden_states, topk_weights, topk_ids)
        return self._gpu_prefill(hidden_states, topk_weights, topk_ids)
