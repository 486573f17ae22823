"""lvllm_b200/runner.py: the caller side of cpu_decode / cpu_prefill / gpu_prefill (reference moe_runner.py:577-664,
routed_experts.py:1824-1899) with a recording stand-in for the lk_moe object and CPU tensors: which entry point is taken for
which LVLLM_* setting, and the buffer handling around each (static fp32 decode buffer, host staging, output dtype)."""
import ctypes

import pytest
import torch

from lvllm_b200 import envs
from lvllm_b200.runner import ExpertsRunner


def _view(ptr, shape, dtype):
    n = int(torch.tensor(shape).prod()) * torch.empty(0, dtype=dtype).element_size()
    return torch.frombuffer((ctypes.c_char * n).from_address(ptr), dtype=dtype).reshape(shape)


class FakeMoe:
    """out[t] = (t + 1) * sum_j w[t, j] for every column: enough to see that the right buffers were wired"""

    def __init__(self, H):
        self.H, self.calls = H, []

    def _fill(self, out, w, M):
        out.copy_(((torch.arange(M).float() + 1) * w.sum(-1)).unsqueeze(1).expand(M, self.H))

    def cpu_decode(self, stream, M, k, hid, ids, w, out):
        self.calls.append(("cpu_decode", stream, M, k))
        self._fill(_view(out, (M, self.H), torch.float32), _view(w, (M, k), torch.float32), M)

    def cpu_prefill(self, M, k, ids, w, hid, out):
        self.calls.append(("cpu_prefill", M, k))
        assert _view(ids, (M, k), torch.int32).min() >= -1
        self._fill(_view(out, (M, self.H), torch.float32), _view(w, (M, k), torch.float32), M)

    def gpu_prefill(self, hid, out, ids, w, M, k, stream):
        self.calls.append(("gpu_prefill", M, k, stream))
        o32 = torch.empty(M, self.H)
        self._fill(o32, _view(w, (M, k), torch.float32), M)
        _view(out, (M, self.H), torch.bfloat16).copy_(o32.bfloat16())


@pytest.fixture()
def clean_env(monkeypatch):
    for k_ in ("LVLLM_MOE_NUMA_ENABLED", "LVLLM_GPU_PREFILL_MIN_BATCH_SIZE", "LVLLM_GPU_RESIDENT_MOE_LAYERS"):
        monkeypatch.delenv(k_, raising=False)
    envs._overrides.clear()
    ExpertsRunner._decode_out.clear()
    yield monkeypatch
    envs._overrides.clear()
    ExpertsRunner._decode_out.clear()


def _mk(H=16, capturing=False, **kw):
    state = {"cap": capturing, "sync": 0}
    moe = FakeMoe(H)
    r = ExpertsRunner("model.layers.5.mlp.experts", moe, 2, H, 8, is_capturing=lambda: state["cap"], stream_ptr=lambda: 77,
                      synchronize=lambda: state.__setitem__("sync", state["sync"] + 1), **kw)
    return r, moe, state


def _inputs(M, H=16, k=2):
    g = torch.Generator().manual_seed(M)
    return (torch.randn(M, H, generator=g).bfloat16(), torch.rand(M, k, generator=g).float(),
            torch.randint(0, 4, (M, k), generator=g, dtype=torch.int32))


def _expected(w, M, H):
    return ((torch.arange(M).float() + 1) * w.sum(-1)).unsqueeze(1).expand(M, H)


def test_entry_point_selection_follows_the_lvllm_flags(clean_env):
    mp = clean_env
    mp.setenv("LVLLM_MOE_NUMA_ENABLED", "1")
    r, moe, st = _mk()
    h, w, ids = _inputs(3)
    # eager, no gpu prefill configured -> the host-pointer entry point, after a stream synchronise, activation dtype out
    y = r.forward(h, w, ids)
    assert moe.calls == [("cpu_prefill", 3, 2)] and st["sync"] == 1 and y.dtype == torch.bfloat16
    torch.testing.assert_close(y.float(), _expected(w, 3, 16).bfloat16().float())
    # under capture -> cpu_decode on the current stream into the static fp32 buffer shared by the process
    st["cap"] = True
    y = r.forward(h, w, ids)
    assert moe.calls[-1] == ("cpu_decode", 77, 3, 2) and y.dtype == torch.bfloat16
    buf = ExpertsRunner._decode_out[("cpu", 16)]
    assert buf.shape == (8, 16) and buf.dtype == torch.float32 and torch.equal(buf[:3], _expected(w, 3, 16))
    r2, moe2, _ = _mk(capturing=True)
    r2.forward(h, w, ids)
    assert ExpertsRunner._decode_out[("cpu", 16)] is buf                      # one buffer for all layers
    with pytest.raises(ValueError):
        r.forward(*_inputs(9))                                                 # beyond max_num_seqs
    # gpu prefill threshold: eager batches at or above it take gpu_prefill, smaller ones cpu_prefill
    st["cap"] = False
    mp.setenv("LVLLM_GPU_PREFILL_MIN_BATCH_SIZE", "4")
    y = r.forward(*_inputs(5))
    assert moe.calls[-1] == ("gpu_prefill", 5, 2, 77) and y.dtype == torch.bfloat16
    r.forward(*_inputs(3))
    assert moe.calls[-1][0] == "cpu_prefill"
    assert r.forward(*_inputs(5), cudagraph_mode_none=False) is not None and moe.calls[-1][0] == "cpu_prefill"
    # a GPU-resident layer is not an lk_moe layer in the reference: gpu_prefill eagerly, cpu_decode under capture
    mp.setenv("LVLLM_GPU_RESIDENT_MOE_LAYERS", "5-6")
    r.forward(*_inputs(2))
    assert moe.calls[-1][0] == "gpu_prefill"
    st["cap"] = True
    r.forward(*_inputs(2))
    assert moe.calls[-1][0] == "cpu_decode"


def test_speculative_tokens_and_argument_checks(clean_env):
    clean_env.setenv("LVLLM_MOE_NUMA_ENABLED", "1")
    r, moe, st = _mk(capturing=True, num_speculative_tokens=2)
    assert r.max_num_seqs == 24
    r.forward(*_inputs(20))                                                    # 8 x (1 + 2) rows fit
    assert moe.calls[-1] == ("cpu_decode", 77, 20, 2)
    h, w, ids = _inputs(3)
    with pytest.raises(ValueError):
        r.forward(h, w, ids.long())
    with pytest.raises(ValueError):
        r.forward(h, w.double(), ids)
    with pytest.raises(ValueError):
        r.forward(h.t().contiguous().t(), w, ids)
    # nan scrubbing like check_nan_in_output (routed_experts.py:130)
    r3, moe3, _ = _mk(capturing=True, check_nan_in_output=True)
    moe3._fill = lambda out, w_, M: out.fill_(float("nan"))
    assert bool((r3.forward(h, w, ids) == 0).all())
