"""Checkpoint -> HBM ingest of the routed experts.

The reference builds stacked CPU parameters `w13 [E, 2I, H]` / `w2 [E, H, I]` with `RoutedExperts.weight_loader`
(vllm/model_executor/layers/fused_moe/routed_experts.py:644-1168: one call per (expert, projection) checkpoint tensor,
`_load_w13` :528-576 / `_load_w2` :578-606 narrow the tensor-parallel slice), hands their `data_ptr()`s to the `lk_moe`
constructor, which copies them again into NUMA memory, and frees them (:1420-1432).  On a B200 the experts go straight to
HBM: this module reads the safetensors shards of a checkpoint through `mmap` (no copy of the file, no stacked host tensor),
cuts exactly the slice `_load_w13` / `_load_w2` would cut for this rank, and feeds ONE expert at a time to
`b200moe_load_experts` (`MOE_X.from_expert_shards`), which stages, re-tiles and quantisation-repacks it on the device.
Host memory in flight: one expert's `w13` (gate and up rows joined) — 88 MB for a DeepSeek-V3 FP8 expert instead of the
1.4 GB per layer the stacked parameter needs.

Host-side only (file parsing, slicing, naming): nothing here touches the GPU until `load_layer` calls the C ABI.
"""
from __future__ import annotations

import json
import mmap
import os
import struct
from dataclasses import dataclass
from typing import Iterator, Sequence

import numpy as np
import torch

# safetensors dtype tags -> (torch dtype, bytes per element); the file format is an 8-byte little-endian header length, a
# JSON header {name: {dtype, shape, data_offsets}}, then the raw little-endian tensor bytes
_ST_DTYPES = {
    "F64": (torch.float64, 8), "F32": (torch.float32, 4), "F16": (torch.float16, 2), "BF16": (torch.bfloat16, 2),
    "I64": (torch.int64, 8), "I32": (torch.int32, 4), "I16": (torch.int16, 2), "I8": (torch.int8, 1),
    "U8": (torch.uint8, 1), "BOOL": (torch.bool, 1), "F8_E4M3": (torch.float8_e4m3fn, 1), "F8_E5M2": (torch.float8_e5m2, 1),
    "F8_E8M0": (torch.uint8, 1),   # ue8m0 scale bytes: handed on as raw bytes
}


class SafetensorsFile:
    """Read-only, zero-copy view of one .safetensors file."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        head = self._f.read(8)
        if len(head) != 8:
            raise ValueError(f"{path}: not a safetensors file (shorter than its length prefix)")
        (n,) = struct.unpack("<Q", head)
        size = os.fstat(self._f.fileno()).st_size
        if n <= 0 or 8 + n > size:
            raise ValueError(f"{path}: not a safetensors file (header length {n} exceeds the file size {size})")
        try:
            self.header = json.loads(self._f.read(n))
        except (UnicodeDecodeError, json.JSONDecodeError) as ex:
            raise ValueError(f"{path}: safetensors header is not JSON: {ex}") from None
        self.header.pop("__metadata__", None)
        self._base = 8 + n
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if size else None
        self._np = np.frombuffer(self._mm, dtype=np.uint8) if self._mm is not None else np.zeros(0, np.uint8)
        for name, info in self.header.items():
            dt, esz = self._dtype(name, info)
            b0, b1 = info["data_offsets"]
            numel = int(np.prod(info["shape"], dtype=np.int64)) if info["shape"] else 1
            if b1 - b0 != numel * esz or self._base + b1 > size or b0 < 0:
                raise ValueError(f"{path}: tensor {name!r} has inconsistent offsets {b0}..{b1} for shape {info['shape']} {info['dtype']}")

    @staticmethod
    def _dtype(name, info):
        try:
            return _ST_DTYPES[info["dtype"]]
        except KeyError:
            raise ValueError(f"tensor {name!r}: unsupported safetensors dtype {info['dtype']!r}") from None

    def names(self):
        return self.header.keys()

    def tensor(self, name: str) -> torch.Tensor:
        """The tensor as a view of the mapped file (read-only memory: slice / copy it, never write to it)."""
        info = self.header[name]
        dt, _ = self._dtype(name, info)
        b0, b1 = info["data_offsets"]
        raw = self._np[self._base + b0:self._base + b1]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)   # "the given NumPy array is not writable": it is never written
            t = torch.from_numpy(raw)
        return t.view(dt).reshape(info["shape"])

    def close(self):
        self._np = None
        if self._mm is not None:
            try:
                self._mm.close()
            except BufferError:
                pass          # tensors handed out still reference the mapping; it is released with them
            self._mm = None
        self._f.close()


def save_safetensors(path: str, tensors: dict) -> None:
    """Write `tensors` (name -> CPU tensor) as one .safetensors file: synthetic checkpoints for tests and tools."""
    tag = {v[0]: k for k, v in _ST_DTYPES.items() if k != "F8_E8M0"}
    header, blobs, off = {}, [], 0
    for name, t in tensors.items():
        t = t.detach().cpu().contiguous()
        raw = t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": tag[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * (-len(hj) % 8)          # the data section starts 8-byte aligned
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


class ExpertCheckpoint:
    """Name -> tensor over the safetensors shards of one checkpoint directory (or an explicit list of files)."""

    def __init__(self, paths: str | Sequence[str]):
        if isinstance(paths, str):
            if os.path.isdir(paths):
                paths = sorted(os.path.join(paths, f) for f in os.listdir(paths) if f.endswith(".safetensors"))
            else:
                paths = [paths]
        if not paths:
            raise ValueError("no .safetensors files")
        self.files = [SafetensorsFile(p) for p in paths]
        self._where = {}
        for f in self.files:
            for n in f.names():
                self._where[n] = f

    def __contains__(self, name: str) -> bool:
        return name in self._where

    def tensor(self, name: str) -> torch.Tensor:
        try:
            return self._where[name].tensor(name)
        except KeyError:
            raise KeyError(f"checkpoint has no tensor {name!r}") from None

    def close(self):
        for f in self.files:
            f.close()


# ------------------------------------------------------------------------------------------------------------
# the reference's tensor-parallel narrowing
# ------------------------------------------------------------------------------------------------------------
def tp_slice(loaded: torch.Tensor, shard_dim: int, tp_rank: int, tp_size: int) -> torch.Tensor:
    """The slice of a checkpoint tensor that `_load_w13` / `_load_w2` copy for this rank (routed_experts.py:551-561,
    :592-602): `shape[shard_dim] // tp_size` entries from `rank * that`; scalars (0-dim) are replicated."""
    if loaded.ndim == 0 or tp_size == 1:
        return loaded
    per = loaded.shape[shard_dim] // tp_size
    start = per * tp_rank
    avail = loaded.shape[shard_dim] - start
    if avail <= 0:
        return loaded.narrow(shard_dim, 0, 0)
    return loaded.narrow(shard_dim, start, min(per, avail))


def join_w13(gate: torch.Tensor | None, up: torch.Tensor, tp_rank: int = 0, tp_size: int = 1) -> torch.Tensor:
    """One expert's w13 parameter slice: rows [0, I_pp) = this rank's gate rows ("w1"), [I_pp, 2 I_pp) = its up rows ("w3")
    (routed_experts.py:564-570; both projections are column-parallel: sharded along their output rows, dim 0).  Non-gated
    experts (`gate is None`) have the single projection only.  Works unchanged for every tensor indexed like the weight
    rows: block scales [I/128, H/128], group scales [I, H/g], packed nibbles [I, H/2]."""
    u = tp_slice(up, 0, tp_rank, tp_size)
    if gate is None:
        return u.contiguous()
    return torch.cat([tp_slice(gate, 0, tp_rank, tp_size), u], dim=0)


def slice_w2(down: torch.Tensor, tp_rank: int = 0, tp_size: int = 1) -> torch.Tensor:
    """One expert's w2 parameter slice: row-parallel, sharded along its input columns, dim 1 (routed_experts.py:578-606)."""
    return tp_slice(down, 1, tp_rank, tp_size).contiguous()


# ------------------------------------------------------------------------------------------------------------
# checkpoint spellings
# ------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ExpertNames:
    """How a checkpoint spells the three projections of expert `e` under a layer prefix — the arguments of the reference's
    `make_expert_params_mapping(ckpt_gate_proj_name, ckpt_down_proj_name, ckpt_up_proj_name)`: DeepSeek / Qwen3 use
    gate_proj / down_proj / up_proj, Mixtral w1 / w2 / w3."""
    gate: str | None = "gate_proj"
    down: str = "down_proj"
    up: str = "up_proj"
    experts: str = "experts"

    def key(self, prefix: str, e, proj: str, suffix: str) -> str:
        if isinstance(e, str):           # a full module path instead of an expert index: e.g. DeepSeek's shared expert
            return f"{e}.{proj}.{suffix}"
        return f"{prefix}.{self.experts}.{e}.{proj}.{suffix}"


def local_expert_ids(ep_size: int, ep_rank: int, global_num_experts: int, strategy: str = "linear") -> list[int]:
    """GLOBAL ids of the experts rank `ep_rank` owns, in local order — the inverse view of the reference's expert map
    (`determine_expert_map`, expert_map_manager.py:22-113: "linear" gives the first `E % ep` ranks one extra expert and
    contiguous blocks, "round_robin" deals ids `rank, rank + ep, ...`)."""
    if ep_size <= 0 or not (0 <= ep_rank < ep_size):
        raise ValueError(f"bad EP coordinates: rank {ep_rank} of {ep_size}")
    base, rem = divmod(global_num_experts, ep_size)
    n_local = base + (1 if ep_rank < rem else 0)
    if strategy == "linear":
        start = ep_rank * base + min(ep_rank, rem)
        return list(range(start, start + n_local))
    if strategy == "round_robin":
        return list(range(ep_rank, global_num_experts, ep_size))
    raise ValueError(f"unsupported expert placement strategy {strategy!r}")


# per weight format: (weight suffix, block / group scale suffix or None, per-tensor global scale suffix or None)
_FORMAT_SUFFIXES = {
    "bf16": ("weight", None, None),
    "fp16": ("weight", None, None),
    "fp8": ("weight", "weight_scale_inv", None),          # DeepSeek-V3 block-128 checkpoints (fp8.py:570-633)
    "wna16": ("weight_packed", "weight_scale", None),     # compressed-tensors int4 (compressed_tensors_moe_wna16.py:155-190)
    "nvfp4": ("weight", "weight_scale", "weight_scale_2"),  # ModelOpt NVFP4 (modelopt.py:1452-1533)
    "mxfp4": ("weight", "weight_scale", None),            # per-expert MXFP4 (mxfp4.py:594-648)
    # compressed-tensors NVFP4: same bytes under other names, and the global scale is stored as its reciprocal — the
    # reference inverts it before the lk_moe constructor (`_process_nvfp4(need_reciprocal_global_scale=True)`, :1684-1688)
    "nvfp4-ct": ("weight_packed", "weight_scale", "weight_global_scale"),
}
_RECIPROCAL_GLOBAL = {"nvfp4-ct"}
# (groupN, groupK) the reference derives from the scale shapes (`_get_quant_params`, routed_experts.py:1440-1453)
FORMAT_GROUPS = {"bf16": (0, 0), "fp16": (0, 0), "fp8": (128, 128), "wna16": (1, 32), "nvfp4": (1, 16), "nvfp4-ct": (1, 16),
                 "mxfp4": (1, 32)}


def layer_config(cfg, fmt: str, num_local_experts: int, top_k: int, hidden_size: int, intermediate_size_per_rank: int,
                 max_num_batched_tokens: int, max_num_seqs: int, has_gate_proj: bool = True, activation_type: int = 0):
    """Fill an `lk_moe.MOEConfigV2` the way `RoutedExperts._process_*` do (routed_experts.py:1490-1511): `cfg` is the empty
    config object, returned filled."""
    cfg.expert_num, cfg.top_k = int(num_local_experts), int(top_k)
    cfg.hidden_size, cfg.intermediate_size = int(hidden_size), int(intermediate_size_per_rank)
    cfg.max_batch_size, cfg.max_num_seqs = int(max_num_batched_tokens), int(max_num_seqs)
    cfg.groupN, cfg.groupK = FORMAT_GROUPS[fmt]
    cfg.has_gate_proj, cfg.activation_type = bool(has_gate_proj), int(activation_type)
    return cfg


def _as_bytes_view(t: torch.Tensor, fmt: str) -> torch.Tensor:
    # compressed-tensors stores int4 weights packed into int32 [N, K/8]; lk_moe receives the same bytes as uint8 [N, K/2]
    # (routed_experts.py:1456-1533); little-endian, so the view is the conversion
    if fmt == "wna16" and t.dtype == torch.int32:
        return t.contiguous().view(torch.uint8)
    return t


def expert_tensors(ckpt: ExpertCheckpoint, prefix: str, fmt: str, expert_ids: Sequence[int], tp_rank: int = 0,
                   tp_size: int = 1, names: ExpertNames = ExpertNames()) -> Iterator[tuple]:
    """Yield `(local_index, w13, w2, s13, s2, g13, g2)` for the experts `expert_ids` (GLOBAL ids of this rank's experts in
    local order: the reference's linear expert map gives rank r the ids [r E/ep, (r+1) E/ep), expert_map_manager.py:65-90),
    each tensor contiguous, in the layout the `lk_moe` constructors take for ONE expert (absent ones None).  An entry of
    `expert_ids` may also be a module path (str): that module's projections become one more local expert — how a shared
    expert of the same shape (`...mlp.shared_experts`, reference runner/shared_experts.py) joins the routed launch as the
    always-on expert of `b200_router_topk(n_shared, shared_local_base, ...)`."""
    if fmt not in _FORMAT_SUFFIXES:
        raise ValueError(f"unknown weight format {fmt!r}")
    w_sfx, s_sfx, g_sfx = _FORMAT_SUFFIXES[fmt]
    for local, e in enumerate(expert_ids):
        def get(proj, sfx):
            return _as_bytes_view(ckpt.tensor(names.key(prefix, e, proj, sfx)), fmt)
        gate_w = get(names.gate, w_sfx) if names.gate else None
        w13 = join_w13(gate_w, get(names.up, w_sfx), tp_rank, tp_size)
        w2 = slice_w2(get(names.down, w_sfx), tp_rank, tp_size)
        s13 = s2 = g13 = g2 = None
        if s_sfx:
            gate_s = get(names.gate, s_sfx) if names.gate else None
            up_s, down_s = get(names.up, s_sfx), get(names.down, s_sfx)
            if up_s.ndim == 0 or up_s.numel() == 1:
                # per-tensor FP8 scales: one value per projection -> [2] for w13 (gate, up), [1] for w2 (fp8.py:570-633)
                vals = ([gate_s.reshape(())] if gate_s is not None else []) + [up_s.reshape(())]
                s13 = torch.stack(vals).to(torch.float32)
                s2 = down_s.reshape(1).to(torch.float32)
            else:
                s13 = join_w13(gate_s, up_s, tp_rank, tp_size)
                s2 = slice_w2(down_s, tp_rank, tp_size)
        if g_sfx:
            vals = ([get(names.gate, g_sfx).reshape(())] if names.gate else []) + [get(names.up, g_sfx).reshape(())]
            g13 = torch.stack(vals).to(torch.float32)
            g2 = get(names.down, g_sfx).reshape(1).to(torch.float32)
            if fmt in _RECIPROCAL_GLOBAL:
                g13, g2 = 1.0 / g13, 1.0 / g2
        yield (local, w13, w2, s13, s2, g13, g2)


def orient_fused(fused: torch.Tensor, shard_id: str, hidden_size: int) -> torch.Tensor:
    """Checkpoints that store all experts of a projection in ONE 3-D tensor come in both orientations; the reference
    normalises them to (intermediate, hidden) for w1 / w3 and (hidden, intermediate) for w2 and only transposes when the
    hidden dimension is definitely on the wrong axis (`_orient_fused_weight`, routed_experts.py:472-493)."""
    hidden_axis, inter_axis = (-2, -1) if shard_id == "w2" else (-1, -2)
    if fused.shape[hidden_axis] != hidden_size and fused.shape[inter_axis] == hidden_size:
        return fused.transpose(-1, -2)
    return fused


def fused_expert_tensors(ckpt: ExpertCheckpoint, prefix: str, expert_ids: Sequence[int], hidden_size: int, tp_rank: int = 0,
                         tp_size: int = 1, gate_up: str = "experts.gate_up_proj", down: str = "experts.down_proj") -> Iterator[tuple]:
    """`expert_tensors` for 16-bit checkpoints whose experts are fused into 3-D tensors `[E, ...]` (Llama-4 / Qwen3-VL-MoE
    style): gate = first half, up = second half of the oriented gate_up rows (`fused_weight.chunk(2, dim=1)`, reference
    load_weights routed_experts.py:988-1001), then the same per-expert slices."""
    gu = orient_fused(ckpt.tensor(f"{prefix}.{gate_up}"), "w1", hidden_size)
    dn = orient_fused(ckpt.tensor(f"{prefix}.{down}"), "w2", hidden_size)
    gate, up = gu.chunk(2, dim=1)
    for local, e in enumerate(expert_ids):
        yield (local, join_w13(gate[e], up[e], tp_rank, tp_size).contiguous(), slice_w2(dn[e], tp_rank, tp_size), None, None,
               None, None)


def shards_from_tensors(tensor_iter) -> Iterator[tuple]:
    """Per-expert tensor tuples (`expert_tensors` / `fused_expert_tensors`) in the form `MOE_X.from_expert_shards` consumes:
    raw pointers of one expert at a time; the tensors stay alive until the generator is advanced (the C ABI has copied them
    to the device by then)."""
    for (local, *ts) in tensor_iter:
        keep = [t.contiguous() if t is not None else None for t in ts]
        yield (local, 1, *[0 if t is None else t.data_ptr() for t in keep])
        del keep


def expert_shards(ckpt: ExpertCheckpoint, prefix: str, fmt: str, expert_ids: Sequence[int], tp_rank: int = 0,
                  tp_size: int = 1, names: ExpertNames = ExpertNames()) -> Iterator[tuple]:
    """`expert_tensors` as the pointer tuples of `MOE_X.from_expert_shards`."""
    return shards_from_tensors(expert_tensors(ckpt, prefix, fmt, expert_ids, tp_rank, tp_size, names))


def load_layer(moe_cls, cfg, ckpt: ExpertCheckpoint, prefix: str, fmt: str, expert_ids: Sequence[int], tp_rank: int = 0,
               tp_size: int = 1, names: ExpertNames = ExpertNames()):
    """Build one MoE layer object straight from the checkpoint: `moe_cls` is one of the `lk_moe.MOE_*` classes, `cfg` its
    `MOEConfigV2` (expert_num = len(expert_ids), intermediate_size = the per-rank size)."""
    if int(cfg.expert_num) != len(expert_ids):
        raise ValueError(f"cfg.expert_num = {cfg.expert_num} but {len(expert_ids)} expert ids were given")
    return moe_cls.from_expert_shards(cfg, expert_shards(ckpt, prefix, fmt, expert_ids, tp_rank, tp_size, names))
