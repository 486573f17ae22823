#!/usr/bin/env python
"""bench.py — decode tok/s of the MoE + decode-attention hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W [--workload NAME] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one decode step of the hot path over the named model configuration with synthetic inputs and
random-init weights of that architecture: for every MoE layer  paged decode attention (MLA or GQA) ->
fused router (router GEMM + top-k + EP id remap, ONE hand-written kernel) -> [EP dispatch] -> routed experts through
the lk_moe ``cpu_decode`` entry point (the call Lvllm makes under CUDA-graph capture) -> [EP combine / all-reduce].
The whole step is captured in one CUDA graph, like the reference's decode path.  Dense projections, norms
and the shared expert are outside SURVEY.md §8 and are not part of the step.

Prints ONE JSON line (see DESIGN.md "Measurement" for every field).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decode tok/s DeepSeek-V3 FP8 @1/2/4/8 B200; expert-GEMM tensor-pipe %"

# public HF config.json values of the named models
WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (needs >= 4 GPUs: 654 GB of experts)
    # configs[3] shapes at decode (the prefill-8192 config itself is a parity/throughput case, not the bench line)
    "dsv3-nvfp4": dict(model="DeepSeek-V3 671B", fmt="nvfp4", layers=58, H=7168, I=2048, E=256, k=8, routing="grouped",
                       n_group=8, topk_group=4, rsf=2.5, attn="mla", Hq=128, Hkv=1, batch=1, seq=4096, page=64),
    "dsv3-fp8": dict(model="DeepSeek-V3 671B", fmt="fp8", layers=58, H=7168, I=2048, E=256, k=8, routing="grouped",
                     n_group=8, topk_group=4, rsf=2.5, attn="mla", Hq=128, Hkv=1, batch=1, seq=4096, page=64),
    # configs[0]: the reference's CPU-runnable case (bf16 experts, 90 GB: fits one B200)
    "mixtral-bf16": dict(model="Mixtral-8x7B", fmt="bf16", layers=32, H=4096, I=14336, E=8, k=2, routing="softmax",
                         attn="gqa", Hq=32, Hkv=8, batch=1, seq=64, page=16),
    # configs[1]
    "mixtral-int4": dict(model="Mixtral-8x7B", fmt="int4", layers=32, H=4096, I=14336, E=8, k=2, routing="softmax",
                         attn="gqa", Hq=32, Hkv=8, batch=64, seq=2048, page=16),
    # configs[4]
    "qwen3-mxfp4": dict(model="Qwen3-235B-A22B", fmt="mxfp4", layers=94, H=4096, I=1536, E=128, k=8,
                        routing="softmax", attn="gqa", Hq=64, Hkv=4, batch=256, seq=512, page=16),
}
IMPLEMENTED_FMTS = ("fp8", "bf16", "int4", "mxfp4", "nvfp4")


def bytes_per_expert(w) -> float:
    n = 3 * w["H"] * w["I"]
    if w["fmt"] == "fp8":
        return n + 4 * (2 * w["I"] // 128 * (w["H"] // 128) + w["H"] // 128 * (w["I"] // 128))
    if w["fmt"] == "bf16":
        return 2.0 * n
    if w["fmt"] == "int4":
        return n * (0.5 + 2 / 32)
    if w["fmt"] == "nvfp4":
        return n * (0.5 + 1 / 16)
    if w["fmt"] == "mxfp4":
        return n * (0.5 + 1 / 32)
    raise ValueError(w["fmt"])


def ncu_traffic(name: str, n_gpus: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the
    .ncu-rep); None where no capture exists."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d.get(name, {}).get(str(n_gpus))
    except Exception:
        return None


def ncu_tensor_pipe():
    """the "expert-GEMM tensor-pipe %" half of BASELINE's metric: ncu `sm__pipe_tensor_cycles_active` of the decode kernel
    this bench times and of the prefill-class grouped GEMM (tools/prefill_bench.py), from the committed captures."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("_tensor_pipe_pct")
    except Exception:
        return None


def default_workload(n_gpus: int) -> str:
    """ONE workload for every N so that the driver's 1/2/4/8 runs form a curve: BASELINE config 5 (Qwen3-235B-A22B MXFP4
    decode batch 256, literally "EP all-to-all sweep 1/2/4/8 GPU"), the largest BASELINE configuration that fits a
    single B200 (118 GB of experts).  The configuration the metric is quoted on (DeepSeek-V3 FP8, 654 GB of experts)
    needs 8 GPUs: the N = 8 line carries it as the sub-object ``dsv3_fp8_ep8``."""
    return "qwen3-mxfp4"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def wait_first(self, timeout: float = 8.0):
        """block until nvidia-smi has delivered its first sample (it can take a second to start on a busy 8-GPU box; a
        short timed region would otherwise end before the first line arrives)"""
        t_end = time.time() + timeout
        while self.proc and not self.lines and time.time() < t_end:
            time.sleep(0.02)

    def stop(self, t0: float, t1: float):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.05:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ model
class HotPathModel:
    """Synthetic decoder: per layer the tensors the hot path touches, all resident in HBM."""

    def __init__(self, w, rank, world, dev, seed=0):
        import torch
        import lk_moe
        from lvllm_b200 import envs
        self.w, self.rank, self.world, self.dev = w, rank, world, dev
        self.torch = torch
        g = torch.Generator(device=dev).manual_seed(seed)  # same weights on every rank; each keeps its experts
        E, H, I, k = w["E"], w["H"], w["I"], w["k"]
        assert E % world == 0
        # token-sharded callers (DP attention) + EP experts -> dispatch/combine all-to-all; a batch that does not
        # split over the ranks (DeepSeek-V3 batch 1) keeps the lk_moe contract "replicated tokens, local partition, sum
        # over ranks" (moe_runner.py:488-494), expert-parallel by default.  BENCH_TP=1 selects the reference's TP partition
        # instead (every rank holds ALL experts at intermediate_size / world — MOEConfigV2's intermediate_size is per
        # partition): balanced where EP at batch 1 is not (8 routed experts land on 8 ranks as 0..4 per rank), but measured
        # slower here: a rank's launch then covers 8 experts x 1/8 width = 16 GEMM1 tiles + 448 two-k-block GEMM2 tiles and
        # the stream-K fix-ups dominate (77 us against 37 us per layer for one rank's shard, profiles/r02_summary.md).
        self.a2a = world > 1 and w["batch"] >= world and w["batch"] % world == 0
        self.tp = (world > 1 and not self.a2a and w["fmt"] in ("fp8", "bf16") and I % world == 0 and (I // world) % 128 == 0
                   and os.environ.get("BENCH_TP") == "1")
        self.E_local = E if self.tp else E // world
        if self.tp:
            I = I // world
        self.I_local = I
        lo = rank * self.E_local
        self.expert_map = None
        if world > 1 and not self.tp:
            em = torch.full((E,), -1, dtype=torch.int32)
            em[lo:lo + self.E_local] = torch.arange(self.E_local, dtype=torch.int32)
            self.expert_map = em.to(dev)
        self.B_global = w["batch"]
        B = w["batch"] // world if self.a2a else w["batch"]
        self.B = B
        self.layers = []
        cfg = lk_moe.MOEConfigV2()
        cfg.num_processes, cfg.process_id, cfg.gpu_id = world, rank, dev.index or 0
        cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = self.E_local, k, H, I
        cfg.max_batch_size, cfg.max_num_seqs = max(self.B_global, 16), max(self.B_global, 16)
        self.graph_sizes = envs.cuda_graph_sizes(max(self.B_global, 8))
        self.raw0 = None   # layer 0's raw checkpoint tensors (multi-rank runs: EP parity check, then dropped)
        self.cfg = cfg
        for li in range(w["layers"]):
            L = {}
            keep = (li == 0 and world > 1)
            L["gate"] = (torch.randn(E, H, device=dev, dtype=torch.bfloat16, generator=g) * 0.02)
            if w["routing"] == "grouped":
                L["bias"] = torch.randn(E, device=dev, generator=g) * 0.1
            if w["fmt"] == "fp8":
                cfg.groupN = cfg.groupK = 128
                w13 = (torch.randn(self.E_local, 2 * I, H, device=dev, dtype=torch.bfloat16, generator=g) / 10).to(torch.float8_e4m3fn)
                w2 = (torch.randn(self.E_local, H, I, device=dev, dtype=torch.bfloat16, generator=g) / 10).to(torch.float8_e4m3fn)
                s13 = torch.rand(self.E_local, 2 * I // 128, H // 128, device=dev, generator=g) * 4e-3 + 1e-3
                s2 = torch.rand(self.E_local, H // 128, I // 128, device=dev, generator=g) * 4e-3 + 1e-3
                L["moe"] = lk_moe.MOE_FP8(cfg, w13.data_ptr(), w2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0,
                                          weights_on_device=True)
                if keep:
                    self.raw0 = ("fp8", w13, w2, s13, s2, None, None)
                del w13, w2, s13, s2
            elif w["fmt"] == "bf16":
                w13 = torch.randn(self.E_local, 2 * I, H, device=dev, dtype=torch.bfloat16, generator=g) / 10
                w2 = torch.randn(self.E_local, H, I, device=dev, dtype=torch.bfloat16, generator=g) / 10
                L["moe"] = lk_moe.MOE_BF16(cfg, w13.data_ptr(), w2.data_ptr(), 0, 0, 0, 0, weights_on_device=True)
                if keep:
                    self.raw0 = ("bf16", w13, w2, None, None, None, None)
                del w13, w2
            elif w["fmt"] in ("int4", "mxfp4", "nvfp4"):
                # raw checkpoint layouts of the 4-bit formats: packed nibbles + group scales
                El = self.E_local
                p13 = torch.randint(0, 256, (El, 2 * I, H // 2), device=dev, dtype=torch.uint8, generator=g)
                p2 = torch.randint(0, 256, (El, H, I // 2), device=dev, dtype=torch.uint8, generator=g)
                if w["fmt"] == "int4":
                    cfg.groupN, cfg.groupK = 1, 32
                    s13 = (torch.rand(El, 2 * I, H // 32, device=dev, generator=g) * 0.01 + 0.002).bfloat16()
                    s2 = (torch.rand(El, H, I // 32, device=dev, generator=g) * 0.01 + 0.002).bfloat16()
                    L["moe"] = lk_moe.MOE_WNA16(cfg, p13.data_ptr(), p2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0,
                                                weights_on_device=True)
                    if keep:
                        self.raw0 = ("int4", p13, p2, s13, s2, None, None)
                elif w["fmt"] == "mxfp4":
                    cfg.groupN, cfg.groupK = 1, 32
                    s13 = torch.randint(117, 122, (El, 2 * I, H // 32), device=dev, dtype=torch.uint8, generator=g)
                    s2 = torch.randint(117, 122, (El, H, I // 32), device=dev, dtype=torch.uint8, generator=g)
                    L["moe"] = lk_moe.MOE_MXFP4(cfg, p13.data_ptr(), p2.data_ptr(), s13.data_ptr(), s2.data_ptr(), 0, 0,
                                                weights_on_device=True)
                    if keep:
                        self.raw0 = ("mxfp4", p13, p2, s13, s2, None, None)
                else:
                    cfg.groupN, cfg.groupK = 1, 16
                    s13 = (torch.rand(El, 2 * I, H // 16, device=dev, generator=g) * 2 + 0.5).to(torch.float8_e4m3fn)
                    s2 = (torch.rand(El, H, I // 16, device=dev, generator=g) * 2 + 0.5).to(torch.float8_e4m3fn)
                    g13 = torch.full((El, 2), 0.004, device=dev)
                    g2 = torch.full((El,), 0.004, device=dev)
                    L["moe"] = lk_moe.MOE_NVFP4(cfg, p13.data_ptr(), p2.data_ptr(), s13.data_ptr(), s2.data_ptr(),
                                                g13.data_ptr(), g2.data_ptr(), weights_on_device=True)
                    if keep:
                        self.raw0 = ("nvfp4", p13, p2, s13, s2, g13, g2)
                    del g13, g2
                del p13, p2, s13, s2
            else:
                raise SystemExit(f"weight format {w['fmt']} is not implemented yet")
            # paged KV cache + this step's query (synthetic; q/k/v projections are outside the hot path)
            S, page = w["seq"], w["page"]
            npg = -(-S // page)
            if w["attn"] == "mla":
                L["kv"] = torch.randn(B * npg, page, 576, device=dev, dtype=torch.bfloat16, generator=g)
                L["qn"] = torch.randn(B, w["Hq"], 512, device=dev, dtype=torch.bfloat16, generator=g)
                L["qp"] = torch.randn(B, w["Hq"], 64, device=dev, dtype=torch.bfloat16, generator=g)
            else:
                L["kc"] = torch.randn(B * npg, page, w["Hkv"], 128, device=dev, dtype=torch.bfloat16, generator=g)
                L["vc"] = torch.randn(B * npg, page, w["Hkv"], 128, device=dev, dtype=torch.bfloat16, generator=g)
                L["q"] = torch.randn(B, w["Hq"], 128, device=dev, dtype=torch.bfloat16, generator=g)
            self.layers.append(L)
            torch.cuda.empty_cache()
        npg = -(-w["seq"] // w["page"])
        self.page_table = torch.arange(B * npg, device=dev, dtype=torch.int32).reshape(B, npg)
        self.seq_lens = torch.full((B,), w["seq"], device=dev, dtype=torch.int32)
        self.hidden_in = torch.zeros(B, H, device=dev, dtype=torch.bfloat16)   # step input (H2D target)
        self.hidden = torch.zeros(B, H, device=dev, dtype=torch.bfloat16)
        self.moe_out = torch.zeros(B, H, device=dev, dtype=torch.float32)      # the lk_moe static fp32 buffer
        self.final = torch.zeros(B, H, device=dev, dtype=torch.bfloat16)
        self.ep = None
        self.last_ids = []
        self.timers = None

    def attach_ep(self, ep):
        self.ep = ep
        if self.a2a:
            ep.a2a_init(self.B, self.w["H"], self.w["k"], self.E_local)

    def route(self, L, hidden):
        """router GEMM + top-k (+ EP id remap) in one kernel (rows a1-a5 / f1 of SURVEY.md 8)."""
        from lvllm_b200 import ops
        w = self.w
        em = None if self.a2a else self.expert_map
        if w["routing"] == "grouped":
            return ops.router_topk(hidden, L["gate"], w["k"], True, "sigmoid", L["bias"], w["rsf"], w["n_group"],
                                   w["topk_group"], em)
        return ops.router_topk(hidden, L["gate"], w["k"], True, "softmax", None, 1.0, 0, 0, em)

    def ep_parity(self):
        """Multi-rank parity where the driver runs N > 1: layer 0's routed-expert output in its EP form (dispatch / combine
        all-to-all, or replicated tokens + all-reduce) against a SINGLE-RANK cpu_decode of the same tokens over all
        experts on rank 0.  Returns max |ep - single| / max |single| (rank 0; None elsewhere)."""
        import torch.distributed as dist
        import lk_moe
        torch = self.torch
        w, dev, world, rank = self.w, self.dev, self.world, self.rank
        H, k, Bg = w["H"], w["k"], self.B_global
        L = self.layers[0]
        st = torch.cuda.current_stream().cuda_stream
        g = torch.Generator(device=dev).manual_seed(1234)
        hid = (torch.randn(Bg, H, device=dev, generator=g) / 10).bfloat16()
        if self.a2a:
            mine = hid[rank * self.B:(rank + 1) * self.B].contiguous()
            tw, ids, _ = self.route(L, mine)
            self.ep.dispatch(mine, ids, tw)
            L["moe"].cpu_decode(st, Bg, k, self.ep.x_ptr, self.ep.ids_ptr, self.ep.w_ptr, self.ep.y_ptr)
            out_l = torch.empty(self.B, H, dtype=torch.float32, device=dev)
            self.ep.combine(ids, out_l)
            out_ep = torch.empty(Bg, H, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(out_ep, out_l)
        else:
            tw, ids, loc = self.route(L, hid)
            part = torch.empty(Bg, H, dtype=torch.float32, device=dev)
            L["moe"].cpu_decode(st, Bg, k, hid.data_ptr(), loc.data_ptr(), tw.data_ptr(), part.data_ptr())
            out_ep = self.ep.allreduce(part).clone()
        torch.cuda.synchronize()
        err = None
        if rank == 0:
            # every rank generated the same local experts, so global expert e = local expert e % E_local
            fmt, a13, a2, b13, b2, c13, c2 = self.raw0
            cfg = self.cfg
            if self.tp:
                # every rank generated the same I/world slice: the unsharded layer is the slice repeated along I
                # (w13 = [gate rows; up rows], w2 along its last dim; block scales likewise)
                def rep13(t):
                    if t is None:
                        return None
                    h = t.shape[1] // 2
                    return torch.cat([t[:, :h]] * world + [t[:, h:]] * world, dim=1).contiguous()
                rep2 = lambda t: None if t is None else torch.cat([t] * world, dim=2).contiguous()
                f13, f2, fs13, fs2, fg13, fg2 = rep13(a13), rep2(a2), rep13(b13), rep2(b2), None, None
                cfg.intermediate_size = w["I"]
            else:
                rep = lambda t: None if t is None else torch.cat([t] * world).contiguous()
                f13, f2, fs13, fs2, fg13, fg2 = rep(a13), rep(a2), rep(b13), rep(b2), rep(c13), rep(c2)
            cfg.expert_num, cfg.num_processes, cfg.process_id = w["E"], 1, 0
            cls = {"fp8": lk_moe.MOE_FP8, "bf16": lk_moe.MOE_BF16, "int4": lk_moe.MOE_WNA16, "mxfp4": lk_moe.MOE_MXFP4,
                   "nvfp4": lk_moe.MOE_NVFP4}[fmt]
            p = lambda t: 0 if t is None else t.data_ptr()
            full = cls(cfg, p(f13), p(f2), p(fs13), p(fs2), p(fg13), p(fg2), weights_on_device=True)
            twg, idg, _ = ops_router_global(self, L, hid)
            ref = torch.empty(Bg, H, dtype=torch.float32, device=dev)
            full.cpu_decode(st, Bg, k, hid.data_ptr(), idg.data_ptr(), twg.data_ptr(), ref.data_ptr())
            torch.cuda.synchronize()
            err = float((out_ep - ref).abs().max() / ref.abs().max().clamp(min=1e-20))
            full.close()
            cfg.expert_num, cfg.intermediate_size = self.E_local, self.I_local
            del f13, f2, fs13, fs2, fg13, fg2
        self.raw0 = None
        torch.cuda.empty_cache()
        dist.barrier()
        return err

    def step(self, record_ids: bool = False):
        """Enqueue one decode step on the current stream (graph-capturable)."""
        torch = self.torch
        from lvllm_b200 import ops
        w = self.w
        st = torch.cuda.current_stream().cuda_stream
        B, k = self.B, w["k"]
        self.hidden.copy_(self.hidden_in)
        if record_ids:
            self.last_ids = []
        tm = self.timers   # None, or {op: [(start_event, end_event), ...]} during the eager breakdown passes

        def mark(op):
            if tm is None:
                return None
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            return (op, e0)

        def done(tok):
            if tok is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                tm.setdefault(tok[0], []).append((tok[1], e1))

        for L in self.layers:
            t = mark("attention")
            if w["attn"] == "mla":
                ops.mla_decode(L["qn"], L["qp"], L["kv"], self.seq_lens, self.page_table, 1.0 / math.sqrt(192),
                               max_seq_len=w["seq"])
            else:
                ops.gqa_decode(L["q"], L["kc"], L["vc"], self.seq_lens, self.page_table, 128 ** -0.5)
            done(t)
            t = mark("router")
            tw, ids, loc = self.route(L, self.hidden)
            done(t)
            if self.a2a:
                # dispatch this rank's rows to the expert owners, run the local experts on the gathered global
                # batch (ids local to this rank, -1 elsewhere), pull + sum the partial rows of this rank's tokens
                ep = self.ep
                t = mark("ep_dispatch")
                ep.dispatch(self.hidden, ids, tw)
                done(t)
                if record_ids:   # eager profiling passes only: the global batch's ids that land on this rank
                    import torch.distributed as dist
                    allids = torch.empty(self.B_global, k, dtype=torch.int32, device=ids.device)
                    dist.all_gather_into_tensor(allids, ids.contiguous())
                    lo = self.rank * self.E_local
                    self.last_ids.append(torch.where((allids >= lo) & (allids < lo + self.E_local), allids - lo, -1))
                t = mark("experts")
                L["moe"].cpu_decode(st, self.B_global, k, ep.x_ptr, ep.ids_ptr, ep.w_ptr, ep.y_ptr)
                done(t)
                # combine all-to-all fused with the norm that follows (one kernel, no fp32 round trip)
                t = mark("ep_combine_norm")
                ep.combine_norm(ids, self.hidden, gain=0.1)
                done(t)
                continue
            else:
                if loc is not None:
                    ids = loc
                if record_ids:
                    self.last_ids.append(ids.clone())
                t = mark("experts")
                L["moe"].cpu_decode(st, B, k, self.hidden.data_ptr(), ids.data_ptr(), tw.data_ptr(), self.moe_out.data_ptr())
                done(t)
                src = self.moe_out
                if self.ep is not None:
                    # replicated tokens: NVLink push all-reduce fused with the norm that follows (one kernel)
                    t = mark("ep_allreduce_norm")
                    self.ep.allreduce_norm(self.moe_out, self.hidden, gain=0.1)
                    done(t)
                    continue
            # the reference casts lk_moe's fp32 output to the activation dtype (routed_experts.py:1855); fused
            # here with an RMS normalisation (the op that follows in the layer) so that chained random-init
            # layers stay O(0.1) and finite
            t = mark("rmsnorm_cast")
            ops.rmsnorm_cast(src, self.hidden, gain=0.1)
            done(t)
        self.final.copy_(self.hidden)


def ops_router_global(model, L, hidden):
    """routing with GLOBAL expert ids (no EP remap): what a single rank holding every expert consumes"""
    from lvllm_b200 import ops
    w = model.w
    if w["routing"] == "grouped":
        return ops.router_topk(hidden, L["gate"], w["k"], True, "sigmoid", L["bias"], w["rsf"], w["n_group"], w["topk_group"], None)
    return ops.router_topk(hidden, L["gate"], w["k"], True, "softmax", None, 1.0, 0, 0, None)


# ------------------------------------------------------------------------------------------------ cpu arm
def _real_lk_moe():
    """try the real ``lk_moe`` wheel from site-packages (never the repo's drop-in of the same name)."""
    import importlib.machinery
    paths = [p for p in sys.path if p and os.path.abspath(p) != ROOT]
    try:
        spec = importlib.machinery.PathFinder.find_spec("lk_moe", paths)
    except Exception:
        spec = None
    return spec is not None and spec.origin is not None and not os.path.abspath(spec.origin).startswith(ROOT)


def _usable_cores() -> int:
    """host threads this process can actually run: the affinity mask, capped by the container's CPU quota (cgroup v2
    cpu.max / v1 cfs quota).  A box whose quota is far below its mask made the calibration below time its first, most
    oversubscribed trial only (128 threads on a handful of CPUs: 0.2 tok/s where another box gave 5)."""
    n = len(os.sched_getaffinity(0))
    try:
        q = None
        if os.path.exists("/sys/fs/cgroup/cpu.max"):
            a, b = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
            if a != "max":
                q = float(a) / float(b)
        elif os.path.exists("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
            a = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            b = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if a > 0:
                q = a / b
        if q:
            n = max(1, min(n, int(math.ceil(q))))
    except Exception:
        pass
    return n


def _host_mem_available() -> int:
    """bytes this process may still allocate: MemAvailable, capped by the cgroup limit when there is one"""
    avail = 0
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
                break
        for lim_f, use_f in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                             ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
            if os.path.exists(lim_f):
                lim = open(lim_f).read().strip()
                if lim != "max" and int(lim) < (1 << 60):
                    avail = min(avail, int(lim) - int(open(use_f).read().strip()))
                break
    except Exception:
        pass
    return max(0, avail)


def cpu_reference_arm(w, steps: int, warmup: int, budget_s: float = 25.0):
    """The reference's CPU expert path on the box's host cores, on a bounded sample of the same workload: timed passes
    of ONE full MoE layer at the REAL batch (no extrapolation over tokens; identical layers are multiplied out).
      bf16     : the reference tree's own CPU fused MoE (csrc/cpu/cpu_fused_moe.cpp, AVX-512 / AMX), kind "reference";
      4-bit    : oracle/moe_ref.c expert-major batched port (rows dequantised once per expert, AVX-512 `omp simd` dots);
      fp8 / M=1: oracle/moe_ref.c token-major GEMV port (the decode algorithm of a DRAM-bound CPU engine).
    The closed lk_moe wheel itself is not installed anywhere (checked at run time and reported)."""
    import torch
    from oracle import c_ref
    H, I, k, B, E = w["H"], w["I"], w["k"], w["batch"], w["E"]
    cores = _usable_cores()
    kind, impl = "port", "oracle/moe_ref.c, OpenMP"
    g = torch.Generator().manual_seed(0)
    # experts resident in DRAM: all of them when a decode batch touches (nearly) all, else a pool that keeps the
    # per-token bytes identical (batch 1 touches k experts per layer)
    pool = E if B * k >= E else max(4 * k, 32)
    pool = min(pool, E)
    if w["fmt"] == "fp8":
        w13 = torch.randint(0, 0x78, (pool, 2 * I, H), dtype=torch.uint8, generator=g).view(torch.float8_e4m3fn)
        w2 = torch.randint(0, 0x78, (pool, H, I), dtype=torch.uint8, generator=g).view(torch.float8_e4m3fn)
        s13 = torch.rand(pool, 2 * I // 128, H // 128, generator=g) * 1e-3
        s2 = torch.rand(pool, H // 128, I // 128, generator=g) * 1e-3
        fn = lambda hid, ids, tw: c_ref.forward_fp8_block(hid, w13, s13, w2, s2, ids, tw)
        impl = "oracle/moe_ref.c token-major GEMV port, OpenMP"
    elif w["fmt"] in ("int4", "nvfp4", "mxfp4"):
        fmt = w["fmt"]
        grp = 16 if fmt == "nvfp4" else 32
        w13 = torch.randint(0, 256, (pool, 2 * I, H // 2), dtype=torch.uint8, generator=g)
        w2 = torch.randint(0, 256, (pool, H, I // 2), dtype=torch.uint8, generator=g)
        if fmt == "int4":
            s13 = (torch.rand(pool, 2 * I, H // grp, generator=g) * 0.01 + 0.002).bfloat16()
            s2 = (torch.rand(pool, H, I // grp, generator=g) * 0.01 + 0.002).bfloat16()
            g13 = g2 = None
        elif fmt == "nvfp4":
            s13 = (torch.rand(pool, 2 * I, H // grp, generator=g) * 2 + 0.5).to(torch.float8_e4m3fn)
            s2 = (torch.rand(pool, H, I // grp, generator=g) * 2 + 0.5).to(torch.float8_e4m3fn)
            g13, g2 = torch.full((pool, 2), 0.004), torch.full((pool,), 0.004)
        else:
            s13 = torch.randint(117, 122, (pool, 2 * I, H // grp), dtype=torch.uint8, generator=g)
            s2 = torch.randint(117, 122, (pool, H, I // grp), dtype=torch.uint8, generator=g)
            g13 = g2 = None
        if B >= 8:
            fn = lambda hid, ids, tw: c_ref.forward_w4_batched(hid, w13, s13, w2, s2, ids, tw, fmt, g13, g2)
            impl = "oracle/moe_ref.c expert-major batched port (rows dequantised once per expert, AVX-512 omp simd), OpenMP"
        else:
            fn = lambda hid, ids, tw: c_ref.forward_w4(hid, w13, s13, w2, s2, ids, tw, fmt, g13, g2, exact=False)
            impl = "oracle/moe_ref.c token-major GEMV port, OpenMP"
    else:
        w13 = (torch.randn(pool, 2 * I, H, generator=g) / 10).bfloat16()
        w2 = (torch.randn(pool, H, I, generator=g) / 10).bfloat16()
        fn = lambda hid, ids, tw: c_ref.forward_bf16(hid, w13, w2, ids, tw)
        try:
            from oracle import ref_moe
            if ref_moe.available():
                rm = ref_moe.RefMoe(w13, w2)
                fn = lambda hid, ids, tw: rm.forward(hid, ids, tw)
                kind, impl = "reference", f"reference csrc/cpu/cpu_fused_moe.cpp, isa {ref_moe.isa()}"
        except Exception:
            pass
    hid = (torch.randn(B, H, generator=g) / 10).bfloat16()
    tw = torch.rand(B, k, generator=g).float()
    mk_ids = lambda n: torch.stack([torch.randperm(pool, generator=g)[:k] for _ in range(n)]).int().contiguous()

    def time_impl(fn, budget):
        """calibrate the thread count, then timed passes of one full layer at the real batch; (s per layer, threads, passes)"""
        # thread count: torchrun exports OMP_NUM_THREADS=1 and a container's CPU quota can be far below its affinity
        # mask, so the count is calibrated on a small pass (fastest wins) and reported as `cores`
        nc = B if B <= 64 else max(64, B // 2)   # calibrate on (nearly) the real batch: parallel efficiency depends on rows per expert
        ids_c = mk_ids(nc)
        trials = []
        t_cal = time.time()
        cands = sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True)
        # every candidate is timed on a small slice first (cheap), so that a slow box cannot spend the whole calibration
        # budget on its first trial; the two best are then re-timed on (nearly) the real batch
        ns = min(nc, 16)
        small = []
        for n_thr in cands:
            c_ref.lib().moe_ref_set_threads(n_thr)
            if not small:
                fn(hid[:4], ids_c[:4], tw[:4])   # page the weights in once
            t0 = time.perf_counter()
            fn(hid[:ns], ids_c[:ns], tw[:ns])
            small.append((time.perf_counter() - t0, n_thr))
        for _, n_thr in sorted(small)[:2]:
            c_ref.lib().moe_ref_set_threads(n_thr)
            t0 = time.perf_counter()
            fn(hid[:nc], ids_c, tw[:nc])
            trials.append((time.perf_counter() - t0, n_thr))
            if time.time() - t_cal > 0.5 * budget:
                break
        best_t = min(t for t, _ in trials)
        best_n = min(n for t, n in trials if t <= 1.05 * best_t)   # fewest threads within 5 % of the best
        c_ref.lib().moe_ref_set_threads(best_n)
        times = []
        t_start = time.time()
        n = 0
        while True:
            ids = mk_ids(B)
            t0 = time.perf_counter()
            fn(hid, ids, tw)
            dt = time.perf_counter() - t0
            n += 1
            if n > min(warmup, 1):
                times.append(dt)
            if len(times) >= max(steps, 3) or time.time() - t_start > budget:
                break
        return sum(times) / max(1, len(times)), best_n, len(times)

    # quantised layers: the reference tree has no CPU kernel for packed 4-bit / block-FP8 experts, but its own CPU fused MoE
    # (AVX-512 / AMX micro-GEMMs) on bf16 weights of the same shapes is the strongest CPU implementation of this layer the
    # tree can offer — a decode batch is compute-bound on the host, so reading 2-4x the bytes costs it little (at batch 1
    # the packed port may win: it streams fewer bytes).  Both are timed and the faster one is the arm (the other is
    # reported next to it).
    cand = [(fn, kind, impl)]
    if w["fmt"] in ("int4", "nvfp4", "mxfp4", "fp8") and not os.environ.get("BENCH_CPU_PORT_ONLY"):
        try:
            from oracle import ref_moe
            need = pool * 3 * H * I * 2 * 3          # bf16 weights, their pre-packed copy, head-room
            if ref_moe.available() and _host_mem_available() > need + (8 << 30):
                one13 = (torch.randn(1, 2 * I, H, generator=g) / 10).bfloat16()
                one2 = (torch.randn(1, H, I, generator=g) / 10).bfloat16()
                wb13, wb2 = one13.repeat(pool, 1, 1), one2.repeat(pool, 1, 1)     # timing does not depend on the values
                rm4 = ref_moe.RefMoe(wb13, wb2)
                del wb13, wb2
                cand.append((lambda hid_, ids_, tw_: rm4.forward(hid_, ids_, tw_), "reference",
                             f"reference csrc/cpu/cpu_fused_moe.cpp (isa {ref_moe.isa()}) on bf16 weights of the layer's shapes: "
                             "the reference tree has no CPU kernel for packed 4-bit / block-FP8 experts"))
        except Exception as ex:   # the port alone is still a valid arm
            sys.stderr.write(f"[cpu arm] reference bf16 candidate unavailable: {ex!r}\n")
    results = []
    for (f_, kind_, impl_) in cand:
        try:
            per, thr, n_pass = time_impl(f_, budget_s / len(cand))
            results.append((per, thr, n_pass, kind_, impl_))
        except Exception as ex:
            sys.stderr.write(f"[cpu arm] candidate failed ({impl_[:40]}...): {ex!r}\n")
    if not results:
        raise RuntimeError("no CPU implementation of the layer could be timed")
    per_layer, cores, n_times, kind, impl = min(results, key=lambda r: r[0])
    step_s = per_layer * w["layers"]
    out = {"value": B / step_s, "unit": "tok/s", "cores": cores, "kind": kind,
           "sample": f"{n_times} timed passes of ONE full MoE layer at the real batch ({B} token(s) x top-{k} over {pool} "
                     f"DRAM-resident experts), x{w['layers']} identical layers ({impl}); real lk_moe wheel installed: "
                     f"{_real_lk_moe()}",
           "ms_per_layer_pass": per_layer * 1e3, "threads": cores, "fits_in_driver_run": True}
    if len(results) > 1:
        out["candidates"] = [{"kind": r[3], "impl": r[4][:80], "ms_per_layer_pass": r[0] * 1e3, "threads": r[1]} for r in results]
    return out


# ------------------------------------------------------------------------------------------------ gpu arm
def measure(name, w, args, rank, world, local_rank, debug_layers=False):
    """Build the model of workload `name` on this rank, check it, time args.steps graph replays.  Returns the fields of
    the JSON line (rank 0) or None."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from lvllm_b200 import _lib
    lib = _lib.lib()
    dev = torch.device("cuda", local_rank)
    warmup = max(args.warmup, 3)
    model = HotPathModel(w, rank, args.ep_shard_of or world, dev)
    if args.ep_shard_of:
        assert world == 1 and not model.a2a, "--ep-shard-of emulates replicated-token EP shards only"
    ep_err = None
    if world > 1:
        from lvllm_b200.ep import EpGroup
        model.attach_ep(EpGroup(rank, world, dev, max_elems=w["batch"] * w["H"]))
        ep_err = model.ep_parity()
        bad = torch.tensor([1.0 if (ep_err is not None and not (ep_err <= 5e-3)) else 0.0], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if bad.item() > 0:
            if rank == 0:
                print(json.dumps({"error": "EP parity check failed", "workload": name, "ep_parity_max_err": ep_err}))
            raise SystemExit(3)
    B, H = w["batch"], w["H"]
    Bl = model.B   # rows this rank feeds / reads per step (B / world under dispatch-combine EP)
    host_in = (torch.randn(Bl, H) / 10).bfloat16().pin_memory()
    host_out = torch.empty(Bl, H, dtype=torch.bfloat16).pin_memory()
    model.hidden_in.copy_(host_in)

    # --- eager warm-up pass with per-kernel timing of the expert kernel (roofline) ----------------------
    lib.b200moe_profile(1)
    for _ in range(2):
        model.step(record_ids=True)
    torch.cuda.synchronize()
    lib.b200moe_profile(1)  # reset the window after the cold pass
    n_prof = 3
    distinct = 0
    for _ in range(n_prof):
        model.step(record_ids=True)
        torch.cuda.synchronize()
        distinct += sum(int((torch.unique(i[i >= 0])).numel()) for i in model.last_ids)
    g1, g2, calls = C.c_double(), C.c_double(), C.c_int64()
    lib.b200moe_profile_read(C.byref(g1), C.byref(g2), C.byref(calls))
    lib.b200moe_profile(0)
    bpe = bytes_per_expert(w) / (model.world if model.tp else 1)   # TP: this rank streams 1/world of each routed expert
    # dominant kernel = the expert GEMMs of one MoE layer: ONE fused persistent kernel for decode batches
    # (moe_fused_kernel: routing table + gather/quant + GEMM1 + GEMM2 + combine), or the GEMM1 + GEMM2 pair of
    # the large-batch path.  Algorithmic bytes per launch = weight bytes of the distinct active local experts.
    moe_bytes_per_launch = distinct / max(1, calls.value) * bpe
    moe_ms = (g1.value + g2.value) / max(1, calls.value)
    fused = g2.value < 0.05 * max(g1.value, 1e-9)

    # --- per-op breakdown of one eager step (CUDA events on the launching stream; eager launches carry ~2-3 us of
    # launch gap each that the graph replay below does not) ----------------------------------------------
    model.timers = {}
    model.step()
    torch.cuda.synchronize()
    breakdown = {op: round(1e3 * sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev)), 2) for op, ev in model.timers.items()}
    model.timers = None

    # --- capture the step in a CUDA graph (the reference's decode path replays a graph) -----------------
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        model.step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        lc1 = lib.b200moe_launch_count()
        with torch.cuda.graph(graph, stream=side):
            model.step()
        launches_per_step = lib.b200moe_launch_count() - lc1
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            if e2e:
                model.hidden_in.copy_(host_in, non_blocking=True)
            graph.replay()
            if e2e:
                host_out.copy_(model.final, non_blocking=True)
                torch.cuda.current_stream().synchronize()   # the caller consumes the step result
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    timed(warmup, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:          # the line is rank 0's; seven more nvidia-smi pollers only slow the box down
        sampler.start()
        sampler.wait_first()
    time.sleep(0.3)
    t0 = time.time()
    ms = timed(args.steps, False)
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    timed(1, True)
    ms_e2e = timed(args.steps, True)
    ok = bool(torch.isfinite(model.final.float()).all().item())
    native_mx = bool(model.layers[0]["moe"].query(0) == 1)
    del graph, model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = moe_bytes_per_launch / (moe_ms * 1e-3) / 1e9 if moe_ms > 0 else 0.0
    step_bytes = distinct / n_prof * bpe   # routed-expert bytes streamed per step on this rank
    return {
        "tok_s": B * args.steps / (ms / 1e3), "tok_s_e2e": B * args.steps / (ms_e2e / 1e3), "ms_per_step": ms / args.steps,
        "clocks": clocks, "launches_per_step": int(launches_per_step), "finite": ok, "ep_parity_max_err": ep_err,
        "native_mx": native_mx, "breakdown_us_per_layer_eager": breakdown,
        "roofline": {"bound": "hbm",
                     "kernel": ("moe_fused_kernel (sort + gather/quant + GEMM1 + SiLU*mul + GEMM2 + combine, stream-K"
                                + ("; native block-scaled MXFP4, W4A8-MX" if native_mx else "") + ")")
                     if fused else "moe_gemm_kernel GEMM1 + GEMM2",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": ncu_traffic(name, args.gpus),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "avg_launch_ms": moe_ms, "algorithmic_bytes_per_launch": moe_bytes_per_launch,
                     "step_expert_gbs": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                     "step_frac_of_peak": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
    }


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--layers", type=int, default=None, help="debug only: override the layer count (invalid as a bench line)")
    ap.add_argument("--ep-shard-of", type=int, default=None,
                    help="debug only: build rank 0's shard of an N-way EP job on ONE GPU without the all-reduce "
                         "(memory / kernel check of the multi-GPU shape; invalid as a bench line)")
    ap.add_argument("--no-sub", action="store_true", help="skip the dsv3_fp8_ep8 sub-measurement of the N = 8 line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    name = args.workload or default_workload(args.ep_shard_of or args.gpus)
    w = dict(WORKLOADS[name])
    debug_layers = args.layers is not None
    if debug_layers:
        w["layers"] = args.layers
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(w, args.steps, warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tok/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * w["batch"] / cb["value"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": w["fmt"],
                "data": "synthetic", "config": _config(name, w, args.gpus),
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "fits_in_driver_run", "candidates")
                                 if k in cb},
                "e2e": {"value": cb["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    m = measure(name, w, args, rank, world, local_rank, debug_layers)
    sub = None
    if world == 8 and args.workload is None and not args.no_sub and not debug_layers:
        # the configuration the metric is quoted on needs all 8 GPUs (654 GB of experts): measured here, reported inside
        # the N = 8 line of the common workload
        ws = dict(WORKLOADS["dsv3-fp8"])
        ms_ = measure("dsv3-fp8", ws, args, rank, world, local_rank)
        if ms_ is not None:
            sub = {"workload": "dsv3-fp8", "config": _config("dsv3-fp8", ws, 8), "tok_s": ms_["tok_s"],
                   "e2e_tok_s": ms_["tok_s_e2e"], "ms_per_step": ms_["ms_per_step"],
                   "step_frac_of_peak": ms_["roofline"]["step_frac_of_peak"], "roofline": ms_["roofline"],
                   "ep_parity_max_err": ms_["ep_parity_max_err"], "finite": ms_["finite"],
                   "breakdown_us_per_layer_eager": ms_["breakdown_us_per_layer_eager"],
                   "launches_per_step": ms_["launches_per_step"]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    B, H = w["batch"], w["H"]
    cb = None
    try:
        cb = cpu_reference_arm(w, 4, 1, budget_s=20.0)
    except Exception as ex:  # the CPU arm must never take the GPU line down
        cb = {"value": None, "unit": "tok/s", "cores": _usable_cores(), "kind": "port",
              "sample": f"failed: {ex!r}", "fits_in_driver_run": False}
    line = {
        "metric": METRIC, "value": m["tok_s"], "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": w["fmt"], "data": "synthetic",
        "config": _config(name, w, args.gpus),
        "clocks": m["clocks"],
        "e2e": {"value": m["tok_s_e2e"], "unit": "tok/s", "h2d_bytes_per_step": B * H * 2, "d2h_bytes_per_step": B * H * 2},
        "gpu_launches": m["launches_per_step"] * args.steps,
        "roofline": m["roofline"],
        "cpu_baseline": {k: cb.get(k) for k in ("value", "unit", "cores", "kind", "sample", "fits_in_driver_run", "candidates")
                         if k in cb},
        "finite": m["finite"],
    }
    line["breakdown_us_per_layer_eager"] = m["breakdown_us_per_layer_eager"]
    line["tensor_pipe_pct"] = ncu_tensor_pipe()
    if world > 1:
        line["ep_parity_max_err"] = m["ep_parity_max_err"]
    if sub is not None:
        line["dsv3_fp8_ep8"] = sub
    if debug_layers:
        line["invalid"] = "debug run with --layers override"
    if args.ep_shard_of:
        line["invalid"] = f"debug run: rank-0 shard of ep{args.ep_shard_of} on one GPU, no all-reduce"
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _config(name, w, n):
    a2a = n > 1 and w["batch"] >= n and w["batch"] % n == 0
    tp = (n > 1 and not a2a and w["fmt"] in ("fp8", "bf16") and w["I"] % n == 0 and (w["I"] // n) % 128 == 0
          and os.environ.get("BENCH_TP") == "1")
    return {"workload": name, "model": w["model"], "weights": w["fmt"], "moe_layers": w["layers"],
            "hidden": w["H"], "intermediate": w["I"], "experts": w["E"], "top_k": w["k"], "batch": w["batch"],
            "kv_seq_len": w["seq"], "attention": w["attn"], "parallelism": (f"tp{n} (experts split along intermediate_size)" if tp else f"ep{n}") if n > 1 else "single",
            "ep_combine": None if n == 1 else
            ("request-sharded attention + NVLink dispatch/combine all-to-all" if a2a
             else "replicated tokens + NVLink all-reduce (lk_moe EP contract)"),
            "router": "fused router kernel (TMA + tcgen05 split-K GEMM + top-k + EP id remap)",
            "l2": "per-step expert+KV traffic (GBs) >> 126 MB L2; no flush needed",
            "graph": "whole step in one CUDA graph"}


if __name__ == "__main__":
    main()
