"""Multi-process tests of the expert-parallel path.
CPU (gloo, world_size 2): handle exchange + expert map / id remap logic the N>1 bench relies on.
GPU (-m gpu, needs >= 2 GPUs): the hand-written NVLink all-reduce against a torch sum, eager and under CUDA graphs,
a 2-rank EP MoE layer against the single-rank result, and the dispatch / combine all-to-all
(token-sharded callers) against the same oracle, eager and under CUDA-graph replay."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lvllm_b200.ep import exchange_handles
    from oracle import moe_oracle as O
    handles = exchange_handles(bytes([rank]) * 128)
    ok = [h == bytes([r]) * 128 for r, h in enumerate(handles)]
    # linear expert map: every global expert is local on exactly one rank; remapped ids partition the slots
    E, k, M = 10, 3, 7
    local, emap = O.determine_expert_map(world, rank, E)
    g = torch.Generator().manual_seed(0)
    ids = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(M)]).int()
    loc = O.global_to_local_expert_ids(ids, emap)
    mine = (loc >= 0).int()
    tot = mine.clone()
    dist.all_reduce(tot)
    q.put((rank, all(ok), bool((tot == 1).all()), int(local)))
    dist.destroy_process_group()


def test_ep_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_cpu_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
    assert all(r[1] and r[2] for r in res)
    assert sum(r[3] for r in res) == 10


def _allreduce_norm_check(ep, dev, world, H, g_private):
    """b200_ep_allreduce_norm against dist.all_reduce + the reference's RMSNorm.forward_native arithmetic (fp32 add of
    the residual, residual_out = cast(x), y = cast(x * rsqrt(var + eps)) * weight); eager epochs and CUDA-graph replay.
    Returns the worst relative error."""
    gs = torch.Generator(device=dev).manual_seed(77)   # replicated tensors: same seed on every rank
    worst = 0.0

    def ref_of(x, res, gamma, gain):
        from oracle import moe_oracle as O   # checker only
        tot = x.clone()
        if world > 1:
            dist.all_reduce(tot)
        # oracle.moe_sum_add_rms_norm = the reference's fused_add_rms_norm (vllm/ir/ops/layernorm.py:44-63) on the fp32
        # sum; pinned to the reference's compiled CPU kernel by tests/test_c_port.py
        return O.moe_sum_add_rms_norm(tot, res, gamma, gain, 1e-6)

    for n_tok, use_res, use_gamma in ((1, True, True), (8, True, True), (5, False, False), (8, True, False)):
        x = torch.randn(n_tok, H, device=dev, generator=g_private)
        res = torch.randn(n_tok, H, device=dev, generator=gs).bfloat16() if use_res else None
        gamma = (torch.rand(H, device=dev, generator=gs) + 0.5).bfloat16() if use_gamma else None
        y_ref, res_ref, sum_ref = ref_of(x, res, gamma, 0.1)
        out = torch.empty(n_tok, H, dtype=torch.bfloat16, device=dev)
        res_io = res.clone() if res is not None else None
        ssum = torch.empty(n_tok, H, device=dev)
        ep.allreduce_norm(x, out, residual=res_io, gamma=gamma, gain=0.1, eps=1e-6, sum_out=ssum)
        torch.cuda.synchronize()
        worst = max(worst, float((out.float() - y_ref.float()).abs().max() / y_ref.float().abs().max()))
        worst = max(worst, float((ssum - sum_ref).abs().max() / sum_ref.abs().max()))
        if res is not None:
            worst = max(worst, float((res_io.float() - res_ref.float()).abs().max() / res_ref.float().abs().max()))
    # graph replay
    x = torch.randn(8, H, device=dev, generator=g_private)
    res0 = torch.randn(8, H, device=dev, generator=gs).bfloat16()
    res_io = res0.clone()
    out = torch.empty(8, H, dtype=torch.bfloat16, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ep.allreduce_norm(x, out, residual=res_io, gain=0.1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            ep.allreduce_norm(x, out, residual=res_io, gain=0.1)
    for _ in range(3):
        x.copy_(torch.randn(8, H, device=dev, generator=g_private))
        res_io.copy_(res0)
        y_ref, _, _ = ref_of(x, res0, None, 0.1)
        gr.replay()
        torch.cuda.synchronize()
        worst = max(worst, float((out.float() - y_ref.float()).abs().max() / y_ref.float().abs().max()))
    return worst


def _gpu_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import lk_moe
    from lvllm_b200 import ops
    from lvllm_b200.ep import EpGroup
    from oracle import moe_oracle as O
    H, M = 1024, 8
    ep = EpGroup(rank, world, dev, max_elems=M * H)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    errs = []
    # eager all-reduce, several epochs, different sizes
    for it in range(5):
        n_tok = [1, 8, 3, 8, 2][it]
        x = torch.randn(n_tok, H, device=dev, generator=g)
        ref = x.clone()
        dist.all_reduce(ref)
        out = ep.allreduce(x).clone()
        errs.append(float((out - ref).abs().max()))
    # bit-identical on every rank
    x = torch.randn(M, H, device=dev, generator=g)
    out = ep.allreduce(x).clone()
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    # CUDA-graph replay (epochs live in device memory)
    xs = torch.randn(M, H, device=dev, generator=g)
    ybuf = torch.zeros(M, H, device=dev, dtype=torch.bfloat16)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ep.allreduce(xs, ybuf)
        torch.cuda.synchronize()
        dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            ep.allreduce(xs, ybuf)
    graph_err = []
    for it in range(4):
        xs.copy_(torch.randn(M, H, device=dev, generator=g))
        ref = xs.clone()
        dist.all_reduce(ref)
        gr.replay()
        torch.cuda.synchronize()
        graph_err.append(float((ybuf.float() - ref).abs().max() / ref.abs().max()))
    # push all-reduce fused with residual add + RMSNorm (replicated residual / gamma, rank-private partials)
    norm_err = _allreduce_norm_check(ep, dev, world, H, g)
    # EP MoE layer: experts split over ranks, tokens replicated, sum over ranks == single-rank oracle
    E, k, Hm, I = 8, 2, 512, 256
    cg = torch.Generator().manual_seed(5)
    hidden = (torch.randn(M, Hm, generator=cg) / 10).bfloat16()
    w13 = (torch.randn(E, 2 * I, Hm, generator=cg) / 10).bfloat16()
    w2 = (torch.randn(E, Hm, I, generator=cg) / 10).bfloat16()
    tw, ids = torch.topk(torch.softmax(torch.randn(M, E, generator=cg), -1), k)
    ref = O.experts_forward_batched(hidden, O.DequantExperts(w13.float(), w2.float()), ids.int(), tw.float())
    local, emap = O.determine_expert_map(world, rank, E)
    lo = int((emap >= 0).nonzero()[0])
    cfg = lk_moe.MOEConfigV2()
    cfg.num_processes, cfg.process_id, cfg.gpu_id = world, rank, rank
    cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = local, k, Hm, I
    cfg.max_batch_size, cfg.max_num_seqs = 64, 16
    w13l, w2l = w13[lo:lo + local].contiguous(), w2[lo:lo + local].contiguous()
    moe = lk_moe.MOE_BF16(cfg, w13l.data_ptr(), w2l.data_ptr(), 0, 0, 0, 0)
    lids = ops.global_to_local_expert_ids(ids.int().to(dev), emap.to(dev))
    part = torch.zeros(M, Hm, device=dev)
    hd, twd = hidden.to(dev), tw.float().to(dev)
    moe.cpu_decode(torch.cuda.current_stream().cuda_stream, M, k, hd.data_ptr(), lids.data_ptr(), twd.data_ptr(), part.data_ptr())
    ep2 = EpGroup(rank, world, dev, max_elems=M * Hm)
    tot = ep2.allreduce(part).clone()
    moe_err = float((tot.cpu() - ref).abs().max())
    # dispatch / combine all-to-all: tokens sharded over ranks (DP attention), experts sharded over ranks (EP);
    # three chained "layers" per CUDA graph exercise the single-buffer reuse protected by the two flag barriers
    m_local = M // world
    ids_g = ids.int().clone()
    ids_g[1, 0] = -1                                   # a padded slot (reference: ids < 0 are skipped)
    tw_pad = tw.float().clone()
    ref2 = O.experts_forward_batched(hidden, O.DequantExperts(w13.float(), w2.float()), ids_g, tw_pad)
    ep2.a2a_init(m_local, Hm, k, local)
    sl = slice(rank * m_local, (rank + 1) * m_local)
    h_loc, ids_loc, tw_loc = hd[sl].contiguous(), ids_g[sl].contiguous().to(dev), tw_pad[sl].contiguous().to(dev)
    outs = [torch.zeros(m_local, Hm, device=dev) for _ in range(3)]

    def a2a_layer(o):
        ep2.dispatch(h_loc, ids_loc, tw_loc)
        moe.cpu_decode(torch.cuda.current_stream().cuda_stream, M, k, ep2.x_ptr, ep2.ids_ptr, ep2.w_ptr, ep2.y_ptr)
        ep2.combine(ids_loc, o)

    a2a_layer(outs[0])
    torch.cuda.synchronize()
    a2a_err = float((outs[0].cpu() - ref2[sl]).abs().max())
    # combine fused with residual + RMSNorm against the unfused chain
    res0 = torch.randn(m_local, Hm, device=dev, generator=torch.Generator(device=dev).manual_seed(3)).bfloat16()
    res_io, on = res0.clone(), torch.empty(m_local, Hm, dtype=torch.bfloat16, device=dev)
    ep2.dispatch(h_loc, ids_loc, tw_loc)
    moe.cpu_decode(torch.cuda.current_stream().cuda_stream, M, k, ep2.x_ptr, ep2.ids_ptr, ep2.w_ptr, ep2.y_ptr)
    ep2.combine_norm(ids_loc, on, residual=res_io, gain=0.5)
    torch.cuda.synchronize()
    xr = outs[0] + res0.float()
    yn = (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6) * 0.5).bfloat16()
    a2a_err = max(a2a_err, float((on.float() - yn.float()).abs().max() / yn.float().abs().max()) * 0.1,
                  float((res_io.float() - xr.bfloat16().float()).abs().max() / xr.abs().max()) * 0.1)
    s2 = torch.cuda.Stream()
    with torch.cuda.stream(s2):
        for o in outs:
            a2a_layer(o)
        torch.cuda.synchronize()
        dist.barrier()
        gr2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr2, stream=s2):
            for o in outs:
                a2a_layer(o)
    for it in range(3):
        for o in outs:
            o.zero_()
        gr2.replay()
        torch.cuda.synchronize()
        a2a_err = max(a2a_err, max(float((o.cpu() - ref2[sl]).abs().max()) for o in outs))
    q.put((rank, max(errs), same, max(graph_err), moe_err, a2a_err, norm_err))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_ep_allreduce_and_moe_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    ps = [ctx.Process(target=_gpu_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in ps:
        p.join(60)
    for rank, err, same, gerr, moe_err, a2a_err, norm_err in res:
        assert norm_err < 1e-2, f"rank {rank}: all-reduce + residual + RMSNorm err {norm_err}"
        assert err < 1e-5, f"rank {rank}: all-reduce err {err}"
        assert same, "all-reduce results differ between ranks"
        assert gerr < 1e-2, f"rank {rank}: graph replay err {gerr}"
        assert moe_err < 5e-3, f"rank {rank}: EP MoE err {moe_err}"
        assert a2a_err < 5e-3, f"rank {rank}: EP dispatch/combine err {a2a_err}"


@pytest.mark.gpu
def test_ep_dispatch_combine_single_rank():
    """world = 1: dispatch pushes every row into this rank's own slots and combine pulls them back, so the
    all-to-all pair around cpu_decode must reproduce a plain cpu_decode (runs on a 1-GPU box)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    sys.path.insert(0, ROOT)
    import lk_moe
    from lvllm_b200.ep import EpGroup
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1)
    try:
        E, k, H, I, M = 8, 2, 512, 256, 6
        g = torch.Generator().manual_seed(3)
        hidden = (torch.randn(M, H, generator=g) / 10).bfloat16().to(dev)
        w13 = (torch.randn(E, 2 * I, H, generator=g) / 10).bfloat16()
        w2 = (torch.randn(E, H, I, generator=g) / 10).bfloat16()
        tw, ids = torch.topk(torch.softmax(torch.randn(M, E, generator=g), -1), k)
        ids = ids.int()
        ids[2, 1] = -1
        ids, tw = ids.to(dev), tw.float().to(dev)
        cfg = lk_moe.MOEConfigV2()
        cfg.num_processes, cfg.process_id, cfg.gpu_id = 1, 0, 0
        cfg.expert_num, cfg.top_k, cfg.hidden_size, cfg.intermediate_size = E, k, H, I
        cfg.max_batch_size, cfg.max_num_seqs = 64, 16
        moe = lk_moe.MOE_BF16(cfg, w13.data_ptr(), w2.data_ptr(), 0, 0, 0, 0)
        st = torch.cuda.current_stream().cuda_stream
        ref = torch.zeros(M, H, device=dev)
        moe.cpu_decode(st, M, k, hidden.data_ptr(), ids.data_ptr(), tw.data_ptr(), ref.data_ptr())
        ep = EpGroup(0, 1, dev, max_elems=M * H)
        ep.a2a_init(M, H, k, E)
        out = torch.zeros(M, H, device=dev)
        for _ in range(3):   # epochs advance, buffers are reused
            out.zero_()
            ep.dispatch(hidden, ids, tw)
            moe.cpu_decode(st, M, k, ep.x_ptr, ep.ids_ptr, ep.w_ptr, ep.y_ptr)
            ep.combine(ids, out)
            torch.cuda.synchronize()
            assert torch.equal(out, ref)
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.gpu
def test_ep_allreduce_norm_single_rank():
    """world = 1: the push all-reduce + residual + RMSNorm kernel degenerates to fused_add_rms_norm of the local
    partial (runs on a 1-GPU box; the 2-rank form is covered by test_ep_allreduce_and_moe_two_gpus and bench.py)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    sys.path.insert(0, ROOT)
    from lvllm_b200.ep import EpGroup
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1)
    try:
        for H in (1024, 7168):
            ep = EpGroup(0, 1, dev, max_elems=8 * H)
            err = _allreduce_norm_check(ep, dev, 1, H, torch.Generator(device=dev).manual_seed(9))
            assert err < 1e-2, f"H={H}: {err}"
    finally:
        if created:
            dist.destroy_process_group()
