"""World-size-2 check of the tensor-parallel host logic on CPU (gloo): the row/column slices that
jlama_b200.model cuts from DistributedContext (model/DistributedContext.java:60-98, LlamaModel.java:120-133) plus an
all-reduce SUM of the per-rank partial products (JlamaService.combine, jlama-net .../JlamaService.java:300-359, which the
GPU path replaces with NCCL) reproduce the unsharded projection.  No GPU, no CUDA calls."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from jlama_b200 import synth  # noqa: E402
from jlama_b200.model import DistributedContext, _slice_cols, _slice_rows  # noqa: E402
from jlama_b200.native import Q4  # noqa: E402
from oracle import oracle as o  # noqa: E402


def _deq(t):
    dt, data, scales = t
    assert dt == Q4
    return o.dequantize_q4(data, scales).astype(np.float64)


def _silu(x):
    return x / (1.0 + np.exp(-x))


def _worker(rank, world, port, cfg_name, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = synth.get_config(cfg_name)
        w = synth.make_weights(cfg, wdtype=Q4, mode="quantize")
        d = DistributedContext(cfg, rank, world)
        rng = np.random.default_rng(7)
        x = rng.standard_normal(cfg["E"])
        att = rng.standard_normal(cfg["E"])  # concatenated head outputs (heads * head_size == E for these configs)
        b = "model.layers.0."
        # MLP: rows of gate/up, columns of down (MLPBlock.java:117-160)
        g = _deq(_slice_rows(w[b + "mlp.gate_proj.weight"], d.hiddenSegmentStart, d.hiddenSegmentLength))
        u = _deq(_slice_rows(w[b + "mlp.up_proj.weight"], d.hiddenSegmentStart, d.hiddenSegmentLength))
        dn = _deq(_slice_cols(w[b + "mlp.down_proj.weight"], d.hiddenSegmentStart, d.hiddenSegmentLength))
        part_mlp = dn @ (_silu(g @ x) * (u @ x))
        # attention output projection: columns of o_proj over this rank's heads (CausalSelfAttention.java:363-378)
        wo = _deq(_slice_cols(w[b + "self_attn.o_proj.weight"], d.attentionSegmentStart, d.attentionSegmentLength))
        part_o = wo @ att[d.attentionSegmentStart:d.attentionSegmentStart + d.attentionSegmentLength]
        t = torch.from_numpy(np.concatenate([part_mlp, part_o]))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        # the segments of all ranks must tile the full ranges exactly
        seg = torch.tensor([d.hiddenSegmentStart, d.hiddenSegmentLength, d.attentionSegmentStart, d.attentionSegmentLength,
                            d.kvSegmentStart, d.kvSegmentLength], dtype=torch.int64)
        segs = [torch.zeros_like(seg) for _ in range(world)]
        dist.all_gather(segs, seg)
        if rank == 0:
            full_mlp = _deq(w[b + "mlp.down_proj.weight"]) @ (
                _silu(_deq(w[b + "mlp.gate_proj.weight"]) @ x) * (_deq(w[b + "mlp.up_proj.weight"]) @ x))
            full_o = _deq(w[b + "self_attn.o_proj.weight"]) @ att
            ref = np.concatenate([full_mlp, full_o])
            err = float(np.abs(t.numpy() - ref).max() / np.abs(ref).max())
            out_q.put(("ok", err, [s.tolist() for s in segs]))
    except Exception as e:  # surfaced in the parent
        if rank == 0:
            out_q.put(("error", repr(e), None))
        raise
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("cfg_name,world", [("tiny", 2), ("small", 4)])
def test_shards_and_allreduce_reproduce_the_unsharded_layer(cfg_name, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, err, segs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert status == "ok", err
    assert err < 1e-12  # float64 arithmetic, only the summation order differs
    cfg = synth.get_config(cfg_name)
    hs = cfg["E"] // cfg["heads"]
    # hidden rows tile [0, H), attention columns tile [0, heads*hs), kv rows tile [0, kv_heads*hs), in rank order
    for col, total in ((0, cfg["H"]), (2, cfg["heads"] * hs), (4, cfg["kv_heads"] * hs)):
        pos = 0
        for seg in segs:
            assert seg[col] == pos
            pos += seg[col + 1]
        assert pos == total
