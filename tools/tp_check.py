"""Tensor-parallel parity check (run under torchrun on N GPUs of one box):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/tp_check.py [config]
Every rank loads its DistributedContext shard (jlama-net model-shard split), the ranks all-reduce the o_proj / down_proj
partial sums over NCCL (replacing JlamaService.combine), and rank 0 compares tokens and logits with the CPU oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    ctx = native.Context(local)
    idbuf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        ctx.check(ctx.lib.jl_comm_unique_id(ctx.h, native.ptr(idbuf)))
    t = torch.from_numpy(idbuf).cuda()
    dist.broadcast(t, 0)
    idbuf = t.cpu().numpy()
    ctx.check(ctx.lib.jl_comm_init(ctx.h, native.ptr(idbuf), rank, world))
    # the collective itself
    buf = np.full(1000, float(rank + 1), dtype=np.float32)
    ctx.check(ctx.lib.jl_comm_allreduce_f32(ctx.h, native.ptr(buf), buf.size))
    assert np.all(buf == world * (world + 1) / 2), buf[:4]
    cfg = synth.get_config(name)
    w = synth.make_weights(cfg)
    prompt = synth.random_prompt(cfg, 19)
    m = LlamaModel(ctx, cfg, w, tp_rank=rank, tp_size=world)
    toks, logits = m.generate(prompt, 16, want_logits=True)
    ok = True
    if rank == 0:
        from oracle import oracle as o
        om = o.OracleLlama(cfg, w, act_q8=True, tp=world)
        ot, ol = om.generate(prompt, 16)
        rel = max(float(np.abs(logits[i] - ol[i]).max() / np.abs(ol[i]).max()) for i in range(16))
        ok = list(toks) == list(ot) and rel <= 1e-2
        print("tp=%d %s: tokens_equal=%s max_logit_rel_err=%.3e -> %s" % (world, name, list(toks) == list(ot), rel, "OK" if ok else "FAIL"), flush=True)
    # all ranks must have produced the same tokens
    tt = torch.from_numpy(np.asarray(toks, dtype=np.int64)).cuda()
    ref = tt.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(tt, ref), "ranks disagree on tokens"
    m.close()
    dist.barrier()
    ctx.lib.jl_comm_destroy(ctx.h)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
