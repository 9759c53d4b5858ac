"""Data-parallel plumbing for the hot path: one process per GPU, batch sharding, ONE collective per
step -- a mean all-reduce of the gradient vector -- plus the rank-0 broadcast of BatchNorm buffers that the reference
gets from DistributedDataParallel (bts_main.py:352: C1 + C3 in SURVEY 2.5).

The drop-in boundary is the module, so an unchanged bts_main.py keeps using torch's DDP.  `FlatGradReducer` is the
B200-native alternative used by bench.py --reducer flat: gradients are copied into ONE flat, persistent fp32 buffer
(a single registered NCCL buffer instead of 25 MB buckets; over NVSwitch the all-reduce cost is launch latency, not
link count) and averaged with a single all_reduce.  Works with the gloo backend on CPU tensors, which is how
tests/test_dist_cpu.py covers it.
"""
import torch
import torch.distributed as dist


def shard_batch(tensors, rank, world):
    """contiguous batch shards, like DistributedSampler + bts_main.py:351 (global batch / ngpus per rank)"""
    out = []
    for t in tensors:
        n = t.shape[0]
        if n % world:
            raise ValueError("global batch %d is not divisible by world size %d" % (n, world))
        per = n // world
        out.append(t[rank * per:(rank + 1) * per])
    return out


class FlatGradReducer:
    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    @torch.no_grad()
    def reduce(self):
        """mean over ranks of every gradient (parameters without a gradient contribute zeros, like DDP with
        find_unused_parameters=True)"""
        world = dist.get_world_size(self.group)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


@torch.no_grad()
def broadcast_buffers(module, src=0, group=None):
    """rank-0 BatchNorm running statistics to every rank (DDP's broadcast_buffers, C3)"""
    for b in module.buffers():
        dist.broadcast(b, src=src, group=group)
