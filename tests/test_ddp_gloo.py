"""Host-side logic of the data-parallel step on CPU with gloo, world_size 2: batch sharding,
the flat gradient bucket (dead Q3 parameters excluded, 16-byte aligned slices) and the single all-reduce,
whose result must equal the mean of the per-rank gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tatt_b200.train import GradBucket, shard_batch


def test_shard_batch_covers_everything_once():
    for n, world in ((1024, 8), (10, 4), (7, 2), (3, 8)):
        got = []
        for r in range(world):
            a, b = shard_batch(n, r, world)
            assert 0 <= a <= b <= n
            got.extend(range(a, b))
        assert got == list(range(n))
        sizes = [shard_batch(n, r, world)[1] - shard_batch(n, r, world)[0] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert shard_batch(1024, 3, 8) == (384, 512)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tatt_b200
        torch.manual_seed(1234)                                   # identical replicas
        net = tatt_b200.TSRN_TL_TRANS(scale_factor=2, width=32, height=16, STN=False)
        dead = {"infoGen.fc_feature_in.weight", "infoGen.fc_feature_in.bias"}
        dead |= {n for n, _ in net.named_parameters() if ".self_attn." in n and "decoder" in n}
        dead |= {n for n, _ in net.named_parameters() if "decoder.layers" in n and ".norm1." in n}
        g = torch.Generator().manual_seed(100 + rank)             # rank-specific synthetic gradients
        for n, p in net.named_parameters():
            p.grad = None if n in dead else torch.randn(p.shape, generator=g)
        before = {n: p.detach().clone() for n, p in net.named_parameters()}
        bucket = GradBucket.from_model_after_backward(net, flatten_params=True)
        assert len(bucket.params) == sum(1 for n, _ in net.named_parameters() if n not in dead)
        assert all(o % 4 == 0 for o in bucket.offsets)
        for n, p in net.named_parameters():                       # flattening must not change values
            assert torch.equal(p.detach(), before[n]), n
        local = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
        bucket.pack()
        bucket.allreduce_mean()
        bucket.unpack_into_grads()
        index_of = {id(q): i for i, q in enumerate(bucket.params)}
        # reference: gather every rank's gradients with plain collectives and average
        for n, p in net.named_parameters():
            if n in dead:
                assert p.grad is None
                continue
            parts = [torch.empty_like(local[n]) for _ in range(world)]
            dist.all_gather(parts, local[n])
            want = torch.stack(parts).mean(0)
            assert torch.allclose(p.grad, want, rtol=1e-6, atol=1e-7), n
            assert p.grad.data_ptr() == bucket.flat_grad.data_ptr() + 4 * bucket.offsets[index_of[id(p)]]
        # parameters are views of one flat buffer (what the fused optimizer kernel updates)
        p0 = bucket.params[0]
        bucket.flat_param[bucket.offsets[0]] += 1.0
        name0 = [n for n, q in net.named_parameters() if q is p0][0]
        assert abs(p0.detach().reshape(-1)[0].item() - (before[name0].reshape(-1)[0].item() + 1.0)) < 1e-5
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
