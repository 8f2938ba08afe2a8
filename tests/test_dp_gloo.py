"""Data-parallel plumbing on CPU with gloo, world_size 2: replicas start identical, the
all-reduced flat gradient equals the single-process full-batch gradient, batch shards tile the global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from erv_b200.parallel import BucketedReducer, FlatParams, shard_slice


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(seed):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.LayerNorm(5), torch.nn.Linear(5, 3))
    m.register_buffer("omega", torch.randn(2, 4))
    return m


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model(seed=100 + rank)  # deliberately different replicas
        fp = FlatParams(model.parameters())
        fp.broadcast(model.buffers(), src=0)
        g = torch.Generator().manual_seed(0)
        x, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
        sl = shard_slice(8, rank, world)
        fp.zero_grad()
        # sum-reduction loss per shard, so the summed gradient is the full-batch sum
        torch.nn.functional.cross_entropy(model(x[sl]), y[sl], reduction="sum").backward()
        fp.allreduce_grad()
        out[rank] = (fp.flat.clone(), fp.grad.clone(), model.omega.clone(), (sl.start, sl.stop))
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_matches_single_process():
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    ref = _model(seed=100)
    fp = FlatParams(ref.parameters())
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
    torch.nn.functional.cross_entropy(ref(x), y, reduction="sum").backward()
    assert out[0][3] == (0, 4) and out[1][3] == (4, 8)
    for r in range(world):
        flat, grad, omega, _ = out[r]
        assert torch.equal(flat, fp.flat), "replicas must equal rank 0 after broadcast"
        assert torch.equal(omega, ref.omega)
        assert torch.allclose(grad, fp.grad, atol=1e-6)


def test_flat_params_are_views_and_shards_validate():
    m = _model(1)
    fp = FlatParams(m.parameters())
    n_par = sum(p.numel() for p in m.parameters())
    assert n_par <= fp.numel() < n_par + 4 * len(fp.params)  # every parameter starts on a 16-byte boundary
    assert all(o % 4 == 0 for o in fp.offsets) and all(p.data_ptr() % 16 == 0 for p in fp.params)
    m[0].weight.data.fill_(2.0)
    assert (fp.flat[:30] == 2.0).all()           # parameter storage is the flat buffer
    m(torch.randn(4, 6)).sum().backward()
    assert fp.grad.abs().sum() > 0               # autograd accumulated into the flat gradient
    fp.zero_grad()
    assert m[0].weight.grad.abs().sum() == 0
    with pytest.raises(ValueError):
        shard_slice(10, 0, 4)


class _TinyViT(torch.nn.Module):
    """Parameter order and cut point of erv_b200.vit.BaseViT: embedding, blocks, head."""

    def __init__(self):
        super().__init__()
        self.pos = torch.nn.Parameter(torch.randn(1, 6) * 0.1)
        self.patch_embedding = torch.nn.Linear(6, 6)
        self.transformer_blocks = torch.nn.ModuleList(
            [torch.nn.Sequential(torch.nn.LayerNorm(6), torch.nn.Linear(6, 6), torch.nn.GELU()) for _ in range(3)])
        self.mlp_head = torch.nn.Sequential(torch.nn.LayerNorm(6), torch.nn.Linear(6, 3))

    def forward(self, x):
        x = self.patch_embedding(x) + self.pos
        for i, b in enumerate(self.transformer_blocks):
            x = x + b(x)
            if getattr(self, "_cut_after", None) == i:
                self._cut_tensor = x
        return self.mlp_head(x)


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ERV_BUCKET_ALLREDUCE="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(5)
        model = _TinyViT()
        fp = FlatParams(model.parameters())
        red = BucketedReducer(model, fp)
        assert red.split is not None and 0 < red.split[1] < fp.numel()
        g = torch.Generator().manual_seed(0)
        x, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
        sl = shard_slice(8, rank, world)
        for _ in range(2):  # twice: the cut tensor is re-recorded by every forward
            fp.zero_grad()
            red.backward_and_reduce(torch.nn.functional.cross_entropy(model(x[sl]), y[sl], reduction="sum"))
        out[rank] = (fp.grad.clone(), red.split)
    finally:
        dist.destroy_process_group()


def test_two_bucket_backward_matches_single_process():
    """The split backward + two all-reduces (first bucket launched before block 0's backward) give the full-batch gradient."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_bucket_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(5)
    ref = _TinyViT()
    fp = FlatParams(ref.parameters())
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
    torch.nn.functional.cross_entropy(ref(x), y, reduction="sum").backward()
    for r in range(world):
        assert torch.allclose(out[r][0], fp.grad, atol=1e-5), (out[r][0] - fp.grad).abs().max()
