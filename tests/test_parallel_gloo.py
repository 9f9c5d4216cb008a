"""World-size-2 gloo test (CPU) of the data-parallel step's host logic: flat-buffer packing,
the single sum-allreduce, global-norm clipping and identical updates on every rank.  The
fused CUDA optimiser kernel is replaced by a torch restatement injected by the test (the
product path has no CPU route); its GPU parity is covered by tests/test_gpu_ops.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def host_update(tr):
    g = tr.grad * (1.0 / tr.world)
    norm = g.norm()
    if tr.max_norm > 0:
        g = g * torch.clamp(tr.max_norm / (norm + 1e-6), max=1.0)
    g = g + tr.weight_decay * tr.flat
    b1, b2 = tr.betas
    tr.exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    tr.exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** tr.steps, 1 - b2 ** tr.steps
    denom = tr.exp_avg_sq.sqrt() / (bc2 ** 0.5) + tr.eps
    tr.flat.addcdiv_(tr.exp_avg, denom, value=-tr.lr / bc1)
    return norm.reshape(1)


class Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.Linear(5, 1)
        self.unused = torch.nn.Parameter(torch.ones(3))

    def forward(self, x):
        return self.b(torch.tanh(self.a(x))).pow(2).mean()


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(4, 6, generator=g) * 3.0


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from eve_b200.parallel import FlatAdamTrainer
    model = Toy()
    tr = FlatAdamTrainer(model, lr=1e-2, weight_decay=1e-3, max_norm=0.05)
    for _ in range(3):
        tr.step(model(_data(rank)), update=host_update)
    packed = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    ret[rank] = (packed.clone(), float(tr.last_grad_norm))
    dist.destroy_process_group()


def _worker_diverged(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from eve_b200.parallel import FlatAdamTrainer
    model = Toy()
    with torch.no_grad():           # ranks built under different seeds / checkpoints
        for p in model.parameters():
            p.add_(float(rank))
    FlatAdamTrainer(model, lr=1e-2)
    ret[rank] = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    dist.destroy_process_group()


def test_trainer_broadcasts_rank0_parameters():
    """Replicas that start from different parameters would silently train diverged models: the
    trainer broadcasts rank 0's flat buffer at construction (what DDP does)."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_diverged, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert torch.equal(ret[0], ret[1])
    want = torch.cat([p.detach().reshape(-1) for p in Toy().parameters()])
    assert torch.equal(ret[0], want)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_ranks_equal_one_rank_with_the_concatenated_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    # every rank holds the identical parameters
    assert torch.equal(ret[0][0], ret[1][0])
    assert ret[0][1] == ret[1][1]

    # single process, loss = mean over the two ranks' losses (equal per-rank batch)
    model = Toy()
    params = [p for p in model.parameters()]
    opt = torch.optim.Adam(params, lr=1e-2, weight_decay=1e-3)
    for _ in range(3):
        opt.zero_grad()
        loss = 0.5 * (model(_data(0)) + model(_data(1)))
        loss.backward()
        model.unused.grad = torch.zeros(3)
        norm = torch.nn.utils.clip_grad_norm_(params, 0.05)
        opt.step()
    want = torch.cat([p.detach().reshape(-1) for p in params])
    assert torch.allclose(ret[0][0], want, rtol=1e-5, atol=1e-6)
    assert abs(ret[0][1] - float(norm)) < 1e-5 * float(norm)


def test_parameters_are_views_of_the_flat_buffer_and_keep_their_names():
    from eve_b200.parallel import FlatAdamTrainer
    model = Toy()
    keys = list(model.state_dict().keys())
    tr = FlatAdamTrainer(model, lr=1e-3, weight_decay=0.0, max_norm=0.0)
    assert list(model.state_dict().keys()) == keys
    tr.flat.zero_()
    assert all(float(p.abs().max()) == 0.0 for p in model.parameters())
    assert tr.flat.numel() >= sum(p.numel() for p in model.parameters())
    assert all((p.data_ptr() - tr.flat.data_ptr()) % 256 == 0 for p in model.parameters())
