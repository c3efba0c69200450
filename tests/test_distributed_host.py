"""world_size-2 gloo test of the only cross-rank step of the path: all-reduce of the energy
moments and of the parameter gradient (fermiflow_b200.VMC)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

torch.set_default_dtype(torch.float64)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fermiflow_b200.VMC import _global_mean_std, _VMCBase
    g = torch.Generator().manual_seed(5)
    full = torch.randn(64, generator=g)
    mine = full[rank::world].clone()
    mean, std, cnt = _global_mean_std(mine)

    class M(_VMCBase):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.zeros(3))
            self.b = torch.nn.Parameter(torch.zeros(2, 2))
    m = M()
    m.a.grad = torch.full((3,), float(rank + 1))
    m.b.grad = torch.full((2, 2), float(10 * (rank + 1)))
    m.allreduce_gradients()
    if rank == 0:
        out.put((mean, std, cnt, float(full.mean()), float(full.std()), m.a.grad.tolist(), m.b.grad.tolist()))
    dist.destroy_process_group()


def test_moments_and_gradient_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    mean, std, cnt, fmean, fstd, ga, gb = res
    assert cnt == 64
    assert abs(mean - fmean) < 1e-14 and abs(std - fstd) < 1e-13
    assert ga == [3.0, 3.0, 3.0] and gb == [[30.0, 30.0], [30.0, 30.0]]
