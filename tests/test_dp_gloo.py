"""CPU, world_size 2 over gloo: the data-parallel path (tools.GradSync inside tools.Optimizer)
gives every rank the gradient of the concatenated batch, so replicas stay bit-identical to each
other and match a single-process step on the global batch."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.SiLU(), torch.nn.Linear(32, 3))


def _data(world):
    g = torch.Generator().manual_seed(1)
    return torch.randn(world * 8, 12, generator=g), torch.randn(world * 8, 3, generator=g)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("dreamerv3-torch_b200")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _model()
    opt = pkg.tools.Optimizer("m", model.parameters(), lr=1e-2, eps=1e-8, clip=100.0, wd=0.0,
                              grad_sync=pkg.tools.GradSync())
    x, y = _data(world)
    xs, ys = x[rank * 8:(rank + 1) * 8], y[rank * 8:(rank + 1) * 8]
    for _ in range(3):
        met = opt(((model(xs) - ys) ** 2).mean())
    q.put((rank, [p.detach().numpy().copy() for p in model.parameters()], float(met["m_grad_norm"])))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_allreduce_matches_global_batch(pkg):
    world, port = 2, 29631
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process on the concatenated batch
    model = _model()
    opt = pkg.tools.Optimizer("m", model.parameters(), lr=1e-2, eps=1e-8, clip=100.0, wd=0.0)
    x, y = _data(world)
    for _ in range(3):
        met = opt(((model(x) - y) ** 2).mean())
    for a, b in zip(got[0][1], got[1][1]):
        assert (a == b).all()                          # replicas identical
    for a, b in zip(got[0][1], model.parameters()):
        assert float((torch.from_numpy(a) - b.detach()).abs().max()) < 1e-6   # == global-batch step
    assert abs(got[0][2] - float(met["m_grad_norm"])) < 1e-5


def test_to_host_batches_scalars(pkg):
    m = {"a": torch.tensor(1.5), "b": torch.ones(2, 3), "c": 0.25, "d": torch.tensor([2.0])}
    out = pkg.tools.to_host(m)
    assert list(out) == ["a", "b", "c", "d"]
    assert float(out["a"]) == 1.5 and out["b"].shape == (2, 3) and out["c"] == 0.25 and float(out["d"]) == 2.0
