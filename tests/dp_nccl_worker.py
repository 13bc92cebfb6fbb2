"""Worker of tests/test_gpu_dp_nccl.py (launched under torch.distributed.run, one rank per GPU).

Every rank builds the same tiny agent, takes its slice of a global replay batch and of the global
noise, and runs one full train step (WorldModel._train + ImagBehavior._train) with the flat fused
Adam and the NCCL ``GradSync``.  Rank 0 then runs the single-GPU step on the concatenated batch
and compares the averaged flat gradients, the grad norms and the updated parameters.
``reward_EMA`` is off: its 5/95 % quantiles are taken per rank, the one
non-linear cross-sample statistic of the step."""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def agent(pkg, device, sync=None):
    cfgs = pkg.configs
    torch.manual_seed(0)
    cfg = cfgs.make_config("dmc_proprio", device=device, dyn_stoch=8, dyn_discrete=8, dyn_deter=64,
                           dyn_hidden=64, units=64, imag_horizon=5, reward_EMA=False,
                           encoder=dict(mlp_units=64, mlp_layers=2), decoder=dict(mlp_units=64, mlp_layers=2))
    wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg, grad_sync=sync)
    beh = pkg.models.ImagBehavior(cfg, wm, grad_sync=sync)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():                       # reward / critic output layers are zero-initialised
        for p in list(wm.parameters()) + list(beh.parameters()):
            p.add_(0.05 * torch.randn(p.shape, generator=g).to(device))
    beh._slow_value.load_state_dict(beh.value.state_dict())
    return cfg, wm, beh


def batch(B, T, A):
    rs = np.random.RandomState(5)
    host = {k: rs.randn(B, T, n).astype(np.float32) for k, n in (("orientations", 14), ("height", 1), ("velocity", 9))}
    host["action"] = rs.uniform(-1, 1, size=(B, T, A)).astype(np.float32)
    host["reward"] = rs.randn(B, T).astype(np.float32)
    host["discount"] = np.ones((B, T), np.float32)
    host["is_terminal"] = np.zeros((B, T), np.float32)
    host["is_first"] = np.zeros((B, T), np.float32)
    host["is_first"][:, 0] = 1
    host["is_first"][1, T // 2] = 1
    return host


def step(wm, beh, data, noise):
    reward = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
    post, _, m1 = wm._train(data, noise=(noise["u_prior"], noise["u_post"]))
    _, _, _, _, m2 = beh._train(post, reward, noise=(noise["act_noise"], noise["u_state"]))
    return {**m1, **m2}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(device))
    pkg = importlib.import_module("dreamerv3-torch_b200")
    Bh, T = 3, 8                                 # per-rank batch; global batch = world * Bh
    B = Bh * world
    cfg, wm, beh = agent(pkg, device, pkg.tools.GradSync())
    A, S, C, H = cfg.num_actions, cfg.dyn_stoch, cfg.dyn_discrete, cfg.imag_horizon
    g = torch.Generator().manual_seed(7)
    u = lambda *s: torch.rand(*s, generator=g).clamp_(1e-30, 1.0)
    noise = dict(u_prior=u(T, B, S, C), u_post=u(T, B, S, C), act_noise=torch.randn(H, B * T, A, generator=g),
                 u_state=u(H, B * T, S, C))
    data = batch(B, T, A)
    rows = slice(rank * Bh, (rank + 1) * Bh)
    flat = slice(rank * Bh * T, (rank + 1) * Bh * T)
    my_noise = dict(u_prior=noise["u_prior"][:, rows], u_post=noise["u_post"][:, rows],
                    act_noise=noise["act_noise"][:, flat], u_state=noise["u_state"][:, flat])
    my_noise = {k: v.contiguous().to(device) for k, v in my_noise.items()}
    m = step(wm, beh, {k: v[rows].copy() for k, v in data.items()}, my_noise)
    torch.cuda.synchronize()
    # replicas identical after the step
    mine = torch.cat([p.detach().reshape(-1) for p in list(wm.parameters()) + list(beh.parameters())])
    other = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(other, mine)
    same = all(torch.equal(other[0], o) for o in other)
    result = {"replicas_identical": bool(same)}
    if rank == 0:
        _, wm1, beh1 = agent(pkg, device, None)
        m1 = step(wm1, beh1, data, {k: v.to(device) for k, v in noise.items()})
        rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-12))
        for name, o, o1 in (("model", wm._model_opt, wm1._model_opt), ("actor", beh._actor_opt, beh1._actor_opt),
                            ("value", beh._value_opt, beh1._value_opt)):
            result[f"{name}_flat_grad_rel"] = rel(o._fg, o1._fg)
            result[f"{name}_param_maxabs"] = float((o._fp - o1._fp).abs().max())
            k = f"{name}_grad_norm"
            result[f"{name}_grad_norm_rel"] = abs(float(m[k]) - float(m1[k])) / abs(float(m1[k]))
    dist.barrier()
    if rank == 0:
        print("DP_RESULT " + json.dumps(result), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
