import importlib, sys, os, time, torch
sys.path.insert(0, '/root/repo')
import numpy as np
pkg = importlib.import_module('dreamerv3-torch_b200')
cfgs = pkg.configs
dev = 'cuda:0'
torch.manual_seed(0)
cfg = cfgs.make_config('dmc_proprio', device=dev, device_metrics=True)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
beh = pkg.models.ImagBehavior(cfg, wm)
reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
rs = np.random.RandomState(0); B, T, A = 16, 64, 6
host = {k: rs.randn(B, T, n).astype(np.float32) for k, n in (("orientations", 14), ("height", 1), ("velocity", 9))}
host["action"] = rs.uniform(-1, 1, size=(B, T, A)).astype(np.float32)
host["reward"] = rs.randn(B, T).astype(np.float32); host["discount"] = np.ones((B, T), np.float32)
host["is_terminal"] = np.zeros((B, T), np.float32); host["is_first"] = np.zeros((B, T), np.float32); host["is_first"][:, 0] = 1
res = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
def step():
    post, _, m1 = wm._train(res)
    beh._train(post, reward_fn)
for _ in range(3): step()
torch.cuda.synchronize()
# phase timing with events
def timed(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("full step ms", timed(step))
print("wm._train ms", timed(lambda: wm._train(res)))
post, _, _ = wm._train(res)
print("beh._train ms", timed(lambda: beh._train(post, reward_fn)))
pd = wm.preprocess(res)
def wm_fwd():
    with torch.no_grad():
        emb = wm.encoder(pd); wm.dynamics.observe(emb, pd["action"], pd["is_first"])
print("encoder+observe fwd (no grad) ms", timed(wm_fwd))
def wm_fb():
    with pkg.tools.RequiresGrad(wm):
        loss, _, _ = wm.loss(pd); loss.backward()
print("wm loss fwd+bwd ms", timed(wm_fb))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type.name == 'CUDA':
        agg[e.name][0] += 1; agg[e.name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print("total GPU kernel time per step (ms):", tot / 2 / 1e3)
groups = collections.defaultdict(float)
for k, v in agg.items():
    gname = ("umma" if "umma" in k else "split" if "split_tf32" in k else "skinny" if "skinny" in k else
             "dv3 other" if "dv3::" in k else "torch gemm" if ("gemm" in k or "cutlass" in k) else "torch other")
    groups[gname] += v[1] / 2 / 1e3
print({k: round(v, 2) for k, v in groups.items()})
print("---- non-library kernels ----")
for k, v in sorted(((k, v) for k, v in agg.items() if "dv3::" not in k), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]/2/1e3:8.3f} ms  n={v[0]//2:5d} avg={v[1]/v[0]:7.1f}us  {k[:150]}")
print("---- all ----")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1]/2/1e3:8.3f} ms  n={v[0]//2:5d} avg={v[1]/v[0]:7.1f}us  {k[:100]}")
