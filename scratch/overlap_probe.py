"""Upper bound for overlapping the behaviour heads (bulk, GPU-filling) with the imagination rollout
(latency-bound): graph of [rollout] then [heads on half the rows] sequentially vs concurrently."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("dreamerv3-torch_b200")
dev = "cuda:0"
cfgs = pkg.configs
torch.manual_seed(0)
cfg = cfgs.make_config("dmc_proprio", device=dev, device_metrics=True)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
beh = pkg.models.ImagBehavior(cfg, wm)
batch = {k: torch.from_numpy(v).to(dev) for k, v in bench.host_batch("dmc_proprio", cfg, 0).items()}
post, _, _ = wm._train(batch)
post = {k: v.detach().clone() for k, v in post.items()}
wm.dynamics.tag_idx(post["stoch"], wm.dynamics._to_idx(post["stoch"]))
H = cfg.imag_horizon
F = cfg.dyn_stoch * cfg.dyn_discrete + cfg.dyn_deter
other = torch.randn(H // 2, 1024, F, device=dev)     # stands for the first half of the previous rollout

def heads(x):
    r = wm.heads["reward"](x).mode()
    c = wm.heads["cont"](x).mean
    v = beh.value(x).mode()
    s = beh._slow_value(x).mode()
    return r, c, v, s

def graph_ms(fn, reps=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    pkg.kernels.invalidate_weight_splits()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s): fn()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

side = torch.cuda.Stream()
def seq():
    with torch.no_grad():
        beh._imagine(post, beh.actor, H); heads(other)
def par():
    with torch.no_grad():
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            heads(other)
        beh._imagine(post, beh.actor, H)
        cur.wait_stream(side)
def only_roll():
    with torch.no_grad(): beh._imagine(post, beh.actor, H)
def only_heads():
    with torch.no_grad(): heads(other)
print("rollout alone %.3f ms, heads(7 x 1024 rows) alone %.3f ms" % (graph_ms(only_roll), graph_ms(only_heads)))
print("sequential %.3f ms, concurrent %.3f ms" % (graph_ms(seq), graph_ms(par)))
