"""Graph-replay times of the stages of one train step (dmc_proprio): world-model step, behaviour step,
imagination forward, behaviour forward (no gradients)."""
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
reward = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()

def graph_ms(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    pkg.kernels.invalidate_weight_splits()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

post, _, _ = wm._train(batch)
post = {k: v.detach().clone() for k, v in post.items()}
wm.dynamics.tag_idx(post["stoch"], wm.dynamics._to_idx(post["stoch"]))
print("world-model step (fwd + bwd + Adam)      %.3f ms" % graph_ms(lambda: wm._train(batch)))
print("behaviour step (imagine, heads, bwd, Adam) %.3f ms" % graph_ms(lambda: beh._train(post, reward)))
def imag():
    with torch.no_grad():
        return beh._imagine(post, beh.actor, cfg.imag_horizon)
print("imagination forward                       %.3f ms" % graph_ms(imag))
def fwd_only():
    with torch.no_grad():
        feat, state, action = beh._imagine(post, beh.actor, cfg.imag_horizon)
        r = reward(feat, state, action)
        t, w, b = beh._compute_target(feat, state, r)
        return t
print("behaviour forward (imagine + heads + target, no grad) %.3f ms" % graph_ms(fwd_only))
def wm_fwd():
    with torch.no_grad():
        data = wm.preprocess(batch)
        return wm.loss(data)[0]
print("world-model forward (no grad)             %.3f ms" % graph_ms(wm_fwd))
