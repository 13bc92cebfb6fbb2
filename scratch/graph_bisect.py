import importlib, sys, os, time, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels
cfgs = pkg.configs; dev = 'cuda:0'
torch.manual_seed(0)
cfg = cfgs.make_config('dmc_proprio', device=dev, device_metrics=True)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
beh = pkg.models.ImagBehavior(cfg, wm)
rs = np.random.RandomState(0); B, T, A = 16, 64, 6
host = {k: rs.randn(B, T, n).astype(np.float32) for k, n in (("orientations", 14), ("height", 1), ("velocity", 9))}
host["action"] = rs.uniform(-1, 1, size=(B, T, A)).astype(np.float32)
host["reward"] = rs.randn(B, T).astype(np.float32); host["discount"] = np.ones((B, T), np.float32)
host["is_terminal"] = np.zeros((B, T), np.float32); host["is_first"] = np.zeros((B, T), np.float32); host["is_first"][:, 0] = 1
res = {k: torch.from_numpy(v).to(dev) for k, v in host.items()}
reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
for _ in range(2):
    post, _, _ = wm._train(res); beh._train(post, reward_fn)
torch.cuda.synchronize()
def attempt(name, fn):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name, flush=True)
    except Exception as e:
        print("FAIL", name, str(e).split("\n")[0][:100], flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass
a = torch.randn(1024, 512, device=dev); w = torch.randn(512, 512, device=dev)
attempt("gemm_tc", lambda: K.gemm_tc(a, w))
attempt("gemm_tc split_k", lambda: K.gemm_tc(a, w, a_t=True, b_t=True, split_k=True))
attempt("ln_grads", lambda: K._ln_grads(a, a))
pd = wm.preprocess(res)
emb = torch.randn(B, T, 1024, device=dev)
def obs():
    with torch.no_grad(): wm.dynamics.observe(emb, pd["action"], pd["is_first"])
attempt("observe fwd (persistent coop)", obs)
def wm_fb():
    with pkg.tools.RequiresGrad(wm):
        loss, _, _ = wm.loss(pd); loss.backward()
attempt("wm loss fwd+bwd", wm_fb)
attempt("wm._train", lambda: wm._train(res))
post, _, _ = wm._train(res)
attempt("beh._imagine", lambda: beh._imagine(post, beh.actor, 15))
attempt("beh.losses", lambda: beh.losses(post, reward_fn))
attempt("beh._train", lambda: beh._train(post, reward_fn))
post, _, _ = wm._train(res)
def both():
    p2, _, _ = wm._train(res); beh._train(p2, reward_fn)
attempt("wm+beh _train", both)
g = pkg.graphs.TrainStepGraph(wm, beh, warmup=1, device_metrics=True)
pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
try:
    for i in range(3): out = g(pinned)
    torch.cuda.synchronize(); print("OK TrainStepGraph", g.captured)
except Exception as e:
    import traceback; traceback.print_exc()
