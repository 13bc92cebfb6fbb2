import importlib, sys, os, time, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
cfgs = pkg.configs; dev = 'cuda:0'
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
for _ in range(3):
    post, _, _ = wm._train(res); beh._train(post, reward_fn)
def gpu_ms(fn, n=5):
    """GPU-side duration via a captured graph replay (no host overhead)."""
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
pd = wm.preprocess(res)
R = pkg.tools.RequiresGrad
def enc():
    with torch.no_grad(): wm.encoder(pd)
emb = wm.encoder(pd).detach()
def obs():
    with torch.no_grad(): wm.dynamics.observe(emb, pd["action"], pd["is_first"])
def wm_fwd():
    with torch.no_grad(): wm.loss(pd)
def wm_fb():
    with R(wm):
        loss, _, _ = wm.loss(pd); loss.backward()
    for p in wm.parameters(): p.grad = None
def obs_fb():
    e = emb.clone().requires_grad_(True)
    with R(wm.dynamics):
        post, prior = wm.dynamics.observe(e, pd["action"], pd["is_first"])
        (post["deter"].sum() + post["stoch"].sum() + prior["logit"].sum() + post["logit"].sum()).backward()
    for p in wm.parameters(): p.grad = None
print("encoder fwd", gpu_ms(enc)); print("observe fwd", gpu_ms(obs)); print("wm loss fwd", gpu_ms(wm_fwd))
print("observe fwd+bwd", gpu_ms(obs_fb)); print("wm loss fwd+bwd", gpu_ms(wm_fb)); print("wm._train", gpu_ms(lambda: wm._train(res)))
post, _, _ = wm._train(res)
def imag():
    with torch.no_grad(): beh._imagine(post, beh.actor, 15)
def imag_fb():
    with R(beh.actor):
        f, s, a = beh._imagine(post, beh.actor, 15)
        (s["deter"].sum() + s["stoch"].sum() + a.sum()).backward()
    for p in beh.actor.parameters(): p.grad = None
def losses_f():
    with torch.no_grad(): beh.losses(post, reward_fn)
def losses_fb():
    al, vl, _, _, _ = beh.losses(post, reward_fn)
    with R(beh):
        al.backward(); vl.backward()
    for p in beh.parameters(): p.grad = None
print("imagine fwd", gpu_ms(imag)); print("imagine fwd+bwd", gpu_ms(imag_fb)); print("beh.losses fwd", gpu_ms(losses_f))
print("beh.losses fwd+bwd", gpu_ms(losses_fb)); print("beh._train", gpu_ms(lambda: beh._train(post, reward_fn)))
