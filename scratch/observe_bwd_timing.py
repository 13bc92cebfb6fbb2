import importlib, sys, os, ctypes, numpy as np, torch
os.environ["DV3_OBSERVE_TIMING"] = "2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
cfgs = pkg.configs; dev = 'cuda:0'
torch.manual_seed(0)
cfg = cfgs.make_config('dmc_proprio', device=dev, device_metrics=True)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
B, T, A = 16, 64, 6
action = torch.rand(B, T, A, device=dev); first = torch.zeros(B, T, device=dev); first[:, 0] = 1
def run():
    e = torch.randn(B, T, 1024, device=dev, requires_grad=True)
    with pkg.tools.RequiresGrad(wm.dynamics):
        post, prior = wm.dynamics.observe(e, action, first)
        (post["deter"].sum() + post["stoch"].sum() + prior["logit"].sum() + post["logit"].sum()).backward()
for _ in range(3): run()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (T * 8))()
pkg._lib.check(pkg._lib.lib().dv3_debug_observe_timing(buf, T), "timing")
a = np.array(buf[:], dtype=np.int64).reshape(T, 8)
d = np.diff(a[:, :7], axis=1)
names = ["1a+bar", "1b+bar", "2+bar", "3a+bar", "3b+bar", "4"]
print("per-step phase medians (ns):", dict(zip(names, np.median(d[2:-2], axis=0).tolist())))
print("step total median ns", float(np.median(a[2:-3, 0] - a[3:-2, 0])))
