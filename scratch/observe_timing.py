import importlib, sys, os, ctypes, numpy as np, torch
os.environ["DV3_OBSERVE_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
cfgs = pkg.configs; dev = 'cuda:0'
torch.manual_seed(0)
cfg = cfgs.make_config('dmc_proprio', device=dev, device_metrics=True)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
B, T, A = 16, 64, 6
embed = torch.randn(B, T, 1024, device=dev); action = torch.rand(B, T, A, device=dev)
first = torch.zeros(B, T, device=dev); first[:, 0] = 1
with torch.no_grad():
    for _ in range(3): wm.dynamics.observe(embed, action, first)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); wm.dynamics.observe(embed, action, first); e1.record(); torch.cuda.synchronize()
print("observe fwd total ms", e0.elapsed_time(e1))
buf = (ctypes.c_ulonglong * (T * 8))()
pkg._lib.check(pkg._lib.lib().dv3_debug_observe_timing(buf, T), "timing")
a = np.array(buf[:], dtype=np.int64).reshape(T, 8)
d = np.diff(a, axis=1)
names = ["A", "bar1", "B", "bar2", "C+D", "bar3", "E"]
med = np.median(d[4:], axis=0)
print("per-step phase medians (ns):", dict(zip(names, med.tolist())))
print("step total median ns", float(np.median(a[5:, 0] - a[4:-1, 0])))
