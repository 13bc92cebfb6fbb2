"""Small workload for `ncu --set full`: the GEMM at three shapes (bulk rows, the large-config GRU
product, an imagination-step product) and one observe forward + backward (persistent kernels)."""
import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; dev = 'cuda:0'
torch.manual_seed(0)
for (M, N, Kd) in [(15360, 512, 1536), (1024, 12288, 5120), (1024, 1536, 1024)]:
    a = K.split(torch.randn(M, Kd, device=dev)); w = K.split(torch.randn(N, Kd, device=dev) / Kd ** 0.5)
    for _ in range(2): K.gemm_tc(a, w)
torch.cuda.synchronize()
cfgs = pkg.configs
cfg = cfgs.make_config('dmc_proprio', device=dev, device_metrics=True)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
B, T, A = 16, 64, 6
action = torch.rand(B, T, A, device=dev); first = torch.zeros(B, T, device=dev); first[:, 0] = 1
for _ in range(2):
    e = torch.randn(B, T, 1024, device=dev, requires_grad=True)
    with pkg.tools.RequiresGrad(wm.dynamics):
        post, prior = wm.dynamics.observe(e, action, first)
        (post["deter"].sum() + post["stoch"].sum() + prior["logit"].sum() + post["logit"].sum()).backward()
torch.cuda.synchronize()
print("ok")
