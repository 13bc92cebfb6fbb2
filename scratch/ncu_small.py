"""ncu target: the M = 1024 products of the imagination steps (single-CTA kernel, two-MMA issue)."""
import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; dev = 'cuda:0'
torch.manual_seed(0)
for (M, N, Kd) in [(1024, 512, 512), (1024, 512, 1536), (1024, 1536, 1024), (1024, 1024, 512)]:
    a = K.split(torch.randn(M, Kd, device=dev)); w = K.split(torch.randn(N, Kd, device=dev) / Kd ** 0.5)
    for _ in range(2): K.gemm_tc(a, w)
torch.cuda.synchronize()
print("ok")
