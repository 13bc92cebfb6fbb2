import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; dev = 'cuda:0'
torch.manual_seed(0)
for (M, N, Kd) in [(256, 128, 32), (256, 128, 256), (1024, 1536, 1024), (15360, 512, 1536), (1000, 1030, 500)]:
    a = torch.randn(M, Kd, device=dev); w = torch.randn(N, Kd, device=dev) / Kd ** 0.5
    ref = a.double() @ w.double().t()
    out = K.gemm_tc(a, w)
    torch.cuda.synchronize()
    print(M, N, Kd, "err", float((out.double() - ref).abs().max() / ref.abs().max()), flush=True)
