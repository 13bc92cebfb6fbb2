"""ncu target: the register-form row kernels at the train step's shapes."""
import importlib, sys, os, ctypes as C, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; L = pkg._lib; dev = 'cuda:0'
torch.manual_seed(0)
M, n = 15360, 512
pre = torch.randn(M, n, device=dev); g = torch.ones(n, device=dev); b = torch.zeros(n, device=dev)
dy = torch.randn(M, n, device=dev)
for _ in range(2):
    K.ln_silu_fwd(pre, g, b, with_split=True)
    K.ln_silu_bwd(pre, g, b, dy, with_split=True)
    K.split(pre)
    K.split(torch.randn(M, 1536, device=dev))
lg = torch.randn(1024, 32, 32, device=dev); u = torch.rand(1024, 32, 32, device=dev).clamp_(1e-30, 1)
for _ in range(2):
    K.onehot_sample(lg, u, 0.01); K.onehot_st_bwd(lg, u, None, 0.01)
D = 512; Mr = 1024
gp = torch.randn(Mr, 3 * D, device=dev); gg = torch.ones(3 * D, device=dev); bb = torch.zeros(3 * D, device=dev)
h = torch.randn(Mr, D, device=dev); dn = torch.randn(Mr, D, device=dev)
o1 = torch.empty(Mr, 3 * D, device=dev); o2 = torch.empty(Mr, 3 * D, device=dev); o3 = torch.empty(Mr, D, device=dev)
for _ in range(2):
    L.check(L.lib().dv3_gru_gates_fwd(L.fptr(gp), 3 * D, L.fptr(gg), L.fptr(bb), 1e-3, L.fptr(h), D, Mr, D,
                                      L.fptr(o3), D, L.stream_ptr()), "f")
    L.check(L.lib().dv3_gru_gates_bwd(L.fptr(gp), 3 * D, L.fptr(gg), L.fptr(bb), 1e-3, L.fptr(h), D, L.fptr(dn), D,
                                      Mr, D, L.fptr(o1), L.fptr(o2), 3 * D, L.fptr(o3), D, L.stream_ptr()), "b")
torch.cuda.synchronize(); print("ok")
