import importlib, sys, torch
sys.path.insert(0, '/root/repo')
pkg = importlib.import_module('dreamerv3-torch_b200')
dev = 'cuda:0'
shapes = [(15360, 512, 1536), (1024, 1536, 1024), (15360, 512, 512), (1024, 512, 512)]
for (M, N, K) in shapes:
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
    for _ in range(2): pkg.kernels.linear_tc(a, w)
    torch.cuda.synchronize()
    import ctypes
    lib = pkg._lib.lib(); lib.dv3_prof_enable(1)
    for _ in range(10): pkg.kernels.linear_tc(a, w)
    torch.cuda.synchronize(); lib.dv3_prof_enable(0)
    pm, pf, pl = (ctypes.c_double * 2)(), (ctypes.c_double * 2)(), (ctypes.c_longlong * 2)()
    lib.dv3_prof_read(pm, pf, pl)
    print(M, N, K, "umma kernel %.1f us  %.1f TFLOP/s" % (pm[1] / pl[1] * 1e3, pf[1] / pm[1] / 1e9))
