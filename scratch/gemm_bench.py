import importlib, sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module('dreamerv3-torch_b200')
dev = 'cuda:0'
K = pkg.kernels
shapes = [(15360, 512, 1536), (1024, 1536, 1024), (15360, 512, 512), (1024, 512, 512), (1024, 1024, 512),
          (1024, 12288, 5120), (1024, 1024, 4096), (1024, 1024, 1024), (15360, 1024, 5120)]
lib = pkg._lib.lib()
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.dv3_prof_enable(1)
    for _ in range(n): fn()
    torch.cuda.synchronize(); lib.dv3_prof_enable(0)
    pm, pf, pl = (ctypes.c_double * 2)(), (ctypes.c_double * 2)(), (ctypes.c_longlong * 2)()
    lib.dv3_prof_read(pm, pf, pl)
    return pm[1] / pl[1] * 1e3, pf[1] / pm[1] / 1e9
for (M, N, Kd) in shapes:
    a = torch.randn(M, Kd, device=dev); w = torch.randn(N, Kd, device=dev) / Kd ** 0.5
    ref = a.double() @ w.double().t()
    As, Ws = K.split(a), K.split(w)
    o1 = K.gemm_tc(As, Ws); o2 = K.linear_tc2(a, w)
    e1 = float((o1.double() - ref).abs().max() / ref.abs().max())
    e2 = float((o2.double() - ref).abs().max() / ref.abs().max())
    t1 = timeit(lambda: K.gemm_tc(As, Ws)); t2 = timeit(lambda: K.linear_tc2(a, w))
    print(f"{M:6d} {N:6d} {Kd:6d}  presplit {t1[0]:8.1f} us {t1[1]:7.1f} TF/s err {e1:.1e} | raw {t2[0]:8.1f} us {t2[1]:7.1f} TF/s err {e2:.1e}", flush=True)
