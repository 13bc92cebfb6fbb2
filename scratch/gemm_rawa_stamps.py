"""Timeline of CTA 0 of the raw-A GEMM (DV3_GEMM_TIMING=2, needs DV3_OBSERVE_TIMING=1 for the buffer)."""
import ctypes as C, importlib, sys, os, torch
os.environ["DV3_GEMM_TIMING"] = "2"; os.environ["DV3_OBSERVE_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; L = pkg._lib; dev = 'cuda:0'; lib = L.lib()
M, N, Kd = 1024, 512, 512
a = torch.randn(M, Kd, device=dev); w = K.split(torch.randn(N, Kd, device=dev))
for _ in range(5):
    out = K.gemm_tc_rawa(a, w)
torch.cuda.synchronize()
nkb = Kd // 32
buf = (C.c_ulonglong * (64 * 8))()
assert lib.dv3_debug_observe_timing(buf, 64) == 0
t0 = buf[0]
names = ["tma issue", "full seen", "aready seen", "mma issued", "conv ready", "afree seen", "conv arrived"]
for kb in range(nkb):
    print(kb, {n: buf[kb * 8 + i] - t0 for i, n in enumerate(names)})
