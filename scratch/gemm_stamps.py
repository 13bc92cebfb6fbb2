"""Where a single-wave GEMM launch spends its time: globaltimer stamps of CTA 0 (DV3_GEMM_TIMING=1)."""
import importlib, sys, os, ctypes, numpy as np, torch
os.environ["DV3_GEMM_TIMING"] = "1"; os.environ["DV3_OBSERVE_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; dev = 'cuda:0'
torch.manual_seed(0)
names = ["entry", "prologue", "tma0 issued", "data0 landed", "last mma issued", "mma done", "sums ready",
         "exit", "cluster sync1", "reduced+stored", "cluster sync2"]
for (M, N, Kd) in [(1024, 512, 512), (1024, 512, 1536)]:
    a = K.split(torch.randn(M, Kd, device=dev)); w = K.split(torch.randn(N, Kd, device=dev)); out = torch.empty(M, N, device=dev)
    for force in ["32,0", "64,0", "128,0", "128,0,2", "128,0,4"]:
        os.environ["DV3_TC_FORCE"] = force
        for _ in range(6): K.gemm_tc(a, w, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): K.gemm_tc(a, w, out=out)
        e1.record(); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)()
        pkg._lib.check(pkg._lib.lib().dv3_debug_observe_timing(buf, 2), "timing")
        t = np.array(buf[:], dtype=np.int64)
        rel = {names[i]: int(t[i] - t[0]) for i in range(len(names)) if t[i] >= t[0] and t[i] - t[0] < 10**8}
        print(f"{M}x{N}x{Kd} force={force:8s} eager {e0.elapsed_time(e1)*50:6.1f} us/launch | ns from entry:", rel, flush=True)
        # clear stamps for the next config
        os.environ["DV3_GEMM_TIMING"] = "1"; os.environ["DV3_OBSERVE_TIMING"] = "1"
