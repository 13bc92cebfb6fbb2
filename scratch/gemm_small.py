"""Fixed cost vs per-k-block cost of the M=1024 products (graph-replayed back to back)."""
import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; dev = 'cuda:0'
def replay_us(fn, reps=50):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for (M, N) in [(1024, 512), (1024, 1536), (1024, 1024)]:
    for Kd in [32, 128, 256, 512, 1024, 2048]:
        a = K.split(torch.randn(M, Kd, device=dev)); w = K.split(torch.randn(N, Kd, device=dev))
        out = torch.empty(M, N, device=dev)
        t = replay_us(lambda: K.gemm_tc(a, w, out=out))
        print(f"M={M} N={N} K={Kd:5d}  {t:6.2f} us/launch", flush=True)
