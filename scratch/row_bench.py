"""Per-shape times of the row kernels (graph-replayed back to back, 20 reps)."""
import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; dev = 'cuda:0'
def replay_us(fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
torch.manual_seed(0)
for (M, n) in [(1024, 512), (1024, 1024), (1024, 1536), (15360, 512), (15360, 1536)]:
    pre = torch.randn(M, n, device=dev); g = torch.ones(n, device=dev); b = torch.zeros(n, device=dev)
    dy = torch.randn(M, n, device=dev)
    r = {}
    r["ln_fwd"] = replay_us(lambda: K.ln_silu_fwd(pre, g, b))
    r["ln_fwd+split"] = replay_us(lambda: K.ln_silu_fwd(pre, g, b, with_split=True))
    r["ln_bwd"] = replay_us(lambda: K.ln_silu_bwd(pre, g, b, dy))
    r["ln_bwd+split"] = replay_us(lambda: K.ln_silu_bwd(pre, g, b, dy, with_split=True))
    r["split"] = replay_us(lambda: K.split(pre))
    r["torch copy"] = replay_us(lambda: pre.clone())
    print(f"{M}x{n}: " + "  ".join(f"{k} {v:6.1f}" for k, v in r.items()), flush=True)
lg = torch.randn(1024, 32, 32, device=dev); u = torch.rand(1024, 32, 32, device=dev).clamp_(1e-30, 1)
print("onehot_sample 1024x32x32: %.1f us" % replay_us(lambda: K.onehot_sample(lg, u, 0.01)))
print("onehot_st_bwd 1024x32x32: %.1f us" % replay_us(lambda: K.onehot_st_bwd(lg, u, None, 0.01)))
