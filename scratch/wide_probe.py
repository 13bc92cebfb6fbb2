"""Two-MMA "wide" issue (DV3_TC_WIDE=1) vs the three-MMA single-CTA path: correctness against fp64, then timing."""
import importlib, sys, os, torch, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(1024, 512, 512), (1024, 1024, 512), (1024, 1536, 1024), (1024, 1024, 1024), (1024, 512, 1024),
          (1024, 512, 1536), (960, 512, 544), (1024, 256, 512), (15360, 512, 1536), (15360, 1536, 512), (1024, 4096, 4096)]
if len(sys.argv) > 1:
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
    torch.manual_seed(0)
    errs, ts = [], []
    for (M, N, Kd) in SHAPES:
        A = torch.randn(M, Kd, device=dev); W = torch.randn(N, Kd, device=dev)
        ref = (A.double() @ W.double().t())
        for (at, bt) in [(False, False), (True, False), (False, True), (True, True)]:
            a = K.split(A.t().contiguous()) if at else K.split(A)
            w = K.split(W.t().contiguous()) if bt else K.split(W)
            out = torch.empty(M, N, device=dev)
            K.gemm_tc(a, w, a_t=at, b_t=bt, out=out)
            errs.append(((out.double() - ref).abs().max() / ref.abs().max()).item())
        a = K.split(A); w = K.split(W); out = torch.empty(M, N, device=dev)
        ts.append(f"{replay_us(lambda: K.gemm_tc(a, w, out=out)):7.2f}")
    print(sys.argv[1].ljust(6), "max rel err %.2e |" % max(errs), " ".join(ts), flush=True)
else:
    print("cfg    " + " ".join("x".join(map(str, s)) for s in SHAPES))
    for cfg in ["w0", "w1", "w0_32", "w1_32", "w0_64", "w1_64", "w0_128", "w1_128"]:
        env = dict(os.environ)
        env["DV3_TC_WIDE"] = cfg[1]
        env["DV3_TC_PAIR"] = "0"
        if "_" in cfg: env["DV3_TC_FORCE"] = cfg.split("_")[1] + ",0"
        try:
            subprocess.run([sys.executable, __file__, cfg], env=env, timeout=120)
        except subprocess.TimeoutExpired:
            print(cfg, "TIMEOUT")
