"""Cluster split-K (128x128 tiles, K over a cluster, DSMEM reduction) vs the other tilings."""
import importlib, sys, os, torch, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(1024, 512, 512), (1024, 512, 1030), (1024, 512, 1536), (1024, 1024, 512), (1024, 1024, 1024),
          (1024, 1536, 1024), (1024, 255, 512), (960, 500, 544), (1024, 512, 128)]
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
        Kp = (Kd + 3) // 4 * 4
        A = torch.zeros(M, Kp, device=dev); A[:, :Kd] = torch.randn(M, Kd, device=dev)
        W = torch.zeros(N, Kp, device=dev); W[:, :Kd] = torch.randn(N, Kd, device=dev)
        ref = (A.double() @ W.double().t())
        for (at, bt) in [(False, False), (True, False), (False, True), (True, True)]:
            if (at and M % 4) or (bt and N % 4): continue
            a = K.split(A.t().contiguous()) if at else K.split(A)
            w = K.split(W.t().contiguous()) if bt else K.split(W)
            out = torch.empty(M, N, device=dev)
            K.gemm_tc(a, w, a_t=at, b_t=bt, out=out)
            errs.append(((out.double() - ref).abs().max() / ref.abs().max()).item())
        # bias + addend + accumulate into a strided view
        bias = torch.randn(N, device=dev); add = torch.randn(M, N, device=dev)
        big = torch.randn(M, N + 8, device=dev); view = big[:, 4:4 + N]; base = view.clone()
        K.gemm_tc(K.split(A), K.split(W), bias=bias, addend=add, out=view, accumulate=True)
        ref2 = ref + bias.double() + add.double() + base.double()
        errs.append(((view.double() - ref2).abs().max() / ref2.abs().max()).item())
        a = K.split(A); w = K.split(W); out = torch.empty(M, N, device=dev)
        o1 = K.gemm_tc(a, w).clone(); o2 = K.gemm_tc(a, w)
        assert torch.equal(o1, o2), "not deterministic"
        ts.append(f"{replay_us(lambda: K.gemm_tc(a, w, out=out)):7.2f}")
    print(sys.argv[1].ljust(10), "max rel err %.2e |" % max(errs), " ".join(ts), flush=True)
else:
    print("cfg        " + " ".join("x".join(map(str, s)) for s in SHAPES))
    for cfg in ["auto", "nocsk", "128,0,2", "128,0,4", "128,0,8"]:
        env = dict(os.environ)
        if cfg == "nocsk": env["DV3_TC_CSK"] = "0"
        elif cfg == "auto": env["DV3_TC_CSK"] = "1"
        elif cfg != "auto": env["DV3_TC_FORCE"] = cfg
        try:
            subprocess.run([sys.executable, __file__, cfg], env=env, timeout=100)
        except subprocess.TimeoutExpired:
            print(cfg, "TIMEOUT")
