import importlib, sys, os, torch, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
    res = []
    for (M, N, Kd) in [(1024, 512, 512), (1024, 1024, 512), (1024, 1536, 1024), (1024, 1024, 1024), (1024, 512, 1024), (1024, 1030, 512), (1024, 512, 1536), (15360, 512, 1536), (15360, 512, 512)]:
        a = K.split(torch.randn(M, Kd, device=dev)); w = K.split(torch.randn(N, Kd, device=dev))
        out = torch.empty(M, N, device=dev)
        res.append(f"{replay_us(lambda: K.gemm_tc(a, w, out=out)):7.2f}")
    print(sys.argv[1].ljust(8), " ".join(res), flush=True)
else:
    print("config   1024x512x512 1024x1024x512 1024x1536x1024 1024x1024x1024 1024x512x1024 1024x1030x512 1024x512x1536 15360x512x1536 15360x512x512")
    for cfg in ["auto", "32,0", "64,0", "128,0", "64,1", "128,1"]:
        env = dict(os.environ)
        if cfg != "auto": env["DV3_TC_FORCE"] = cfg
        subprocess.run([sys.executable, __file__, cfg], env=env)
