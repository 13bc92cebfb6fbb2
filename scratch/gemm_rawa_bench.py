"""Raw-A (TMEM-fed) GEMM vs the pre-split kernel on the in-loop shapes: bit equality and time per
launch (graph-replayed back to back)."""
import ctypes as C
import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels; L = pkg._lib; dev = 'cuda:0'
lib = L.lib()

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

def rawa(a, a2, wsp, out, transposed=False, addend=None):
    M, K1 = a.shape
    K2 = a2.shape[1] if a2 is not None else 0
    N = out.shape[1]
    op = wsp.operand(transposed)
    L.check(lib.dv3_gemm_tc_rawa(K._raw(a), a.stride(0), K1, K._raw(a2),
                                 a2.stride(0) if a2 is not None else 0, K2, C.byref(op), None,
                                 K._raw(addend), addend.stride(0) if addend is not None else 0,
                                 K._raw(out), out.stride(0), M, N, L.stream_ptr()), "gemm_tc_rawa")

torch.manual_seed(0)
for (M, N, K1, K2, tr) in [(1024, 512, 512, 0, False), (1024, 1536, 512, 512, False), (1024, 1024, 512, 0, False),
                           (1024, 512, 1024, 0, True), (1024, 512, 512, 0, True), (1024, 1024, 1536, 0, True),
                           (1024, 1030, 512, 0, True), (256, 512, 512, 0, False), (1024, 1024, 1024, 0, False),
                           (15360, 512, 512, 0, False)]:
    Kt = K1 + K2
    a = torch.randn(M, K1, device=dev); a2 = torch.randn(M, K2, device=dev) if K2 else None
    w = torch.randn(Kt, N, device=dev) if tr else torch.randn(N, Kt, device=dev)
    asp = K.split(a); a2sp = K.split(a2) if K2 else None; wsp = K.split(w)
    out0 = torch.empty(M, N, device=dev); out1 = torch.empty(M, N, device=dev)
    K.gemm_tc(asp, wsp, b_t=tr, A2=a2sp, out=out0)
    rawa(a, a2, wsp, out1, tr)
    ref = (torch.cat([a, a2], 1) if K2 else a).double() @ (w.double() if tr else w.double().t())
    err = ((out1.double() - ref).abs().max() / ref.abs().max()).item()
    same = torch.equal(out0, out1)
    t0 = replay_us(lambda: K.gemm_tc(asp, wsp, b_t=tr, A2=a2sp, out=out0))
    t1 = replay_us(lambda: rawa(a, a2, wsp, out1, tr))
    line = f"M={M} N={N} K={K1}+{K2} Bt={int(tr)}: planes {t0:6.2f} us, raw-A {t1:6.2f} us, bit-equal {same}, err vs fp64 {err:.2e}"
    for bn in ("32", "64", "96"):
        os.environ["DV3_TCT_FORCE"] = bn; lib.dv3_reload_env()
        try:
            line += f" | BN={bn}: {replay_us(lambda: rawa(a, a2, wsp, out1, tr)):6.2f}"
        except Exception as e:
            line += f" | BN={bn}: err"
    os.environ.pop("DV3_TCT_FORCE"); lib.dv3_reload_env()
    print(line, flush=True)
