import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels
orig = K.gemm_tc
bad = []
def checked(A, B, a_t=False, b_t=False, A2=None, bias=None, addend=None, out=None, accumulate=False, split_k=False):
    As, Bs = K.split(A), K.split(B)
    prev = out.clone() if (out is not None and accumulate) else None
    r = orig(As, Bs, a_t=a_t, b_t=b_t, A2=A2, bias=bias, addend=addend, out=out, accumulate=accumulate, split_k=split_k)
    a = (As.hi + As.lo)[:, :As.cols].double(); b = (Bs.hi + Bs.lo)[:, :Bs.cols].double()
    if a_t: a = a.t()
    if b_t: b = b.t()
    if A2 is not None:
        A2s = K.split(A2); a2 = (A2s.hi + A2s.lo)[:, :A2s.cols].double()
        a = torch.cat([a, a2.t() if a_t else a2], 1)
    ref = a @ b.t()
    if bias is not None: ref = ref + bias.double()
    if addend is not None: ref = ref + addend.double()
    if prev is not None: ref = ref + prev.double()
    err = float((r.double() - ref).abs().max() / (ref.abs().max() + 1e-30))
    if err > 1e-5:
        bad.append((tuple(a.shape), tuple(b.shape), a_t, b_t, split_k, err))
    return r
K.gemm_tc = checked
import test_gpu_golden as T
try:
    T.test_train_steps_vs_reference(pkg, 'cuda:0', 'normal')
    print("PASS")
except AssertionError as e:
    print("FAIL", str(e)[:200])
print(len(bad), bad[:20])
