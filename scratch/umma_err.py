import importlib, sys, torch
sys.path.insert(0, '/root/repo')
pkg = importlib.import_module('dreamerv3-torch_b200')
dev = 'cuda:0'
g = torch.Generator().manual_seed(0)
for K in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    a = torch.randn(1024, K, generator=g).to(dev)
    w = (torch.randn(512, K, generator=g) / K ** 0.5).to(dev)
    ref = a.double() @ w.double().t()
    out = pkg.kernels.linear_tc_fwd(a, w)
    f32 = a @ w.t()
    d = (out.double() - ref)
    print(K, "tc max rel %.2e  mean signed %.2e  rms %.2e | fp32 max rel %.2e rms %.2e" % (
        float(d.abs().max() / ref.abs().max()), float(d.mean() / ref.abs().mean()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
        float((f32.double() - ref).abs().max() / ref.abs().max()), float((f32.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())))
# positive-only data exposes truncation bias
a = torch.rand(1024, 1024, generator=g).to(dev); w = torch.rand(512, 1024, generator=g).to(dev) / 32
ref = a.double() @ w.double().t(); out = pkg.kernels.linear_tc_fwd(a, w)
print("positive data: mean signed rel err %.3e (fp32 %.3e)" % (float(((out.double() - ref) / ref).mean()), float((((a @ w.t()).double() - ref) / ref).mean())))
import time
for (M, N, K) in ((1024, 1536, 1024), (1024, 512, 512), (1024, 1024, 512)):
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
    for _ in range(3): pkg.kernels.linear_tc_fwd(a, w)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): pkg.kernels.linear_tc_fwd(a, w)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20): pkg.kernels.linear_fwd(a, w)
    e1.record(); torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20): a @ w.t()
    e1.record(); torch.cuda.synchronize()
    t3 = e0.elapsed_time(e1) / 20
    print(M, N, K, "tc (incl. split) %.1f us, simt tiled %.1f us, torch fp32 %.1f us" % (t * 1e3, t2 * 1e3, t3 * 1e3))
