"""-m gpu: the tcgen05 3xTF32 GEMM against an fp64 matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 64), (1024, 1536, 1024), (1024, 512, 512),
                                   (1000, 1030, 512), (24, 48, 48), (257, 130, 100), (15, 32, 1), (70, 9, 30)])
def test_linear_tc_matches_fp64(pkg, device, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(device)
    bias = torch.randn(N, generator=g).to(device)
    add = torch.randn(M, N, generator=g).to(device)
    ref = (a.double() @ w.double().t() + bias.double() + add.double())
    out = pkg.kernels.linear_tc_fwd(a, w, bias, add)
    torch.cuda.synchronize()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    fp32 = float(((a @ w.t() + bias + add).double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-6, (err, fp32)
    out2 = pkg.kernels.linear_tc_fwd(a, w)
    ref2 = a.double() @ w.double().t()
    assert float((out2.double() - ref2).abs().max() / ref2.abs().max()) < 5e-6


def test_gemm_config4_gru_product_matches_fp64(pkg, device):
    """The GRU product of BASELINE configs[3]: [x|h] (1024 x 5120) times W_gru^T (12288 x 5120),
    two K segments, pre-split planes -- the launch the tensor-pipe figures are quoted on."""
    g = torch.Generator().manual_seed(4)
    M, N, K1, K2 = 1024, 12288, 1024, 4096
    x = torch.randn(M, K1, generator=g).to(device)
    h = torch.tanh(torch.randn(M, K2, generator=g)).to(device)
    w = (torch.randn(N, K1 + K2, generator=g) / (K1 + K2) ** 0.5).to(device)
    out = pkg.kernels.gemm_tc(x, w, A2=h)
    ref = torch.cat([x, h], 1).double() @ w.double().t()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-6, err
    dx = pkg.kernels.gemm_tc(out, w, b_t=True)                 # dy W, K = 12288
    refx = out.double() @ w.double()
    assert float((dx.double() - refx).abs().max() / refx.abs().max()) < 5e-6


@pytest.mark.parametrize("M,N,K", [(96, 80, 40), (512, 1536, 1000), (14, 1024, 15)])
def test_linear_tc_transposed_operands(pkg, device, M, N, K):
    g = torch.Generator().manual_seed(1)
    at = torch.randn(K, M, generator=g).to(device)      # stored [K,M]
    wt = torch.randn(K, N, generator=g).to(device)      # stored [K,N]
    ref = at.double().t() @ wt.double()
    out = pkg.kernels.linear_tc(at, wt, trans_a=True, trans_w=True)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6
    out = pkg.kernels.linear_tc(at.t().contiguous(), wt, trans_w=True)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 64), (1024, 1536, 1024), (1024, 512, 512),
                                   (1000, 1032, 500), (24, 48, 48), (257, 130, 100), (15, 32, 4),
                                   (70, 9, 36), (15360, 512, 1536), (2000, 640, 4096)])
def test_linear_tc2_matches_fp64(pkg, device, M, N, K):
    """raw-operand persistent kernel (in-SM hi/lo split) against fp64."""
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(device)
    bias = torch.randn(N, generator=g).to(device)
    add = torch.randn(M, N, generator=g).to(device)
    ref = (a.double() @ w.double().t() + bias.double() + add.double())
    out = pkg.kernels.linear_tc2(a, w, bias=bias, addend=add)
    torch.cuda.synchronize()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-6, err
    ref2 = a.double() @ w.double().t()
    out2 = pkg.kernels.linear_tc2(a, w)
    assert float((out2.double() - ref2).abs().max() / ref2.abs().max()) < 5e-6
    out3 = pkg.kernels.linear_tc2(a, w, out=out2.clone(), accumulate=True)
    assert float((out3.double() - 2 * ref2).abs().max() / ref2.abs().max()) < 1e-5


@pytest.mark.parametrize("M,N,K1,K2", [(1024, 1536, 512, 512), (300, 200, 100, 60), (64, 64, 32, 4)])
def test_linear_tc2_two_segments_and_strides(pkg, device, M, N, K1, K2):
    """[A1|A2] concatenation along K with row-strided views (the GRU's [x|h] product)."""
    g = torch.Generator().manual_seed(5)
    big1 = torch.randn(M, K1 + 8, generator=g).to(device)
    big2 = torch.randn(M, K2 + 12, generator=g).to(device)
    a1, a2 = big1[:, 4:4 + K1], big2[:, 8:8 + K2]
    w = (torch.randn(N, K1 + K2, generator=g) / (K1 + K2) ** 0.5).to(device)
    ref = torch.cat([a1, a2], 1).double() @ w.double().t()
    out = pkg.kernels.linear_tc2(a1, w, a2=a2)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (1024, 512, 512), (1000, 1032, 500), (257, 130, 100),
                                   (70, 9, 36), (15360, 512, 1536), (512, 1536, 15360), (40, 24, 2000)])
@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_tc_all_storage_orders(pkg, device, M, N, K, a_t, b_t):
    """pre-split operands read K-major or MN-major (transposed) through the UMMA descriptors:
    y = x W^T, dx = dy W, dW = dy^T x from the same planes."""
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    a = torch.randn(M, K, generator=g).to(device)
    b = (torch.randn(N, K, generator=g) / K ** 0.5).to(device)
    ref = a.double() @ b.double().t()
    As = pkg.kernels.split(a.t().contiguous() if a_t else a)
    Bs = pkg.kernels.split(b.t().contiguous() if b_t else b)
    out = pkg.kernels.gemm_tc(As, Bs, a_t=a_t, b_t=b_t)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err < 5e-6, err


def test_gemm_tc_segments_bias_accumulate(pkg, device):
    g = torch.Generator().manual_seed(9)
    M, N, K1, K2 = 1024, 1536, 512, 512
    a1 = torch.randn(M, K1, generator=g).to(device)
    a2 = torch.randn(M, K2, generator=g).to(device)
    w = (torch.randn(N, K1 + K2, generator=g) / 32).to(device)
    bias = torch.randn(N, generator=g).to(device)
    add = torch.randn(M, N, generator=g).to(device)
    ref = torch.cat([a1, a2], 1).double() @ w.double().t() + bias.double() + add.double()
    out = pkg.kernels.gemm_tc(a1, w, A2=a2, bias=bias, addend=add)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6
    out2 = pkg.kernels.gemm_tc(a1, w, A2=a2, out=out.clone(), accumulate=True)
    ref2 = ref + torch.cat([a1, a2], 1).double() @ w.double().t()
    assert float((out2.double() - ref2).abs().max() / ref2.abs().max()) < 5e-6
    # hi + lo reproduces the input exactly
    sp = pkg.kernels.split(a1)
    assert torch.equal(sp.hi[:, :K1] + sp.lo[:, :K1], a1)


@pytest.mark.parametrize("M,N,K", [(512, 512, 15360), (1536, 1024, 1024), (512, 6, 1024), (200, 1030, 5000)])
def test_gemm_tc_split_k(pkg, device, M, N, K):
    """dW-shaped products (deep K, few output tiles): K partitioned over CTAs, fp32 atomics."""
    g = torch.Generator().manual_seed(M + N + K)
    dy = torch.randn(K, M, generator=g).to(device)          # [rows, out]
    x = (torch.randn(K, N, generator=g) / K ** 0.5).to(device)
    ref = dy.double().t() @ x.double()
    out = pkg.kernels.gemm_tc(dy, x, a_t=True, b_t=True, split_k=True)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6
    # strided output view + accumulate
    big = torch.zeros(M, N + 8, device=device)
    pkg.kernels.gemm_tc(dy, x, a_t=True, b_t=True, split_k=True, out=big[:, 4:4 + N])
    pkg.kernels.gemm_tc(dy, x, a_t=True, b_t=True, split_k=True, out=big[:, 4:4 + N], accumulate=True)
    assert float((big[:, 4:4 + N].double() - 2 * ref).abs().max() / ref.abs().max()) < 1e-5
    assert float(big[:, :4].abs().max()) == 0.0 and float(big[:, 4 + N:].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1024, 1536, 1024), (15360, 512, 1536), (300, 640, 200),
                                   (4096, 256, 4096)])
@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_pair_kernel(pkg, device, M, N, K, a_t, b_t):
    """Shapes the cost model sends to the cta_group::2 pair kernel (256 x BN tiles, B tile shared by
    two SMs), in all four storage orders, against fp64; and the same product with the pair kernel
    switched off must agree to fp32 rounding."""
    import os
    g = torch.Generator().manual_seed(2 * M + N + K)
    a = torch.randn(M, K, generator=g).to(device)
    b = (torch.randn(N, K, generator=g) / K ** 0.5).to(device)
    bias = torch.randn(N, generator=g).to(device)
    ref = a.double() @ b.double().t() + bias.double()
    As = pkg.kernels.split(a.t().contiguous() if a_t else a)
    Bs = pkg.kernels.split(b.t().contiguous() if b_t else b)
    out = pkg.kernels.gemm_tc(As, Bs, a_t=a_t, b_t=b_t, bias=bias)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6


def test_gemm_pair_kernel_is_used(pkg, device):
    """The bulk-row product of the train step must run on the pair kernel (kernel name check
    through the profiler), so that the test above really covers it."""
    from torch.profiler import profile, ProfilerActivity
    a = pkg.kernels.split(torch.randn(15360, 1536, device=device))
    w = pkg.kernels.split(torch.randn(512, 1536, device=device))
    pkg.kernels.gemm_tc(a, w)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        pkg.kernels.gemm_tc(a, w)
        torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type.name == "CUDA"]
    assert any("umma2x_gemm_kernel" in n for n in names), names


@pytest.mark.parametrize("force", ["32,0", "64,0", "128,0", "128,0,2", "128,0,4", "128,0,8"])
@pytest.mark.parametrize("M,N,K", [(1024, 512, 512), (1000, 255, 1030), (1024, 1024, 96)])
def test_gemm_forced_tilings_agree(pkg, device, knob, force, M, N, K):
    """Every tiling of the single-CTA kernel -- the two-MMA issue at BN = 32 / 64 / 128 and the
    cluster split-K mode (K over 2 / 4 / 8 CTAs of a cluster, partial tiles reduced through
    distributed shared memory in a fixed order) -- against fp64, with bias + addend + accumulate
    into a strided view, and bit-identical from run to run."""
    knob("DV3_TC_FORCE", force)
    g = torch.Generator().manual_seed(M + 3 * N + 5 * K)
    Kp = (K + 3) // 4 * 4
    a = torch.zeros(M, Kp); a[:, :K] = torch.randn(M, K, generator=g)
    b = torch.zeros(N, Kp); b[:, :K] = torch.randn(N, K, generator=g) / K ** 0.5
    a, b = a.to(device), b.to(device)
    bias = torch.randn(N, generator=g).to(device)
    add = torch.randn(M, N, generator=g).to(device)
    ref = a.double() @ b.double().t()
    As, Bs = pkg.kernels.split(a), pkg.kernels.split(b)
    out = pkg.kernels.gemm_tc(As, Bs).clone()
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6
    assert torch.equal(out, pkg.kernels.gemm_tc(As, Bs))
    big = torch.randn(M, N + 8, generator=g).to(device)
    view = big[:, 4:4 + N]
    base = view.clone()
    pkg.kernels.gemm_tc(As, Bs, bias=bias, addend=add, out=view, accumulate=True)
    ref2 = ref + bias.double() + add.double() + base.double()
    assert float((view.double() - ref2).abs().max() / ref2.abs().max()) < 5e-6
    # transposed storage of both operands
    if M % 4 == 0 and N % 4 == 0:
        At, Bt = pkg.kernels.split(a.t().contiguous()), pkg.kernels.split(b.t().contiguous())
        out = pkg.kernels.gemm_tc(At, Bt, a_t=True, b_t=True)
        assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6


@pytest.mark.parametrize("M,N,K1,K2,b_t,force", [
    (1024, 512, 512, 0, False, None), (1024, 1536, 512, 512, False, None), (1024, 1024, 512, 0, False, None),
    (1024, 512, 512, 0, True, None), (1024, 1030, 512, 0, True, None), (1000, 130, 100, 0, False, None),
    (64, 8, 4, 0, False, None), (257, 520, 96, 36, False, "64"), (1024, 1536, 512, 512, False, "32"),
    (300, 96, 1024, 0, True, "64"), (2048, 1536, 512, 512, False, "96")])
def test_gemm_rawa_tmem_operand(pkg, device, knob, M, N, K1, K2, b_t, force):
    """The raw-A kernel (fp32 A split in the SM, tcgen05.mma with A from tensor memory) against fp64
    and, where the pre-split kernel picks the same 128 x BN single-CTA tile, bit for bit against it."""
    if force:
        knob("DV3_TCT_FORCE", force)
    g = torch.Generator().manual_seed(M + 3 * N + 5 * K1 + K2)
    K = K1 + K2
    a = torch.randn(M, K1, generator=g).to(device)
    a2 = torch.randn(M, K2, generator=g).to(device) if K2 else None
    w = (torch.randn(K, N, generator=g) if b_t else torch.randn(N, K, generator=g)).to(device) / K ** 0.5
    bias = torch.randn(N, generator=g).to(device)
    add = torch.randn(M, N + 4, generator=g).to(device)[:, :N]          # strided addend
    full = torch.cat([a, a2], 1) if K2 else a
    ref = full.double() @ (w.double() if b_t else w.double().t())
    out = pkg.kernels.gemm_tc_rawa(a, w, b_t=b_t, a2=a2)
    assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6
    out2 = pkg.kernels.gemm_tc_rawa(a, w, b_t=b_t, a2=a2, bias=bias, addend=add)
    ref2 = ref + bias.double() + add.double()
    assert float((out2.double() - ref2).abs().max() / ref2.abs().max()) < 5e-6
    if (M, N, K1, K2) == (1024, 512, 512, 0) and not force:
        knob("DV3_TC_FORCE", "32,0")
        same = pkg.kernels.gemm_tc(a, w, b_t=b_t)
        assert torch.equal(same, out)
