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
