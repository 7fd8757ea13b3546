"""GPU: tcgen05 split-bf16 radial GEMM (csrc/radial_gemm.cu) vs an fp32 reference of the same Linear.
Tolerance: the 3-term bf16 split keeps ~16 mantissa bits per operand -> 3e-5 of the output's max magnitude."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("E,K,N", [(1000, 144, 7128), (128, 144, 312), (77, 96, 312), (4099, 144, 2784), (300, 48, 500)])
def test_radial_gemm_matches_fp32_linear(built_lib, E, K, N):
    from diffdock_b200.radial import build_b_images, radial_gemm
    g = torch.Generator().manual_seed(E + N)
    h = torch.relu(torch.randn(E, K, generator=g)).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    img, bp, nt = build_b_images(W, b)
    out = radial_gemm(h, img, bp, nt)
    torch.cuda.synchronize()
    ref = torch.nn.functional.linear(h.double(), W.double(), b.double())
    err = (out[:, :N].double() - ref).abs().max() / ref.abs().max()
    assert err < 3e-5, float(err)
    assert torch.all(out[:, N:] == 0) or out.shape[1] == N     # padded columns: zero weights + zero bias
