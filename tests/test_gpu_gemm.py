"""tcgen05 GEMM bring-up: the TMA/TMEM/UMMA pipeline vs a torch fp32 reference of the same
fp16-rounded operands (tolerance: fp32 accumulation-order noise only) and vs the SIMT check
kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from relax_vqa_b200 import ops
    c = ops.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("mnk", [(128, 128, 64), (128, 256, 128), (256, 256, 768), (197 * 3, 768, 768), (1000, 2304, 768),
                                 (394, 3072, 768), (394, 768, 3072), (130, 48, 192), (64, 16, 64), (777, 1000, 200)])
def test_gemm_matches_fp32_reference(ctx, mnk):
    from relax_vqa_b200 import ops
    M, N, K = mnk
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ B.float().t() + bias
    got = ops.gemm_f16(ctx, A, B, bias, impl=0)
    torch.cuda.synchronize()
    tol = 1e-5 * K ** 0.5 * 4 + 1e-6          # fp32 accumulation-order noise, |a|,|b| ~ 0.5
    err = (got - ref).abs().max().item()
    assert err <= max(tol, 2e-3 * ref.abs().max().item() * 1e-2), (mnk, err)
    chk = ops.gemm_f16(ctx, A, B, bias, impl=1)
    assert (chk - ref).abs().max().item() <= max(tol, 1e-4)


@pytest.mark.parametrize("mnk", [(256, 256, 64), (512, 768, 768), (197 * 5, 2304, 768), (1000, 768, 3072), (300, 3072, 768)])
def test_gemm_2cta_matches_fp32_reference(ctx, mnk):
    """cta_group::2 kernel (256 x 256 tile per SM pair, cluster launch, multicast commits)."""
    from relax_vqa_b200 import ops
    M, N, K = mnk
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ B.float().t() + bias
    got = ops.gemm_f16(ctx, A, B, bias, impl=2)
    torch.cuda.synchronize()
    tol = 1e-5 * K ** 0.5 * 4 + 1e-6
    assert (got - ref).abs().max().item() <= max(tol, 1e-4), mnk


@pytest.mark.parametrize("mnk", [(256, 256, 64), (512, 768, 768), (700, 256, 128), (197 * 5, 2304, 768), (1000, 768, 3072), (300, 3072, 768),
                                 (197 * 40, 768, 768)])
def test_gemm_4cta_matches_fp32_reference(ctx, mnk):
    """4-CTA clusters: two SM pairs, the weight tile multicast between them (odd numbers of row-tile pairs leave the second
    pair of the last cluster on zero-filled rows)."""
    from relax_vqa_b200 import ops
    M, N, K = mnk
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ B.float().t() + bias
    got = ops.gemm_f16(ctx, A, B, bias, impl=3)
    torch.cuda.synchronize()
    tol = 1e-5 * K ** 0.5 * 4 + 1e-6
    assert (got - ref).abs().max().item() <= max(tol, 1e-4), mnk
    two = ops.gemm_f16(ctx, A, B, bias, impl=2)
    assert torch.equal(got, two), "same MMA order per tile as the 2-CTA kernel: results must be bit-identical"
