"""GPU tests of the tcgen05 tensor-core path: building-block self-test (descriptor encodings, TMA
swizzle, TMEM loads) and the fused DAMSM kernels against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def agb():
    import attention_gan_b200 as pkg
    return pkg


@pytest.mark.parametrize("N,K", [(128, 64), (256, 256), (128, 256), (256, 128)])
@pytest.mark.parametrize("bf16", [0, 1])
@pytest.mark.parametrize("manual_a", [0, 1])
def test_tcgen05_building_blocks(agb, N, K, bf16, manual_a):
    g = torch.Generator().manual_seed(N + K + bf16)
    dt = torch.bfloat16 if bf16 else torch.float16
    A = torch.randn(128, K, generator=g).to(dt).cuda()
    B = torch.randn(N, K, generator=g).to(dt).cuda()
    C = torch.zeros(128, N, device="cuda")
    lib = agb.native.lib()
    rc = lib.agb_tc_selftest(A.data_ptr(), B.data_ptr(), C.data_ptr(), N, K, bf16, manual_a,
                             torch.cuda.current_stream().cuda_stream)
    agb.native.check(rc, "agb_tc_selftest")
    torch.cuda.synchronize()
    ref = A.double().cpu() @ B.double().cpu().T
    err = (C.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, f"tcgen05 GEMM mismatch: rel err {err:.3e}"
