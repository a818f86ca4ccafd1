"""tcgen05 building block: split-fp16 GEMM on the tensor cores vs fp64 (through the C ABI)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(128, 128, 128), (256, 128, 64), (128, 256, 128), (128, 64, 32), (640, 128, 128)])
def test_split_fp16_umma_gemm(shape):
    from umma_probe import run
    err, scale = run(*shape)
    assert err is not None, scale
    # 3-term fp16 split: ~2^-21 relative per product, fp32 accumulation
    assert err < 2e-5 * max(scale, 1.0), (err, scale)


@pytest.mark.parametrize("B,n,H,C", [(3, 128, 1, 32), (2, 64, 1, 64), (2, 37, 2, 64),          # narrow kernel (tiny / small)
                                     (3, 128, 2, 128), (2, 100, 2, 128), (2, 64, 1, 128),      # wide kernel, base block 0
                                     (3, 64, 4, 256), (2, 33, 4, 256), (5, 1, 2, 128)])        # base block 1, degenerate n
def test_attention_core_tensor_core_and_simt(B, n, H, C):
    """softmax(scale q k^T) v for every head geometry of tiny / small / base, tcgen05 and SIMT kernels against float64
    (layers/blocks.py:44-63: heads are full width, no mask on the scores)."""
    import torch
    from efficientspeech_b200 import _cabi
    dev = "cuda:0"
    g = torch.Generator().manual_seed(B * 1000 + n + C)
    qkv = torch.randn(B, n, 3 * H * C, generator=g) * 0.7
    scale = float((C // H) ** -0.5)
    q, k, v = qkv.double().reshape(B, n, 3, H, C).permute(2, 0, 3, 1, 4).unbind(0)
    want = (((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1) @ v).transpose(1, 2).reshape(B, n, H * C)
    x = qkv.to(dev)
    lib = _cabi.load()
    for tc in (1, 0):
        out = torch.full((B, n, H * C), float("nan"), device=dev)
        with torch.cuda.device(dev):
            _cabi.check(lib.es_selftest_attention(torch.cuda.current_stream().cuda_stream, B, n, C, H, scale, x.data_ptr(),
                                                  out.data_ptr(), tc))
            _cabi.check(lib.es_check_async_errors(torch.cuda.current_stream().cuda_stream))
        err = float((out.cpu().double() - want).abs().max())
        assert err <= 2e-5 * max(1.0, float(want.abs().max())), (tc, err)


def test_attention_envelope_is_reported():
    import torch
    from efficientspeech_b200 import _cabi
    x = torch.zeros(1, 128, 3 * 256, device="cuda:0")
    out = torch.zeros(1, 128, 256, device="cuda:0")
    rc = _cabi.load().es_selftest_attention(torch.cuda.current_stream().cuda_stream, 1, 128, 256, 1, 1.0, x.data_ptr(), out.data_ptr(), 1)
    assert rc == 2                                   # C = 256 with n > 64 does not fit: the model falls back to SIMT there
