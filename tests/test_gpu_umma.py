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
