"""Run the tcgen05 GEMM self-test for a few shapes and print the error vs fp64 (run under gpurun)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientspeech_b200 import _cabi  # noqa: E402


def run(M, N, K, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c = torch.full((M, N), float("nan"), device="cuda")
    rc = _cabi.load().es_selftest_umma_gemm(torch.cuda.current_stream().cuda_stream, M, N, K,
                                            a.data_ptr(), b.data_ptr(), c.data_ptr())
    if rc != 0:
        return None, _cabi.load().es_last_error().decode()
    torch.cuda.synchronize()
    got = c.cpu().numpy().astype(np.float64)
    return float(np.abs(got - want).max()), float(np.abs(want).max())


if __name__ == "__main__":
    print("variant", os.environ.get("ES_UMMA_VARIANT", "0"))
    for shape in [(128, 128, 128), (256, 128, 64), (128, 256, 128), (128, 64, 32), (384, 128, 16)]:
        print(shape, run(*shape))
