"""tcgen05 plumbing self-test: TF32 MMA with hand-built descriptors vs numpy.

Only K-major operands are exercised: on this B200 / CUDA 12.9 the MN-major no-swizzle
descriptor form produced all-zero accumulators (selftest variants 1-3), so the production
kernels re-tile the transposed operands instead (idrlnet_b200/csrc/wide_dw.cu).  Variant 4
reads the A operand from TMEM (tcgen05.st + the TS form of tcgen05.mma)."""
import numpy as np
import pytest
import torch

from idrlnet_b200 import native as nv

pytestmark = pytest.mark.gpu


def _tf32(a):
    """round-to-nearest-away to 10 mantissa bits (cvt.rna.tf32.f32)."""
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("variant,N,K", [(0, 16, 8), (0, 64, 32), (0, 240, 104), (0, 256, 64), (0, 112, 128),
                                         (4, 16, 8), (4, 64, 32), (4, 112, 160), (4, 256, 64)])  # 4: A operand from TMEM
def test_tc_gemm(variant, N, K):
    rng = np.random.default_rng(N + K + variant)
    amn, bmn = variant in (1, 2), variant in (1, 2, 3)
    A = rng.standard_normal((K, 128) if amn else (128, K)).astype(np.float32)
    B = rng.standard_normal((K, N) if bmn else (N, K)).astype(np.float32)
    a64 = _tf32(A).astype(np.float64)
    b64 = _tf32(B).astype(np.float64)
    ref = (a64.T if amn else a64) @ (b64 if bmn else b64.T)
    dA, dB = torch.as_tensor(A, device="cuda"), torch.as_tensor(B, device="cuda")
    D = torch.zeros(128, N, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    nv.check(nv.lib().pinn_tc_selftest(variant, N, K, dA.data_ptr(), dB.data_ptr(), D.data_ptr(), status.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(status[0]) == 0, "MMA completion barrier timed out"
    got = D.cpu().numpy()
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err < 1e-5, (err, got[:2, :4], ref[:2, :4])
