"""tcgen05 layer-chain machinery vs a numpy emulation (fp16 operands, fp32 accumulate)."""
import ctypes

import numpy as np
import pytest
import torch

from njf_b200 import _lib

pytestmark = pytest.mark.gpu


def f16(x):
    return x.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("ntiles,grid", [(1, 1), (2, 1), (5, 2), (9, 3)])
def test_chain_matches_numpy(ntiles, grid):
    rng = np.random.default_rng(ntiles * 7 + grid)
    w0 = (rng.standard_normal((128, 64)) * 0.2).astype(np.float32)
    w1 = (rng.standard_normal((128, 128)) * 0.15).astype(np.float32)
    w2 = (rng.standard_normal((128, 128)) * 0.15).astype(np.float32)
    w3 = (rng.standard_normal((16, 128)) * 0.15).astype(np.float32)
    bias = (rng.standard_normal((3, 128)) * 0.3).astype(np.float32)
    rows = ntiles * 128
    a = rng.standard_normal((rows, 64)).astype(np.float32)
    tz = rng.standard_normal((rows, 128)).astype(np.float32)

    # biases ride in the weight images (fp16) and are accumulated by the tensor core
    x = f16(a) @ f16(w0).T + f16(bias[0])
    v = x + f16(tz)
    net = f16(np.maximum(v, 0)) @ f16(w1).T + f16(bias[1])
    x2 = v + f16(np.maximum(net, 0)) @ f16(w2).T + f16(bias[2])
    y = f16(np.maximum(x2, 0)) @ f16(w3).T

    dev = torch.device("cuda:0")
    a_d = torch.from_numpy(a).to(dev)
    tz_d = torch.from_numpy(tz).to(dev)
    x_d = torch.full((rows, 128), float("nan"), device=dev)
    y_d = torch.full((rows, 16), float("nan"), device=dev)
    L = _lib.lib()
    hp = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    _lib.check(
        L.njf_selftest_chain(hp(w0), hp(w1), hp(w2), hp(w3), hp(bias), _lib.ptr(a_d), _lib.ptr(tz_d),
                             _lib.ptr(x_d), _lib.ptr(y_d), ntiles, grid,
                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    )
    torch.cuda.synchronize()
    xg, yg = x_d.cpu().numpy(), y_d.cpu().numpy()
    assert np.isfinite(xg).all() and np.isfinite(yg).all()
    # fp16 operand rounding can flip on an accumulation-order ulp: allow isolated 1-ulp-of-fp16 flips
    np.testing.assert_allclose(xg, x2, rtol=4e-3, atol=8e-3)
    np.testing.assert_allclose(yg, y, rtol=4e-3, atol=8e-3)
    assert np.mean(np.abs(xg - x2) > 2e-3) < 1e-3
