"""Fit of the exact-erf GELU used by xf_kernel (csrc/xf_head.cu: xf_gelu2).

gelu(v) = v * Phi(v) = relu(v) - |v|/2 * erfc(|v| / sqrt(2)); erfc(a / sqrt(2)) = 2^-Q(a) with Q a degree-6
polynomial on [0, 6] (|v| is clamped to 6 inside Q only; erfc(6/sqrt2) = 2e-9).  Weighted least squares on
Chebyshev nodes, weight = sensitivity of gelu to an error in Q.  Prints the fp32 coefficients c0..c6."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erfc

AMAX, DEG, N = 6.0, 6, 20000


def fit():
    a = (np.cos(np.pi * (np.arange(N) + 0.5) / N) + 1) / 2 * AMAX
    q = -np.log2(erfc(a / np.sqrt(2)))
    wgt = np.maximum(a * erfc(a / np.sqrt(2)), 1e-4)
    v = C.chebvander(2 * a / AMAX - 1, DEG)
    coef, *_ = np.linalg.lstsq(v * wgt[:, None], q * wgt, rcond=None)
    mono = P.Polynomial(C.cheb2poly(coef))(P.Polynomial([-1, 2 / AMAX]))
    return mono.coef.astype(np.float32)


if __name__ == "__main__":
    print(", ".join("%.9ef" % c for c in fit()))
