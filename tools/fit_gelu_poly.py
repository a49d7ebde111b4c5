"""Coefficients of the MUFU-free GELU / GELU' evaluation in csrc/common.cuh (poly_cdf2 / poly_gelu_grad2): Chebyshev least-squares fits
of (f(u) - 0.5) / u in s = u^2 * (2 / U^2) - 1 on |u| <= U, converted to monomials in s, and their error when evaluated with fp32
Horner steps including the clamp.  Runs on the CPU (numpy + scipy)."""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf

U = 4.5


def fit(deg, fn):
    u = np.linspace(1e-6, U, 80001)
    return C.cheb2poly(C.chebfit(2 * u * u / U ** 2 - 1, (fn(u) - 0.5) / u, deg))


def fp32_error(mono, fn):
    uu = np.linspace(-1.5 * U, 1.5 * U, 400001).astype(np.float32)
    uc = np.clip(uu, -U, U).astype(np.float32)
    s = ((uc * uc).astype(np.float32) * np.float32(2 / U ** 2) + np.float32(-1)).astype(np.float32)
    p = np.full_like(s, np.float32(mono[-1]))
    for k in range(len(mono) - 2, -1, -1):
        p = (p * s + np.float32(mono[k])).astype(np.float32)
    return np.abs((uc * p + np.float32(0.5)).astype(np.float32) - fn(uu.astype(np.float64))).max()


def phi(u):
    return 0.5 * (1 + erf(u / np.sqrt(2)))


def dgelu(u):
    return phi(u) + u * np.exp(-u * u / 2) / np.sqrt(2 * np.pi)


if __name__ == '__main__':
    for name, deg, fn in (('Phi', 10, phi), ("gelu'", 11, dgelu)):
        m = fit(deg, fn)
        print('%s: degree %d in s = u^2 * %.10f - 1, max abs error (fp32 Horner, clamp at %.1f) %.2e' % (name, deg, 2 / U ** 2, U, fp32_error(m, fn)))
        print('   coefficients (s^0 .. s^%d): %s' % (deg, ', '.join('%.9e' % v for v in m)))
