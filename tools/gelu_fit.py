"""Coefficients of the GELU epilogue (csrc/gemm_sm100.cu): GELU(x) = max(x, 0) - |x| h(|x|), h(a) = 0.5 erfc(a / sqrt 2) =
exp2(Q(min(a, 6))), Q = weighted least-squares polynomial fit of log2 h on Chebyshev nodes of [0, 6].  Prints the fp32
coefficients and the error of an fp32 evaluation against the exact function (also after bf16 rounding), next to the
Abramowitz-Stegun 7.1.26 form used before.  Needs numpy / scipy / torch (CPU)."""
import numpy as np
import torch
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf, erfc

TM, DEG, WPOW = 6.0, 6, 0.75


def gelu(x):
    return 0.5 * x * (1 + erf(x / np.sqrt(2)))


def fit():
    t = np.cos(np.pi * (np.arange(8000) + 0.5) / 8000) * TM / 2 + TM / 2
    y = np.log2(0.5 * erfc(t / np.sqrt(2)))
    w = (0.5 * erfc(t / np.sqrt(2)) * np.maximum(t, 0.05)) ** WPOW      # ~ contribution of an error in Q to the GELU value
    pu = C.cheb2poly(C.chebfit(2 * t / TM - 1, y, DEG, w=w))
    lin, acc = np.array([-1.0, 2 / TM]), np.zeros(1)
    for k, coef in enumerate(pu):
        acc = P.polyadd(acc, coef * P.polypow(lin, k))
    return acc.astype(np.float32)


def new(xs, c):
    a = np.minimum(np.abs(xs), np.float32(TM)).astype(np.float32)
    q = np.full_like(a, c[-1])
    for cf in c[-2::-1]:
        q = (q * a + cf).astype(np.float32)
    h = np.exp2(q.astype(np.float64)).astype(np.float32)
    return (np.maximum(xs, 0) - np.abs(xs) * h).astype(np.float32)


def old(x):
    x = x.astype(np.float64)
    ax = np.abs(x)
    t = 1 / (1 + 0.3275911 * 0.70710678118654752440 * ax)
    p = 0.5 * 1.061405429 * t + 0.5 * -1.453152027
    for cf in (1.421413741, -0.284496736, 0.254829592):
        p = p * t + 0.5 * cf
    return (-ax * (p * t) * np.exp2((x * -0.72134752044448170368) * x) + np.maximum(x, 0)).astype(np.float32)


if __name__ == "__main__":
    c = fit()
    print("coefficients a^0..a^6:", ", ".join(repr(float(v)) + "f" for v in c))
    for lo, hi in ((-3, 3), (-5, -3), (-8, 8)):
        xs = np.linspace(lo, hi, 1000001).astype(np.float32)
        ref = gelu(xs.astype(np.float64))
        rb = torch.tensor(ref).float().bfloat16()
        for name, g in (("exp2(Q)", new(xs, c)), ("A-S 7.1.26", old(xs))):
            mism = (torch.tensor(g).bfloat16() != rb).float().mean().item()
            print(f"[{lo}, {hi}] {name:11s}: max |err| {np.abs(g - ref).max():.2e}, bf16 roundings that differ {mism:.2e}")
