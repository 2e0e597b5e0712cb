"""CPU feasibility study for the next round: the metric build G = V . KR2(X) emulated with INT8 slices
(Ozaki scheme) -- how many slice products does the 1e-9 per-step parity need?

Rows of V (per chain) and columns of KR2(X) (per pair) are scaled by a power of two to |.| <= 1 and split into S signed
7-bit slices; the slice products with i + j < S are exact in int32 (max |sum| 1.3e7 for N = 1000) and are recombined in
FP64.  German-shaped data, 64 random positions:
    S = 4: 10 INT8 GEMMs, max rel err 2.2e-09      S = 6: 21 INT8 GEMMs, 1.4e-13
    S = 5: 15 INT8 GEMMs,             1.6e-11      S = 7: 28 INT8 GEMMs, 1.5e-15
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from riemannhamiltonianmontecarlo_b200 import datasets  # noqa: E402


def slices(a, axis, n_slices, bits=7):
    m = np.abs(a).max(axis=axis, keepdims=True)
    scale = 2.0 ** np.ceil(np.log2(m))
    r = a / scale
    out = []
    for _ in range(n_slices):
        q = np.round(r * 2 ** bits)
        out.append(q.astype(np.int64))
        r = r * 2 ** bits - q
    return out, scale


def main():
    xx, _ = datasets.shaped("german")
    d = xx.shape[1]
    theta = np.random.default_rng(0).normal(0, 0.3, (64, d))
    p = 1 / (1 + np.exp(-(theta @ xx.T)))
    v = p * (1 - p)
    ia, ib = np.triu_indices(d)
    kr = xx[:, ia] * xx[:, ib]
    g = v @ kr
    for s in (4, 5, 6, 7):
        vs, vsc = slices(v, 1, s)
        ks, ksc = slices(kr, 0, s)
        acc, n_prod = np.zeros_like(g), 0
        for i in range(s):
            for j in range(s - i):
                acc += (vs[i] @ ks[j]).astype(np.float64) * 2.0 ** (-7 * (i + j + 2))
                n_prod += 1
        err = np.abs(acc * vsc * ksc - g).max() / np.abs(g).max()
        print(f"S={s}: {n_prod} int8 GEMMs, max rel err {err:.2e}, max |int32 sum| {np.abs(vs[0] @ ks[0]).max():.3g}")


if __name__ == "__main__":
    main()
