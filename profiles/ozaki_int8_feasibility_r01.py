"""CPU feasibility study for round 2: fp64-grade Legendre GEMM on the int8 tensor pipe (Ozaki-type error-free
slicing), run on the REAL operands of the transform (oracle Legendre tables, synthetic spectra of SURVEY 8d).

C[lat][r] = sum_k P[k][lat] S[k][r].  Row i of A = P^T (one latitude) is scaled by a power of two so that
|a| < 1 and cut into `nsl` signed slices of `w` bits, a = sum_s a_s 2^{-w (s+1)}, a_s integer in [-2^{w-1}, 2^{w-1}];
likewise every column of B = S.  All slice products a_s b_t are exact integer GEMMs (int32 accumulation is exact:
K * 2^{2w-2} < 2^31 for K <= 641, w <= 7); products with s + t >= nsl are dropped.  Reported: max error relative
to max |C| (the transform's own tolerance is 1e-13 rms / 1e-12 max) and the number of int8 GEMMs.

Usage: python profiles/ozaki_int8_feasibility_r01.py   (needs oracle/liboracle.so; a few seconds)
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import helpers as H  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def slices(x, axis, w, nsl):
    """Error-free split along `axis` (the contraction axis): returns integer slices and the power-of-two scales."""
    amax = np.abs(x).max(axis=axis, keepdims=True)
    amax[amax == 0] = 1.0
    e = np.ceil(np.log2(amax)) + 1          # |x| / 2^e < 1/2
    r = x / np.exp2(e)
    out = []
    for _ in range(nsl):
        r = r * (1 << w)
        q = np.rint(r)                       # |q| <= 2^{w-1}
        out.append(q.astype(np.int64))
        r = r - q
    return out, e


def ozaki_gemm(A, B, w, nsl):
    """A: (M, K) sliced along axis 1; B: (K, N) sliced along axis 0."""
    As, ea = slices(A, 1, w, nsl)
    Bs, eb = slices(B, 0, w, nsl)
    C = np.zeros((A.shape[0], B.shape[1]))
    ngemm = 0
    for d in range(nsl):                     # anti-diagonals s + t = d share one weight (one int32 accumulator)
        acc = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
        for s in range(d + 1):
            acc += As[s] @ Bs[d - s]
            ngemm += 1
        assert np.abs(acc).max() < 2 ** 31
        C += acc.astype(np.float64) * np.exp2(-w * (d + 2))
    return C * np.exp2(ea) * np.exp2(eb), ngemm


def main():
    N, T, nf = 400, 399, 16
    lat, wts = po.gaussian_quadrature(N)
    nx = np.array([20 + 4 * j for j in range(N)] + [20 + 4 * j for j in range(N - 1, -1, -1)], dtype=np.int32)
    plan = po.OraclePlan(nx, lat, T, weights=wts)
    sym, asym, sb, ab = plan.tables()
    nlat0 = plan.nlat0()
    sp = H.synthetic_spectra(T, nf).reshape(-1, 2, nf)
    print(f"grid O{N}, T{T}, {nf} fields; columns = 2 x fields")
    print(" m   K  lats |  w nsl GEMMs  max err / max|C| (inverse)   (direct, quadrature-weighted)")
    for m in (0, 1, 50, 200, 350):
        K = (T + 1 - m + 2) // 2             # symmetric block of the table built for T+1: n = T+1, T-1, ... (k = 0 highest n)
        blk = sym[int(sb[m]):int(sb[m]) + K * N].reshape(N, K)   # [lat][k]
        j0 = int(nlat0[m])
        P = blk[j0:, :]                      # (lats, K)
        n_of_k = (T + 1) - 2 * np.arange(K)
        S = np.zeros((K, 2 * nf))
        for k, n in enumerate(n_of_k):
            if m <= n <= T and (m < T):
                c = (2 * T + 3 - m) * m // 2 + (n - m)
                S[k, 0::2] = sp[c, 0]
                S[k, 1::2] = sp[c, 1]
        ref = P.astype(np.longdouble) @ S.astype(np.longdouble)
        # direct: X[k][r] = sum_lat w P[k][lat] G[lat][r] with G = the inverse's output (band limited)
        G = np.asarray(ref, dtype=np.float64) * wts[j0:N, None]
        refd = P.T.astype(np.longdouble) @ G.astype(np.longdouble)
        for w, nsl in ((7, 6), (7, 7), (7, 8), (6, 8), (6, 9)):
            C, ng = ozaki_gemm(P, S, w, nsl)
            err = float(np.abs(C - ref).max() / np.abs(ref).max())
            Cd, _ = ozaki_gemm(np.ascontiguousarray(P.T), G, w, nsl)
            errd = float(np.abs(Cd - refd).max() / np.abs(refd).max())
            print(f"{m:3d} {K:3d} {P.shape[0]:4d} | {w:2d} {nsl:3d} {ng:4d}   {err:10.2e}                  {errd:10.2e}")
        e64 = float(np.abs(P @ S - ref).max() / np.abs(ref).max())
        print(f"            fp64 GEMM (what DMMA computes):            {e64:10.2e}")


if __name__ == "__main__":
    main()
