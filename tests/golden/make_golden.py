"""Generates the committed golden fixtures from the reference tree (run in the build container only,
where /root/reference exists; the GPU box never needs it).

  gaussian_latitudes_N{32,400,1280}.npy : the reference's tabulated Gaussian latitudes (12 decimals),
        parsed from src/atlas/grid/detail/spacing/gaussian/N{32,400,1280}.cc (DEFINE_GAUSSIAN_LATITUDES lists)
  legendre_ref_T33.npz : output of the UNMODIFIED reference compute_legendre_polynomials_lat
        (src/atlas/trans/local/LegendrePolynomials.cc, compiled in place into oracle/_ref) at truncation 33
        for four latitudes, plus compute_legendre_polynomials tables for a 4-latitude set.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/atlas"
sys.path.insert(0, REPO)


def parse_table(N):
    txt = open(f"{REF}/grid/detail/spacing/gaussian/N{N}.cc").read()
    body = txt[txt.index("LIST(") + 5:]
    vals = [float(x) for x in re.findall(r"[-+]?\d+\.\d+", body)]
    assert len(vals) == N, (N, len(vals))
    return np.array(vals)


def main():
    for N in (32, 400, 1280):
        np.save(os.path.join(HERE, f"gaussian_latitudes_N{N}.npy"), parse_table(N))
    from oracle import pyoracle as po

    ref = po.load_ref_legendre()
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    trc = 33
    lats = np.deg2rad(np.array([87.863798839233, 45.0, 1.395306910819, 89.9999999]))
    pol = np.stack([po.legendre_lat(trc, float(l), ref=True) for l in lats])
    np.savez(os.path.join(HERE, "legendre_ref_T33.npz"), trc=trc, lats=lats, legpol=pol)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
