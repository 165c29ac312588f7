"""Generates the committed golden fixtures of the VOF+CMOM path (tests/golden/*.npz).

The reference (TzuYaoHuang/InterfaceAdvection.jl) is 100 % Julia and no Julia toolchain exists in this image, so the
fixtures cannot come from the reference itself: they are produced by the CPU oracle (oracle/, the C++ restatement that
tests/test_oracle_kat.py pins against every known-answer test of the reference, test/maintests.jl) and frozen here so
that (a) a later change of the oracle cannot silently move the target and (b) the GPU parity tests have fixed
input/output vectors that travel to the GPU box.  The reference's own known-answer values for the path are transcribed
(values only) into reference_kats.json.

    python tests/golden/make_golden.py        # rewrites tests/golden/cmom_*.npz, vof_*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pyoracle as O  # noqa: E402
from tests.helpers import make_state, oracle_mom_advect_step  # noqa: E402

# (name, N, kind, perdir, uBC, dirO)
CMOM = [
    ("cmom_walls_f64", (20, 12, 10), "C3", (), (0.0, 0.0, 0.0), (3, 1, 2), np.float64),
    ("cmom_walls_f32", (20, 12, 10), "C3", (), (0.0, 0.0, 0.0), (3, 1, 2), np.float32),
    ("cmom_periodic_f64", (16, 12, 12), "C4", (1, 2), (0.0, 0.0, 0.0), (2, 3, 1), np.float64),
    ("cmom_periodic_f32", (16, 12, 12), "C4", (1, 2), (0.0, 0.0, 0.0), (2, 3, 1), np.float32),
    ("cmom_inflow_f64", (18, 10, 8), "C3", (3,), (0.3, 0.0, 0.0), (1, 2, 3), np.float64),
    ("cmom_2d_f64", (24, 20), "C1", (), (0.0, 0.0), (2, 1), np.float64),
]
VOF = [
    ("vof_enright_f64", (16, 16, 16), "C2", (), (1, 2, 3), np.float64),
    ("vof_zalesak_f64", (32, 32), "C1", (), (2, 1), np.float64),
]


def gen_cmom(name, N, kind, perdir, uBC, dirO, T):
    st = make_state(N, kind, T, perdir=perdir, uBC=uBC)
    f = st["f"].copy(order="F")
    rhou = oracle_mom_advect_step(st, f, st["u"], 1.0, dirO)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), N=np.array(N), perdir=np.array(perdir, dtype=np.int64), uBC=np.array(uBC),
                        dirO=np.array(dirO), lam_rho=st["lam_rho"], kind=kind, f_in=st["f"], u=st["u"], f_out=f, rhou_out=rhou)


def gen_vof(name, N, kind, perdir, dirO, T):
    st = make_state(N, kind, T, perdir=perdir)
    D = len(N)
    Ng = st["Ng"]
    f = st["f"].copy(order="F")
    z = lambda *s: O.zeros(s, T)
    ff, al, nh, rhouf = z(*Ng), z(*Ng), z(*Ng, D), z(*Ng, D)
    cbar = np.zeros(Ng, dtype=np.int8, order="F")
    for _ in range(3):
        O.advectVOF(f, ff, al, nh, st["u"], st["u"], 1.0, cbar, rhouf, st["lam_rho"], "WH", st["perdir"], dirO)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), N=np.array(N), perdir=np.array(perdir, dtype=np.int64), dirO=np.array(dirO),
                        lam_rho=st["lam_rho"], kind=kind, f_in=st["f"], u=st["u"], f_out=f, rhouf_out=rhouf, steps=3)


if __name__ == "__main__":
    O.build()
    for c in CMOM:
        gen_cmom(*c)
    for c in VOF:
        gen_vof(*c)
    print("golden fixtures written to", HERE)
