import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.dev.debug_sweep import case
for T, tol in [(np.float64, 1e-12)]:
    for N, kind, per in [((16, 12, 10), "C2", ()), ((16, 12, 10), "C3", (1, 2, 3)), ((40, 24, 44), "C3", ()), ((70, 40, 36), "C3", (2,)), ((33, 20, 9), "C4", (1, 2))]:
        for dirO in [(1, 2, 3), (3, 1, 2)]:
            case(N, kind, per, T, dirO, tol)
