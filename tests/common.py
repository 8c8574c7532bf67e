"""Shared helpers of the test-suite."""
import os

import numpy as np

from kestrel_b200 import capi
from kestrel_b200.host.inputfile import read_input_file
from kestrel_b200.host.run import Simulation
from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUTS = os.path.join(ROOT, "tests", "inputs")

FIELD_NAMES = ["w", "rhoHnu", "rhoHnv", "Hnpsi", "Hn", "u", "v", "psi", "rho", "b0", "bt", "dbdx", "dbdy"]


def rel_linf(a: np.ndarray, b: np.ndarray) -> float:
    """Relative L-infinity distance used for the 1e-10 parity criterion (BASELINE.json)."""
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def domain_stepper(lib, rs, q4, b0v, btv=None):
    p, keep = rs.to_c()
    st = capi.Stepper(lib, p, keep)
    st.upload_domain(q4, b0v, btv)
    return st


def run_input(lib, path, **over):
    rs = read_input_file(path)
    for k, v in over.items():
        setattr(rs, k, v)
    rs.finalize()
    return Simulation(rs, lib).run()


def compare_snapshots(sa, sb, fields=range(13)):
    """max rel-Linf per field over all tiles of two snapshot dicts {tile: {'u':..}}."""
    assert sorted(sa) == sorted(sb), f"active sets differ: {sorted(sa)} vs {sorted(sb)}"
    out = {}
    for d in fields:
        A = np.concatenate([sa[k]["u"][..., d].ravel() for k in sorted(sa)])
        B = np.concatenate([sb[k]["u"][..., d].ravel() for k in sorted(sb)])
        out[FIELD_NAMES[d]] = (rel_linf(A, B), bool(np.array_equal(A, B)))
    return out
