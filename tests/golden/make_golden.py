"""Regenerate tests/golden/oracle_golden.npz from the CPU oracle.

The reference ships no golden vectors and cannot be built here, so these fixtures are
OUTPUTS OF THE ORACLE (parity unpinned, see oracle/kestrel_oracle.cpp): they freeze the
oracle's answers bit for bit so that (a) the CPU suite notices any drift of the oracle and
(b) the GPU suite has committed vectors to compare with on a box without the oracle's history.
Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import INPUTS, domain_stepper, run_input  # noqa: E402
from kestrel_b200 import capi  # noqa: E402
from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state  # noqa: E402


def cases(lib):
    out = {}
    rs = dambreak_runset(1, 32)
    q4, b0v = dambreak_state(rs)
    st = domain_stepper(lib, rs, q4, b0v)
    info = st.integrate_to(1e9, 40)
    out["dambreak32_q4"] = st.download_domain()
    out["dambreak32_t"] = np.array([info.t, info.dt_last])
    rs = dambreak_runset(1, 24, morpho=True)
    q4, b0v = dambreak_state(rs)
    st = domain_stepper(lib, rs, q4, b0v)
    info = st.integrate_to(1e9, 12)
    q, bt = st.download_domain(True)
    out["morpho24_q4"], out["morpho24_bt"] = q, bt
    out["morpho24_t"] = np.array([info.t, info.dt_last, info.nrefines])
    sim = run_input(lib, os.path.join(INPUTS, "case_1d_cap_constslope.txt"), tend=10.0, Nout=1)
    for k, t in sim.snapshots[-1].items():
        out[f"cap1d_tile{k}_u"] = t["u"][..., :4]
    out["cap1d_active"] = sim.stepper.active_tiles()
    sim = run_input(lib, os.path.join(INPUTS, "case_cap_morpho.txt"), tend=2.0, Nout=1)
    for k, t in sim.snapshots[-1].items():
        out[f"capmorpho_tile{k}_u"] = t["u"][..., :4]
        out[f"capmorpho_tile{k}_bt"] = t["bt"]
    out["capmorpho_active"] = sim.stepper.active_tiles()
    sim = run_input(lib, os.path.join(INPUTS, "case_lake_at_rest_hydro_2d.txt"), tend=1.0, Nout=1)
    out["lake2d_w"] = sim.snapshots[-1][1]["u"][..., 0]
    return out


def main():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = capi.Library(os.path.join(ROOT, "oracle", "libkestrel_oracle.so"), "kor_")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"), **cases(lib))


if __name__ == "__main__":
    main()
