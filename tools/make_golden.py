#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the CPU oracle (oracle/, the C restatement of the reference
routines; the Fortran reference itself cannot be built in this image) for small seeded cases.  The
fixtures pin the oracle against drift (tests/test_oracle_golden.py, CPU) and give the GPU tests a
checker that does not need the oracle library (tests/test_gpu_golden.py).
   python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity import *  # noqa: F401,F403,E402

CASES = {
    # config-4 flavour: tripole, variable biharmonic mixing, KPP-shaped coefficients, P-CSI
    "tripole_del4_pcsi_40x28x6": dict(nx=40, ny=28, km=6, seed=91, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                                      hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21,
                                      am=-27.0e21, given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0),
    # config-1/2 flavour: closed, del2, constant vmix with convective diffusion, upwind3 on T,S, ChronGear
    "closed_del2_upwind3_chrongear_36x24x5": dict(nx=36, ny=24, km=5, seed=92, tadvect=c.TADVECT_UPWIND3,
                                                  convergence_criterion=1e-12),
    # gx1v7 flavour: GM/Redi in its general skew-flux form, KPP-shaped coefficients, P-CSI, a passive tracer;
    # Euler step, then two leapfrog steps closed by the Robert-Asselin-Williams filter
    "gm_general_robert_36x28x6": dict(nx=36, ny=28, km=6, nt=3, seed=93, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_GM,
                                      ah_gm=0.6e7, ah_bolus=0.4e7, ah_bkg_srfbl=0.5e7, slm_b=0.2, given_vmix=True,
                                      solver_choice=c.SOLVER_PCSI, dtt=1800.0,
                                      steps=(c.TS_EULER, c.TS_ROBERT, c.TS_ROBERT)),
    # convection_type = 'adjustment' (two convad passes), GM with the cancelling terms dropped, ChronGear
    "gm_convad_chrongear_32x24x6": dict(nx=32, ny=24, km=6, seed=94, hmix_tracer_itype=c.HMIX_GM, convection_diff=0,
                                        nconvad=2, convergence_criterion=1e-12),
}
FIELDS_OUT = ("TRACER", "UVEL", "VVEL", "RHO", "PSURF", "UBTROP", "VBTROP")
STEPS = (c.TS_EULER, c.TS_LEAPFROG)


def steps_of(name):
    return CASES[name].get("steps", STEPS)


def build_case(name):
    kw = dict(CASES[name])
    kw.pop("steps", None)
    return make_case(kw.pop("nx"), kw.pop("ny"), kw.pop("km"), **kw)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]          # optional: names of the cases to (re)generate
    for name in CASES:
        if only and name not in only:
            continue
        cs = build_case(name)
        o = load_oracle(cs)
        its = []
        for ts in steps_of(name):
            assert o.step(ts) == 0
            its.append(o.solver_diag()[0])
        data = {n: oracle_global(o, n, c.TIME_CUR) for n in FIELDS_OUT}
        data["solver_iterations"] = np.array(its, dtype=np.int64)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **data)
        print(name, its, {n: data[n].shape for n in FIELDS_OUT})


if __name__ == "__main__":
    main()
