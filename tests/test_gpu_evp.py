"""GPU parity of the EVP block preconditioner (preconditionerChoice = 'evp', the production default:
POP_SolversMod.F90 preconditioner :2273-2369, EvpPre :2434-2506, ExplicitBlockEvpPre :2508-2616, ExplicitEvp
:2618-2696, EvpBlockPartition :2992-3040) with each of the three solvers and the Lanczos eigenvalue estimate.

The test grids use dt = 7200 s: on a 100 x 52 grid the cells are ~400 km wide, and with a short time step the
diagonal term TAREA/(beta c2dtp dtp g) dwarfs the corner weights, which makes the north-east marching of the EVP method
amplify by |centre/corner| per cell (the reference's own failure mode: 'Check EVP sub-block size!').  With the
production ratio (~4, as on the tx0.1v3 grid with dt = 288 s) the influence matrices invert to ~1e-12."""
import numpy as np
import pytest

from parity import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu
PROG = ("TRACER", "UVEL", "VVEL", "RHO", "PSURF", "UBTROP", "VBTROP", "GRADPX", "GRADPY")
SOLVERS = {"pcsi": c.SOLVER_PCSI, "chrongear": c.SOLVER_CHRONGEAR, "pcg": c.SOLVER_PCG}


def evp_case(solver, precond=c.PRECOND_EVP, nx=100, ny=52, **kw):
    return make_case(nx, ny, 6, seed=81, ns=c.BNDY_TRIPOLE, given_vmix=True, solver_choice=solver, dtt=7200.0,
                     preconditioner_choice=precond, max_lanczos_step=100, lanczos_convergence_criterion=0.15, **kw)


def prep_both(cs, o, p):
    if cs.cfg.solver_choice != c.SOLVER_PCSI:      # load_* already ran POP_SolversPrep for P-CSI
        assert o.solvers_prep() == 0
        p.solvers_prep()


@pytest.mark.parametrize("solver,sums", [("pcsi", "r16"), ("chrongear", "r16"), ("pcg", "r16"), ("pcsi", "r8")])
def test_evp_steps_match_oracle(solver, sums):
    """r16: the oracle in the reference's REPRODUCIBLE build, every field bit-identical.  r8 (default sums, 1e-12 per step)
    only for P-CSI, whose recurrence contains no dot product: with this dt the conjugate-gradient recurrences of
    ChronGear / pcg amplify the r8 summation-order rounding of their dot products beyond 1e-12 (as in
    test_gpu_benchmark_shapes.py::test_config2_gx3v7_shape), which is a property of r8 sums, not of the kernels."""
    cs = evp_case(SOLVERS[solver])
    o, p = load_oracle(cs, reproducible=(sums == "r16")), load_pop(cs)
    try:
        prep_both(cs, o, p)
        for i, ts in enumerate([c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG]):
            assert o.step(ts) == 0
            p.step(ts)
            (it_o, res_o), (it_p, res_p) = o.solver_diag(), p.solvers_get_diagnostics()
            assert it_o == it_p, "%s step %d: iterations %d vs %d" % (solver, i, it_o, it_p)
            for n in PROG:
                a, b = oracle_global(o, n, c.TIME_CUR), pop_global(p, n, c.TIME_CUR)
                if sums == "r16":
                    assert res_o == res_p
                    assert np.array_equal(a, b), "%s step %d: %s not bit-identical (max rel %.2e)" % (solver, i, n, relerr(b, a))
                else:
                    assert relerr(b, a) <= 1.0e-12 * (i + 1), (solver, i, n, relerr(b, a))
                    assert np.array_equal(a == 0.0, b == 0.0)
    finally:
        p.finalize()


def test_evp_setup_matches_oracle_and_reference_selfcheck():
    """The influence-matrix inverses pass the reference's own self-test (max |rinv*rin - I| <= 1e-8,
    POP_SolversMod.F90:2594-2614), the value and the land / ocean sub-block counts equal the oracle's, and so do the
    Lanczos eigenvalue bounds of the EVP-preconditioned operator."""
    cs = evp_case(c.SOLVER_PCSI)
    o, p = load_oracle(cs, reproducible=True), load_pop(cs)
    try:
        nsub, nland, err = p.solvers_get_evp_diagnostics()
        o.L.oracle_evp_selfcheck.restype = C.c_double
        ns_o, nl_o = C.c_int(), C.c_int()
        assert o.L.oracle_evp_counts(C.byref(ns_o), C.byref(nl_o)) == 0
        assert (nsub, nland) == (ns_o.value, nl_o.value)
        assert 0 < nland < nsub
        assert err == o.L.oracle_evp_selfcheck()
        assert err <= 1.0e-8
        mn, mx = C.c_double(), C.c_double()
        o.L.oracle_get_eigs(C.byref(mn), C.byref(mx))
        assert p.solvers_get_eigs() == (mn.value, mx.value)
    finally:
        p.finalize()


def test_evp_needs_fewer_iterations_than_the_diagonal():
    its = {}
    for name, pc in (("diagonal", c.PRECOND_DIAGONAL), ("evp", c.PRECOND_EVP)):
        cs = evp_case(c.SOLVER_PCSI, precond=pc, nx=200, ny=120)
        p = load_pop(cs)
        try:
            p.step(c.TS_EULER)
            p.step(c.TS_LEAPFROG)
            its[name] = p.solvers_get_diagnostics()[0]
        finally:
            p.finalize()
    assert its["evp"] < its["diagonal"], its


def test_evp_without_prep_fails_loudly():
    cs = evp_case(c.SOLVER_CHRONGEAR)
    p = load_pop(cs)      # ChronGear: load_pop does not call POP_SolversPrep
    try:
        with pytest.raises(P.api.PopError, match="pop_solvers_prep"):
            p.step(c.TS_EULER)
    finally:
        p.finalize()
