"""GPU parity at the BENCHMARK configurations (BASELINE.json configs 2, 3, 4): the named grid shapes or their full
column depth with the real vertical grids (`input_templates/{gx3v7,gx1v7,tx0.1v3}_vert_grid`: km = 60 / 62), the
production time steps and option sets -- the cases the small synthetic grids of test_gpu_step.py do not reach: the
8-level chunk tails of the Thomas kernels at km = 60 and 62, the 3-deep TMA ring wrapping over 62 levels, tiles in
both directions.  `bench.py` repeats the tx0.1v3 check at 1200 x 800 x 62 next to its CPU baseline (`parity_check`).

Bars: slab operators bit-exact.  Steps are compared twice: against the oracle with the reference's default r8 global
sums (1e-12 per step in the field-max norm, exact zero masks, equal solver iteration counts; the pointwise error with
an absolute floor of 1e-3 of the field maximum is reported) and against the oracle in the reference's REPRODUCIBLE
build (r16 sums), where every field of every step must be BIT-IDENTICAL (the library's sums are correctly rounded
too, so nothing is left to differ).  gx3v7 + ChronGear is compared in the REPRODUCIBLE mode only: on that grid the
conjugate-gradient recurrence amplifies the r8 summation-order rounding of the dot products to ~1e-9 in PSURF (both
sides converge to the same residual, 4.45e-13; the emulated CPU build of the library shows the same numbers), which
says something about r8 dot products, nothing about the kernels."""
import numpy as np
import pytest

from parity import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu
vp, ci = C.c_void_p, C.c_int
PROG = ("TRACER", "UVEL", "VVEL", "RHO", "PSURF", "UBTROP", "VBTROP", "GRADPX", "GRADPY")


def run_steps(cs, steps, tag, rtol=1.0e-12, block=None, sums="r16"):
    """sums = "r16": oracle in the REPRODUCIBLE build, bit-identical fields required; "r8": default sums, rtol per step"""
    o, p = load_oracle(cs, block_size=block, reproducible=(sums == "r16")), load_pop(cs)
    worst = worst_pt = 0.0
    try:
        for i, ts in enumerate(steps):
            assert o.step(ts) == 0, "%s: oracle step %d failed (solver did not converge?)" % (tag, i)
            p.step(ts)
            (it_o, res_o), (it_p, res_p) = o.solver_diag(), p.solvers_get_diagnostics()
            assert it_o == it_p, "%s step %d: solver iterations %d (oracle) vs %d" % (tag, i, it_o, it_p)
            if sums == "r16":
                assert res_o == res_p, "%s step %d: rms residual %r vs %r" % (tag, i, res_o, res_p)
            for n in PROG:
                a, b = oracle_global(o, n, c.TIME_CUR), pop_global(p, n, c.TIME_CUR)
                if sums == "r16":
                    assert np.array_equal(a, b), "%s step %d: %s is not bit-identical (%d cells, max rel %.3e)" % (
                        tag, i, n, np.count_nonzero(a != b), relerr(b, a))
                    continue
                e, ept = relerr(b, a), relerr_pointwise(b, a)
                worst, worst_pt = max(worst, e), max(worst_pt, ept)
                assert e <= rtol * (i + 1), "%s step %d: %s relative error %.3e" % (tag, i, n, e)
                assert np.array_equal(a == 0.0, b == 0.0), "%s step %d: %s zero-mask differs" % (tag, i, n)
        print("%s [%s sums]: %d steps, solver iterations %d, %s" % (
            tag, sums, len(steps), it_o, "every field bit-identical" if sums == "r16" else
            "max relerr field-max %.2e pointwise(floor 1e-3) %.2e" % (worst, worst_pt)))
    finally:
        p.finalize()


@pytest.mark.parametrize("vmix", ["const", "rich"])
def test_config2_gx3v7_shape(vmix):
    """BASELINE config 2: gx3v7 shape 100 x 116 x 60, gx3v7 vertical grid, third-order upwind advection, Laplacian
    mixing (ah = 1e7), const / Richardson-number vertical mixing applied implicitly, ChronGear, dt = 86400/12 s."""
    cs = make_case(100, 116, 60, seed=201, vgrid="gx3v7", tadvect=c.TADVECT_UPWIND3, ah=1.0e7, am=1.0e8,
                   vmix_itype=c.VMIX_RICH if vmix == "rich" else c.VMIX_CONST, solver_choice=c.SOLVER_CHRONGEAR,
                   convergence_criterion=1.0e-12, dtt=7200.0)
    run_steps(cs, [c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG], "gx3v7/" + vmix, block=(25, 29), sums="r16")


@pytest.mark.parametrize("sums", ["r16", "r8"])
def test_config3_gx1v7_shape(sums):
    """BASELINE config 3: gx1v7 shape 320 x 384 x 60, gx1v7 vertical grid, GM/Redi (constant kappa, notanh), KPP-shaped
    given coefficients, P-CSI with the production criterion 1e-13, dt = 3600 s.  The oracle runs 8 x 8 blocks of 40 x 48."""
    cs = make_case(320, 384, 60, seed=202, vgrid="gx1v7", hmix_tracer_itype=c.HMIX_GM, given_vmix=True,
                   solver_choice=c.SOLVER_PCSI, convergence_criterion=1.0e-13, max_lanczos_step=100,
                   lanczos_convergence_criterion=0.15, dtt=3600.0, am=0.6e8)
    run_steps(cs, [c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG], "gx1v7", block=(40, 48), sums=sums)


@pytest.mark.parametrize("sums", ["r16", "r8"])
def test_config4_tx01v3_columns(sums):
    """BASELINE config 4 at full column depth: tx0.1v3 vertical grid (km = 62), tripole, centred advection, variable
    biharmonic mixing of tracers and momentum, KPP-shaped given coefficients, P-CSI 1e-13, on a 384 x 240 grid (several
    tiles of every kernel in both directions, a partial last tile row); dt scaled with the grid spacing."""
    scale = 3600.0 / 384
    cs = make_case(384, 240, 62, seed=203, vgrid="tx0.1v3", ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                   hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e17 * scale ** 3,
                   am=-27.0e17 * scale ** 3, given_vmix=True, solver_choice=c.SOLVER_PCSI,
                   convergence_criterion=1.0e-13, max_lanczos_step=100, lanczos_convergence_criterion=0.15,
                   dtt=288.0 * scale)
    run_steps(cs, [c.TS_EULER, c.TS_LEAPFROG, c.TS_AVG, c.TS_LEAPFROG], "tx0.1v3-columns", block=(48, 40), sums=sums)


@pytest.mark.parametrize("km,vgrid", [(60, "gx1v7"), (62, "tx0.1v3")])
def test_thomas_solves_at_full_depth(km, vgrid):
    """impvmixt / impvmixt_correct / impvmixu (vertical_mix.F90:1164, :1460, :1679) at km = 60 and 62 with the real
    layer thicknesses: bit-exact, including the partial last chunk of the level loops, on a row wider than one CTA."""
    cs = make_case(272, 12, km, nt=3, seed=210 + km, vgrid=vgrid, given_vmix=True, dtt=1800.0)
    o, p = load_oracle(cs), load_pop(cs)
    try:
        phys = lambda a: a[..., 2:-2, 2:-2]
        Tc = o.view("TRACER", c.TIME_CUR, (o.nt, o.km))[0].copy()
        To = o.view("TRACER", c.TIME_OLD, (o.nt, o.km))[0].copy()
        Uo = o.view("UVEL", c.TIME_OLD, (o.km,))[0].copy()
        Vo = o.view("VVEL", c.TIME_OLD, (o.km,))[0].copy()
        Ps = o.view("PSURF", c.TIME_CUR)[0].copy()
        o.set_timestep(c.TS_LEAPFROG)
        p.set_timestep(c.TS_LEAPFROG)
        rhs = np.ascontiguousarray(1.0e-3 * (Tc - To))
        f_t = osig(o.L, "o_impvmixt", [vp, vp, vp, ci, ci, ci])
        f_c = osig(o.L, "o_impvmixt_correct", [vp, vp, vp, ci, ci, ci])
        f_u = osig(o.L, "o_impvmixu", [vp, vp, ci])
        a, b = rhs.copy(), rhs.copy()
        f_t(op(a), op(To), op(Ps), 1, o.nt, 0)
        p.impvmixt(b, To, Ps, 1, o.nt)
        assert np.array_equal(phys(a), phys(b))
        assert np.abs(phys(a) - phys(To)).max() > 0
        R1 = np.ascontiguousarray(1.0e-2 * Tc[:, 0])
        a, b = Tc.copy(), Tc.copy()
        f_c(op(a), op(Ps), op(R1), 1, 2, 0)
        p.impvmixt_correct(b, Ps, R1, 1, 2)
        assert np.array_equal(phys(a), phys(b))
        ua, va, ub, vb = Uo.copy(), Vo.copy(), Uo.copy(), Vo.copy()
        f_u(op(ua), op(va), 0)
        p.impvmixu(ub, vb)
        assert np.array_equal(phys(ua), phys(ub)) and np.array_equal(phys(va), phys(vb))
        # the deepest columns reach the last level, so the chunk tail was exercised
        assert cs.kmt.max() == km
    finally:
        p.finalize()
