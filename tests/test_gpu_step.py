"""GPU parity of the fused drivers (pop_step = dhdt + baroclinic_driver + barotropic_driver +
baroclinic_correct_adjust + halo updates + time-level rotation) against the CPU oracle.

Tolerances (BASELINE.json north_star): fields agree within a relative tolerance of 1e-12 after one
step; KMT/KMU masking (exact zeros below the bottom) and solver iteration counts are bit-exact.
Everything before the elliptic solve is bit-identical; the only source of rounding differences is
the order of the global dot products (double-double tree on the GPU, sequential on the CPU)."""
import numpy as np
import pytest

from parity import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu
RTOL_1STEP = 1.0e-12
PROG = ("TRACER", "UVEL", "VVEL", "RHO", "PSURF", "UBTROP", "VBTROP", "GRADPX", "GRADPY")


def compare(o, p, rtol, tag):
    """field-max norm against rtol (north_star: 1e-12 after one step); the pointwise norm with an absolute floor of 1e-3
    of the field maximum is held to 1e3 x rtol, i.e. every cell larger than a thousandth of the maximum agrees to
    rtol x 1e3 of ITS OWN magnitude at worst, and is printed (pytest -s / -rP)."""
    worst = worst_pt = 0.0
    for n in PROG + ("PGUESS",):
        for t in (c.TIME_OLD, c.TIME_CUR, c.TIME_NEW):
            if n == "PGUESS" and t != c.TIME_CUR:
                continue
            a = oracle_global(o, n, t)
            b = pop_global(p, n, t)
            e, ept = relerr(b, a), relerr_pointwise(b, a)
            worst, worst_pt = max(worst, e), max(worst_pt, ept)
            assert e <= rtol, "%s: %s[%d] relative error %.3e > %.1e" % (tag, n, t, e, rtol)
            assert ept <= 1.0e3 * rtol, "%s: %s[%d] pointwise relative error %.3e" % (tag, n, t, ept)
            # masking is exact: wherever the oracle has an exact zero (land / below the bottom) so do we
            assert np.array_equal(a == 0.0, b == 0.0), "%s: %s[%d] zero-mask differs" % (tag, n, t)
    print("%s: relerr field-max %.2e, pointwise (floor 1e-3) %.2e" % (tag, worst, worst_pt))
    return worst


CASES = {
    # config 1/2 flavour: centred advection + del2 + const vmix with convective diffusion, ChronGear
    "del2_chrongear": dict(nx=48, ny=36, km=8, seed=21, convergence_criterion=1e-12),
    # config 4 flavour: tripole, variable biharmonic mixing, KPP-shaped given coefficients, P-CSI
    "tripole_del4_pcsi": dict(nx=48, ny=36, km=8, seed=22, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                              hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21,
                              am=-27.0e21, given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0),
    # gx flavour: upwind3 advection, pcg, extra passive tracer, cyclic north-south, no convective diffusion
    "upwind3_pcg": dict(nx=40, ny=32, km=7, nt=3, seed=23, ns=c.BNDY_CYCLIC, tadvect=c.TADVECT_UPWIND3,
                        solver_choice=c.SOLVER_PCG, convection_diff=0),
    # config 1 flavour (input_templates/test_pop2_in): Richardson-number vertical mixing, pcg, del2, centred advection
    "rich_pcg": dict(nx=48, ny=32, km=8, seed=25, vmix_itype=c.VMIX_RICH, solver_choice=c.SOLVER_PCG,
                     convergence_criterion=1e-12),
    "rich_noconv_chrongear": dict(nx=40, ny=32, km=7, seed=26, vmix_itype=c.VMIX_RICH, convection_diff=0,
                                  convergence_criterion=1e-12),
    # gx1v7 flavour (BASELINE config 3): GM/Redi tracer mixing, KPP-shaped given coefficients, P-CSI, extra tracer
    "gm_pcsi": dict(nx=48, ny=36, km=8, nt=3, seed=27, hmix_tracer_itype=c.HMIX_GM, given_vmix=True,
                    solver_choice=c.SOLVER_PCSI, dtt=1800.0),
    # GM in its general skew-flux form (ah_bolus != ah, slm_b != slm_r) on the tripole grid, convective diffusion
    "gm_general_tripole": dict(nx=40, ny=32, km=7, nt=2, seed=28, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_GM,
                               ah_gm=0.6e7, ah_bolus=0.4e7, ah_bkg_srfbl=0.5e7, slm_b=0.2,
                               convergence_criterion=1e-12),
    # convection_type = 'adjustment': two convad passes (vertical_mix.F90:1888-2027) instead of convective diffusion
    "convad_chrongear": dict(nx=40, ny=32, km=8, nt=3, seed=29, convection_diff=0, nconvad=2, convergence_criterion=1e-12),
    # partial bottom cells (grid.F90:917-1022 and the DZT/DZU branches of every operator): the production variant of
    # config 4 (tripole, variable del4, KPP-shaped coefficients, P-CSI) ...
    "pbc_tripole_del4_pcsi": dict(nx=48, ny=36, km=8, seed=32, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                                  hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21,
                                  am=-27.0e21, given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0,
                                  partial_bottom_cells=1),
    # ... with del2 mixing, const coefficients, an extra tracer and convective adjustment ...
    "pbc_del2_convad_chrongear": dict(nx=40, ny=32, km=8, nt=3, seed=33, convection_diff=0, nconvad=2,
                                      convergence_criterion=1e-12, partial_bottom_cells=1),
    # ... and with explicit vertical mixing (vdifft / vdiffu carry the cell thicknesses)
    "pbc_explicit_vmix": dict(nx=40, ny=32, km=6, seed=34, implicit_vertical_mix=0, convection_diff=0,
                              partial_bottom_cells=1),
    # lw_lim advection (advection.F90:2684-3280 + comp_flux_vel_ghost :1014-1120) for every tracer ...
    "lw_lim_chrongear": dict(nx=48, ny=36, km=8, nt=3, seed=35, tadvect=c.TADVECT_LW_LIM, convergence_criterion=1e-12),
    # ... mixed with the centred scheme on the tripole grid (config 5: passive tracers on lw_lim) ...
    "lw_lim_mixed_tripole_pcsi": dict(nx=48, ny=36, km=8, nt=4, seed=36, ns=c.BNDY_TRIPOLE,
                                      tadvect=[c.TADVECT_CENTERED, c.TADVECT_CENTERED, c.TADVECT_LW_LIM, c.TADVECT_LW_LIM],
                                      hmix_tracer_itype=c.HMIX_DEL4, hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1,
                                      lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21, given_vmix=True,
                                      solver_choice=c.SOLVER_PCSI, dtt=600.0),
    # ... mixed with upwind3 (an odd tracer count: a half-filled last pass), cyclic north-south ...
    "lw_lim_upwind3_cyclic": dict(nx=40, ny=32, km=7, nt=3, seed=37, ns=c.BNDY_CYCLIC,
                                  tadvect=[c.TADVECT_UPWIND3, c.TADVECT_LW_LIM, c.TADVECT_LW_LIM], solver_choice=c.SOLVER_PCG),
    # ... and with partial bottom cells
    "pbc_lw_lim": dict(nx=40, ny=32, km=8, nt=3, seed=38, tadvect=c.TADVECT_LW_LIM, given_vmix=True,
                       convergence_criterion=1e-12, partial_bottom_cells=1),
    # explicit vertical mixing, rigid-lid-free options off: no pressure averaging, no implicit Coriolis
    "explicit_options": dict(nx=40, ny=32, km=6, seed=24, implicit_vertical_mix=0, convection_diff=0,
                             lpressure_avg=0, impcor=0, lbouss_correct=0, state_range_iopt=c.STATE_RANGE_IGNORE),
}


def compare_bitwise(o, p, tag):
    for n in PROG + ("PGUESS",):
        for t in (c.TIME_OLD, c.TIME_CUR, c.TIME_NEW):
            if n == "PGUESS" and t != c.TIME_CUR:
                continue
            a, b = oracle_global(o, n, t), pop_global(p, n, t)
            assert np.array_equal(a, b), "%s: %s[%d] not bit-identical: %d cells, max rel %.3e" % (
                tag, n, t, np.count_nonzero(a != b), relerr(b, a))


@pytest.mark.parametrize("sums", ["r8", "r16"])
@pytest.mark.parametrize("name", list(CASES))
def test_step_sequence_matches_oracle(name, sums):
    """sums = r8: the oracle with the reference's default r8 global sums, 1e-12 per step; sums = r16: the oracle in the
    reference's REPRODUCIBLE build (r16 sums, rounded once), every field of every time level bit-identical."""
    kw = dict(CASES[name])
    cs = make_case(kw.pop("nx"), kw.pop("ny"), kw.pop("km"), **kw)
    o, p = load_oracle(cs, reproducible=(sums == "r16")), load_pop(cs)
    try:
        # first step is forward Euler, then leapfrog, an averaging step, leapfrog again
        for i, ts in enumerate([c.TS_EULER, c.TS_LEAPFROG, c.TS_AVG, c.TS_LEAPFROG]):
            assert o.step(ts) == 0
            p.step(ts)
            it_o, res_o = o.solver_diag()
            it_p, res_p = p.solvers_get_diagnostics()
            assert it_o == it_p, "%s step %d: solver iterations %d (oracle) vs %d" % (name, i, it_o, it_p)
            if sums == "r16":
                assert res_o == res_p
                compare_bitwise(o, p, "%s step %d" % (name, i))
            else:
                compare(o, p, RTOL_1STEP * (i + 1), "%s step %d" % (name, i))
    finally:
        p.finalize()


@pytest.mark.parametrize("name,alpha", [("del2_chrongear", 0.53), ("tripole_del4_pcsi", 0.53), ("upwind3_pcg", 1.0)])
def test_robert_filter_steps_match_oracle(name, alpha):
    """Leapfrog steps closed by the Robert-Asselin-Williams filter (step_RF, step_mod.F90:919-1354).
    alpha = 0.53 filters both the new and the current level; alpha = 1 (plain Robert-Asselin) only the
    current one and averages the conservation factor with the previous step's (:1176-1182)."""
    kw = dict(CASES[name])
    cs = make_case(kw.pop("nx"), kw.pop("ny"), kw.pop("km"), robert_alpha=alpha, **kw)
    o, p = load_oracle(cs), load_pop(cs)
    try:
        for i, ts in enumerate([c.TS_EULER, c.TS_ROBERT, c.TS_ROBERT, c.TS_ROBERT]):
            assert o.step(ts) == 0
            p.step(ts)
            assert o.solver_diag()[0] == p.solvers_get_diagnostics()[0]
            compare(o, p, RTOL_1STEP * (i + 1), "%s robert step %d" % (name, i))
    finally:
        p.finalize()


def test_robert_filter_needs_variable_thickness_surface_layer():
    cs = make_case(32, 24, 6, seed=3, sfc_layer_type=c.SFC_RIGID)
    p = load_pop(cs)
    try:
        p.step(c.TS_EULER)
        with pytest.raises(Exception, match="varthick"):
            p.step(c.TS_ROBERT)
    finally:
        p.finalize()


def test_ten_steps_stay_within_solver_tolerance():
    cs = make_case(48, 36, 8, seed=31, given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=900.0)
    o, p = load_oracle(cs), load_pop(cs)
    try:
        seq = [c.TS_EULER, c.TS_AVG] + [c.TS_LEAPFROG] * 8
        for ts in seq:
            assert o.step(ts) == 0
            p.step(ts)
            assert o.solver_diag()[0] == p.solvers_get_diagnostics()[0]
        # "stay within the solver convergence tolerance over 10 steps": criterion 1e-13 (relative rms)
        compare(o, p, 1.0e-10, "10 steps")
    finally:
        p.finalize()


def test_product_is_decomposition_independent_of_oracle_blocks():
    """The oracle run as 4x3 blocks of 12x12 (the reference's multi-block layout) and the library's
    single strip give the same physical-domain answer."""
    cs = make_case(48, 36, 6, seed=41)
    o, p = load_oracle(cs, block_size=(12, 12)), load_pop(cs)
    try:
        for ts in (c.TS_EULER, c.TS_LEAPFROG):
            assert o.step(ts) == 0
            p.step(ts)
            assert o.solver_diag()[0] == p.solvers_get_diagnostics()[0]
        compare(o, p, 2.0e-12, "blocks")
    finally:
        p.finalize()


def test_conservation_constant_preservation_and_masking():
    """Size-independent properties (bench.py repeats the oracle comparison at 1200 x 800 x 62, `parity_check`):
    (1) cells below KMT stay exactly 0; (2) with a rigid lid and closed/cyclic boundaries the flux-form
    advection + mixing + implicit vertical mixing conserve the volume integral of every tracer;
    (3) with the ocean at rest a spatially constant tracer stays constant."""
    cs = make_case(64, 40, 10, nt=3, seed=51, sfc_layer_type=c.SFC_RIGID, lpressure_avg=0, given_vmix=True)
    kmask = (np.arange(1, cs.km + 1)[:, None, None] <= cs.kmt[None]).astype(np.float64)
    for f in ("STF", "TFW"):
        cs.forcing[f][:] = 0.0
    vol = kmask * cs.dz[:, None, None] * (cs.grid["DXT"] * cs.grid["DYT"])[None]
    p = load_pop(cs)
    try:
        Told = pop_global(p, "TRACER", c.TIME_OLD).reshape(3, cs.km, cs.ny, cs.nx)
        p.step(c.TS_EULER)
        T = pop_global(p, "TRACER", c.TIME_CUR).reshape(3, cs.km, cs.ny, cs.nx)
        assert np.array_equal(T[:, kmask == 0.0], np.zeros_like(T[:, kmask == 0.0]))
        for n in range(3):
            before, after = np.sum(Told[n] * vol), np.sum(T[n] * vol)
            assert abs(after - before) <= 1.0e-11 * np.sum(np.abs(Told[n]) * vol), (n, before, after)
    finally:
        p.finalize()
    for lev in ("cur", "old"):
        cs.state["TRACER_" + lev][2] = 3.25 * kmask
        cs.state["UVEL_" + lev][:] = 0.0
        cs.state["VVEL_" + lev][:] = 0.0
    p = load_pop(cs)
    try:
        p.step(c.TS_EULER)
        T = pop_global(p, "TRACER", c.TIME_CUR).reshape(3, cs.km, cs.ny, cs.nx)
        assert np.max(np.abs(T[2][kmask == 1.0] - 3.25)) <= 3.25 * 1.0e-12
    finally:
        p.finalize()


def test_step_coupled_host_buffers_match_resident_step():
    """pop_step_coupled (forcing in from host buffers, surface state out to a host buffer, copies overlapped with the
    step on their own stream) against the resident step fed the same forcing, with forcing that CHANGES every step
    (a copy that arrived late or an output that left early would show) and every kind of time step."""
    cs = make_case(72, 40, 6, seed=61, ns=c.BNDY_TRIPOLE, given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0)
    steps = [c.TS_EULER, c.TS_LEAPFROG, c.TS_AVG, c.TS_LEAPFROG, c.TS_ROBERT, c.TS_LEAPFROG]
    rng = np.random.default_rng(5)
    ocean, oceanu = (cs.kmt > 0), (cs.kmu > 0)
    forcing = [dict(STF=np.ascontiguousarray(1.0e-4 * rng.standard_normal((cs.nt, cs.ny, cs.nx)) * ocean),
                    SMF=np.ascontiguousarray(0.5 * rng.standard_normal((2, cs.ny, cs.nx)) * oceanu),
                    SHF_QSW=np.ascontiguousarray(1.0e-3 * rng.standard_normal((cs.ny, cs.nx)) * ocean),
                    FW=np.ascontiguousarray(1.0e-6 * rng.standard_normal((cs.ny, cs.nx)) * ocean)) for _ in steps]
    ref = []
    p = load_pop(cs)
    try:
        for ts, f in zip(steps, forcing):
            for name in ("STF", "SMF", "SHF_QSW", "FW"):
                p.scatter(name, 0, f[name])
            # the coupled entry point fills the ghost cells of FW (the one flux read off its own cell: DH - FW is averaged
            # to U points), as update_ghost_cells_coupler_fluxes does in the reference
            p.halo_field("FW", 0, c.LOC_CENTER, c.KIND_SCALAR)
            p.step(ts)
            ref.append({n: pop_global(p, n, c.TIME_CUR) for n in ("TRACER", "PSURF", "UVEL", "VVEL")})
    finally:
        p.finalize()
    p = load_pop(cs)
    try:
        for i, (ts, f) in enumerate(zip(steps, forcing)):
            out = np.full((5, cs.ny, cs.nx), np.nan)
            p.step_coupled(ts, f["STF"], f["SMF"], f["SHF_QSW"], f["FW"], out)
            r = ref[i]
            assert np.array_equal(out[0], r["TRACER"][0]) and np.array_equal(out[1], r["TRACER"][cs.km]), i
            assert np.array_equal(out[2], r["PSURF"][0]), i
            assert np.array_equal(out[3], r["UVEL"][0]) and np.array_equal(out[4], r["VVEL"][0]), i
            # the resident state is the same too
            assert np.array_equal(pop_global(p, "TRACER", c.TIME_CUR), r["TRACER"]), i
    finally:
        p.finalize()


def test_pcsi_two_iterations_per_pass_is_bitwise_the_single_pass_solver(monkeypatch):
    """The temporally blocked P-CSI kernel (two iterations per pass, pcsi_iter2_kernel) must produce
    the same bits and the same iteration count as one iteration per pass, on a grid with several tiles
    in both directions, the tripole seam and the cyclic east-west seam (edge tiles are partial)."""
    import os
    cs = make_case(200, 90, 5, seed=71, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                   hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21,
                   given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0)
    res = {}
    for tag, env in (("blocked", "0"), ("single", "1")):
        monkeypatch.setenv("POP_B200_NO_PCSI_BLOCKING", env)
        p = load_pop(cs)
        try:
            its = []
            for ts in (c.TS_EULER, c.TS_LEAPFROG):
                p.step(ts)
                its.append(p.solvers_get_diagnostics()[0])
            res[tag] = (its, pop_global(p, "PSURF", c.TIME_CUR), pop_global(p, "UVEL", c.TIME_CUR))
        finally:
            p.finalize()
    assert res["blocked"][0] == res["single"][0]
    assert min(res["blocked"][0]) >= 60
    assert np.array_equal(res["blocked"][1], res["single"][1])
    assert np.array_equal(res["blocked"][2], res["single"][2])
    # and against the oracle
    o = load_oracle(cs)
    for ts in (c.TS_EULER, c.TS_LEAPFROG):
        assert o.step(ts) == 0
    assert o.solver_diag()[0] == res["blocked"][0][-1]
    assert relerr(res["blocked"][1], oracle_global(o, "PSURF", c.TIME_CUR)) <= 2.0e-12


@pytest.mark.parametrize("ns,ny,depth", [(c.BNDY_TRIPOLE, 90, "12"), (c.BNDY_TRIPOLE, 87, "22"), (c.BNDY_CLOSED, 64, "12"),
                                         (c.BNDY_TRIPOLE, 90, "8")])
def test_pcsi_deep_strip_layout_is_bitwise_the_plain_solver(monkeypatch, ns, ny, depth):
    """The P-CSI passes of the multi-rank solver run on strips with a deep ghost zone (one strip exchange per `depth`
    iterations instead of one per pass; ghost cells with a source in the same strip -- east-west wrap, tripole fold --
    written by the pass itself).  POP_B200_DEEP_HALO_FORCE runs that layout on one rank, where everything but the peer
    exchange is exercised: same bits, same iteration counts as the plain layout."""
    cs = make_case(200, ny, 5, seed=73, ns=ns, hmix_tracer_itype=c.HMIX_DEL4,
                   hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21,
                   given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0)
    res = {}
    monkeypatch.setenv("POP_B200_DEEP_HALO", depth)
    monkeypatch.setenv("POP_B200_DEEP_OVERLAP", "1" if depth == "22" else "0")   # the opt-in split of the pass after an exchange
    for tag, env in (("plain", "0"), ("deep", "1"), ("lagged", "0")):
        monkeypatch.setenv("POP_B200_DEEP_HALO_FORCE", env)
        # the multi-rank solver also reads the verdict of a convergence check one check period late (X_m of the check
        # is kept aside meanwhile): same answer, same iteration count
        monkeypatch.setenv("POP_B200_LAGGED_CHECKS", "1" if tag != "plain" else "0")
        p = load_pop(cs)
        try:
            its = []
            for ts in (c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG):
                p.step(ts)
                its.append(p.solvers_get_diagnostics()[0])
            res[tag] = (its, {n: pop_global(p, n, c.TIME_CUR) for n in PROG})
        finally:
            p.finalize()
    assert res["plain"][0] == res["deep"][0] == res["lagged"][0]
    assert min(res["plain"][0]) >= 60
    for n in PROG:
        assert np.array_equal(res["plain"][1][n], res["deep"][1][n]), n
        assert np.array_equal(res["plain"][1][n], res["lagged"][1][n]), n


@pytest.mark.parametrize("flags", [("POP_B200_NO_TMA",), ("POP_B200_NO_FAST_TRACER",), ("POP_B200_NO_OVERLAP",),
                                   ("POP_B200_NO_TMA", "POP_B200_NO_OVERLAP"), ("POP_B200_OVERLAP_FINISH",),
                                   ("POP_B200_THOMAS_TMA",)])
def test_staging_and_overlap_variants_are_bitwise_identical(monkeypatch, flags):
    """The TMA-staged kernels, the specialised leapfrog tracer kernel, the placement of the velocity finish (fused after
    the barotropic solve with the barotropic add, inside baroclinic_driver, or on a side stream under the solve) and the
    opt-in TMA-staged Thomas kernels are variants of one computation: switching between them must not change a single
    bit, on a tripole grid with partial edge tiles."""
    cs = make_case(104, 52, 6, seed=72, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                   hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21,
                   given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0)
    res = []
    for on in (False, True):
        for f in flags:
            monkeypatch.setenv(f, "1" if on else "0")
        p = load_pop(cs)
        try:
            for ts in (c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG):
                p.step(ts)
            res.append({n: pop_global(p, n, c.TIME_CUR) for n in PROG})
        finally:
            p.finalize()
    for n in PROG:
        assert np.array_equal(res[0][n], res[1][n]), n


def test_partial_bottom_cell_fast_tracer_kernel_is_bitwise_the_general_kernel(monkeypatch):
    """Partial bottom cells: the leapfrog-specialised tracer kernel (PBC template branch) against the general column
    kernel (POP_B200_NO_PBC_FAST=1), tripole, variable del4, partial edge tiles."""
    cs = make_case(104, 52, 6, seed=74, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                   hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21, am=-27.0e21,
                   given_vmix=True, solver_choice=c.SOLVER_PCSI, dtt=600.0, partial_bottom_cells=1)
    res = []
    for flag in ("0", "1"):
        monkeypatch.setenv("POP_B200_NO_PBC_FAST", flag)
        p = load_pop(cs)
        try:
            for ts in (c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG):
                p.step(ts)
            res.append({n: pop_global(p, n, c.TIME_CUR) for n in PROG})
        finally:
            p.finalize()
    for n in PROG:
        assert np.array_equal(res[0][n], res[1][n]), n


@pytest.mark.parametrize("kw", [
    dict(nx=45, ny=33, km=5, seed=73, convergence_criterion=1e-12),                      # odd nx: no TMA descriptors
    dict(nx=37, ny=29, km=3, seed=74, ns=c.BNDY_CYCLIC, hmix_tracer_itype=c.HMIX_DEL4, hmix_momentum_itype=c.HMIX_DEL4,
         ah=-3.0e21, am=-27.0e21, solver_choice=c.SOLVER_PCSI, dtt=600.0, nt=3),        # minimum depth, odd extents
    dict(nx=72, ny=30, km=4, seed=76, ns=c.BNDY_CYCLIC, solver_choice=c.SOLVER_PCSI, dtt=600.0),   # P-CSI across a cyclic north-south seam
    dict(nx=33, ny=8, km=4, seed=75, hmix_tracer_itype=c.HMIX_GM, given_vmix=True, convergence_criterion=1e-12),  # one tile row
])
def test_ragged_extents_match_oracle(kw):
    """Extents that are not multiples of any tile size, odd nx (the non-TMA staging paths), the minimum km = 3 and a
    strip thinner than one tile: same parity bar as the regular cases."""
    kw = dict(kw)
    cs = make_case(kw.pop("nx"), kw.pop("ny"), kw.pop("km"), **kw)
    o, p = load_oracle(cs), load_pop(cs)
    try:
        for i, ts in enumerate([c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG]):
            assert o.step(ts) == 0
            p.step(ts)
            assert o.solver_diag()[0] == p.solvers_get_diagnostics()[0]
            compare(o, p, RTOL_1STEP * (i + 1), "ragged step %d" % i)
    finally:
        p.finalize()


def test_config1_shape_test_grid_options():
    """BASELINE.json config 1: the reference's own CPU-runnable case (input_templates/test_pop2_in and
    test_domain_size.F90: 192 x 128 x 20, 16 x 16 blocks, cyclic/closed, pcg + diagonal preconditioner,
    Richardson-number vertical mixing, centred advection, del2, MWJF 'enforce', pressure averaging,
    Boussinesq correction) -- the oracle runs it in the reference's multi-block layout, the library as
    one strip; synthetic bathymetry and fields of the same shape (no input files exist in the image)."""
    cs = make_case(192, 128, 20, seed=111, vmix_itype=c.VMIX_RICH, solver_choice=c.SOLVER_PCG,
                   convergence_criterion=1e-12, dtt=3600.0)
    o, p = load_oracle(cs, block_size=(16, 16)), load_pop(cs)
    try:
        for i, ts in enumerate([c.TS_EULER, c.TS_LEAPFROG, c.TS_LEAPFROG]):
            assert o.step(ts) == 0
            p.step(ts)
            assert o.solver_diag()[0] == p.solvers_get_diagnostics()[0]
        compare(o, p, 3.0e-12, "config 1 shape")
    finally:
        p.finalize()
