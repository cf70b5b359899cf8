"""CPU: invariants that pin the parts of the oracle for which the reference tree holds no golden data
(SURVEY 8c): symmetry and negative definiteness of the barotropic operator, the residual of the
tridiagonal solves against an independently assembled matrix, the solver's reported residual against the
convergence criterion, constant preservation of the advection operator (flux form + comp_flux_vel), and the
biharmonic operator as two passes of the harmonic one."""
import ctypes as C

import numpy as np
import pytest

from parity import *  # noqa: F401,F403

ci, vp = C.c_int, C.c_void_p
GRAV = 980.6


def _case(**kw):
    base = dict(seed=5)
    base.update(kw)
    cs = make_case(40, 32, 7, **base)
    return cs, load_oracle(cs)


def test_btrop_operator_is_symmetric_and_negative_definite():
    cs, o = _case(ns=c.BNDY_CYCLIC)
    o.set_timestep(c.TS_LEAPFROG)
    osig(o.L, "o_solvers_prep", [], ci)()
    f_b = osig(o.L, "o_btrop_operator", [vp, vp, ci])
    rng = np.random.default_rng(1)
    mask = o.view("mMaskTropic", 1)[0]

    def vec():
        g = rng.standard_normal((cs.ny, cs.nx)) * (cs.kmt > 0)
        a = np.zeros((o.nyb, o.nxb))
        a[2:-2, 2:-2] = g
        A = np.ascontiguousarray(a[None])
        o.halo_array(A, c.LOC_CENTER, c.KIND_SCALAR, 0.0)
        return A[0]

    x, y = vec(), vec()
    ax, ay = np.zeros_like(x), np.zeros_like(x)
    f_b(op(ax), op(x), 0)
    f_b(op(ay), op(y), 0)
    dot = lambda u, w: float(np.sum((u * w * mask)[2:-2, 2:-2]))
    assert abs(dot(x, ay) - dot(ax, y)) <= 1e-12 * max(abs(dot(x, ay)), 1e-300)
    assert dot(x, ax) < 0.0 and dot(y, ay) < 0.0      # -(div H grad) - diagonal: negative definite on ocean points


def test_tridiagonal_solves_satisfy_their_matrix():
    cs, o = _case(given_vmix=True)
    o.set_timestep(c.TS_LEAPFROG)
    km, nt = o.km, o.nt
    To = o.view("TRACER", c.TIME_OLD, (nt, km))[0].copy()
    Ps = o.view("PSURF", c.TIME_CUR)[0].copy()
    VDC = o.view("VDC", 0, o.inner_shape("VDC"))[0]          # (2, km+2, nyb, nxb), level index 0..km+1
    kmt = o.view("KMT", 1, (), np.int32)[0]
    rng = np.random.default_rng(2)
    rhs = np.ascontiguousarray(rng.standard_normal(To.shape) * 1e-3)
    x = rhs.copy()
    osig(o.L, "o_impvmixt", [vp, vp, vp, ci, ci, ci])(op(x), op(To), op(Ps), 1, nt, 0)
    dz = cs.dz
    dzw = np.concatenate([[0.5 * dz[0]], 0.5 * (dz[:-1] + dz[1:])])   # dzw(0..km-1): between k and k+1 is dzw[k]
    c2dtt = 2.0 * cs.cfg.dtt
    worst = 0.0
    for n in range(nt):
        slot = min(n, 1)
        for (j, i) in [(5, 7), (12, 20), (20, 31), (25, 11), (9, 3)]:
            K = kmt[j, i]
            if K < 2:
                continue
            # (h_k + A_k + A_{k-1}) f_k - A_k f_{k+1} - A_{k-1} f_{k-1} = h_k rhs_k,  A_k = VDC(k)/dzw(k), A_K = 0
            A = np.array([VDC[slot, k + 1, j, i] / dzw[k + 1] for k in range(K)])  # A[k]: between level k+1 and k+2 (1-based)
            A[K - 1] = 0.0
            h = dz[:K] / c2dtt
            h[0] = dz[0] / c2dtt + Ps[j, i] / (GRAV * c2dtt)
            hr = dz[:K] / c2dtt
            f = x[n, :K, j, i] - To[n, :K, j, i]
            res = np.zeros(K)
            for k in range(K):
                up = A[k - 1] if k > 0 else 0.0
                res[k] = (h[k] + A[k] + up) * f[k] - (A[k] * f[k + 1] if k + 1 < K else 0.0) - (up * f[k - 1] if k > 0 else 0.0) \
                         - hr[k] * rhs[n, k, j, i]
            scale = np.max(np.abs(hr * rhs[n, :K, j, i])) + 1e-300
            worst = max(worst, np.max(np.abs(res)) / scale)
    assert worst <= 1e-10, worst


@pytest.mark.parametrize("solver", [c.SOLVER_PCG, c.SOLVER_CHRONGEAR, c.SOLVER_PCSI])
def test_solver_meets_its_convergence_criterion(solver):
    cs, o = _case(solver_choice=solver, convergence_criterion=1e-12, given_vmix=True)
    for ts in (c.TS_EULER, c.TS_LEAPFROG):
        assert o.step(ts) == 0
        its, rms = o.solver_diag()
        assert 0 < its < cs.cfg.max_iterations
        assert rms <= 1e-12        # rmsResidual = sqrt(rr * residualNorm) < convergenceCriterion (POP_SolversMod.F90:895-906)


def test_advection_preserves_a_constant_tracer():
    """L(1) = 0 by construction of WTKB: with T == const the advective tendency cancels to rounding (closed
    column: the fluxes through the top (rigid lid) and the bottom vanish)."""
    cs, o = _case(sfc_layer_type=c.SFC_RIGID)
    km, nt = o.km, o.nt
    U = o.view("UVEL", c.TIME_CUR, (km,))[0].copy()
    V = o.view("VVEL", c.TIME_CUR, (km,))[0].copy()
    kmt = o.view("KMT", 1, (), np.int32)[0]
    T = np.zeros((nt, km, o.nyb, o.nxb))
    for k in range(km):
        T[:, k] = 3.5 * (kmt > k)
    T = np.ascontiguousarray(T)
    f_a = osig(o.L, "o_advt", [ci, vp, vp, vp, vp, vp, vp, ci])
    WTK = np.zeros((o.nyb, o.nxb))
    speed = max(np.abs(U).max(), np.abs(V).max())
    for k in range(1, km + 1):
        L = np.zeros((nt, o.nyb, o.nxb))
        f_a(k, op(L), op(WTK), op(T), op(T), op(U), op(V), 0)
        inner = (kmt[3:-3, 3:-3] >= k)
        # interior ocean cells whose four neighbours are ocean at this level see T == const everywhere
        ok = inner & (kmt[3:-3, 2:-4] >= k) & (kmt[3:-3, 4:-2] >= k) & (kmt[2:-4, 3:-3] >= k) & (kmt[4:-2, 3:-3] >= k)
        ok &= (kmt[3:-3, 3:-3] > k)   # not the bottom cell of the column: there WTKB = 0 closes the column instead
        if ok.any():
            assert np.max(np.abs(L[0, 3:-3, 3:-3][ok])) <= 1e-12 * 3.5 * speed / cs.dz.min()


def test_del4_is_two_passes_of_the_del2_stencil():
    """hdifft_del4 = ah * L(AHF * L(T)) with the same masked 5-point weights as hdifft_del2 (ah = 1): compare
    on cells whose whole 2-ring is ocean and where AHF == 1 (constant coefficient)."""
    kw = dict(hmix_tracer_itype=c.HMIX_DEL4, hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=0, lvariable_hmixu=0, ah=-2.0, am=-2.0)
    cs4, o4 = _case(**kw)
    T = o4.view("TRACER", c.TIME_CUR, (o4.nt, o4.km))[0].copy()
    kmt = o4.view("KMT", 1, (), np.int32)[0].copy()      # the oracle is re-initialised below: no views into its memory
    f_h = osig(o4.L, "o_hdifft", [ci, vp, vp, vp, vp, ci])
    k = 2
    H4 = np.zeros((o4.nt, o4.nyb, o4.nxb))
    f_h(k, op(H4), op(T), None, None, 0)
    cs2, o2 = _case(hmix_tracer_itype=c.HMIX_DEL2, ah=1.0)
    f_h2 = osig(o2.L, "o_hdifft", [ci, vp, vp, vp, vp, ci])
    L1 = np.zeros((o2.nt, o2.nyb, o2.nxb))
    f_h2(k, op(L1), op(T), None, None, 0)                    # L(T) on the physical cells
    T2 = T.copy()
    T2[:, k - 1] = 0.0
    T2[:, k - 1, 2:-2, 2:-2] = L1[:, 2:-2, 2:-2]
    A = np.ascontiguousarray(T2[:, k - 1][None].reshape(1, o2.nt, o2.nyb, o2.nxb))
    o2.halo_array(A, c.LOC_CENTER, c.KIND_SCALAR, 0.0)
    T2[:, k - 1] = A[0]
    L2 = np.zeros((o2.nt, o2.nyb, o2.nxb))
    f_h2(k, op(L2), op(np.ascontiguousarray(T2)), None, None, 0)
    ocean = (kmt >= k)
    deep = ocean.copy()
    for dj in range(-2, 3):
        for di in range(-2, 3):
            deep &= np.roll(np.roll(ocean, dj, 0), di, 1)
    deep[:4] = deep[-4:] = False
    deep[:, :4] = deep[:, -4:] = False
    assert deep.sum() > 20
    ref = -2.0 * L2[0][deep]
    assert np.max(np.abs(H4[0][deep] - ref)) <= 1e-11 * np.max(np.abs(ref))


def _content(o, cs, tlev):
    """volume integrals of the tracers (variable-thickness surface layer) and the area integral of PSURF
    over the physical ocean cells of time level tlev"""
    T = oracle_global(o, "TRACER", tlev).reshape(o.nt, o.km, cs.ny, cs.nx)
    P = oracle_global(o, "PSURF", tlev).reshape(cs.ny, cs.nx)
    tarea = oracle_global(o, "TAREA", 0).reshape(cs.ny, cs.nx)
    dz = np.asarray(cs.dz, dtype=np.float64)
    lev = np.arange(1, o.km + 1)[:, None, None]
    thick = np.broadcast_to(dz[:, None, None], (o.km, cs.ny, cs.nx)).copy()
    thick[0] = dz[0] + P / GRAV
    vol = tarea[None] * thick * (cs.kmt[None] >= lev)
    return [float(np.sum(vol * T[n])) for n in range(o.nt)], float(np.sum(tarea * (cs.kmt >= 1) * P)), float(np.sum(vol))


@pytest.mark.parametrize("alpha", [0.53, 1.0])
def test_robert_filter_conserves_tracer_content_and_mean_surface_pressure(alpha):
    """step_RF's conservation terms (step_mod.F90:1114-1205): filtering changes the values of the current and
    new levels but neither their volume-integrated tracer content nor the area mean of PSURF, so a leapfrog
    step closed by the filter and a plain leapfrog step from the same state agree in those integrals."""
    res = {}
    for ts in (c.TS_LEAPFROG, c.TS_ROBERT):
        cs, o = _case(ns=c.BNDY_TRIPOLE, robert_alpha=alpha, convergence_criterion=1e-13)
        assert o.step(c.TS_EULER) == 0 and o.step(ts) == 0
        res[ts] = [_content(o, cs, t) for t in (c.TIME_OLD, c.TIME_CUR)]   # = (cur, new) of the step just taken
        fields = {t: oracle_global(o, "TRACER", t).copy() for t in (c.TIME_OLD, c.TIME_CUR)}
        if ts == c.TS_LEAPFROG:
            plain = fields
    for lvl in (0, 1):
        (ta, pa, va), (tb, pb, vb) = res[c.TS_LEAPFROG][lvl], res[c.TS_ROBERT][lvl]
        if alpha == 1.0 and lvl == 1:      # plain Robert-Asselin leaves the new level untouched
            assert np.array_equal(plain[c.TIME_CUR], fields[c.TIME_CUR])
        # the reference normalises by the CURRENT level's volume (:1162-1171, "placeholder for
        # lrf_nonzero_newtime; this is not correct" :1178), so the new level is conserved only as far as the
        # two surface volumes agree
        tol = 1e-12 if lvl == 0 else 1e-9
        for n in range(len(ta)):
            assert abs(ta[n] - tb[n]) <= tol * abs(ta[n]), (lvl, n, ta[n], tb[n])
        assert abs(pa - pb) <= 1e-12 * max(abs(pa), va * 1e-6)
    # ... while the fields themselves did change at the filtered level
    assert not np.array_equal(plain[c.TIME_OLD], fields[c.TIME_OLD])


def test_convective_adjustment_mixes_unstable_pairs_and_conserves_column_content():
    """convad (vertical_mix.F90:1888-2027): pairs that are statically unstable are replaced by their thickness-weighted
    mean, so the dz-weighted column integral of every tracer is unchanged."""
    new = {}
    for nconvad in (0, 2):
        cs, o = _case(nt=3, convection_diff=0, nconvad=nconvad, convergence_criterion=1e-13)
        assert o.step(c.TS_EULER) == 0
        new[nconvad] = oracle_global(o, "TRACER", c.TIME_CUR).reshape(o.nt, o.km, cs.ny, cs.nx).copy()
    a, b = new[0], new[2]
    changed = np.any(a != b, axis=(0,))
    assert changed.sum() > 50                                       # the synthetic state has unstable pairs
    dz = np.asarray(cs.dz)[None, :, None, None]
    lev = np.arange(1, cs.km + 1)[None, :, None, None]
    w = dz * (lev <= cs.kmt[None, None])
    ia, ib = np.sum(a * w, axis=1), np.sum(b * w, axis=1)
    assert np.max(np.abs(ia - ib)) <= 1e-12 * np.max(np.abs(ia))
    assert np.all(b[:, :, cs.kmt == 0] == 0.0)


def test_partial_bottom_cells_conserve_tracer_content():
    """Partial bottom cells (grid.F90:917-1022; DZT/DZU branches of comp_flux_vel, advt_centered, hdifft, vdifft,
    impvmixt): with a rigid lid, closed/cyclic boundaries and no surface fluxes the flux form conserves the volume
    integral of every tracer, the volumes being TAREA * DZT -- which only holds if every branch uses the cell
    thicknesses consistently (faces as thick as the thinner neighbour, vertical fluxes over the local distances)."""
    cs = make_case(64, 40, 10, nt=3, seed=51, sfc_layer_type=c.SFC_RIGID, lpressure_avg=0, given_vmix=True,
                   partial_bottom_cells=1)
    for f in ("STF", "TFW"):
        cs.forcing[f][:] = 0.0
    o = load_oracle(cs, block_size=(16, 20))
    lev = np.arange(1, cs.km + 1)[:, None, None]
    dzt = np.where(lev == cs.kmt[None], cs.dzbc[None], cs.dz[:, None, None])
    vol = (lev <= cs.kmt[None]) * dzt * (cs.grid["DXT"] * cs.grid["DYT"])[None]
    told = oracle_global(o, "TRACER", c.TIME_OLD).reshape(3, cs.km, cs.ny, cs.nx)
    assert o.step(c.TS_EULER) == 0
    t = oracle_global(o, "TRACER", c.TIME_CUR).reshape(3, cs.km, cs.ny, cs.nx)
    for n in range(3):
        before, after = np.sum(told[n] * vol), np.sum(t[n] * vol)
        assert abs(after - before) <= 1.0e-12 * np.sum(np.abs(told[n]) * vol), (n, before, after)
    # the depths follow the partial cells: HU = zw(KMU-1) + DZU(KMU) <= zw(KMU), HT likewise (grid.F90:1001-1015)
    zw = np.concatenate([[0.0], np.cumsum(cs.dz)])
    ht = oracle_global(o, "HT", 1)[0]
    ocean = cs.kmt > 0
    assert np.all(ht[ocean] <= zw[cs.kmt[ocean]] + 1e-9) and np.all(ht[ocean] > zw[cs.kmt[ocean] - 1])
    np.testing.assert_array_equal(ht[ocean], zw[cs.kmt[ocean] - 1] + cs.dzbc[ocean])


def _lw_case(**kw):
    base = dict(nt=3, seed=61, tadvect=c.TADVECT_LW_LIM, sfc_layer_type=c.SFC_RIGID, lpressure_avg=0, given_vmix=True)
    base.update(kw)
    cs = make_case(64, 40, 10, **base)
    for f in ("STF", "TFW"):
        cs.forcing[f][:] = 0.0
    return cs


@pytest.mark.parametrize("pbc", [0, 1])
def test_lw_lim_conserves_tracer_content_and_is_block_independent(pbc):
    """lw_lim (advection.F90:2684-3280) is in flux form: with a rigid lid and no surface fluxes the volume integral of
    every tracer is conserved -- across block boundaries too, which needs the flux velocities of the outermost ghost
    cells (comp_flux_vel_ghost, :1014-1120) to be the neighbouring block's own.  And the answer must not depend on the
    block decomposition at all: 4 x 2 blocks give the bits of a single block."""
    cs = _lw_case(partial_bottom_cells=pbc)
    lev = np.arange(1, cs.km + 1)[:, None, None]
    dzt = np.where(lev == cs.kmt[None], cs.dzbc[None], cs.dz[:, None, None]) if pbc else cs.dz[:, None, None] + 0 * lev
    vol = (lev <= cs.kmt[None]) * dzt * (cs.grid["DXT"] * cs.grid["DYT"])[None]
    res = []
    for bs in ((16, 20), None):
        # (REPRODUCIBLE sums: the solver's dot products do not depend on the block order either)
        o = load_oracle(cs, block_size=bs, reproducible=True) if bs else load_oracle(cs, reproducible=True)
        told = oracle_global(o, "TRACER", c.TIME_OLD).reshape(3, cs.km, cs.ny, cs.nx).copy()
        for ts in (c.TS_EULER, c.TS_LEAPFROG):
            assert o.step(ts) == 0
        t = oracle_global(o, "TRACER", c.TIME_CUR).reshape(3, cs.km, cs.ny, cs.nx).copy()
        res.append(t)
        if bs:
            o2 = load_oracle(cs, block_size=bs)
            assert o2.step(c.TS_EULER) == 0
            t1 = oracle_global(o2, "TRACER", c.TIME_CUR).reshape(3, cs.km, cs.ny, cs.nx)
            for n in range(3):
                before, after = np.sum(told[n] * vol), np.sum(t1[n] * vol)
                assert abs(after - before) <= 1.0e-12 * np.sum(np.abs(told[n]) * vol), (n, before, after)
    assert np.array_equal(res[0], res[1])


def test_lw_lim_limits_where_centered_advection_overshoots():
    """The one-dimensional flux limiters keep a front between two plateaus within its bounds, where the centred scheme
    (advection.F90:2139-2306) over- and undershoots: T - dt L(T) of one forward step."""
    out = {}
    for scheme in (c.TADVECT_LW_LIM, c.TADVECT_CENTERED):
        cs = _lw_case(tadvect=scheme, nt=2, flat=True)
        o = load_oracle(cs)
        o.L.oracle_set_timestep(c.TS_EULER)
        km, nt = o.km, o.nt
        U = o.view("UVEL", c.TIME_CUR, (km,))[0].copy()
        V = o.view("VVEL", c.TIME_CUR, (km,))[0].copy()
        # a strong zonal flow with a uniform mass flux U*DYU (no divergence, hence no vertical velocity) that moves the
        # front a good fraction of a cell: CFL <= 0.3 in x
        dxt, dyu = o.view("DXT", 1, ())[0], o.view("DYU", 1, ())[0]
        kmu = o.view("KMU", 1, (), np.int32)[0]
        dt = cs.cfg.dtt
        flux = 0.3 * np.min(dxt * dyu) / dt
        for k in range(km):
            U[k] = flux / dyu * (kmu > k)
            V[k] = 0.0
        o.view("UVEL", c.TIME_CUR, (km,))[0][:] = U
        o.view("VVEL", c.TIME_CUR, (km,))[0][:] = V
        kmt = o.view("KMT", 1, (), np.int32)[0]
        T = np.zeros((nt, km, o.nyb, o.nxb))
        ii = np.arange(o.nxb)[None, :] * np.ones((o.nyb, 1))
        for k in range(km):
            T[:, k] = np.where((ii > 20) & (ii < 40), 1.0, 0.0) * (kmt > k)
        T = np.ascontiguousarray(T)
        if scheme == c.TADVECT_LW_LIM:
            osig(o.L, "o_comp_flux_vel_ghost", [])()
        f_a = osig(o.L, "o_advt", [ci, vp, vp, vp, vp, vp, vp, ci])
        WTK = np.zeros((o.nyb, o.nxb))
        new = np.zeros((km, o.nyb, o.nxb))
        for k in range(1, km + 1):
            L = np.zeros((nt, o.nyb, o.nxb))
            f_a(k, op(L), op(WTK), op(T), op(T), op(U), op(V), 0)
            new[k - 1] = T[0, k - 1] - dt * L[0]
        out[scheme] = new[:, 2:-2, 2:-2]
    lw, cen = out[c.TADVECT_LW_LIM], out[c.TADVECT_CENTERED]
    assert lw.min() >= -1e-12 and lw.max() <= 1.0 + 1e-12
    assert cen.min() < -1e-3 and cen.max() > 1.0 + 1e-3        # the unlimited scheme rings
    assert np.abs(lw - cen).max() > 1e-3
    assert abs(lw[0].sum() - cen[0].sum()) <= 1e-9 * cen[0].sum()   # both move the same amount of tracer
