"""GPU parity of the reference-signature slab operators (called through the C ABI with HOST arrays,
as a Fortran caller would) against the CPU oracle on the same seeded inputs.  Integer masks and
every fp64 field must agree BIT-EXACTLY: both sides evaluate the reference's expressions in the
reference's order without FMA contraction."""
import numpy as np
import pytest

from parity import *  # noqa: F401,F403

pytestmark = pytest.mark.gpu
vp, ci = C.c_void_p, C.c_int


def both(cs):
    o = load_oracle(cs)
    p = load_pop(cs)
    return o, p


def phys(a):
    return a[..., 2:-2, 2:-2]


# The oracle and the library are singletons (like the Fortran modules they mirror): one live case at a
# time, re-created when a test asks for the other configuration.
_CUR = {"key": None, "val": None}


def get_case(key):
    if _CUR["key"] == key:
        return _CUR["val"]
    if _CUR["val"] is not None:
        _CUR["val"][2].finalize()
    if key == "del2":
        cs = make_case(40, 28, 7, nt=3, seed=11, given_vmix=True,
                       tadvect=[c.TADVECT_CENTERED, c.TADVECT_UPWIND3, c.TADVECT_CENTERED])
    elif key == "gm":          # gx1v7 flavour: GM with the terms that cancel dropped (ah == ah_bolus)
        cs = make_case(40, 28, 7, nt=3, seed=13, hmix_tracer_itype=c.HMIX_GM, given_vmix=True)
    elif key == "pbc":         # partial bottom cells: every operator takes its DZT / DZU branch (del4, given coefficients)
        cs = make_case(36, 24, 7, nt=3, seed=15, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                       hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21,
                       am=-27.0e21, given_vmix=True, partial_bottom_cells=1)
    elif key == "pbc_del2":
        cs = make_case(40, 28, 7, nt=2, seed=16, given_vmix=True, partial_bottom_cells=1)
    elif key == "lw_lim":      # limited advection next to the centred scheme, tripole
        cs = make_case(36, 24, 7, nt=3, seed=17, ns=c.BNDY_TRIPOLE, given_vmix=True,
                       tadvect=[c.TADVECT_LW_LIM, c.TADVECT_CENTERED, c.TADVECT_LW_LIM])
    elif key == "pbc_lw_lim":
        cs = make_case(40, 28, 7, nt=2, seed=18, given_vmix=True, partial_bottom_cells=1, tadvect=c.TADVECT_LW_LIM)
    elif key == "gm_general":  # general skew-flux form: different thickness diffusivity and slope limits
        cs = make_case(36, 24, 6, nt=3, seed=14, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_GM,
                       ah_gm=0.6e7, ah_bolus=0.4e7, ah_bkg_srfbl=0.5e7, slm_r=0.3, slm_b=0.2)
    else:
        cs = make_case(36, 24, 6, nt=2, seed=12, ns=c.BNDY_TRIPOLE, hmix_tracer_itype=c.HMIX_DEL4,
                       hmix_momentum_itype=c.HMIX_DEL4, lvariable_hmixt=1, lvariable_hmixu=1, ah=-3.0e21,
                       am=-27.0e21)
    o, p = both(cs)
    _CUR["key"], _CUR["val"] = key, (cs, o, p)
    return _CUR["val"]


@pytest.fixture
def case_del2():
    return get_case("del2")


@pytest.fixture
def case_del4():
    return get_case("del4")


def fields(o, t):
    T = o.view("TRACER", t, (o.nt, o.km))[0].copy()
    U = o.view("UVEL", t, (o.km,))[0].copy()
    V = o.view("VVEL", t, (o.km,))[0].copy()
    R = o.view("RHO", t, (o.km,))[0].copy()
    return T, U, V, R


@pytest.mark.parametrize("which", ["del4", "pbc"])
def test_grid_masks_and_metrics_bit_exact(which):
    cs, o, p = get_case(which)
    if which == "pbc":   # grid.F90:917-1022: DZT, DZU (levels 0..km+1), and the depths built from them
        for n in ("DZT", "DZU"):
            a = o.view(n, 1, (o.km + 2,))[0]
            assert np.array_equal(a, p.get_padded(n, 1).reshape(a.shape)), n
        for n in ("HU", "HUR", "HT"):
            assert np.array_equal(o.view(n, 1)[0], p.get_padded(n, 1)[0]), n
    for n in ("KMT", "KMU", "CHECKER", "CONSTNT"):
        assert np.array_equal(o.view(n, 1, (), np.int32)[0], p.get_padded(n, 1, np.int32)[0]), n
    for n in ("DXU", "DYU", "TAREA_R", "UAREA_R", "HUR", "FCOR", "AU0", "AUNE", "KXU", "KYU", "DTN", "DTS", "DTE",
              "DTW", "AHF", "DUC", "DUN", "DUS", "DUE", "DUW", "DMC", "DMN", "DME", "DUM", "AMF", "btropWgtNorth",
              "btropWgtEast", "btropWgtNE", "centerWgtClinicIndep", "mMaskTropic"):
        assert np.array_equal(o.view(n, 1)[0], p.get_padded(n, 1)[0]), n
    for n in ("TRACER", "UVEL", "VVEL", "RHO", "GRADPX", "GRADPY"):
        for t in (c.TIME_OLD, c.TIME_CUR):
            a = o.view(n, t, o.inner_shape(n)).reshape(-1, o.nyb, o.nxb)
            assert np.array_equal(a, p.get_padded(n, t)), (n, t)


def test_state_mwjf_all_outputs(case_del2):
    cs, o, p = case_del2
    T, U, V, R = fields(o, c.TIME_CUR)
    f = osig(o.L, "o_state", [ci, ci, vp, vp, ci, vp, vp, vp, vp])
    for k in (1, 3, o.km):
        outs_o = [np.zeros((o.nyb, o.nxb)) for _ in range(4)]
        outs_p = [np.zeros((o.nyb, o.nxb)) for _ in range(4)]
        kk = min(k + 1, o.km)
        f(k, kk, op(T[0, k - 1]), op(T[1, k - 1]), 0, *[op(x) for x in outs_o])
        p.state(k, kk, T[0, k - 1], T[1, k - 1], *outs_p)
        for a, b in zip(outs_o, outs_p):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("which", ["del2", "del4", "pbc", "pbc_del2"])
def test_hdifft_advt_vdifft(which):
    cs, o, p = get_case(which)
    Tc, Uc, Vc, _ = fields(o, c.TIME_CUR)
    To = fields(o, c.TIME_OLD)[0]
    nt, km = o.nt, o.km
    f_h = osig(o.L, "o_hdifft", [ci, vp, vp, vp, vp, ci])
    f_a = osig(o.L, "o_advt", [ci, vp, vp, vp, vp, vp, vp, ci])
    f_v = osig(o.L, "o_vdifft", [ci, vp, vp, vp, ci])
    STF = o.view("STF", 0, (nt,))[0].copy()
    DH = np.ascontiguousarray(0.01 * np.sin(np.arange(o.nyb * o.nxb)).reshape(o.nyb, o.nxb))
    wo, wp = DH.copy(), DH.copy()
    for k in range(1, km + 1):
        ho, hp = np.zeros((nt, o.nyb, o.nxb)), np.zeros((nt, o.nyb, o.nxb))
        f_h(k, op(ho), op(To), None, None, 0)
        p.hdifft(k, hp, To)
        assert np.array_equal(phys(ho), phys(hp)), ("hdifft", k)
        lo, lp = np.zeros_like(ho), np.zeros_like(ho)
        f_a(k, op(lo), op(wo), op(To), op(Tc), op(Uc), op(Vc), 0)
        p.advt(k, lp, wp, To, Tc, Uc, Vc)
        assert np.array_equal(phys(lo), phys(lp)), ("advt", k)
        assert np.array_equal(phys(wo), phys(wp)), ("WTK", k)
        vo, vq = np.zeros_like(ho), np.zeros_like(ho)
        f_v(k, op(vo), op(To), op(STF), 0)
        p.vdifft(k, vq, To, STF)
        assert np.array_equal(phys(vo), phys(vq)), ("vdifft", k)
    assert np.abs(phys(lo)).max() >= 0.0 and np.abs(phys(ho)).max() > 0.0


@pytest.mark.parametrize("which", ["lw_lim", "pbc_lw_lim"])
def test_advt_lw_lim_level_by_level(which):
    """advt with tadvect_itype = lw_lim through the reference's own calling sequence: comp_flux_vel_ghost(DH) once
    (baroclinic.F90:667), then advt(k) for k = 1..km with WTK carried by the caller (baroclinic.F90:1989-2010).  TMIX
    (the advected field of lw_lim) differs from TRCR (the advected field of the centred scheme)."""
    cs, o, p = get_case(which)
    Tc, Uc, Vc, _ = fields(o, c.TIME_CUR)
    To = fields(o, c.TIME_OLD)[0]
    nt, km = o.nt, o.km
    f_a = osig(o.L, "o_advt", [ci, vp, vp, vp, vp, vp, vp, ci])
    kmt = o.view("KMT", 1, (), np.int32)[0]
    DH = np.ascontiguousarray(0.01 * np.sin(np.arange(o.nyb * o.nxb)).reshape(o.nyb, o.nxb) * (kmt > 0))
    o.view("DH", 0, ())[0][:] = DH
    osig(o.L, "o_comp_flux_vel_ghost", [])()
    p.comp_flux_vel_ghost(DH)
    wo, wp = DH.copy(), DH.copy()
    seen = 0.0
    for k in range(1, km + 1):
        lo, lp = np.zeros((nt, o.nyb, o.nxb)), np.zeros((nt, o.nyb, o.nxb))
        f_a(k, op(lo), op(wo), op(To), op(Tc), op(Uc), op(Vc), 0)
        p.advt(k, lp, wp, To, Tc, Uc, Vc)
        assert np.array_equal(phys(lo), phys(lp)), ("advt", k, np.abs(phys(lo) - phys(lp)).max())
        assert np.array_equal(phys(wo), phys(wp)), ("WTK", k)
        seen = max(seen, np.abs(phys(lo)[0]).max())
    assert seen > 0.0


@pytest.mark.parametrize("which", ["gm", "gm_general"])
def test_hdifft_gm_slabs_and_vdc_side_effect(which):
    """hdifft with hmix_tracer_itype = GM (hmix_gm.F90:1102-2219 after tracer_diffs_and_isopyc_slopes), called
    k = 1,2,3,... like the reference: every level's tendency and the VDC the calls leave behind are bit-exact."""
    cs, o, p = get_case(which)
    To = fields(o, c.TIME_OLD)[0]
    nt, km = o.nt, o.km
    f_h = osig(o.L, "o_hdifft", [ci, vp, vp, vp, vp, ci])
    shape = o.inner_shape("VDC")
    v0 = o.view("VDC", 0, shape)[0].copy()
    assert np.array_equal(v0, p.get_padded("VDC", 0).reshape(v0.shape))
    big = 0.0
    for k in range(1, km + 1):
        ho, hp = np.zeros((nt, o.nyb, o.nxb)), np.zeros((nt, o.nyb, o.nxb))
        f_h(k, op(ho), op(To), None, None, 0)
        p.hdifft(k, hp, To)
        assert np.array_equal(phys(ho), phys(hp)), ("hdifft_gm", k)
        big = max(big, np.abs(phys(ho)).max())
    assert big > 0.0
    v1o, v1p = o.view("VDC", 0, shape)[0], p.get_padded("VDC", 0).reshape(v0.shape)
    assert np.array_equal(phys(v1o), phys(v1p))
    assert (phys(v1o) - phys(v0)).max() > 0.0                   # VDC_GM was added
    o.view("VDC", 0, shape)[0][...] = v0                        # leave the case as it was found
    p.set_padded("VDC", 0, v0)


@pytest.mark.parametrize("which", ["del2", "del4", "pbc", "pbc_del2"])
def test_advu_hdiffu_gradp_vdiffu(which):
    cs, o, p = get_case(which)
    _, Uc, Vc, Rc = fields(o, c.TIME_CUR)
    _, Uo, Vo, Ro = fields(o, c.TIME_OLD)
    Rn = np.ascontiguousarray(0.5 * (Rc + Ro))
    km = o.km
    SMF = o.view("SMF", 0, (2,))[0].copy()
    f_a = osig(o.L, "o_advu", [ci, vp, vp, vp, vp, vp, ci])
    f_h = osig(o.L, "o_hdiffu", [ci, vp, vp, vp, vp, ci])
    f_g = osig(o.L, "o_gradp", [ci, vp, vp, vp, vp, vp, ci])
    f_v = osig(o.L, "o_vdiffu", [ci, vp, vp, vp, vp, vp, ci])
    o.set_timestep(c.TS_LEAPFROG)
    p.set_timestep(c.TS_LEAPFROG)
    DHU = np.ascontiguousarray(0.01 * np.cos(np.arange(o.nyb * o.nxb)).reshape(o.nyb, o.nxb))
    wo, wp = DHU.copy(), DHU.copy()
    z = lambda: np.zeros((o.nyb, o.nxb))
    for k in range(1, km + 1):
        a1, a2, b1, b2 = z(), z(), z(), z()
        f_a(k, op(a1), op(a2), op(wo), op(Uc), op(Vc), 0)
        p.advu(k, b1, b2, wp, Uc, Vc)
        assert np.array_equal(phys(a1), phys(b1)) and np.array_equal(phys(a2), phys(b2)), ("advu", k)
        assert np.array_equal(phys(wo), phys(wp)), ("WUK", k)
        a1, a2, b1, b2 = z(), z(), z(), z()
        f_h(k, op(a1), op(a2), op(Uo[k - 1]), op(Vo[k - 1]), 0)
        p.hdiffu(k, b1, b2, np.ascontiguousarray(Uo[k - 1]), np.ascontiguousarray(Vo[k - 1]))
        assert np.array_equal(phys(a1), phys(b1)) and np.array_equal(phys(a2), phys(b2)), ("hdiffu", k)
        a1, a2, b1, b2 = z(), z(), z(), z()
        f_g(k, op(a1), op(a2), op(Ro[k - 1]), op(Rc[k - 1]), op(Rn[k - 1]), 0)
        p.gradp(k, b1, b2, np.ascontiguousarray(Ro[k - 1]), np.ascontiguousarray(Rc[k - 1]), np.ascontiguousarray(Rn[k - 1]))
        assert np.array_equal(phys(a1), phys(b1)) and np.array_equal(phys(a2), phys(b2)), ("gradp", k)
        a1, a2, b1, b2 = z(), z(), z(), z()
        f_v(k, op(a1), op(a2), op(Uo), op(Vo), op(SMF), 0)
        p.vdiffu(k, b1, b2, Uo, Vo, SMF)
        assert np.array_equal(phys(a1), phys(b1)) and np.array_equal(phys(a2), phys(b2)), ("vdiffu", k)


@pytest.mark.parametrize("which", ["del2", "pbc"])
def test_impvmix_tridiagonal_solves(which):
    cs, o, p = get_case(which)
    Tc = fields(o, c.TIME_CUR)[0]
    To, Uo, Vo, _ = fields(o, c.TIME_OLD)
    Ps = o.view("PSURF", c.TIME_CUR)[0].copy()
    nt = o.nt
    o.set_timestep(c.TS_LEAPFROG)
    p.set_timestep(c.TS_LEAPFROG)
    rhs = np.ascontiguousarray(1.0e-3 * (Tc - To))
    f_t = osig(o.L, "o_impvmixt", [vp, vp, vp, ci, ci, ci])
    f_c = osig(o.L, "o_impvmixt_correct", [vp, vp, vp, ci, ci, ci])
    f_u = osig(o.L, "o_impvmixu", [vp, vp, ci])
    a, b = rhs.copy(), rhs.copy()
    f_t(op(a), op(To), op(Ps), 1, nt, 0)
    p.impvmixt(b, To, Ps, 1, nt)
    assert np.array_equal(phys(a), phys(b))
    assert np.abs(phys(a) - phys(To)).max() > 0
    # partial tracer range and the empty range (nfirst > nlast is a no-op, vertical_mix.F90:1232)
    a, b = rhs.copy(), rhs.copy()
    f_t(op(a), op(To), op(Ps), 3, nt, 0)
    p.impvmixt(b, To, Ps, 3, nt)
    assert np.array_equal(phys(a), phys(b))
    b2 = rhs.copy()
    p.impvmixt(b2, To, Ps, 3, 2)
    assert np.array_equal(b2, rhs)
    R1 = np.ascontiguousarray(1.0e-2 * Tc[:, 0])
    a, b = Tc.copy(), Tc.copy()
    f_c(op(a), op(Ps), op(R1), 1, 2, 0)
    p.impvmixt_correct(b, Ps, R1, 1, 2)
    assert np.array_equal(phys(a), phys(b))
    ua, va, ub, vb = Uo.copy(), Vo.copy(), Uo.copy(), Vo.copy()
    f_u(op(ua), op(va), 0)
    p.impvmixu(ub, vb)
    assert np.array_equal(phys(ua), phys(ub)) and np.array_equal(phys(va), phys(vb))


def test_grad_div_btrop_operator(case_del4):
    cs, o, p = case_del4
    rng = np.random.default_rng(5)
    F = np.ascontiguousarray(rng.standard_normal((o.nyb, o.nxb)))
    Gm = np.ascontiguousarray(rng.standard_normal((o.nyb, o.nxb)))
    z = lambda: np.zeros((o.nyb, o.nxb))
    f_g = osig(o.L, "o_grad", [ci, vp, vp, vp, ci])
    f_d = osig(o.L, "o_div", [ci, vp, vp, vp, ci])
    f_b = osig(o.L, "o_btrop_operator", [vp, vp, ci])
    a1, a2, b1, b2 = z(), z(), z(), z()
    f_g(2, op(a1), op(a2), op(F), 0)
    p.grad(2, b1, b2, F)
    assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
    a1, b1 = z(), z()
    f_d(1, op(a1), op(F), op(Gm), 0)
    p.div(1, b1, F, Gm)
    assert np.array_equal(a1, b1)
    osig(o.L, "o_solvers_prep", [], ci)()
    a1, b1 = z(), z()
    f_b(op(a1), op(F), 0)
    p.btrop_operator(b1, F)
    assert np.array_equal(a1, b1)
    # operator symmetry <x, A y> = <A x, y> over the interior with zero boundary data
    X, Y = z(), z()
    X[2:-2, 2:-2] = rng.standard_normal((o.nyb - 4, o.nxb - 4)) * (o.view("KMT", 1, (), np.int32)[0][2:-2, 2:-2] > 0)
    Y[2:-2, 2:-2] = rng.standard_normal((o.nyb - 4, o.nxb - 4)) * (o.view("KMT", 1, (), np.int32)[0][2:-2, 2:-2] > 0)
    p.halo_update(X, c.LOC_CENTER, c.KIND_SCALAR)
    p.halo_update(Y, c.LOC_CENTER, c.KIND_SCALAR)
    AX, AY = z(), z()
    p.btrop_operator(AX, X)
    p.btrop_operator(AY, Y)
    s1, s2 = np.sum(phys(X) * phys(AY)), np.sum(phys(AX) * phys(Y))
    assert abs(s1 - s2) <= 1e-10 * max(abs(s1), abs(s2), 1e-300)


def test_global_sums(case_del4):
    cs, o, p = case_del4
    i = np.arange(o.nxb)[None, :]
    j = np.arange(o.nyb)[:, None]
    A = np.ascontiguousarray(((i + j) * 10).astype(np.float64))  # integer-valued: the sum is exact
    M = o.view("mMaskTropic", 1)[0].copy()
    for loc in (c.LOC_CENTER, c.LOC_NECORNER, c.LOC_NFACE):
        assert p.global_sum(A, loc) == o.global_sum(A, loc)
        assert p.global_sum(A, loc, M) == o.global_sum(A, loc, M)
    two = np.ascontiguousarray(np.stack([A, 2 * A]))
    s = p.global_sum(two, c.LOC_CENTER, M)
    assert s[0] == o.global_sum(A, c.LOC_CENTER, M) and s[1] == 2 * s[0]
    # a wide-dynamic-range sum: the double-double result is the correctly rounded exact sum
    rng = np.random.default_rng(1)
    B = np.ascontiguousarray(rng.standard_normal((o.nyb, o.nxb)) * 10.0 ** rng.integers(-8, 8, (o.nyb, o.nxb)))
    import math
    exact = math.fsum(phys(B).ravel().tolist())
    assert p.global_sum(B, c.LOC_CENTER) == exact


def test_error_convention(case_del2):
    cs, o, p = case_del2
    api = P.api
    z = np.zeros((o.nyb, o.nxb))
    with pytest.raises(api.PopError):
        p.state(1, o.km + 1, z, z, z)       # kk out of range -> POP_Fail with a message
    with pytest.raises(api.PopError):
        p.hdifft(0, np.zeros((o.nt, o.nyb, o.nxb)), np.zeros((o.nt, o.km, o.nyb, o.nxb)))
    with pytest.raises(api.PopError):
        p.solvers_diagonal(z, 2)            # only one block per rank
    assert p.L.pop_last_error()
