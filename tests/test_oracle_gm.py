"""CPU: the oracle's Gent-McWilliams restatement (oracle/o_gm.c; hmix_gm.F90:1102-2219,
hmix_gm_submeso_share.F90:149-434). The reference tree holds no golden data for GM, so the
restatement is pinned by properties of the scheme: flux form (the volume integral of the tendency vanishes),
reduction to Laplacian diffusion for level isopycnals, equality of the 'cancellation' short cut with the general
skew-flux form when ah = ah_bolus, taper bounds, and the positive-definite vertical coefficient."""
import ctypes as C

import numpy as np
import pytest

from parity import *  # noqa: F401,F403

ci, vp = C.c_int, C.c_void_p


def _gm_case(**kw):
    base = dict(nt=3, seed=31, hmix_tracer_itype=c.HMIX_GM, ns=c.BNDY_CYCLIC)
    base.update(kw)
    return make_case(40, 32, 8, **base)


def _tendency(o, cs, tmix=None):
    """GTK of all levels: (nblocks, km, nt, nyb, nxb); also returns VDC before/after."""
    T = o.view("TRACER", c.TIME_CUR, (o.nt, o.km))
    f = osig(o.L, "o_hdifft", [ci, vp, vp, vp, vp, ci])
    out = np.zeros((o.nblocks, o.km, o.nt, o.nyb, o.nxb))
    for b in range(o.nblocks):
        for k in range(1, o.km + 1):
            H = np.zeros((o.nt, o.nyb, o.nxb))
            f(k, op(H), op(T[b]), None, None, b)
            out[b, k - 1] = H
    return out


def _physical(o, cs, blockfield):
    """(nblocks, ..., nyb, nxb) with one block -> (..., ny, nx)"""
    assert o.nblocks == 1
    return blockfield[0][..., 2:-2, 2:-2]


def test_gm_tendency_is_a_flux_divergence():
    cs = _gm_case()
    o = load_oracle(cs)
    G = _physical(o, cs, _tendency(o, cs))                         # (km, nt, ny, nx)
    tarea = oracle_global(o, "TAREA", 0)[0]
    dz = np.asarray(cs.dz)
    vol = tarea[None] * dz[:, None, None] * (np.arange(1, o.km + 1)[:, None, None] <= cs.kmt[None])
    for n in range(o.nt):
        total = float(np.sum(vol * G[:, n]))
        scale = float(np.sum(np.abs(vol * G[:, n])))
        assert scale > 0.0 and abs(total) <= 1e-12 * scale, (n, total, scale)
    assert np.all(G[:, :, cs.kmt == 0] == 0.0)                       # land columns untouched
    below = np.arange(1, o.km + 1)[:, None, None] > cs.kmt[None]
    assert np.all(G.transpose(1, 0, 2, 3)[:, below] == 0.0)          # nothing below the bottom


def test_gm_with_level_isopycnals_is_laplacian_diffusion():
    """T,S horizontally uniform -> all slopes vanish, tapers are 1, and a passive tracer feels
    div(ah grad) only: compare with hdifft_del2 at ah on cells whose neighbourhood is deep ocean and whose
    quarter cells are all interior (3 <= k < KMT)."""
    cs = _gm_case(flat=True)
    prof_t = 2.0 + 20.0 * np.exp(-np.cumsum(cs.dz) / 1.0e5)
    for lev in ("cur", "old"):
        T = cs.state["TRACER_" + lev]
        T[0] = prof_t[:, None, None]
        T[1] = 0.0347
    ogm = load_oracle(cs)
    G = _physical(ogm, cs, _tendency(ogm, cs))[:, 2].copy()
    assert np.all(ogm.view("SLX", 1, (ogm.km, 2, 2)) == 0.0) and np.all(ogm.view("SLY", 1, (ogm.km, 2, 2)) == 0.0)
    cs2 = _gm_case(flat=True, hmix_tracer_itype=c.HMIX_DEL2, ah=cs.cfg.ah_gm)
    cs2.state = cs.state
    od = load_oracle(cs2)
    D = _physical(od, cs2, _tendency(od, cs2))[:, 2]
    k = 3                                                           # 1-based level 4: all quarter cells interior
    deep = cs.kmt > k + 1
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            deep &= np.roll(np.roll(cs.kmt > k + 1, dj, 0), di, 1)
    assert deep.sum() > 100
    ref, got = D[k][deep], G[k][deep]
    assert np.max(np.abs(ref)) > 0.0
    assert np.max(np.abs(got - ref)) <= 1e-11 * np.max(np.abs(ref))


def test_general_skew_flux_branch_equals_the_cancellation_short_cut():
    """ah == ah_bolus and slm_r == slm_b: the reference drops the terms that cancel (hmix_gm.F90:970-983);
    the general branch with the same coefficients must agree to rounding."""
    from oracle.oracle import lib as lib_oracle
    f = osig(lib_oracle(), "o_gm_force_general", [ci], None)
    res = []
    try:
        for force in (0, 1):
            f(force)
            cs = _gm_case()
            o = load_oracle(cs)
            res.append(_physical(o, cs, _tendency(o, cs)).copy())
            assert osig(o.L, "o_gm_cancellation", [], ci)() == 1 - force
    finally:
        f(0)
    a, b = res
    assert np.max(np.abs(a)) > 0.0
    assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(a))


def test_tapers_and_vertical_coefficient_bounds():
    cs = _gm_case(ah_bolus=0.5e7, slm_b=0.2)                         # general branch, differential tapering
    o = load_oracle(cs)
    v0 = o.view("VDC", 0, o.inner_shape("VDC")).copy()
    _tendency(o, cs)
    assert osig(o.L, "o_gm_cancellation", [], ci)() == 0
    ki, kt, hd = (o.view(n, 1, (o.km, 2)) for n in ("KAPPA_ISOP", "KAPPA_THIC", "HOR_DIFF"))
    assert ki.min() >= 0.0 and ki.max() <= cs.cfg.ah_gm
    assert kt.min() >= 0.0 and kt.max() <= cs.cfg.ah_bolus
    assert hd.min() >= 0.0 and hd.max() <= cs.cfg.ah_bkg_srfbl
    assert np.any((ki > 0.0) & (ki < cs.cfg.ah_gm))                  # some cells sit on the taper ramp
    assert np.all(ki[:, 0, 0] == 0.0) and np.all(kt[:, 0, 0] == 0.0)   # top quarter cells (B.C. at the surface)
    kmt = o.view("KMT", 1, (), np.int32)
    for b in range(o.nblocks):
        for k in range(1, o.km + 1):
            assert np.all(ki[b, k - 1, 1][kmt[b] == k] == 0.0)      # bottom half of the bottom cell
    dv = o.view("VDC", 0, o.inner_shape("VDC")) - v0
    assert dv.min() >= 0.0 and dv.max() > 0.0                       # VDC_GM = K * slope^2 >= 0
    assert np.all(dv[:, :, -1] == 0.0)                              # no interface below level km
