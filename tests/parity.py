"""Shared parity harness: builds the same seeded synthetic case for the CPU oracle (checker) and the
CUDA library (product) and offers helpers to compare them.  Test infrastructure only."""
import ctypes as C

import numpy as np

from oracle.oracle import Oracle, P, _p as op

c = P.config
syn = P.synthetic


class Case:
    pass


def make_case(nx, ny, km, nt=2, seed=1, ns=c.BNDY_CLOSED, ew=c.BNDY_CYCLIC, vgrid="stretched",
              given_vmix=False, flat=False, **cfg_kw):
    cs = Case()
    tripole = ns == c.BNDY_TRIPOLE
    kw = dict(nx_global=nx, ny_global=ny, km=km, nt=nt, ew_boundary_type=ew, ns_boundary_type=ns)
    if given_vmix:
        kw.update(vmix_itype=c.VMIX_GIVEN, vdc_kdim_halo=1, vdc_ndim=2)
    kw.update(cfg_kw)
    cs.cfg = c.make_config(**kw)
    cs.nx, cs.ny, cs.km, cs.nt = nx, ny, km, nt
    cs.grid = syn.horiz_grid(nx, ny, tripole=tripole)
    cs.dz = syn.vert_grid(vgrid, km)
    cs.kmt = syn.bathymetry(nx, ny, km, seed, south_land_rows=0 if ns == c.BNDY_CYCLIC else 2, flat=flat)
    if tripole:  # mirror-symmetric bathymetry on the top rows
        cs.kmt[-3:, :] = np.minimum(cs.kmt[-3:, :], cs.kmt[-3:, ::-1])
    cs.kmu = syn.kmu_from_kmt(cs.kmt, ew_cyclic=(ew == c.BNDY_CYCLIC), ns_type=ns)
    cs.dzbc = None
    if cs.cfg.partial_bottom_cells:
        cs.dzbc = syn.bottom_cells(cs.kmt, cs.dz, seed + 11)
        if tripole:
            cs.dzbc[-3:, :] = np.minimum(cs.dzbc[-3:, :], cs.dzbc[-3:, ::-1])
    cs.state = syn.state(nx, ny, km, nt, cs.dz, cs.kmt, cs.kmu, seed + 100)
    rng = np.random.default_rng(seed + 7)
    cs.state["UBTROP_cur"] = 2.0 * rng.standard_normal((ny, nx)) * (cs.kmu > 0)
    cs.state["VBTROP_cur"] = 2.0 * rng.standard_normal((ny, nx)) * (cs.kmu > 0)
    cs.state["UBTROP_old"] = cs.state["UBTROP_cur"] + 0.1 * rng.standard_normal((ny, nx)) * (cs.kmu > 0)
    cs.state["VBTROP_old"] = cs.state["VBTROP_cur"] + 0.1 * rng.standard_normal((ny, nx)) * (cs.kmu > 0)
    cs.forcing = dict(
        STF=1.0e-4 * rng.standard_normal((nt, ny, nx)) * (cs.kmt > 0),
        SMF=0.5 * rng.standard_normal((2, ny, nx)) * (cs.kmu > 0),
        FW=1.0e-6 * rng.standard_normal((ny, nx)) * (cs.kmt > 0),
        FW_OLD=1.0e-6 * rng.standard_normal((ny, nx)) * (cs.kmt > 0),
        TFW=1.0e-5 * rng.standard_normal((nt, ny, nx)) * (cs.kmt > 0),
    )
    if given_vmix:
        cs.vdc, cs.vvc = syn.kpp_shaped_vdc(nx, ny, km, cs.kmt, seed + 5, ndim=2, halo=True)
    else:
        cs.vdc = cs.vvc = None
    return cs


LEVELS = (("cur", c.TIME_CUR), ("old", c.TIME_OLD))
FIELDS = (("TRACER", c.LOC_CENTER, c.KIND_SCALAR), ("UVEL", c.LOC_NECORNER, c.KIND_VECTOR),
          ("VVEL", c.LOC_NECORNER, c.KIND_VECTOR), ("PSURF", c.LOC_CENTER, c.KIND_SCALAR),
          ("UBTROP", c.LOC_NECORNER, c.KIND_VECTOR), ("VBTROP", c.LOC_NECORNER, c.KIND_VECTOR))


FORCING_LOC = {"STF": (c.LOC_CENTER, c.KIND_SCALAR), "TFW": (c.LOC_CENTER, c.KIND_SCALAR),
               "FW": (c.LOC_CENTER, c.KIND_SCALAR), "FW_OLD": (c.LOC_CENTER, c.KIND_SCALAR),
               "SMF": (c.LOC_NECORNER, c.KIND_VECTOR)}


def load_oracle(cs, block_size=None, reproducible=False):
    """reproducible: the reference's REPRODUCIBLE build (global sums accumulated in real(r16), rounded once:
    mpi/POP_ReductionsMod.F90:279-283,357-363).  The library's sums are double-double with one final rounding as well,
    so in that mode every solver scalar -- and with it the whole step -- must agree BIT FOR BIT, whatever the block
    decomposition.  With the default r8 sums the two sides differ by the rounding of the dot products (1e-12 bar)."""
    from oracle import oracle as _O
    cfg = cs.cfg if block_size is None else c.copy_config(cs.cfg, block_size_x=block_size[0], block_size_y=block_size[1])
    _O.lib().oracle_set_reproducible(1 if reproducible else 0)
    o = Oracle(cfg)
    if cs.dzbc is not None:
        o.set_bottom_cells(cs.dzbc)
    o.set_grid(cs.grid, cs.kmt, cs.dz)
    for lev, t in LEVELS:
        for name, loc, kind in FIELDS:
            o.scatter(name, t, cs.state[name + "_" + lev])
            o.halo(name, t, loc, kind)
        o.state_all(t)
        o.grad_psurf(t)
    if cfg.hmix_tracer_itype == c.HMIX_GM:
        o.scatter("TLAT", 0, cs.grid["TLAT"])
        o.halo("TLAT", 0, c.LOC_CENTER, c.KIND_SCALAR)
    o.scatter("PGUESS", c.TIME_CUR, cs.state["PSURF_cur"])
    o.halo("PGUESS", c.TIME_CUR, c.LOC_CENTER, c.KIND_SCALAR)
    for name in ("STF", "SMF", "FW", "FW_OLD", "TFW"):
        o.scatter(name, 0, cs.forcing[name])
        o.halo(name, 0, *FORCING_LOC[name])
    if cs.vdc is not None:
        o.scatter("VDC", 0, cs.vdc)
        o.halo("VDC", 0, c.LOC_CENTER, c.KIND_SCALAR)
        o.scatter("VVC", 0, cs.vvc)
        o.halo("VVC", 0, c.LOC_NECORNER, c.KIND_SCALAR)
    if cfg.solver_choice == c.SOLVER_PCSI:
        assert o.solvers_prep() == 0
    return o


def load_pop(cs, cfg=None, comm_id=None):
    api = P.api
    p = api.Pop(cfg if cfg is not None else cs.cfg, comm_id)
    if cs.dzbc is not None:
        p.set_bottom_cells(cs.dzbc)
    p.set_grid(cs.grid, cs.kmt, cs.dz)
    n2 = p.nxb * p.nyb
    for lev, t in LEVELS:
        for name, loc, kind in FIELDS:
            p.scatter(name, t, cs.state[name + "_" + lev])
            p.halo_field(name, t, loc, kind)
        T, R = p.dptr("TRACER", t), p.dptr("RHO", t)
        for k in range(1, p.km + 1):
            p.state(k, k, T + 8 * n2 * (k - 1), T + 8 * n2 * (p.km + k - 1), R + 8 * n2 * (k - 1))
        p.grad(1, p.dptr("GRADPX", t), p.dptr("GRADPY", t), p.dptr("PSURF", t))
        p.halo_field("GRADPX", t, c.LOC_NECORNER, c.KIND_VECTOR)
        p.halo_field("GRADPY", t, c.LOC_NECORNER, c.KIND_VECTOR)
    if p.cfg.hmix_tracer_itype == c.HMIX_GM:
        p.scatter("TLAT", 0, cs.grid["TLAT"])
        p.halo_field("TLAT", 0, c.LOC_CENTER, c.KIND_SCALAR)
    p.scatter("PGUESS", 0, cs.state["PSURF_cur"])
    p.halo_field("PGUESS", 0, c.LOC_CENTER, c.KIND_SCALAR)
    for name in ("STF", "SMF", "FW", "FW_OLD", "TFW"):
        p.scatter(name, 0, cs.forcing[name])
        p.halo_field(name, 0, *FORCING_LOC[name])
    if cs.vdc is not None:
        p.scatter("VDC", 0, cs.vdc)
        p.halo_field("VDC", 0, c.LOC_CENTER, c.KIND_SCALAR)
        p.scatter("VVC", 0, cs.vvc)
        p.halo_field("VVC", 0, c.LOC_NECORNER, c.KIND_SCALAR)
    if p.cfg.solver_choice == c.SOLVER_PCSI:
        p.solvers_prep()
    return p


def oracle_global(o, name, tlev=c.TIME_CUR):
    """physical domain of an oracle field as (nz, ny, nx)."""
    g = o.gather(name, tlev)
    return g.reshape(-1, o.cfg.ny_global, o.cfg.nx_global)


def pop_global(p, name, tlev=c.TIME_CUR):
    return p.gather(name, tlev)


def relerr(a, b):
    """max |a-b| / max|b| (field-relative), 0 if both are identically zero."""
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = np.max(np.abs(b)) if b.size else 0.0
    return 0.0 if d == 0.0 else d / max(s, 1e-300)


def relerr_pointwise(a, b, floor=1.0e-3):
    """max over cells of |a-b| / max(|b|, floor * max|b|): the relative error of every cell against its own magnitude,
    with an absolute floor (cells smaller than `floor` of the field maximum are measured against that floor).  Reported
    next to the field-max norm: the differences this suite sees are absolute perturbations from the summation order of
    the solver's dot products, so cells near a zero crossing carry the same absolute error as the large ones."""
    if not a.size:
        return 0.0
    s = np.max(np.abs(b))
    if s == 0.0:
        return 0.0 if not np.any(a) else float("inf")
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * s)))


def osig(L, name, argtypes, restype=None):
    f = getattr(L, name)
    f.argtypes = argtypes
    f.restype = restype
    return f


__all__ = ["P", "Oracle", "make_case", "load_oracle", "load_pop", "oracle_global", "pop_global", "relerr", "relerr_pointwise",
           "c", "op", "C", "osig"]
