"""ctypes mirror of `pop_config` / `pop_block` (include/pop_b200.h) and option ids.

The struct replaces the reference's namelist groups (advect_nml advection.F90:234, hmix_nml
horizontal_mix.F90:140, vertical_mix_nml vertical_mix.F90:224, pressure_grad_nml
pressure_grad.F90:110, domain_nml domain.F90:145, solvers POP_SolversMod.F90:578-676).
Defaults follow the module defaults / bld/namelist_files/namelist_defaults_pop.xml.
"""
import ctypes as C

POP_MAX_NT = 64
LOC_CENTER, LOC_NECORNER, LOC_NFACE, LOC_EFACE = 1, 2, 3, 4
KIND_SCALAR, KIND_VECTOR, KIND_ANGLE = 1, 2, 3
BNDY_CLOSED, BNDY_CYCLIC, BNDY_TRIPOLE = 0, 1, 2
TADVECT_CENTERED, TADVECT_UPWIND3, TADVECT_LW_LIM = 1, 2, 3
HMIX_DEL2, HMIX_DEL4, HMIX_GM = 1, 2, 3
VMIX_CONST, VMIX_RICH, VMIX_GIVEN = 1, 2, 3
SFC_VARTHICK, SFC_RIGID, SFC_OLDFREE = 1, 2, 3
STATE_MWJF, STATE_LINEAR = 2, 4
STATE_RANGE_IGNORE, STATE_RANGE_ENFORCE = 1, 3
SOLVER_PCG, SOLVER_CHRONGEAR, SOLVER_PCSI = 1, 2, 3
PRECOND_DIAGONAL, PRECOND_EVP = 0, 1
TS_LEAPFROG, TS_EULER, TS_AVG, TS_ROBERT = 1, 2, 3, 4
TIME_OLD, TIME_CUR, TIME_NEW = 0, 1, 2


class PopConfig(C.Structure):
    _fields_ = [
        ("nx_global", C.c_int), ("ny_global", C.c_int), ("km", C.c_int), ("nt", C.c_int),
        ("ew_boundary_type", C.c_int), ("ns_boundary_type", C.c_int),
        ("block_size_x", C.c_int), ("block_size_y", C.c_int),
        ("tadvect_itype", C.c_int * POP_MAX_NT),
        ("hmix_tracer_itype", C.c_int), ("hmix_momentum_itype", C.c_int),
        ("ah", C.c_double), ("am", C.c_double),
        ("lvariable_hmixt", C.c_int), ("lvariable_hmixu", C.c_int),
        ("lauto_hmixt", C.c_int), ("lauto_hmixu", C.c_int),
        ("ah_gm", C.c_double), ("ah_bolus", C.c_double), ("ah_bkg_srfbl", C.c_double),
        ("slm_r", C.c_double), ("slm_b", C.c_double),
        ("vmix_itype", C.c_int), ("implicit_vertical_mix", C.c_int),
        ("vdc_kdim_halo", C.c_int), ("vdc_ndim", C.c_int),
        ("aidif", C.c_double), ("bottom_drag", C.c_double),
        ("const_vdc", C.c_double), ("const_vvc", C.c_double),
        ("bckgrnd_vdc", C.c_double), ("bckgrnd_vvc", C.c_double), ("rich_mix", C.c_double),
        ("convection_diff", C.c_int),
        ("convect_diff", C.c_double), ("convect_visc", C.c_double),
        ("sfc_layer_type", C.c_int), ("partial_bottom_cells", C.c_int),
        ("lpressure_avg", C.c_int), ("lbouss_correct", C.c_int), ("impcor", C.c_int),
        ("state_itype", C.c_int), ("state_range_iopt", C.c_int),
        ("solver_choice", C.c_int), ("max_iterations", C.c_int),
        ("convergence_check_freq", C.c_int), ("convergence_check_start", C.c_int),
        ("max_lanczos_step", C.c_int),
        ("convergence_criterion", C.c_double), ("lanczos_convergence_criterion", C.c_double),
        ("dtt", C.c_double),
        ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int),
        ("robert_alpha", C.c_double), ("robert_nu", C.c_double),
        ("nconvad", C.c_int),
        ("preconditioner_choice", C.c_int),
    ]


class PopBlock(C.Structure):
    _fields_ = [
        ("block_id", C.c_int), ("local_id", C.c_int),
        ("ib", C.c_int), ("ie", C.c_int), ("jb", C.c_int), ("je", C.c_int),
        ("iblock", C.c_int), ("jblock", C.c_int),
        ("i_glob", C.POINTER(C.c_int)), ("j_glob", C.POINTER(C.c_int)),
    ]


def make_config(**kw):
    """pop_config with reference defaults, overridden by keyword arguments.

    `tadvect` may be an int (all tracers) or a per-tracer list.
    """
    c = PopConfig()
    d = dict(
        nx_global=0, ny_global=0, km=0, nt=2,
        ew_boundary_type=BNDY_CYCLIC, ns_boundary_type=BNDY_CLOSED,
        block_size_x=0, block_size_y=0,
        hmix_tracer_itype=HMIX_DEL2, hmix_momentum_itype=HMIX_DEL2,
        ah=1.0e7, am=1.0e7, lvariable_hmixt=0, lvariable_hmixu=0, lauto_hmixt=0, lauto_hmixu=0,
        ah_gm=0.8e7, ah_bolus=0.8e7, ah_bkg_srfbl=0.8e7, slm_r=0.3, slm_b=0.3,
        vmix_itype=VMIX_CONST, implicit_vertical_mix=1, vdc_kdim_halo=0, vdc_ndim=1,
        aidif=1.0, bottom_drag=1.0e-3, const_vdc=0.25, const_vvc=0.25,
        bckgrnd_vdc=0.1, bckgrnd_vvc=1.0, rich_mix=50.0,
        convection_diff=1, convect_diff=1000.0, convect_visc=1000.0,
        sfc_layer_type=SFC_VARTHICK, partial_bottom_cells=0,
        lpressure_avg=1, lbouss_correct=1, impcor=1,
        state_itype=STATE_MWJF, state_range_iopt=STATE_RANGE_ENFORCE,
        solver_choice=SOLVER_CHRONGEAR, max_iterations=1000, convergence_check_freq=10,
        convergence_check_start=60, max_lanczos_step=20,
        convergence_criterion=1.0e-13, lanczos_convergence_criterion=0.1,
        dtt=3600.0, rank=0, nranks=1, device=0, robert_alpha=0.53, robert_nu=0.20, nconvad=0,
        preconditioner_choice=PRECOND_DIAGONAL,
    )
    tadv = kw.pop("tadvect", TADVECT_CENTERED)
    d.update(kw)
    for k, v in d.items():
        setattr(c, k, v)
    if c.block_size_x == 0:
        c.block_size_x = c.nx_global
    if c.block_size_y == 0:
        c.block_size_y = c.ny_global
    nt = c.nt
    if isinstance(tadv, int):
        tadv = [tadv] * nt
    for n in range(POP_MAX_NT):
        c.tadvect_itype[n] = tadv[n] if n < len(tadv) else TADVECT_CENTERED
    return c


def copy_config(c, **kw):
    c2 = PopConfig()
    C.memmove(C.byref(c2), C.byref(c), C.sizeof(PopConfig))
    for k, v in kw.items():
        setattr(c2, k, v)
    return c2
