/*
 * pop_b200.h -- C ABI of the B200-native POP2 baroclinic/barotropic time-step hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / ints / doubles and returns
 * an int error code following the reference's convention (POP_Success = 0, POP_Fail = -1,
 * reference source/POP_ErrorMod.F90:45-47).  A message for the last failure is available
 * from pop_last_error().  The Fortran side binds these with ISO_C_BINDING interface blocks
 * (pop2-cesm_b200/fortran/pop_b200_bind.F90).
 *
 * Conventions (same as the reference):
 *   - arrays are Fortran column-major, i fastest: (nx_block, ny_block [,km [,nt]]), one block per
 *     rank including the 2-cell ghost ring (source/blocks.F90:51-56)
 *   - level index k, tracer indices nfirst/nlast, block index are 1-based
 *   - a slab routine "must be called successively with k = 1,2,3,..." for a block; k==1 resets the
 *     carried state (source/advection.F90:1160, pressure_grad.F90:206, vertical_mix.F90:703)
 *   - pointer arguments may be DEVICE pointers (zero-copy) or HOST pointers; host pointers are
 *     staged through HBM inside the call (literal drop-in from Fortran host arrays).
 *
 * There is no CPU fallback: every compute entry fails (-1) if no CUDA device is usable.
 */
#ifndef POP_B200_H
#define POP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define POP_SUCCESS 0
#define POP_FAIL (-1)

/* grid location / field kind of POP_HaloUpdate and POP_GlobalSum
 * (source/POP_GridHorzMod.F90:52-57, source/POP_FieldMod.F90:106-108) */
enum { POP_LOC_CENTER = 1, POP_LOC_NECORNER = 2, POP_LOC_NFACE = 3, POP_LOC_EFACE = 4 };
enum { POP_KIND_SCALAR = 1, POP_KIND_VECTOR = 2, POP_KIND_ANGLE = 3 };
/* boundary types of create_blocks (source/blocks.F90:181-257) */
enum { POP_BNDY_CLOSED = 0, POP_BNDY_CYCLIC = 1, POP_BNDY_TRIPOLE = 2 };
/* option ids */
enum { POP_TADVECT_CENTERED = 1, POP_TADVECT_UPWIND3 = 2, POP_TADVECT_LW_LIM = 3 };  /* advection.F90:73-76 */
enum { POP_HMIX_DEL2 = 1, POP_HMIX_DEL4 = 2, POP_HMIX_GM = 3 };         /* horizontal_mix.F90 */
enum { POP_VMIX_CONST = 1, POP_VMIX_RICH = 2, POP_VMIX_GIVEN = 3 };   /* vertical_mix.F90:393-419;
                                   GIVEN = KPP-shaped VDC(0:km+1,2)/VVC(km) supplied by the caller */
enum { POP_SFC_VARTHICK = 1, POP_SFC_RIGID = 2, POP_SFC_OLDFREE = 3 };/* grid.F90 sfc_layer_type */
enum { POP_STATE_MWJF = 2, POP_STATE_LINEAR = 4 };                    /* state_mod.F90:66-70 */
enum { POP_STATE_RANGE_IGNORE = 1, POP_STATE_RANGE_ENFORCE = 3 };     /* state_mod.F90:79-82 */
enum { POP_SOLVER_PCG = 1, POP_SOLVER_CHRONGEAR = 2, POP_SOLVER_PCSI = 3 };
enum { POP_PRECOND_DIAGONAL = 0, POP_PRECOND_EVP = 1 };               /* POP_SolversMod.F90:120-126 */
enum { POP_TS_LEAPFROG = 1, POP_TS_EULER = 2, POP_TS_AVG = 3,       /* step_mod.F90:302-320,663 */
       POP_TS_ROBERT = 4 };  /* leapfrog step closed by the Robert-Asselin-Williams filter, step_mod.F90:798,919-1354 */
enum { POP_TIME_OLD = 0, POP_TIME_CUR = 1, POP_TIME_NEW = 2 };        /* prognostic.F90:63-68 */

#define POP_MAX_NT 64

/* One plain struct replaces the reference's namelist groups (SURVEY.md section 5 "Config"). */
typedef struct pop_config {
  /* domain_size / domain_nml */
  int nx_global, ny_global, km, nt;
  int ew_boundary_type, ns_boundary_type;
  /* decomposition: the product uses 1 x nranks j-strips (one block per rank);
     block_size_x/y are only honoured by the CPU oracle (multi-block serial back-end). */
  int block_size_x, block_size_y;
  /* advect_nml: per-tracer scheme (advection.F90:234) */
  int tadvect_itype[POP_MAX_NT];
  /* hmix_nml + hmix_del2{u,t}_nml / hmix_del4{u,t}_nml */
  int hmix_tracer_itype, hmix_momentum_itype;
  double ah, am;
  int lvariable_hmixt, lvariable_hmixu;
  int lauto_hmixt, lauto_hmixu;
  /* hmix_gm_nml (hmix_gm.F90:407-428): constant kappa_isop (ah_gm) / kappa_thic (ah_bolus), slope_control 'notanh',
     use_const_ah_bkg_srfbl = .true., ah_bkg_bottom = 0, transition_layer_on = .false.; needs the field "TLAT" */
  double ah_gm, ah_bolus, ah_bkg_srfbl, slm_r, slm_b;
  /* vertical_mix_nml */
  int vmix_itype, implicit_vertical_mix;
  int vdc_kdim_halo; /* GIVEN: 1 -> VDC has levels 0:km+1 (KPP shape), 0 -> 1:km */
  int vdc_ndim;      /* GIVEN: 1 or 2 tracer slots */
  double aidif, bottom_drag;
  double const_vdc, const_vvc;
  double bckgrnd_vdc, bckgrnd_vvc, rich_mix;
  int convection_diff; /* 1: convection_type='diffusion' */
  double convect_diff, convect_visc;
  /* grid_nml */
  int sfc_layer_type, partial_bottom_cells;
  /* pressure_grad_nml / time_manager_nml */
  int lpressure_avg, lbouss_correct, impcor;
  /* state_nml */
  int state_itype, state_range_iopt;
  /* solvers */
  int solver_choice, max_iterations, convergence_check_freq, convergence_check_start;
  int max_lanczos_step;
  double convergence_criterion, lanczos_convergence_criterion;
  /* time step (s): dt(k)=dtt, dtu=dtp=dtt (time_management.F90:962-964) */
  double dtt;
  /* ranks */
  int rank, nranks, device;
  /* time_manager_nml: Robert filter coefficients (time_management.F90:461-464,897-898; Williams 2009) */
  double robert_alpha, robert_nu;
  /* vertical_mix_nml: convection_diff = 0 is convection_type = 'adjustment' with nconvad passes of convad
     (vertical_mix.F90:239,1888-2027; the reference default is 2). 0 passes = no convective adjustment. */
  int nconvad;
  /* solvers: preconditionerChoice (POP_SolversMod.F90:120-126,709-722). 'diagonal' or 'evp' (the block preconditioner
     of Hu et al.: 8 x 8 sub-blocks inverted by error-vector propagation, :2273-2369,2434-2696; production default
     outside gx3v7, namelist_defaults_pop.xml:262-263); 'file' is not supported. */
  int preconditioner_choice;
} pop_config;

/* block descriptor handed to slab routines: mirrors `type block`, source/blocks.F90:30-39 */
typedef struct pop_block {
  int block_id, local_id;
  int ib, ie, jb, je;
  int iblock, jblock;
  const int* i_glob;
  const int* j_glob;
} pop_block;

const char* pop_last_error(void);
void pop_config_defaults(pop_config* cfg);

/* ---- lifecycle (replaces pop_init_phase1/2 for this path, source/initial.F90:133,464) ---- */
int pop_init(const pop_config* cfg);
int pop_finalize(void);
int pop_is_initialized(void);
/* multi-rank bootstrap: rank 0 creates an id (128 bytes), the caller broadcasts it (MPI_Bcast /
   torch.distributed), every rank passes it to pop_comm_init before pop_init. */
int pop_comm_unique_id(char id128[128]);
int pop_comm_init(int rank, int nranks, const char id128[128]);
/* block geometry of this rank (get_block, source/blocks.F90:282) */
int pop_get_block(pop_block* blk);
int pop_local_shape(int* nx_block, int* ny_block, int* j_start_global, int* ny_local);

/* ---- grid (a4): primaries are this rank's PHYSICAL strip (nx_global x ny_local), i fastest.
   The library scatters them into the padded block, fills ghost cells by halo update with the
   field location the reference uses (grid.F90:1420-1520), and derives every other grid array
   (grid.F90:587-647,786-803,978-1041,2537-2596,2882-2932; hmix_del2/del4 init; advection init;
   POP_SolversInit). */
/* partial bottom cells (grid_nml partial_bottom_cells, grid.F90:917-960): DZBC = thickness (cm) of the bottom cell of
   every column, this rank's physical strip -- what the reference reads from bottom_cell_file (read_bottom_cell,
   grid.F90:2116).  Call BEFORE pop_set_grid, which derives DZT, DZU, HT, HU and HUR from it. */
int pop_set_bottom_cells(const double* DZBC);
int pop_set_grid(const double* ULAT, const double* HTN, const double* HTE, const double* HUS,
                 const double* HUW, const double* DXU, const double* DYU, const double* DXT,
                 const double* DYT, const int* KMT, const double* dz /* km, cm */);

/* ---- fields: name in {TRACER,UVEL,VVEL,RHO,PSURF,GRADPX,GRADPY,UBTROP,VBTROP,PGUESS,
   VDC,VVC,STF,SMF,SHF_QSW,FW,FW_OLD,TFW, and every grid array}.  tlev = POP_TIME_* for
   prognostic fields (resolved through the rotating indices), ignored otherwise. */
void* pop_device_ptr(const char* name, int tlev);
long pop_field_size(const char* name);   /* number of elements of the padded local array */
/* copy a padded local array host<->device */
int pop_set_field(const char* name, int tlev, const void* host);
int pop_get_field(const char* name, int tlev, void* host);
/* scatter/gather this rank's physical strip (nx_global x ny_local [x nz]) */
int pop_scatter_field(const char* name, int tlev, const void* host_strip);
int pop_gather_field(const char* name, int tlev, void* host_strip);
/* levels z0 .. z0+nz-1 only (0-based; level index runs over km, or n*km+k for TRACER); `strip` may be
   a host or a device pointer */
int pop_scatter_field_levels(const char* name, int tlev, int z0, int nz, const void* strip);

/* ---- time-step scalars (a29): step_mod.F90:302-320 ---- */
int pop_set_timestep(int ts_type);

/* ---- slab operators: reference argument lists. ---- */
/* advection.F90:1577 (tracers with tadvect_itype = lw_lim: call pop_comp_flux_vel_ghost once per step first, and
   levels in ascending order) */
int pop_advt(int k, double* LTK, double* WTK, const double* TMIX, const double* TRCR,
             const double* UUU, const double* VVV, const pop_block* blk);
/* advection.F90:1014 comp_flux_vel_ghost(DH, errorCode): flux velocities of UVEL/VVEL(curtime) of every level with
   halo-updated ghost cells, for the lw_lim scheme; DH (nx_block, ny_block) host or device; no-op without lw_lim */
int pop_comp_flux_vel_ghost(const double* DH);
/* advection.F90:1127 */
int pop_advu(int k, double* LUK, double* LVK, double* WUK, const double* UUU, const double* VVV,
             const pop_block* blk);
/* horizontal_mix.F90:486 / :427 */
int pop_hdifft(int k, double* HDTK, const double* TMIX, const double* UMIX, const double* VMIX,
               const pop_block* blk);
int pop_hdiffu(int k, double* HDUK, double* HDVK, const double* UMIXK, const double* VMIXK,
               const pop_block* blk);
/* pressure_grad.F90:187 */
int pop_gradp(int k, double* PKX, double* PKY, const double* RHOK_OLD, const double* RHOK_CUR,
              const double* RHOK_NEW, const pop_block* blk);
/* operators.F90:126 / :49 */
int pop_grad(int k, double* GRADX, double* GRADY, const double* F, const pop_block* blk);
int pop_div(int k, double* DIV_OUT, const double* UX, const double* UY, const pop_block* blk);
/* vertical_mix.F90:691 / :853 / :1164 / :1460 / :1679 */
int pop_vdifft(int k, double* VDTK, const double* TOLD, const double* STF, const pop_block* blk);
int pop_vdiffu(int k, double* VDUK, double* VDVK, const double* UOLD, const double* VOLD,
               const double* SMF, const pop_block* blk);
int pop_impvmixt(double* TNEW, const double* TOLD, const double* PSFC, int nfirst, int nlast,
                 const pop_block* blk);
int pop_impvmixt_correct(double* TNEW, const double* PSFC, const double* RHS, int nfirst,
                         int nlast, const pop_block* blk);
int pop_impvmixu(double* UNEW, double* VNEW, const pop_block* blk);
/* vertical_mix.F90:518: const (vmix_const.F90:144), rich (vmix_rich.F90:179, full cells); GIVEN is a no-op */
int pop_vmix_coeffs(int k, const double* TMIX, const double* UMIX, const double* VMIX,
                    const double* RHOMIX, const pop_block* blk);
/* state_mod.F90:258: optional outputs are nullable */
int pop_state(int k, int kk, const double* TEMPK, const double* SALTK, const pop_block* blk,
              double* RHOOUT, double* RHOFULL, double* DRHODT, double* DRHODS);

/* ---- barotropic solver (POP_SolversMod.F90:327,1110,1158,2376) ---- */
int pop_solvers_run(double* sfcPressure, const double* rhsClinic);
int pop_solvers_diagonal(const double* diagonalCorrection, int blockIndx);
int pop_solvers_get_diagnostics(int* iterationCount, double* residual);
int pop_btrop_operator(double* AX, const double* X, int bid);
int pop_solvers_prep(void); /* POP_SolversPrep: Lanczos eigenvalue bounds for PCSI */
int pop_solvers_get_eigs(double* mineig, double* maxeig);
/* EVP preconditioner set-up (POP_SolversPrep :252-292, EvpPre :2434-2506): number of sub-blocks of this rank's block,
   how many of them fall back to the diagonal (landIndx = 1), and the largest max|rinv*rin - I| of the influence-matrix
   inverses -- the quantity the reference checks against 1e-8 (:2594-2614; pop_solvers_prep fails above that bound) */
int pop_solvers_get_evp_diagnostics(int* subBlocks, int* landSubBlocks, double* maxInverseError);

/* ---- communication (mpi/POP_HaloMod.F90:1732,2766,4122; mpi/POP_ReductionsMod.F90:144,823) ----
   fillValue: the reference writes it into ghost cells that face an ELIMINATED land block (POP_HaloMod.F90:1911-1913).
   The strip decomposition of this library eliminates no blocks, so no ghost cell ever receives it; ghost cells on a
   closed boundary are left untouched exactly as in the reference (no message targets them).  The argument is kept for
   the reference's signature and has no effect. */
int pop_halo_update_2d_r8(double* array, int fieldLoc, int fieldKind, double fillValue);
int pop_halo_update_3d_r8(double* array, int nz, int fieldLoc, int fieldKind, double fillValue);
int pop_halo_update_4d_r8(double* array, int nz, int nt, int fieldLoc, int fieldKind,
                          double fillValue);
int pop_halo_update_2d_i4(int* array, int fieldLoc, int fieldKind, int fillValue);
/* the other members of the generic POP_HaloUpdate interface (mpi/POP_HaloMod.F90:79-89: 2DR4 :2078, 3DR4 :3218,
   3DI4 :3670, 4DR4 :4592, 4DI4 :5062) */
int pop_halo_update_3d_i4(int* array, int nz, int fieldLoc, int fieldKind, int fillValue);
int pop_halo_update_4d_i4(int* array, int nz, int nt, int fieldLoc, int fieldKind, int fillValue);
int pop_halo_update_2d_r4(float* array, int fieldLoc, int fieldKind, float fillValue);
int pop_halo_update_3d_r4(float* array, int nz, int fieldLoc, int fieldKind, float fillValue);
int pop_halo_update_4d_r4(float* array, int nz, int nt, int fieldLoc, int fieldKind, float fillValue);
int pop_global_sum_2d_r8(const double* array, int fieldLoc, const double* mMask, double* sum);
int pop_global_sum_nfields_2d_r8(const double* array, int nfields, int fieldLoc,
                                 const double* mMask, double* sums);

/* ---- fused drivers on the library-resident state ---- */
int pop_dhdt(void);                               /* surface_hgt.F90:131 */
int pop_baroclinic_driver(void);                  /* baroclinic.F90:578 */
int pop_barotropic_driver(void);                  /* barotropic.F90:267 */
int pop_baroclinic_correct_adjust(void);          /* baroclinic.F90:1217 */
int pop_step(int ts_type);                        /* step_mod.F90:126 */
/* end-to-end step used by a coupled host: uploads this step's surface forcing from HOST
   buffers (STF nt fields, SMF 2 fields, FW; physical strips), advances one step,
   returns the new surface state (SST, SSS, PSURF, U1, V1: physical strips) to HOST buffers.
   SHF_QSW is accepted for the coupler's argument list but IGNORED (may be NULL): penetrative short-wave absorption
   (add_sw_absorb, baroclinic.F90:2176) is not part of this path, so no kernel reads it and it is not copied. */
int pop_step_coupled(int ts_type, const double* STF, const double* SMF, const double* SHF_QSW,
                     const double* FW, double* sfc_out /* 5 strips */);

/* ---- instrumentation ---- */
long pop_kernel_launch_count(void);               /* kernels launched by this library so far */
int pop_timer_get(const char* name, double* ms, long* calls); /* CUDA-event timers, reference names */
int pop_timers_reset(void);
int pop_timers_enable(int on);
int pop_sync(void);
void* pop_stream(void);                            /* cudaStream_t the library launches on */

#ifdef __cplusplus
}
#endif
#endif /* POP_B200_H */
