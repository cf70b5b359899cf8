/*
 * pop_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the POP2 baroclinic/barotropic time-step hot path, following the
 * reference's loop bounds, operation order and masks routine by routine (each function cites
 * the reference file:line it follows).  It is a multi-block single-process implementation with
 * the semantics of the reference's `serial/` communication back-end.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (pop2-cesm_b200/) never links, imports or calls it.
 *
 * Pinning status: halo update, global sums, block decomposition and the MWJF known answer are
 * pinned against the reference's own unit-test expectations (tests/test_oracle_*.py).  The
 * operators advt, advu, hdifft, hdiffu, vdifft, vdiffu, impvmixt/u, gradp, btropOperator, solvers have no
 * golden data in the reference tree: for those this oracle is "parity unpinned" and relies on
 * line-by-line fidelity plus invariant tests.
 */
#ifndef POP_ORACLE_H
#define POP_ORACLE_H
#include <stddef.h>
#include "../include/pop_b200.h" /* pop_config struct + enums only (no product code) */

#define NGHOST 2 /* source/blocks.F90:51 */

/* physical constants, standalone (non-CCSMCOUPLED) values: source/pop_constants.F90:235-241 */
#define O_GRAV 980.6
#define O_OMEGA 7.292123625e-5
#define O_RADIUS 6370.0e5
#define O_PI 3.14159265358979323846

typedef struct {
  pop_config cfg;
  int nxb, nyb, km, nt;
  int nbx, nby, nblocks; /* all blocks (global ids 1..nblocks); eliminated blocks: active=0 */
  int *ib, *ie, *jb, *je, *iblk, *jblk, *active;
  int *i_glob, *j_glob; /* [nblocks][nxb], [nblocks][nyb] */
  size_t n2, n3;        /* nxb*nyb, nxb*nyb*km */
  /* vertical grid (index 0..km+1 where the reference has 0:km) */
  double *dz, *dzw, *dzr, *dz2r, *dzwr, *c2dz, *zt, *zw;
  /* horizontal grid, each [nblocks][nyb][nxb] */
  double *TLAT; /* T-point latitude (radians): input, only GM reads it (FCORT, grid.F90:1159) */
  double *ULAT, *HTN, *HTE, *HUS, *HUW, *DXU, *DYU, *DXT, *DYT;
  double *DXUR, *DYUR, *DXTR, *DYTR, *UAREA, *TAREA, *UAREA_R, *TAREA_R;
  double *HU, *HUR, *HT, *FCOR, *RCALCT, *RCALCU;
  double *AU0, *AUN, *AUE, *AUNE, *AT0, *ATS, *ATW, *ATSW;
  double *KXU, *KYU;
  int *KMT, *KMU, *KMTN, *KMTS, *KMTE, *KMTW, *KMTEE, *KMTNN;
  /* partial bottom cells (grid.F90:917-960): DZBC input (thickness of the bottom cell of each column), DZT / DZU
     [nblocks][0:km+1][nyb][nxb] */
  double *DZBC, *DZT, *DZU;
  double uarea_equator;
  /* hmix coefficients (hmix_del2.F90 / hmix_del4.F90 init) */
  double *DTN, *DTS, *DTE, *DTW, *AHF;
  double *DUC, *DUN, *DUS, *DUE, *DUW, *DMC, *DMN, *DMS, *DME, *DMW, *DUM, *AMF;
  double ah, am;
  /* upwind3 tables (advection.F90:420-562) */
  double *talfzp, *tbetzp, *tgamzp, *talfzm, *tbetzm, *tdelzm;
  double *TALFXP, *TBETXP, *TGAMXP, *TALFYP, *TBETYP, *TGAMYP;
  double *TALFXM, *TBETXM, *TDELXM, *TALFYM, *TBETYM, *TDELYM;
  /* state */
  double *pressz, *tmin, *tmax, *smin, *smax, *bouss;
  /* lw_lim advection (advection.F90:566-694): grid coefficients, flux velocities of the outermost ghost cells
     (comp_flux_vel_ghost, :1014-1120), flux velocities of level k+1 kept between levels (FLUX_VEL_prev) */
  int use_lw_lim;
  double *p5_dz_ph_r, *p5_DXT_ph_R, *p5_DYT_ph_R, *UTE_to_UVEL_E, *VTN_to_VVEL_N; /* the last two: [nblocks][km or 1][n2] */
  double *UTE_jbm2, *WTKB_jbm2, *WTKB_jep2, *WTKB_ibm2, *WTKB_iep2;               /* [nblocks][km][nxb or nyb] */
  double *FLUX_VEL_prev;                                                          /* [nblocks][5][n2] */
  /* vertical mixing */
  double *VDC, *VVC; /* VDC: [nblocks][vdc_nd][vdc_nk][nyb][nxb], level offset vdc_k0 */
  int vdc_nk, vdc_k0, vdc_nd, vvc_nk;
  double *afac_t, *afac_u;
  /* carried state */
  double *VTF, *VUF, *VVF, *SUMX, *SUMY, *RHOKMX, *RHOKMY, *AUX, *UTK, *VTK;
  /* prognostic (3 time levels; index via oldtime/curtime/newtime) */
  double *TRACER[3], *UVEL[3], *VVEL[3], *RHO[3];
  double *PSURF[3], *GRADPX[3], *GRADPY[3], *UBTROP[3], *VBTROP[3];
  double *PGUESS;
  int oldtime, curtime, newtime, mixtime;
  /* forcing */
  double *STF, *TFW, *SMF, *SHF_QSW, *FW, *FW_OLD;
  /* time step scalars */
  double dtt, dtu, dtp, c2dtu, c2dtp, beta, alpha, theta, gamma_;
  double *c2dtt, *dt;
  int leapfrogts, f_euler_ts, avg_ts, mix_pass;
  /* solver */
  double *centerWgtClinicIndep, *centerWgtClinic;
  double *btropWgtCenter, *btropWgtNorth, *btropWgtEast, *btropWgtNE, *mMaskTropic;
  double residualNorm, convergenceCriterion, rmsResidual;
  int numIterations;
  double PcsiMaxEigs, PcsiMinEigs;
  int lanczos_steps;
  /* barotropic */
  int *CHECKER, *CONSTNT;
  double rcheck, rconst;
  /* Robert filter (step_mod.F90:81-103,1558-1600): S terms, budget areas, previous conservation factors */
  double *STORE_RF, *bgtarea_t_k;
  double rf_volume_2_km, rf_S_prev[POP_MAX_NT];
  int rf_ready, rf_S_prev_valid[POP_MAX_NT];
  /* work */
  double *DH, *DHU, *ZX, *ZY;
  /* timers (seconds), indices below */
  double timer[16];
} omodel;

enum { OT_STEP = 0, OT_BAROCLINIC, OT_BAROTROPIC, OT_HALO, OT_ADVT, OT_HDIFFT, OT_VMIXT,
       OT_ADVU, OT_HDIFFU, OT_VMIXU, OT_STATE, OT_SOLVER, OT_N };

extern omodel M;

/* partial-bottom-cell thickness of level k (0..km+1) of block b */
#define PBC (M.cfg.partial_bottom_cells)
#define DZT3(b, k) (M.DZT + ((size_t)(b) * (M.km + 2) + (k)) * M.n2)
#define DZU3(b, k) (M.DZU + ((size_t)(b) * (M.km + 2) + (k)) * M.n2)
/* index macros: 1-based like the reference */
#define IX2(i, j) ((size_t)((j)-1) * M.nxb + ((i)-1))
#define B2(a, b) ((a) + (size_t)(b)*M.n2)                               /* block b (0-based) of 2-d field */
#define B3(a, b) ((a) + (size_t)(b)*M.n3)                               /* 3-d field */
#define B4(a, b) ((a) + (size_t)(b)*M.n3 * M.nt)                        /* tracer field */
#define K3(a, k) ((a) + (size_t)((k)-1) * M.n2)                         /* level k of a block 3-d */
#define KN4(a, k, n) ((a) + ((size_t)((n)-1) * M.km + ((k)-1)) * M.n2) /* level k, tracer n */

/* ---- domain / comm (o_domain.c) ---- */
int oracle_init(const pop_config* cfg);
void oracle_finalize(void);
int oracle_set_active(const int* active); /* land-block elimination for halo tests */
void oracle_halo_2d(double* a, int loc, int kind, double fill);
void oracle_halo_3d(double* a, int nz, int loc, int kind, double fill);
void oracle_halo_4d(double* a, int nz, int nt, int loc, int kind, double fill);
void oracle_halo_2d_i4(int* a, int loc, int kind, int fill);
double oracle_global_sum(const double* a, int loc, const double* mask);
void oracle_global_sum_n(const double* a, int nf, int loc, const double* mask, double* out);
void oracle_scatter_2d(double* dst, const double* glob, int loc, int kind);
void oracle_scatter_2d_i4(int* dst, const int* glob);
void oracle_scatter(double* dst, const double* glob, int nz); /* physical cells only */
void oracle_gather(double* glob, const double* src, int nz);
void oracle_gather_i4(int* glob, const int* src);
int oracle_distribution_cartesian(int nprocs, int nbx, int nby, const int* work, int* loc);
int oracle_block_info(int b, int* out8, int* iglob, int* jglob);

/* ---- GM (o_gm.c) ---- */
void o_gm_reset(void);
void o_gm_begin_step(void);
void* o_gm_field(const char* name);
int o_gm_cancellation(void);

/* ---- grid + init (o_grid.c) ---- */
int oracle_set_bottom_cells(const double* DZBC_G); /* before oracle_set_grid when partial_bottom_cells */
int oracle_set_grid(const double* ULAT, const double* HTN, const double* HTE, const double* HUS,
                    const double* HUW, const double* DXU, const double* DYU, const double* DXT,
                    const double* DYT, const int* KMT, const double* dz);
void oracle_set_timestep(int ts_type);

/* ---- operators (o_ops.c): slab routines, block index b is 0-based, k 1-based ---- */
void o_state(int k, int kk, const double* T, const double* S, int b, double* RHOOUT,
             double* RHOFULL, double* DRHODT, double* DRHODS);
void o_comp_flux_vel(int k, const double* UUU, const double* VVV, const double* WTK, double* UTE,
                     double* UTW, double* VTN, double* VTS, double* WTKB, int b);
void o_comp_flux_vel_ghost(void); /* baroclinic.F90:667 */
void o_advt(int k, double* LTK, double* WTK, const double* TMIX, const double* TRCR,
            const double* UUU, const double* VVV, int b);
void o_advu(int k, double* LUK, double* LVK, double* WUK, const double* UUU, const double* VVV,
            int b);
void o_hdifft(int k, double* HDTK, const double* TMIX, const double* UMIX, const double* VMIX,
              int b);
void o_hdiffu(int k, double* HDUK, double* HDVK, const double* UMIXK, const double* VMIXK, int b);
void o_grad(int k, double* GX, double* GY, const double* F, int b);
void o_div(int k, double* D, const double* UX, const double* UY, int b);
void o_gradp(int k, double* PKX, double* PKY, const double* RO, const double* RC,
             const double* RN, int b);
void o_vdifft(int k, double* VDTK, const double* TOLD, const double* STF, int b);
void o_vdiffu(int k, double* VDUK, double* VDVK, const double* UOLD, const double* VOLD,
              const double* SMF, int b);
void o_impvmixt(double* TNEW, const double* TOLD, const double* PSFC, int nfirst, int nlast,
                int b);
void o_impvmixt_correct(double* TNEW, const double* PSFC, const double* RHS, int nfirst,
                        int nlast, int b);
void o_impvmixu(double* UNEW, double* VNEW, int b);
void o_vmix_coeffs(int k, const double* TMIX, const double* UMIX, const double* VMIX,
                   const double* RHOMIX, int b);
void o_tgrid_to_ugrid(double* AU, const double* AT, int b);
void o_ugrid_to_tgrid(double* AT, const double* AU, int b);

/* ---- solver (o_solver.c) ---- */
int o_solvers_init(void);
void o_solvers_diagonal(const double* diagCorr, int b);
void o_btrop_operator(double* AX, const double* X, int b);
int o_solvers_run(double* sfcPressure, const double* rhs);
int o_solvers_prep(void);

/* ---- drivers (o_step.c) ---- */
void o_dhdt(void);
int o_baroclinic_driver(void);
int o_barotropic_driver(void);
void o_baroclinic_correct_adjust(void);
int oracle_step(int ts_type);
void o_init_barotropic(void);

/* field registry for the python harness */
void* oracle_field(const char* name, int tlev);
double oracle_timer(int id);
void oracle_timers_reset(void);
double o_now(void);
#endif
