// pop_ctx.h -- library context of the B200-native POP2 hot path (one process / one block per GPU).
//
// Data layout in HBM: exactly the reference's block layout (source/prognostic.F90:38-61): every
// array is (nx_block, ny_block [,km [,nt]]) with i fastest and a 2-cell ghost ring
// (source/blocks.F90:51-56); one block per rank = a 1 x P strip in j, so
// nx_block = nx_global+4 and ny_block = ny_global/P+4.  Three time levels are separate
// allocations addressed through rotating indices (step_mod.F90:827-830: zero bytes moved).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <map>
#include <utility>
#include <string>
#include <vector>
#include "../../include/pop_b200.h"

#define POP_NGHOST 2
#define POP_KMAX 128  // capacity of the per-level constant tables

// physical constants: standalone (non-CCSMCOUPLED) values, source/pop_constants.F90:235-241
#define POP_GRAV 980.6
#define POP_OMEGA 7.292123625e-5
#define POP_RADIUS 6370.0e5
#define POP_PI 3.14159265358979323846

// per-level tables, uploaded to __constant__ memory (uniform index across a warp)
struct VertConst {
  double dz[POP_KMAX + 2], dzw[POP_KMAX + 2], dzr[POP_KMAX + 2], dz2r[POP_KMAX + 2];
  double dzwr[POP_KMAX + 2], c2dz[POP_KMAX + 2], zt[POP_KMAX + 2], zw[POP_KMAX + 2];
  double c2dtt[POP_KMAX + 2], afac_t[POP_KMAX + 2], afac_u[POP_KMAX + 2];
  double pressz[POP_KMAX + 2], tmin[POP_KMAX + 2], tmax[POP_KMAX + 2], smin[POP_KMAX + 2];
  double smax[POP_KMAX + 2], bouss[POP_KMAX + 2];
  // pressure-dependent MWJF coefficient sets of each level (state_mod.F90:433-452), evaluated once on the host
  double eos_n0t0[POP_KMAX + 2], eos_n0t2[POP_KMAX + 2], eos_n1t0[POP_KMAX + 2];
  double eos_d0t0[POP_KMAX + 2], eos_d0t1[POP_KMAX + 2], eos_d0t3[POP_KMAX + 2];
  double hfac_t[POP_KMAX + 2], hfac_u[POP_KMAX + 2];  // dz(k)/c2dtt(k), dz(k)/c2dtu (vertical_mix.F90:1240,1755)
  double talfzp[POP_KMAX + 2], tbetzp[POP_KMAX + 2], tgamzp[POP_KMAX + 2];
  double talfzm[POP_KMAX + 2], tbetzm[POP_KMAX + 2], tdelzm[POP_KMAX + 2];
};

struct DevField {
  void* p = nullptr;
  size_t elems = 0;  // number of elements of the padded local array
  int nz = 1;        // product of the trailing dims (km, nt*km, ...)
  bool is_int = false;
};

struct Timer {
  double ms = 0.0;
  long calls = 0;
};

struct Ctx {
  pop_config cfg;
  bool initialized = false, grid_set = false;
  int nxg = 0, nyg = 0, km = 0, nt = 0;
  int rank = 0, nranks = 1;
  int ny_local = 0, j0 = 1;  // physical rows of this strip: global rows j0 .. j0+ny_local-1
  int nxb = 0, nyb = 0, ib = 3, ie = 0, jb = 3, je = 0;  // 1-based like the reference
  size_t n2 = 0, n3 = 0;
  std::vector<int> i_glob, j_glob;
  std::vector<double> dzbc_strip;  // pop_set_bottom_cells: thickness of the bottom cell, physical strip
  int *d_iglob = nullptr, *d_jglob = nullptr;  // device copies
  cudaStream_t stream = nullptr;
  // side stream: with finish_mode 1 (opt-in) the velocity finish (impvmixu + barotropic-mean removal) of step n runs
  // here, concurrently with the barotropic solve on `stream` (the solve only needs ZX,ZY of the column kernel)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_chk[2] = {nullptr, nullptr};  // lagged P-CSI convergence checks
  cudaEvent_t ev_chk_go = nullptr;
  cudaStream_t stream_chk = nullptr;           // ... whose reductions run beside the passes
  bool no_overlap = false;  // POP_B200_NO_OVERLAP=1
  // velocity finish (impvmixu + Uold + mean removal + mask): 0 = one kernel after the barotropic solve that also adds
  // UBTROP/VBTROP(new) (default); 1 = on the side stream under the solve, separate add-barotropic kernel
  // (POP_B200_OVERLAP_FINISH=1, the round-1 layout); 2 = inside baroclinic_driver (POP_B200_NO_OVERLAP=1)
  int finish_mode = 0;
  bool overlap_exchange = false;  // POP_B200_OVERLAP_EXCHANGE=1 (opt-in): P-CSI boundary tiles first, exchange under the interior
  // P-CSI on deep strips (P > 1): ghost depth in rows (0: off; POP_B200_DEEP_HALO=<rows>, POP_B200_NO_DEEP_HALO=1);
  // deep_force (POP_B200_DEEP_HALO_FORCE=1) runs the deep layout on one rank too (test aid)
  int deep_halo = 12;
  bool deep_force = false;
  cudaStream_t stream_x = nullptr;
  cudaEvent_t ev_xb = nullptr, ev_xx = nullptr;
  bool no_pdl = false;      // POP_B200_NO_PDL=1: plain stream-ordered launches in the P-CSI loop
  // pop_step_coupled: host<->device copies of the surface forcing / surface state on their own stream, overlapped
  // with the step (pop_abi.cu); the step calls the three hooks below at the points where the data are needed / final
  cudaStream_t stream_cp = nullptr;
  cudaEvent_t ev_cp_in = nullptr, ev_cp_a = nullptr, ev_cp_b = nullptr;
  struct {
    bool forcing_pending = false;  // SMF/SHF_QSW/FW still in flight on stream_cp
    bool fw_halo_pending = false;  // FW arrived as a physical strip: its ghost cells are filled before the first reader
    double* out = nullptr;         // host buffer of the surface state, or null
    bool early = false;            // outputs may leave as soon as they are final (not on averaging / filtered steps)
    bool psurf_sent = false, ts_sent = false;
    // U1/V1: when the host buffer is pinned (device-accessible) the velocity-finish kernel stores the surface level
    // straight into it while it runs; uv_dev = device alias of out + 3*strip (V follows one strip later), or null
    double* uv_dev = nullptr;
    bool uv_sent = false;
    // STF: with a pinned host buffer the tracer kernel reads the surface flux of its columns straight from the host
    // (once per column, at level 1) while the copy into the device field runs on the copy stream, off the critical path
    const double* stf_dev = nullptr;
  } cio;
  std::map<std::string, DevField> fields;
  VertConst vc;  // host copy
  // rotating time indices (prognostic.F90:63-68)
  int oldtime = 0, curtime = 1, newtime = 2, mixtime = 0;
  // time-step scalars (step_mod.F90:302-320, time_management.F90:434-439)
  double dtt = 0, dtu = 0, dtp = 0, c2dtu = 0, c2dtp = 0, beta = 0;
  double alpha = 1.0 / 3.0, theta = 0.5, gamma = 1.0 - 2.0 * (1.0 / 3.0);
  bool leapfrogts = true, f_euler_ts = false, avg_ts = false;
  // Robert filter state (step_mod.F90:81-103, init_step :1558-1600)
  bool rf_ready = false;
  double rf_volume_2_km = 0, rf_bgtarea1 = 0;
  double rf_S_prev[POP_MAX_NT] = {0};
  bool rf_S_prev_valid[POP_MAX_NT] = {false};
  // GM (pop_gm.cu): RBR/metric ratios need rebuilding (TLAT or the grid changed); VDC_BASE needs re-saving
  bool gm_dirty = true, gm_vdc_dirty = true, gm_force_general = false;
  // hmix / misc scalars
  double ah = 0, am = 0, uarea_equator = 0;
  int vdc_nk = 0, vdc_k0 = 1, vdc_nd = 1, vvc_nk = 0;
  bool use_upwind3 = false, use_centered = false, use_lw_lim = false;
  bool lw_flux_ready = false;  // comp_flux_vel_ghost has run for the current velocities
  bool lw_coef_dirty = true;   // grid coefficients of lw_lim follow pop_set_grid
  // solver
  double residualNorm = 0, convergenceCriterion = 0, rmsResidual = 0;
  int numIterations = 0;
  double pcsiMaxEigs = 0, pcsiMinEigs = 0;
  int lanczosSteps = 0;
  double rcheck = 0, rconst = 0;
  // reduction scratch (pop_reduce.cu)
  double* d_partials = nullptr;  // [POP_RED_NF][red_blocks][2] double-double block partials
  double* d_partials_big = nullptr;  // [n][2] partials of kernels with their own (fixed) grid
  int partials_big_n = 0;
  double* d_sums = nullptr;      // device results: [POP_RED_NF] doubles
  double* h_sums = nullptr;      // pinned mirror of d_sums
  double* d_local = nullptr;     // [POP_RED_NF][2] this rank's dd sums (all-gather send buffer)
  double* d_gather = nullptr;    // [nranks][POP_RED_NF][2] all-gathered dd sums
  void* d_scal = nullptr;        // SolverScalars on the device
  int red_blocks = 0;            // fixed grid of every reduction (determinism)
  double* d_tripole = nullptr;   // bufTripole(nxGlobal, haloWidth+1, nz) of the tripole fold
  size_t tripole_elems = 0;
  // halo message buffers (multi-rank): rows to/from the south and north neighbours
  double* d_sendS = nullptr;
  double* d_sendN = nullptr;
  double* d_recvS = nullptr;
  double* d_recvN = nullptr;
  size_t halo_buf_elems = 0;
  void* nccl_comm = nullptr;
  // peer-memory halo (pop_halo.cu): mailbox in HBM that the neighbouring ranks write over NVLink
  bool p2p_on = false;
  double* p2p_mbox = nullptr;                       // [2 parities][2 sides][p2p_cap] + 2 flags
  double *p2p_peerS = nullptr, *p2p_peerN = nullptr;  // the neighbours' mailboxes (CUDA IPC mappings)
  size_t p2p_cap = 0;                               // doubles per message slot
  unsigned long long p2p_seq = 0;                   // halo sequence number (same on every rank)
  unsigned int* p2p_counter = nullptr;              // CTA arrival counter of the fused kernel
  int* p2p_err = nullptr;                           // set by a kernel whose wait timed out
  // staging buffers for host-pointer arguments of the slab API
  std::map<std::string, std::pair<void*, size_t>> stage;
  // instrumentation
  long launches = 0;
  bool timers_on = false;
  int timer_suppress = 0;  // >0 inside the solver iterations (no per-iteration events)
  std::map<std::string, Timer> timers;
  int sm_count = 148;
  bool no_pbc_fast = false;       // POP_B200_NO_PBC_FAST=1: partial bottom cells on the general tracer column kernel
  bool no_pcsi_blocking = false;  // POP_B200_NO_PCSI_BLOCKING=1: one P-CSI iteration per pass (debugging aid)
  bool no_fast_tracer = false;  // POP_B200_NO_FAST_TRACER=1: always the general tracer column kernel
  bool no_tma = false;  // POP_B200_NO_TMA=1: force the plain-load column kernels (debugging aid)
  // POP_B200_THOMAS_TMA=1: the TMA-staged Thomas kernels (opt-in: measured slower than the register-ring kernels,
  // 33.6 vs 20.2 ms for the two tracer solves at tx0.1v3 -- profiles/r2_ncu_summary.md)
  bool thomas_tma = false;
};

extern Ctx G;

// ---- error handling (POP_ErrorMod.F90:45-47 convention) ----
void pop_set_error(const char* fmt, ...);
#define POP_CHECK_CUDA(call)                                                                 \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      pop_set_error("%s:%d CUDA error %s in %s", __FILE__, __LINE__, cudaGetErrorString(e_), \
                    #call);                                                                  \
      return POP_FAIL;                                                                       \
    }                                                                                        \
  } while (0)
#define POP_REQUIRE(cond, ...)    \
  do {                            \
    if (!(cond)) {                \
      pop_set_error(__VA_ARGS__); \
      return POP_FAIL;            \
    }                             \
  } while (0)
#define POP_TRY(call)                   \
  do {                                  \
    int rc_ = (call);                   \
    if (rc_ != POP_SUCCESS) return rc_; \
  } while (0)

// ---- field registry ----
double* fld(const char* name);                  // device pointer (double field), nullptr if absent
int* fldi(const char* name);                    // device pointer (int field)
double* fld_t(const char* base, int storage);   // prognostic field at storage index 0..2
int alloc_field(const char* name, int nz, bool is_int);
bool resolve_name(const char* name, int tlev, std::string* out);

// launch bookkeeping: every kernel launch of this library goes through POP_LAUNCH
#ifndef POP_LAUNCH
#define POP_LAUNCH(kernel, grid, block, smem, ...)                    \
  do {                                                                \
    kernel<<<(grid), (block), (smem), G.stream>>>(__VA_ARGS__);       \
    G.launches++;                                                     \
  } while (0)
#endif
// launches of the P-CSI loop (pass kernel <-> ghost fill / strip exchange, ~600 dependent launches per step): the
// next kernel's launch overlaps the tail of the previous one (programmatic dependent launch); see pdl_wait()
#ifndef POP_EMUL
template <typename... KArgs, typename... Args>
inline void pop_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define POP_LAUNCH_PDL(kernel, grid, block, smem, ...)                                              \
  do {                                                                                              \
    if (G.no_pdl) kernel<<<(grid), (block), (smem), G.stream>>>(__VA_ARGS__);                       \
    else pop_launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), G.stream, __VA_ARGS__);    \
    G.launches++;                                                                                   \
  } while (0)
#else
#define POP_LAUNCH_PDL POP_LAUNCH
#endif
int pop_post_launch(const char* what);  // cudaGetLastError -> POP error

// scoped CUDA-event timer with the reference timer names (timers.F90)
struct ScopedTimer {
  const char* name;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool on;
  explicit ScopedTimer(const char* n);
  ~ScopedTimer();
};

// device-side view of everything a kernel needs (passed by value as a kernel parameter)
struct GridView {
  int nxb, nyb, km, nt, ib, ie, jb, je;  // ib..je 1-based
  size_t n2, n3;
  const int *KMT, *KMU;
  const double *DXU, *DYU, *DXUR, *DYUR, *UAREA_R, *TAREA_R, *TAREA, *HUR, *HU, *FCOR, *KXU, *KYU;
  const double *DTN, *DTS, *DTE, *DTW, *AHF;
  const double *DUC, *DUN, *DUS, *DUE, *DUW, *DMC, *DMN, *DMS, *DME, *DMW, *DUM, *AMF;
  const double *AU0, *AUN, *AUE, *AUNE, *RCALCT, *RCALCU;
  const double *TALFXP, *TBETXP, *TGAMXP, *TALFYP, *TBETYP, *TGAMYP;
  const double *TALFXM, *TBETXM, *TDELXM, *TALFYM, *TBETYM, *TDELYM;
  const double *VDC, *VVC;
  int vdc_nk, vdc_k0, vdc_nd, vvc_nk;
  // partial bottom cells (grid.F90:917-960): thickness of the T / U cells, levels 0..km+1 (level k at + k*n2); null
  // when partial_bottom_cells is off
  const double *DZT, *DZU;
  double aidif;  // vertical_mix_nml aidif (afac = aidif/dzw is rebuilt per cell with partial bottom cells)
};
GridView grid_view();

// ---- internal entry points implemented across the .cu files ----
int upload_vert_const();
int set_timestep(int ts_type);
// halo (pop_halo.cu)
int halo_update(double* a, int nz, int loc, int kind, double fill);
int halo_update_i4(int* a, int nz, int loc, int kind, int fill);
int halo_update_r4(float* a, int nz, int loc, int kind, float fill);
int halo_rows_only(double* a, int nz);
int halo_exchange_deep(double* a, int nz, int gd, size_t n2d);  // strip exchange of the solver's deep strips
int halo_exchange_rows(double* a, int nz);  // peer-memory exchange of centre-scalar rows without the wrap of own rows
int halo_ew_own_rows(double* a, int nz);    // the deferred east-west wrap
int comm_init(int rank, int nranks, const char* id128);
int comm_unique_id(char* id128);
int comm_finalize();
int p2p_setup();     // after the sizes are known (pop_init); falls back to NCCL send/recv when unavailable
int p2p_teardown();
int p2p_check();     // POP_FAIL if a peer wait timed out
int comm_allreduce_min(double* host_value);
// reductions (pop_reduce.cu): masked physical-domain sums of nfields 2-d fields, accumulated in
// double-double; results (rounded to double) land in out_host[nfields] when out_host != nullptr
// and always in G.d_sums[0..nfields-1] on the device (stream ordered).
#define POP_RED_NF 4
enum { RED_POST_NONE = 0, RED_POST_CG_INIT, RED_POST_CG_ITER, RED_POST_PCG_ETA1, RED_POST_PCG_ETA2,
       RED_POST_RR };
int global_sum_dev(const double* a, int nfields, size_t field_stride, int loc, const double* mask,
                   double* out_host);
int reduce_alloc();
// second stage for kernels that wrote their own block partials into G.d_partials with a grid of
// G.red_blocks blocks: combine blocks (fixed order), combine ranks (rank order), apply `postop` to
// the device-resident SolverScalars, optionally copy the sums to the host (synchronises).
int reduce_finish(int nfields, int postop, double* out_host);
// the same for block partials in `partials` ([nfields][nblocks][2]) written by a kernel with its own grid
int reduce_finish_n(int nfields, int postop, double* out_host, const double* partials, int nblocks);
int reduce_reserve_partials(int nblocks);  // sizes G.d_partials_big for POP_RED_NF fields x nblocks
// state (pop_state.cu)
int state_slab(int k, int kk, const double* T, const double* S, double* RHOOUT, double* RHOFULL,
               double* DRHODT, double* DRHODS, size_t n);
int state_3d(const double* TRACER, double* RHO);
// GM (pop_gm.cu)
// hooks of the overlapped coupled step (pop_abi.cu); no-ops outside pop_step_coupled
int coupled_wait_forcing();      // before the first reader of SMF / SHF_QSW / FW
int coupled_after_barotropic();  // PSURF(new) is final
int coupled_after_corrector();   // T,S(new) of the physical cells are final
int gm_alloc_fields();
int gm_begin_step();
int gm_tendency_dev(const double* TMIX);  // fills the field GM_HDT (nxb,nyb,km,nt), adds VDC_GM to VDC
// tracer path (pop_tracer.cu)
enum { TR_FULL = 0, TR_ADVT = 1, TR_HDIFFT = 2, TR_VDIFFT = 3 };
struct TracerIO {
  const double *TCUR, *TMIX, *TOLD, *UCUR, *VCUR, *STF, *TFW, *DH, *POLD, *PCUR;
  bool stf_strip = false;  // STF is a (nt, ny_local, nx_global) strip of physical cells instead of a padded field
  double* TNEW;  // FULL: (nxb,nyb,km,nt); slab modes: OUT(nxb,nyb,nt)
  double* WTK;   // slab modes: carried vertical velocity (in/out); FULL: unused
};
int tracer_column(int mode, int k, const TracerIO& io);
int impvmixt_dev(double* TNEW, const double* TOLD, const double* PSFC, const double* RHS,
                 int nfirst, int nlast, int correct);
int impvmixu_dev(double* UNEW, double* VNEW);
int vmix_coeffs_dev(int k0, int k1, const double* TMIX, const double* UMIX, const double* VMIX,
                    const double* RHOMIX);
// momentum path (pop_momentum.cu)
enum { MO_FULL = 0, MO_ADVU = 1, MO_HDIFFU = 2, MO_GRADP = 3, MO_VDIFFU = 4 };
struct MomentumIO {
  const double *UCUR, *VCUR, *UOLD, *VOLD, *UMIX, *VMIX, *RHOOLD, *RHOCUR, *RHONEW, *SMF, *DHU;
  double *UNEW, *VNEW, *ZX, *ZY;  // FULL outputs; slab modes: UNEW/VNEW = OUT1/OUT2 (nxb,nyb)
  double* WUK;                    // slab modes: carried (in/out)
};
int momentum_column(int mode, int k, const MomentumIO& io);
// UB/VB: also add the barotropic velocity (step_mod.F90:581-592) except on array row bt_skip_row (0-based; the tripole
// seam row, which the NE-corner halo update symmetrises non-linearly: the caller adds there after the halo update)
int momentum_finish(double* UNEW, double* VNEW, const double* UOLD, const double* VOLD, const double* UB = nullptr,
                    const double* VB = nullptr, int bt_skip_row = -1);
int grad_dev(int k, double* GX, double* GY, const double* F);
int div_dev(int k, double* D, const double* UX, const double* UY);
// barotropic + solver (pop_barotropic.cu)
int solvers_init_dev();
int solvers_prep_dev();
int solvers_evp_diagnostics(int* nsub, int* nland, double* selfcheck);
void evp_release();
void deep_release();
// lw_lim advection (pop_lwlim.cu)
void lw_release();
int lw_flux_prepare_dev(const double* U, const double* V, const double* DH);  // comp_flux_vel_ghost
int lw_lim_dev(const int* slots, const double* TMIX, int k0, int k1, double* out, size_t tstride, size_t lstride);
int solvers_diagonal_dev(const double* diagCorr);
int solvers_run_dev(double* X, const double* B);
int btrop_operator_dev(double* AX, const double* X);
int init_barotropic_dev();
int barotropic_driver_dev();
// drivers (pop_step.cu)
int dhdt_dev();
int baroclinic_driver_dev(bool defer_finish = false);  // defer_finish: the caller launches momentum_finish
int momentum_finish_new(bool add_barotropic);         // the deferred velocity finish on UVEL/VVEL(newtime)
int baroclinic_correct_adjust_dev();
int step_dev(int ts_type);
// grid (pop_grid.cu)
int set_grid_host(const double* ULAT, const double* HTN, const double* HTE, const double* HUS,
                  const double* HUW, const double* DXU, const double* DYU, const double* DXT,
                  const double* DYT, const int* KMT, const double* dz);
// staging of host pointers (pop_abi.cu)
bool is_device_ptr(const void* p);
