// pop_core.cu -- context, error stack, field registry, lifecycle, time-step scalars.
// Replaces (for this path) pop_init_phase1/2 bookkeeping of source/initial.F90:133,464 and the
// time-step switches of source/step_mod.F90:302-320 / source/time_management.F90:434-439.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "pop_dev.cuh"

Ctx G;
__constant__ VertConst c_vc;

static thread_local char g_err[1024] = "";
static char g_err_shared[1024] = "";

void pop_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  strncpy(g_err_shared, g_err, sizeof(g_err_shared) - 1);
}

extern "C" const char* pop_last_error(void) { return g_err[0] ? g_err : g_err_shared; }

int pop_post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pop_set_error("kernel launch failed (%s): %s", what, cudaGetErrorString(e));
    return POP_FAIL;
  }
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ timers
// CUDA-event timers under the reference's timer names (timers.F90).  Events are recorded on the
// library stream without synchronising; elapsed times are resolved lazily when a timer is read, so
// timers can stay enabled inside a timed region.
struct PendingTimer {
  const char* name;
  cudaEvent_t e0, e1;
};
static std::vector<PendingTimer> g_pending;
static void resolve_timers() {
  for (auto& t : g_pending) {
    cudaEventSynchronize(t.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.e0, t.e1);
    Timer& tm = G.timers[t.name];
    tm.ms += ms;
    tm.calls++;
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  g_pending.clear();
}
ScopedTimer::ScopedTimer(const char* n) : name(n), on(G.timers_on && G.timer_suppress == 0) {
  if (!on) return;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, G.stream);
}
ScopedTimer::~ScopedTimer() {
  if (!on) return;
  cudaEventRecord(e1, G.stream);
  g_pending.push_back(PendingTimer{name, e0, e1});
  if (g_pending.size() > 4096) resolve_timers();
}

// ------------------------------------------------------------------ registry
static const char* kPrognostic[] = {"TRACER", "UVEL",   "VVEL",   "RHO",   "PSURF",
                                    "GRADPX", "GRADPY", "UBTROP", "VBTROP"};

bool resolve_name(const char* name, int tlev, std::string* out) {
  for (const char* p : kPrognostic)
    if (!strcmp(name, p)) {
      int s = (tlev == POP_TIME_OLD) ? G.oldtime : (tlev == POP_TIME_NEW) ? G.newtime : G.curtime;
      *out = std::string(name) + "#" + std::to_string(s);
      return true;
    }
  *out = name;
  return G.fields.count(*out) > 0;
}

double* fld(const char* name) {
  auto it = G.fields.find(name);
  return (it == G.fields.end() || it->second.is_int) ? nullptr : (double*)it->second.p;
}
int* fldi(const char* name) {
  auto it = G.fields.find(name);
  return (it == G.fields.end() || !it->second.is_int) ? nullptr : (int*)it->second.p;
}
double* fld_t(const char* base, int storage) {
  return fld((std::string(base) + "#" + std::to_string(storage)).c_str());
}

int alloc_field(const char* name, int nz, bool is_int) {
  if (G.fields.count(name)) return POP_SUCCESS;
  DevField f;
  f.nz = nz;
  f.is_int = is_int;
  f.elems = G.n2 * (size_t)nz;
  size_t bytes = f.elems * (is_int ? sizeof(int) : sizeof(double));
  cudaError_t e = cudaMalloc(&f.p, bytes);
  if (e != cudaSuccess) {
    pop_set_error("cudaMalloc(%s, %zu bytes) failed: %s", name, bytes, cudaGetErrorString(e));
    return POP_FAIL;
  }
  POP_CHECK_CUDA(cudaMemsetAsync(f.p, 0, bytes, G.stream));
  G.fields[name] = f;
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ config defaults
extern "C" void pop_config_defaults(pop_config* c) {
  memset(c, 0, sizeof(*c));
  c->nt = 2;
  c->ew_boundary_type = POP_BNDY_CYCLIC;
  c->ns_boundary_type = POP_BNDY_CLOSED;
  for (int n = 0; n < POP_MAX_NT; n++) c->tadvect_itype[n] = POP_TADVECT_CENTERED;
  c->hmix_tracer_itype = POP_HMIX_DEL2;
  c->hmix_momentum_itype = POP_HMIX_DEL2;
  c->ah = 1.0e7;
  c->am = 1.0e7;
  c->ah_gm = c->ah_bolus = c->ah_bkg_srfbl = 0.8e7;
  c->slm_r = c->slm_b = 0.3;
  c->vmix_itype = POP_VMIX_CONST;
  c->implicit_vertical_mix = 1;
  c->vdc_ndim = 1;
  c->aidif = 1.0;
  c->bottom_drag = 1.0e-3;
  c->const_vdc = 0.25;
  c->const_vvc = 0.25;
  c->bckgrnd_vdc = 0.1;
  c->bckgrnd_vvc = 1.0;
  c->rich_mix = 50.0;
  c->convection_diff = 1;
  c->convect_diff = 1000.0;
  c->convect_visc = 1000.0;
  c->sfc_layer_type = POP_SFC_VARTHICK;
  c->lpressure_avg = 1;
  c->lbouss_correct = 1;
  c->impcor = 1;
  c->state_itype = POP_STATE_MWJF;
  c->state_range_iopt = POP_STATE_RANGE_ENFORCE;
  c->solver_choice = POP_SOLVER_CHRONGEAR;
  c->max_iterations = 1000;
  c->convergence_check_freq = 10;
  c->convergence_check_start = 60;
  c->max_lanczos_step = 20;
  c->convergence_criterion = 1.0e-13;
  c->lanczos_convergence_criterion = 0.1;
  c->dtt = 3600.0;
  c->nranks = 1;
  c->robert_alpha = 0.53;  // time_management.F90:461-462 (Williams 2009)
  c->robert_nu = 0.20;
  c->nconvad = 0;
  c->preconditioner_choice = POP_PRECOND_DIAGONAL;
}

// ------------------------------------------------------------------ lifecycle
extern "C" int pop_is_initialized(void) { return G.initialized ? 1 : 0; }

static int alloc_all_fields() {
  const pop_config& c = G.cfg;
  const int km = G.km, nt = G.nt;
  for (int s = 0; s < 3; s++) {
    std::string sfx = "#" + std::to_string(s);
    POP_TRY(alloc_field(("TRACER" + sfx).c_str(), km * nt, false));
    POP_TRY(alloc_field(("UVEL" + sfx).c_str(), km, false));
    POP_TRY(alloc_field(("VVEL" + sfx).c_str(), km, false));
    POP_TRY(alloc_field(("RHO" + sfx).c_str(), km, false));
    for (const char* n2d : {"PSURF", "GRADPX", "GRADPY", "UBTROP", "VBTROP"})
      POP_TRY(alloc_field((std::string(n2d) + sfx).c_str(), 1, false));
  }
  POP_TRY(alloc_field("PGUESS", 1, false));
  POP_TRY(alloc_field("STF", nt, false));
  POP_TRY(alloc_field("TFW", nt, false));
  POP_TRY(alloc_field("SMF", 2, false));
  for (const char* n2d : {"SHF_QSW", "FW", "FW_OLD", "DH", "DHU", "ZX", "ZY", "VUF", "VVF", "SUMX",
                          "SUMY", "RHOKMX", "RHOKMY", "WTK_C", "WUK_C", "RHS_BT", "UH_BT", "VH_BT",
                          "DIAGC", "W2A", "W2B", "W2C", "W2D"})
    POP_TRY(alloc_field(n2d, 1, false));
  POP_TRY(alloc_field("VTF", nt, false));
  POP_TRY(alloc_field("AUX", nt, false));
  POP_TRY(alloc_field("RHS1", nt, false));
  // vertical_mix.F90:383-419: VDC/VVC shapes
  if (c.vmix_itype == POP_VMIX_GIVEN) {
    G.vdc_nk = c.vdc_kdim_halo ? km + 2 : km;
    G.vdc_k0 = c.vdc_kdim_halo ? 0 : 1;
    G.vdc_nd = c.vdc_ndim > 0 ? c.vdc_ndim : 1;
    G.vvc_nk = km;
  } else if (c.vmix_itype == POP_VMIX_CONST) {
    G.vdc_nk = c.implicit_vertical_mix ? km : 1;
    G.vdc_k0 = 1;
    G.vdc_nd = 1;
    G.vvc_nk = c.implicit_vertical_mix ? km : 1;
  } else {
    G.vdc_nk = c.implicit_vertical_mix ? km : 1;
    G.vdc_k0 = 1;
    G.vdc_nd = 1;
    G.vvc_nk = km;
  }
  POP_TRY(alloc_field("VDC", G.vdc_nk * G.vdc_nd, false));
  POP_TRY(alloc_field("VVC", G.vvc_nk, false));
  // work: Thomas E coefficients; momentum Thomas E
  POP_TRY(alloc_field("WORK3D_E", km, false));
  if (c.hmix_tracer_itype == POP_HMIX_GM) POP_TRY(gm_alloc_fields());
  return POP_SUCCESS;
}

extern "C" int pop_init(const pop_config* cfg) {
  if (G.initialized) pop_finalize();
  POP_REQUIRE(cfg != nullptr, "pop_init: null config");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  POP_REQUIRE(e == cudaSuccess && ndev > 0,
              "pop_init: no usable CUDA device (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  POP_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "pop_init: device %d out of range (%d)",
              cfg->device, ndev);
  POP_CHECK_CUDA(cudaSetDevice(cfg->device));
  G.cfg = *cfg;
  G.nxg = cfg->nx_global;
  G.nyg = cfg->ny_global;
  G.km = cfg->km;
  G.nt = cfg->nt;
  G.rank = cfg->rank;
  G.nranks = cfg->nranks > 0 ? cfg->nranks : 1;
  POP_REQUIRE(G.nxg > 0 && G.nyg > 0 && G.km >= 3 && G.km <= POP_KMAX, "pop_init: bad sizes %d x %d x %d",
              G.nxg, G.nyg, G.km);
  POP_REQUIRE(G.nt >= 2 && G.nt <= POP_MAX_NT, "pop_init: nt=%d out of range", G.nt);
  POP_REQUIRE(G.nyg % G.nranks == 0, "pop_init: ny_global=%d not divisible by nranks=%d", G.nyg,
              G.nranks);
  if (cfg->partial_bottom_cells) {
    // partial bottom cells are built for the option set of the tx0.1v3 production configuration and its neighbours:
    // centred or lw_lim advection, del2 / del4 mixing, const / given vertical mixing coefficients
    for (int n = 0; n < cfg->nt; n++)
      POP_REQUIRE(cfg->tadvect_itype[n] != POP_TADVECT_UPWIND3,
                  "pop_init: partial_bottom_cells needs centred or lw_lim advection (tracer %d uses upwind3, whose "
                  "vertical interpolation weights become 3-d with partial cells: not built)", n + 1);
    POP_REQUIRE(cfg->hmix_tracer_itype != POP_HMIX_GM, "pop_init: partial_bottom_cells with GM is not built");
    POP_REQUIRE(cfg->vmix_itype != POP_VMIX_RICH, "pop_init: partial_bottom_cells with vmix 'rich' is not built");
  }
  POP_REQUIRE(G.nranks == 1 || G.nccl_comm != nullptr,
              "pop_init: nranks=%d but pop_comm_init was not called", G.nranks);
  G.ny_local = G.nyg / G.nranks;
  POP_REQUIRE(G.ny_local >= 2 * POP_NGHOST, "pop_init: strip of %d rows is thinner than the halo",
              G.ny_local);
  G.j0 = G.rank * G.ny_local + 1;
  G.nxb = G.nxg + 2 * POP_NGHOST;
  G.nyb = G.ny_local + 2 * POP_NGHOST;
  G.ib = POP_NGHOST + 1;
  G.ie = G.nxb - POP_NGHOST;
  G.jb = POP_NGHOST + 1;
  G.je = G.nyb - POP_NGHOST;
  G.n2 = (size_t)G.nxb * G.nyb;
  G.n3 = G.n2 * G.km;
  // i_glob / j_glob: source/blocks.F90:166-268 for one block spanning the strip
  G.i_glob.assign(G.nxb, 0);
  G.j_glob.assign(G.nyb, 0);
  for (int i = 1; i <= G.nxb; i++) {
    int v = i - POP_NGHOST;
    if (v < 1) v = (cfg->ew_boundary_type == POP_BNDY_CYCLIC) ? v + G.nxg : 0;
    else if (v > G.nxg) v = (cfg->ew_boundary_type == POP_BNDY_CYCLIC) ? v - G.nxg : 0;
    G.i_glob[i - 1] = v;
  }
  for (int j = 1; j <= G.nyb; j++) {
    int v = G.j0 - POP_NGHOST + j - 1;
    if (v < 1) v = (cfg->ns_boundary_type == POP_BNDY_CYCLIC) ? v + G.nyg : 0;
    else if (v > G.nyg) {
      if (cfg->ns_boundary_type == POP_BNDY_CYCLIC) v -= G.nyg;
      else if (cfg->ns_boundary_type == POP_BNDY_TRIPOLE) v = -v;
      else v = 0;
    }
    G.j_glob[j - 1] = v;
  }
  cudaDeviceProp prop;
  POP_CHECK_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  G.sm_count = prop.multiProcessorCount;
  int prio_least = 0, prio_greatest = 0;
  POP_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  // with the opt-in exchange overlap: exchange stream > main stream > side streams (lower number = higher priority)
  const bool want_x = getenv("POP_B200_OVERLAP_EXCHANGE") != nullptr && getenv("POP_B200_OVERLAP_EXCHANGE")[0] == '1';
  // (the strip exchange of the solver's deep strips and the convergence-check reductions run on streams of their own that
  // must get their few CTAs scheduled ahead of the pass kernels: main stream one level below the greatest priority)
  (void)want_x;
  const int prio_main = (prio_least - prio_greatest >= 2) ? prio_greatest + 1 : prio_greatest;
  if (!G.stream) POP_CHECK_CUDA(cudaStreamCreateWithPriority(&G.stream, cudaStreamNonBlocking, prio_main));
  if (!G.stream2) {
    // low priority: its CTAs take the slots the main stream's kernels leave free (tails, launch gaps)
    POP_CHECK_CUDA(cudaStreamCreateWithPriority(&G.stream2, cudaStreamNonBlocking, prio_least));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_fork, cudaEventDisableTiming));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_join, cudaEventDisableTiming));
  }
  if (!G.stream_x) {
    POP_CHECK_CUDA(cudaStreamCreateWithPriority(&G.stream_x, cudaStreamNonBlocking, prio_greatest));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_xb, cudaEventDisableTiming));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_xx, cudaEventDisableTiming));
  }
  G.overlap_exchange = getenv("POP_B200_OVERLAP_EXCHANGE") != nullptr && getenv("POP_B200_OVERLAP_EXCHANGE")[0] == '1';
  if (!G.stream_cp) {
    POP_CHECK_CUDA(cudaStreamCreateWithPriority(&G.stream_cp, cudaStreamNonBlocking, prio_least));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_cp_in, cudaEventDisableTiming));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_cp_a, cudaEventDisableTiming));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_cp_b, cudaEventDisableTiming));
  }
  G.no_overlap = getenv("POP_B200_NO_OVERLAP") != nullptr && getenv("POP_B200_NO_OVERLAP")[0] == '1';
  G.finish_mode = G.no_overlap ? 2 : 0;
  if (!G.no_overlap && getenv("POP_B200_OVERLAP_FINISH") != nullptr && getenv("POP_B200_OVERLAP_FINISH")[0] == '1')
    G.finish_mode = 1;
  // programmatic dependent launch pays when the per-rank kernels are short (A/B in profiles/r1_ncu_summary.md: -7 %
  // solver time on a 3600x300 strip, +1 % on the full 3600x2400 grid where the early CTAs compete with the
  // overlapped velocity-finish kernel): on for strips up to 2.5 M points, POP_B200_PDL=1 / POP_B200_NO_PDL=1 override
  // (round 2: with the velocity finish after the solve nothing competes any more, and the full grid gains 3 % too)
  G.no_pdl = (G.finish_mode == 1) && (size_t)G.nxg * (size_t)(G.nyg / G.nranks) > 2500000;
  G.deep_halo = 12;
  if (const char* e = getenv("POP_B200_DEEP_HALO")) G.deep_halo = atoi(e);
  if (getenv("POP_B200_NO_DEEP_HALO") != nullptr && getenv("POP_B200_NO_DEEP_HALO")[0] == '1') G.deep_halo = 0;
  G.deep_force = getenv("POP_B200_DEEP_HALO_FORCE") != nullptr && getenv("POP_B200_DEEP_HALO_FORCE")[0] == '1';
  if (getenv("POP_B200_PDL") != nullptr && getenv("POP_B200_PDL")[0] == '1') G.no_pdl = false;
  if (getenv("POP_B200_NO_PDL") != nullptr && getenv("POP_B200_NO_PDL")[0] == '1') G.no_pdl = true;
  POP_REQUIRE(cfg->ns_boundary_type != POP_BNDY_TRIPOLE || cfg->ew_boundary_type == POP_BNDY_CYCLIC,
              "pop_init: a tripole grid needs a cyclic east-west boundary");
  POP_CHECK_CUDA(cudaMalloc(&G.d_iglob, sizeof(int) * G.nxb));
  POP_CHECK_CUDA(cudaMalloc(&G.d_jglob, sizeof(int) * G.nyb));
  POP_CHECK_CUDA(cudaMemcpy(G.d_iglob, G.i_glob.data(), sizeof(int) * G.nxb, cudaMemcpyHostToDevice));
  POP_CHECK_CUDA(cudaMemcpy(G.d_jglob, G.j_glob.data(), sizeof(int) * G.nyb, cudaMemcpyHostToDevice));
  G.oldtime = 0;
  G.curtime = 1;
  G.newtime = 2;
  G.mixtime = 0;
  G.use_upwind3 = G.use_centered = G.use_lw_lim = false;
  G.lw_flux_ready = false;
  G.lw_coef_dirty = true;
  for (int n = 0; n < G.nt; n++) {
    if (cfg->tadvect_itype[n] == POP_TADVECT_UPWIND3) G.use_upwind3 = true;
    else if (cfg->tadvect_itype[n] == POP_TADVECT_CENTERED) G.use_centered = true;
    else if (cfg->tadvect_itype[n] == POP_TADVECT_LW_LIM) G.use_lw_lim = true;
    else POP_REQUIRE(false, "pop_init: tadvect_itype[%d]=%d not supported", n, cfg->tadvect_itype[n]);
  }
  POP_REQUIRE(cfg->hmix_tracer_itype == POP_HMIX_DEL2 || cfg->hmix_tracer_itype == POP_HMIX_DEL4 ||
                  cfg->hmix_tracer_itype == POP_HMIX_GM,
              "pop_init: hmix_tracer_itype=%d not supported (del2, del4, gm)", cfg->hmix_tracer_itype);
  POP_REQUIRE(cfg->hmix_tracer_itype != POP_HMIX_GM || cfg->implicit_vertical_mix,
              "pop_init: implicit vertical mixing must be used with GM horiz mixing (hmix_gm.F90:1192-1194)");
  POP_REQUIRE(cfg->hmix_momentum_itype == POP_HMIX_DEL2 || cfg->hmix_momentum_itype == POP_HMIX_DEL4,
              "pop_init: hmix_momentum_itype=%d not supported", cfg->hmix_momentum_itype);
  POP_REQUIRE(cfg->vmix_itype == POP_VMIX_CONST || cfg->vmix_itype == POP_VMIX_GIVEN ||
                  cfg->vmix_itype == POP_VMIX_RICH,
              "pop_init: vmix_itype=%d is not implemented (const, rich, given)", cfg->vmix_itype);
  POP_REQUIRE(cfg->preconditioner_choice == POP_PRECOND_DIAGONAL || cfg->preconditioner_choice == POP_PRECOND_EVP,
              "pop_init: unknown preconditioner choice %d (diagonal, evp)", cfg->preconditioner_choice);
  POP_REQUIRE(cfg->vmix_itype != POP_VMIX_RICH || cfg->implicit_vertical_mix,
              "pop_init: vmix_itype=rich needs implicit_vertical_mix (the coefficients of all levels are built before the column kernels)");
  POP_TRY(alloc_all_fields());
  POP_TRY(reduce_alloc());
  POP_TRY(p2p_setup());
  // time constants: time_management.F90:962-964,434-439
  G.dtt = G.dtu = G.dtp = cfg->dtt;
  G.alpha = 1.0 / 3.0;
  G.theta = 0.5;
  G.gamma = 1.0 - 2.0 * G.alpha;
  memset(&G.vc, 0, sizeof(G.vc));
  G.launches = 0;
  G.no_tma = getenv("POP_B200_NO_TMA") != nullptr && getenv("POP_B200_NO_TMA")[0] == '1';
  G.thomas_tma = getenv("POP_B200_THOMAS_TMA") != nullptr && getenv("POP_B200_THOMAS_TMA")[0] == '1';
  G.no_fast_tracer = getenv("POP_B200_NO_FAST_TRACER") != nullptr && getenv("POP_B200_NO_FAST_TRACER")[0] == '1';
  G.no_pbc_fast = getenv("POP_B200_NO_PBC_FAST") != nullptr && getenv("POP_B200_NO_PBC_FAST")[0] == '1';
  G.no_pcsi_blocking = getenv("POP_B200_NO_PCSI_BLOCKING") != nullptr && getenv("POP_B200_NO_PCSI_BLOCKING")[0] == '1';
  G.timers.clear();
  G.grid_set = false;
  G.initialized = true;
  return POP_SUCCESS;
}

extern "C" int pop_finalize(void) {
  if (G.stream) cudaStreamSynchronize(G.stream);
  if (G.stream2) cudaStreamSynchronize(G.stream2);
  if (G.stream_cp) cudaStreamSynchronize(G.stream_cp);
  if (G.stream_x) cudaStreamSynchronize(G.stream_x);
  p2p_teardown();
  evp_release();
  deep_release();
  lw_release();
  for (auto& kv : G.fields) cudaFree(kv.second.p);
  G.fields.clear();
  for (auto& kv : G.stage) cudaFree(kv.second.first);
  G.stage.clear();
  cudaFree(G.d_partials);
  cudaFree(G.d_partials_big);
  G.d_partials_big = nullptr;
  G.partials_big_n = 0;
  cudaFree(G.d_sums);
  cudaFree(G.d_gather);
  cudaFree(G.d_local);
  cudaFree(G.d_scal);
  cudaFree(G.d_tripole);
  cudaFree(G.d_iglob);
  cudaFree(G.d_jglob);
  if (G.h_sums) cudaFreeHost(G.h_sums);
  G.d_partials = G.d_sums = G.d_gather = G.d_local = G.d_tripole = nullptr;
  G.d_scal = nullptr;
  G.d_iglob = G.d_jglob = nullptr;
  G.tripole_elems = 0;
  G.h_sums = nullptr;
  cudaFree(G.d_sendS);
  cudaFree(G.d_sendN);
  cudaFree(G.d_recvS);
  cudaFree(G.d_recvN);
  G.d_sendS = G.d_sendN = G.d_recvS = G.d_recvN = nullptr;
  G.halo_buf_elems = 0;
  G.initialized = false;
  G.grid_set = false;
  return POP_SUCCESS;
}

extern "C" int pop_get_block(pop_block* blk) {
  POP_REQUIRE(G.initialized, "pop_get_block: not initialized");
  blk->block_id = G.rank + 1;
  blk->local_id = 1;
  blk->ib = G.ib;
  blk->ie = G.ie;
  blk->jb = G.jb;
  blk->je = G.je;
  blk->iblock = 1;
  blk->jblock = G.rank + 1;
  blk->i_glob = G.i_glob.data();
  blk->j_glob = G.j_glob.data();
  return POP_SUCCESS;
}

extern "C" int pop_local_shape(int* nx_block, int* ny_block, int* j_start_global, int* ny_local) {
  POP_REQUIRE(G.initialized, "pop_local_shape: not initialized");
  if (nx_block) *nx_block = G.nxb;
  if (ny_block) *ny_block = G.nyb;
  if (j_start_global) *j_start_global = G.j0;
  if (ny_local) *ny_local = G.ny_local;
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ time step scalars
int upload_vert_const() {
  POP_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_vc, &G.vc, sizeof(VertConst), 0, cudaMemcpyHostToDevice,
                                         G.stream));
  // the host copy may be modified right after this call: make the copy complete first
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}

// step_mod.F90:302-320
int set_timestep(int ts_type) {
  POP_REQUIRE(G.grid_set, "pop_set_timestep: grid not set");
  bool leap = (ts_type == POP_TS_LEAPFROG || ts_type == POP_TS_AVG || ts_type == POP_TS_ROBERT);
  bool changed = (leap != G.leapfrogts) || G.vc.c2dtt[1] == 0.0;
  G.leapfrogts = leap;
  G.f_euler_ts = (ts_type == POP_TS_EULER);
  G.avg_ts = (ts_type == POP_TS_AVG);
  if (leap) {
    G.mixtime = G.oldtime;
    G.beta = G.alpha;
    for (int k = 1; k <= G.km; k++) G.vc.c2dtt[k] = 2.0 * G.dtt;
    G.c2dtu = 2.0 * G.dtu;
    G.c2dtp = 2.0 * G.dtp;
  } else {
    G.mixtime = G.curtime;
    G.beta = G.theta;
    for (int k = 1; k <= G.km; k++) G.vc.c2dtt[k] = G.dtt;
    G.c2dtu = G.dtu;
    G.c2dtp = G.dtp;
  }
  for (int k = 1; k <= G.km; k++) {
    G.vc.hfac_t[k] = G.vc.dz[k] / G.vc.c2dtt[k];
    G.vc.hfac_u[k] = G.vc.dz[k] / G.c2dtu;
  }
  if (changed) POP_TRY(upload_vert_const());
  return POP_SUCCESS;
}

extern "C" int pop_set_timestep(int ts_type) {
  POP_REQUIRE(G.initialized, "pop_set_timestep: not initialized");
  return set_timestep(ts_type);
}

GridView grid_view() {
  GridView g;
  memset(&g, 0, sizeof(g));
  g.nxb = G.nxb; g.nyb = G.nyb; g.km = G.km; g.nt = G.nt;
  g.ib = G.ib; g.ie = G.ie; g.jb = G.jb; g.je = G.je;
  g.n2 = G.n2; g.n3 = G.n3;
  g.KMT = fldi("KMT"); g.KMU = fldi("KMU");
#define GV(n) g.n = fld(#n)
  GV(DXU); GV(DYU); GV(DXUR); GV(DYUR); GV(UAREA_R); GV(TAREA_R); GV(TAREA); GV(HUR); GV(HU);
  GV(FCOR); GV(KXU); GV(KYU); GV(DTN); GV(DTS); GV(DTE); GV(DTW); GV(AHF);
  GV(DUC); GV(DUN); GV(DUS); GV(DUE); GV(DUW); GV(DMC); GV(DMN); GV(DMS); GV(DME); GV(DMW);
  GV(DUM); GV(AMF); GV(AU0); GV(AUN); GV(AUE); GV(AUNE); GV(RCALCT); GV(RCALCU);
  GV(TALFXP); GV(TBETXP); GV(TGAMXP); GV(TALFYP); GV(TBETYP); GV(TGAMYP);
  GV(TALFXM); GV(TBETXM); GV(TDELXM); GV(TALFYM); GV(TBETYM); GV(TDELYM);
  GV(VDC); GV(VVC);
#undef GV
  if (G.cfg.partial_bottom_cells) { g.DZT = fld("DZT"); g.DZU = fld("DZU"); }
  g.aidif = G.cfg.aidif;
  g.vdc_nk = G.vdc_nk; g.vdc_k0 = G.vdc_k0; g.vdc_nd = G.vdc_nd; g.vvc_nk = G.vvc_nk;
  return g;
}

// ------------------------------------------------------------------ instrumentation
extern "C" long pop_kernel_launch_count(void) { return G.launches; }
extern "C" int pop_timer_get(const char* name, double* ms, long* calls) {
  resolve_timers();
  auto it = G.timers.find(name);
  if (it == G.timers.end()) {
    if (ms) *ms = 0.0;
    if (calls) *calls = 0;
    return POP_FAIL;
  }
  if (ms) *ms = it->second.ms;
  if (calls) *calls = it->second.calls;
  return POP_SUCCESS;
}
extern "C" int pop_timers_reset(void) {
  resolve_timers();
  G.timers.clear();
  return POP_SUCCESS;
}
extern "C" int pop_timers_enable(int on) {
  G.timers_on = on != 0;
  return POP_SUCCESS;
}
extern "C" int pop_sync(void) {
  POP_REQUIRE(G.stream != nullptr, "pop_sync: not initialized");
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
extern "C" void* pop_stream(void) { return (void*)G.stream; }

// ------------------------------------------------------------------ TMA tensor maps
#ifndef POP_EMUL
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}
bool make_tmap(PopTmap* out, const double* field, int nlev) {
  EncodeTiledFn enc = get_encode();
  if (!enc || !field || (G.nxb % 2) != 0 || G.cfg.device < 0) return false;
  cuuint64_t dims[3] = {(cuuint64_t)G.nxb, (cuuint64_t)G.nyb, (cuuint64_t)nlev};
  cuuint64_t strides[2] = {(cuuint64_t)G.nxb * 8, (cuuint64_t)G.n2 * 8};
  cuuint32_t box[3] = {POP_TW, POP_TH, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&out->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)field, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
bool make_tmap_box(PopTmap* out, const double* field, int nlev, int boxw, int boxh) {
  EncodeTiledFn enc = get_encode();
  if (!enc || !field || (G.nxb % 2) != 0 || (boxw % 2) != 0 || boxw > 256 || boxh > 256) return false;
  cuuint64_t dims[3] = {(cuuint64_t)G.nxb, (cuuint64_t)G.nyb, (cuuint64_t)nlev};
  cuuint64_t strides[2] = {(cuuint64_t)G.nxb * 8, (cuuint64_t)G.n2 * 8};
  cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&out->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)field, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
bool make_tmap_2d(PopTmap* out, const double* field, int boxw, int boxh, int nrows) {
  // encoded as a one-level 3-d tensor (the form the column kernels use): box boxw x boxh x 1
  EncodeTiledFn enc = get_encode();
  if (!enc || !field || (G.nxb % 2) != 0 || (boxw % 2) != 0 || boxw > 256 || boxh > 256) return false;
  cuuint64_t dims[3] = {(cuuint64_t)G.nxb, (cuuint64_t)(nrows > 0 ? nrows : G.nyb), 1};
  cuuint64_t strides[2] = {(cuuint64_t)G.nxb * 8, (cuuint64_t)G.n2 * 8};
  cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&out->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)field, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
#else
bool make_tmap(PopTmap* out, const double* field, int nlev) {
  if (!field || (G.nxb % 2) != 0) return false;
  out->p = field; out->nx = G.nxb; out->ny = G.nyb; out->nz = nlev;
  out->bw = POP_TW; out->bh = POP_TH;
  return true;
}
bool make_tmap_box(PopTmap* out, const double* field, int nlev, int boxw, int boxh) {
  if (!field || (G.nxb % 2) != 0 || (boxw % 2) != 0) return false;
  out->p = field; out->nx = G.nxb; out->ny = G.nyb; out->nz = nlev;
  out->bw = boxw; out->bh = boxh;
  return true;
}
bool make_tmap_2d(PopTmap* out, const double* field, int boxw, int boxh, int nrows) {
  if (!field || (G.nxb % 2) != 0 || (boxw % 2) != 0) return false;
  out->p = field; out->nx = G.nxb; out->ny = nrows > 0 ? nrows : G.nyb; out->nz = 1;
  out->bw = boxw; out->bh = boxh;
  return true;
}
#endif
