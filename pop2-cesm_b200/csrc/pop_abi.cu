// pop_abi.cu -- the C ABI (include/pop_b200.h): reference-signature slab operators, field access,
// communication entry points and the fused drivers.  Pointer arguments of the slab operators may be
// device pointers (zero-copy) or host pointers; host arrays are staged through HBM inside the call
// so that a Fortran caller can hand over its module arrays unchanged (SURVEY 8b "Ownership").
#include <cstring>
#include <string>
#include <vector>
#include "pop_dev.cuh"

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

namespace {
// stages host arguments of one call; device arguments pass through untouched
struct Stage {
  struct Out {
    void* host;
    void* dev;
    size_t bytes;
  };
  std::vector<Out> outs;
  bool ok = true;
  void* buf(const char* key, size_t bytes) {
    auto& s = G.stage[key];
    if (s.second < bytes) {
      cudaFree(s.first);
      s.first = nullptr;
      s.second = 0;
      if (cudaMalloc(&s.first, bytes) != cudaSuccess) {
        pop_set_error("staging buffer %s: cudaMalloc(%zu) failed", key, bytes);
        ok = false;
        return nullptr;
      }
      s.second = bytes;
    }
    return s.first;
  }
  template <typename T>
  const T* in(const char* key, const T* p, size_t n) {
    if (!p || is_device_ptr(p)) return p;
    void* d = buf(key, n * sizeof(T));
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, G.stream) != cudaSuccess) ok = false;
    return (const T*)d;
  }
  template <typename T>
  T* inout(const char* key, T* p, size_t n, bool copy_in) {
    if (!p || is_device_ptr(p)) return p;
    void* d = buf(key, n * sizeof(T));
    if (!d) return nullptr;
    if (copy_in) {
      if (cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, G.stream) != cudaSuccess) ok = false;
    } else {
      cudaMemsetAsync(d, 0, n * sizeof(T), G.stream);
    }
    outs.push_back(Out{(void*)p, d, n * sizeof(T)});
    return (T*)d;
  }
  int finish(int rc) {
    if (!ok && rc == POP_SUCCESS) {
      if (!*pop_last_error()) pop_set_error("host<->device staging failed");
      rc = POP_FAIL;
    }
    if (rc == POP_SUCCESS)
      for (auto& o : outs)
        if (cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, G.stream) != cudaSuccess) {
          pop_set_error("staging copy-back failed");
          rc = POP_FAIL;
        }
    if (!outs.empty() || rc != POP_SUCCESS) cudaStreamSynchronize(G.stream);
    // multi-rank: a peer-memory exchange that timed out inside this call must surface here, not at the next pop_step
    if (rc == POP_SUCCESS && G.p2p_on) rc = p2p_check();
    return rc;
  }
};
// the drivers that exchange halos without going through Stage
static int with_p2p_check(int rc) { return (rc == POP_SUCCESS && G.p2p_on) ? p2p_check() : rc; }

int check_ready(const char* what, const pop_block* blk, bool need_blk) {
  POP_REQUIRE(G.initialized, "%s: pop_init has not been called", what);
  POP_REQUIRE(G.grid_set, "%s: pop_set_grid has not been called", what);
  if (need_blk) {
    POP_REQUIRE(blk != nullptr, "%s: null this_block", what);
    POP_REQUIRE(blk->local_id == 1, "%s: block local_id=%d, this rank owns exactly one block", what, blk->local_id);
  }
  return POP_SUCCESS;
}
}  // namespace

// ------------------------------------------------------------------ lifecycle extras
extern "C" int pop_comm_unique_id(char id128[128]) { return comm_unique_id(id128); }
extern "C" int pop_comm_init(int rank, int nranks, const char id128[128]) { return comm_init(rank, nranks, id128); }

extern "C" int pop_set_bottom_cells(const double* DZBC) {
  POP_REQUIRE(G.initialized, "pop_set_bottom_cells: pop_init has not been called");
  POP_REQUIRE(G.cfg.partial_bottom_cells, "pop_set_bottom_cells: partial_bottom_cells is off in pop_config");
  POP_REQUIRE(DZBC != nullptr, "pop_set_bottom_cells: null argument");
  G.dzbc_strip.assign(DZBC, DZBC + (size_t)G.nxg * G.ny_local);
  return POP_SUCCESS;
}
extern "C" int pop_set_grid(const double* ULAT, const double* HTN, const double* HTE, const double* HUS,
                            const double* HUW, const double* DXU, const double* DYU, const double* DXT,
                            const double* DYT, const int* KMT, const double* dz) {
  POP_REQUIRE(G.initialized, "pop_set_grid: pop_init has not been called");
  POP_REQUIRE(ULAT && HTN && HTE && HUS && HUW && DXU && DYU && DXT && DYT && KMT && dz, "pop_set_grid: null argument");
  return with_p2p_check(set_grid_host(ULAT, HTN, HTE, HUS, HUW, DXU, DYU, DXT, DYT, KMT, dz));
}

// ------------------------------------------------------------------ field access
extern "C" void* pop_device_ptr(const char* name, int tlev) {
  std::string key;
  if (!G.initialized || !name || !resolve_name(name, tlev, &key)) return nullptr;
  auto it = G.fields.find(key);
  return it == G.fields.end() ? nullptr : it->second.p;
}
extern "C" long pop_field_size(const char* name) {
  std::string key;
  if (!G.initialized || !name || !resolve_name(name, POP_TIME_CUR, &key)) return -1;
  auto it = G.fields.find(key);
  return it == G.fields.end() ? -1 : (long)it->second.elems;
}
static int find_field(const char* what, const char* name, int tlev, DevField** f) {
  POP_REQUIRE(G.initialized, "%s: not initialized", what);
  std::string key;
  POP_REQUIRE(name && resolve_name(name, tlev, &key) && G.fields.count(key), "%s: unknown field '%s'", what, name ? name : "(null)");
  *f = &G.fields[key];
  return POP_SUCCESS;
}
// caller-supplied inputs that GM caches (pop_gm.cu)
static void note_field_written(const char* name) {
  if (!strcmp(name, "TLAT")) G.gm_dirty = true;
  if (!strcmp(name, "VDC")) G.gm_vdc_dirty = true;
}
extern "C" int pop_set_field(const char* name, int tlev, const void* host) {
  DevField* f;
  POP_TRY(find_field("pop_set_field", name, tlev, &f));
  note_field_written(name);
  POP_REQUIRE(host != nullptr, "pop_set_field: null host pointer");
  POP_CHECK_CUDA(cudaMemcpyAsync(f->p, host, f->elems * (f->is_int ? sizeof(int) : sizeof(double)), cudaMemcpyHostToDevice, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
extern "C" int pop_get_field(const char* name, int tlev, void* host) {
  DevField* f;
  POP_TRY(find_field("pop_get_field", name, tlev, &f));
  POP_REQUIRE(host != nullptr, "pop_get_field: null host pointer");
  POP_CHECK_CUDA(cudaMemcpyAsync(host, f->p, f->elems * (f->is_int ? sizeof(int) : sizeof(double)), cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
// physical strip (nx_global x ny_local x nz) <-> padded block; ghost cells are not touched
static int strip_copy(DevField* f, void* host, bool to_device, int z0, int nz, cudaStream_t st = nullptr) {
  const size_t es = f->is_int ? sizeof(int) : sizeof(double);
  if (!st) st = G.stream;
  for (int z = z0; z < z0 + nz; z++) {
    char* d = (char*)f->p + ((size_t)z * G.n2 + (size_t)POP_NGHOST * G.nxb + POP_NGHOST) * es;
    char* h = (char*)host + (size_t)(z - z0) * G.nxg * G.ny_local * es;
    // cudaMemcpyDefault: the strip may live in host (pageable / pinned) or device memory
    if (to_device)
      POP_CHECK_CUDA(cudaMemcpy2DAsync(d, G.nxb * es, h, G.nxg * es, G.nxg * es, G.ny_local, cudaMemcpyDefault, st));
    else
      POP_CHECK_CUDA(cudaMemcpy2DAsync(h, G.nxg * es, d, G.nxb * es, G.nxg * es, G.ny_local, cudaMemcpyDefault, st));
  }
  return POP_SUCCESS;
}
extern "C" int pop_scatter_field(const char* name, int tlev, const void* host_strip) {
  DevField* f;
  POP_TRY(find_field("pop_scatter_field", name, tlev, &f));
  note_field_written(name);
  POP_REQUIRE(host_strip != nullptr, "pop_scatter_field: null host pointer");
  POP_TRY(strip_copy(f, (void*)host_strip, true, 0, f->nz));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
extern "C" int pop_scatter_field_levels(const char* name, int tlev, int z0, int nz, const void* strip) {
  DevField* f;
  POP_TRY(find_field("pop_scatter_field_levels", name, tlev, &f));
  note_field_written(name);
  POP_REQUIRE(strip != nullptr && z0 >= 0 && nz >= 1 && z0 + nz <= f->nz,
              "pop_scatter_field_levels: bad level range %d..%d of %d", z0, z0 + nz - 1, f->nz);
  POP_TRY(strip_copy(f, (void*)strip, true, z0, nz));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
extern "C" int pop_gather_field(const char* name, int tlev, void* host_strip) {
  DevField* f;
  POP_TRY(find_field("pop_gather_field", name, tlev, &f));
  POP_REQUIRE(host_strip != nullptr, "pop_gather_field: null host pointer");
  POP_TRY(strip_copy(f, host_strip, false, 0, f->nz));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ slab operators
extern "C" int pop_state(int k, int kk, const double* TEMPK, const double* SALTK, const pop_block* blk,
                         double* RHOOUT, double* RHOFULL, double* DRHODT, double* DRHODS) {
  POP_TRY(check_ready("state", blk, true));
  POP_REQUIRE(TEMPK && SALTK, "state: null TEMPK/SALTK");
  Stage s;
  const size_t n = G.n2;
  const double* T = s.in("st_T", TEMPK, n);
  const double* S = s.in("st_S", SALTK, n);
  double* r0 = s.inout("st_r0", RHOOUT, n, false);
  double* r1 = s.inout("st_r1", RHOFULL, n, false);
  double* r2 = s.inout("st_r2", DRHODT, n, false);
  double* r3 = s.inout("st_r3", DRHODS, n, false);
  return s.finish(s.ok ? state_slab(k, kk, T, S, r0, r1, r2, r3, n) : POP_FAIL);
}

extern "C" int pop_advt(int k, double* LTK, double* WTK, const double* TMIX, const double* TRCR,
                        const double* UUU, const double* VVV, const pop_block* blk) {
  POP_TRY(check_ready("advt", blk, true));
  POP_REQUIRE(LTK && WTK && TRCR && UUU && VVV, "advt: null argument");
  (void)TMIX;
  ScopedTimer tm("ADVECTION_TRACER");
  Stage s;
  TracerIO io;
  memset(&io, 0, sizeof(io));
  io.TCUR = s.in("a_trcr", TRCR, G.n3 * G.nt);
  io.TMIX = io.TCUR;
  io.TOLD = io.TCUR;
  if (G.use_lw_lim) {  // lw_lim advects the mix-time field (advection.F90:1704, :2806)
    POP_REQUIRE(TMIX, "advt: lw_lim needs TMIX");
    io.TMIX = s.in("a_tmix", TMIX, G.n3 * G.nt);
  }
  io.UCUR = s.in("a_u", UUU, G.n3);
  io.VCUR = s.in("a_v", VVV, G.n3);
  io.TNEW = s.inout("a_ltk", LTK, G.n2 * G.nt, false);
  io.WTK = s.inout("a_wtk", WTK, G.n2, true);
  return s.finish(s.ok ? tracer_column(TR_ADVT, k, io) : POP_FAIL);
}

extern "C" int pop_hdifft(int k, double* HDTK, const double* TMIX, const double* UMIX, const double* VMIX,
                          const pop_block* blk) {
  POP_TRY(check_ready("hdifft", blk, true));
  POP_REQUIRE(HDTK && TMIX, "hdifft: null argument");
  (void)UMIX; (void)VMIX;
  ScopedTimer tm("HMIX_TRACER");
  Stage s;
  TracerIO io;
  memset(&io, 0, sizeof(io));
  io.TMIX = s.in("h_tmix", TMIX, G.n3 * G.nt);
  io.TCUR = io.TMIX;
  io.TOLD = io.TMIX;
  io.TNEW = s.inout("h_hdtk", HDTK, G.n2 * G.nt, false);
  if (s.ok && G.cfg.hmix_tracer_itype == POP_HMIX_GM) {
    // hdifft_gm must be called with k = 1,2,3,...: the k == 1 call does the work of every level (slopes,
    // diffusivities, VDC += VDC_GM, tendency), later calls hand out their level (hmix_gm.F90:1109)
    POP_REQUIRE(k >= 1 && k <= G.km, "hdifft: k=%d out of range", k);
    if (k == 1) POP_TRY(gm_tendency_dev(io.TMIX));
    for (int n = 0; n < G.nt; n++)
      POP_CHECK_CUDA(cudaMemcpyAsync(io.TNEW + (size_t)n * G.n2, fld("GM_HDT") + ((size_t)n * G.km + (k - 1)) * G.n2,
                                     sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
    return s.finish(POP_SUCCESS);
  }
  return s.finish(s.ok ? tracer_column(TR_HDIFFT, k, io) : POP_FAIL);
}

extern "C" int pop_vdifft(int k, double* VDTK, const double* TOLD, const double* STF, const pop_block* blk) {
  POP_TRY(check_ready("vdifft", blk, true));
  POP_REQUIRE(VDTK && TOLD && STF, "vdifft: null argument");
  ScopedTimer tm("VMIX_TRACER_EXPLICIT");
  Stage s;
  TracerIO io;
  memset(&io, 0, sizeof(io));
  io.TOLD = s.in("v_told", TOLD, G.n3 * G.nt);
  io.TCUR = io.TOLD;
  io.TMIX = io.TOLD;
  io.STF = s.in("v_stf", STF, G.n2 * G.nt);
  io.TNEW = s.inout("v_vdtk", VDTK, G.n2 * G.nt, false);
  return s.finish(s.ok ? tracer_column(TR_VDIFFT, k, io) : POP_FAIL);
}

extern "C" int pop_advu(int k, double* LUK, double* LVK, double* WUK, const double* UUU, const double* VVV,
                        const pop_block* blk) {
  POP_TRY(check_ready("advu", blk, true));
  POP_REQUIRE(LUK && LVK && WUK && UUU && VVV, "advu: null argument");
  ScopedTimer tm("ADVECTION_MOMENTUM");
  Stage s;
  MomentumIO io;
  memset(&io, 0, sizeof(io));
  io.UCUR = s.in("m_u", UUU, G.n3);
  io.VCUR = s.in("m_v", VVV, G.n3);
  io.UNEW = s.inout("m_o1", LUK, G.n2, false);
  io.VNEW = s.inout("m_o2", LVK, G.n2, false);
  io.WUK = s.inout("m_wuk", WUK, G.n2, true);
  return s.finish(s.ok ? momentum_column(MO_ADVU, k, io) : POP_FAIL);
}

extern "C" int pop_hdiffu(int k, double* HDUK, double* HDVK, const double* UMIXK, const double* VMIXK,
                          const pop_block* blk) {
  POP_TRY(check_ready("hdiffu", blk, true));
  POP_REQUIRE(HDUK && HDVK && UMIXK && VMIXK, "hdiffu: null argument");
  ScopedTimer tm("HMIX_MOMENTUM");
  Stage s;
  MomentumIO io;
  memset(&io, 0, sizeof(io));
  io.UMIX = s.in("m_u", UMIXK, G.n2);
  io.VMIX = s.in("m_v", VMIXK, G.n2);
  io.UNEW = s.inout("m_o1", HDUK, G.n2, false);
  io.VNEW = s.inout("m_o2", HDVK, G.n2, false);
  return s.finish(s.ok ? momentum_column(MO_HDIFFU, k, io) : POP_FAIL);
}

extern "C" int pop_gradp(int k, double* PKX, double* PKY, const double* RHOK_OLD, const double* RHOK_CUR,
                         const double* RHOK_NEW, const pop_block* blk) {
  POP_TRY(check_ready("gradp", blk, true));
  POP_REQUIRE(PKX && PKY && RHOK_CUR, "gradp: null argument");
  Stage s;
  MomentumIO io;
  memset(&io, 0, sizeof(io));
  io.RHOOLD = s.in("m_ro", RHOK_OLD, G.n2);
  io.RHOCUR = s.in("m_rc", RHOK_CUR, G.n2);
  io.RHONEW = s.in("m_rn", RHOK_NEW, G.n2);
  io.UNEW = s.inout("m_o1", PKX, G.n2, false);
  io.VNEW = s.inout("m_o2", PKY, G.n2, false);
  return s.finish(s.ok ? momentum_column(MO_GRADP, k, io) : POP_FAIL);
}

extern "C" int pop_vdiffu(int k, double* VDUK, double* VDVK, const double* UOLD, const double* VOLD,
                          const double* SMF, const pop_block* blk) {
  POP_TRY(check_ready("vdiffu", blk, true));
  POP_REQUIRE(VDUK && VDVK && UOLD && VOLD && SMF, "vdiffu: null argument");
  ScopedTimer tm("VMIX_MOMENTUM_EXPLICIT");
  Stage s;
  MomentumIO io;
  memset(&io, 0, sizeof(io));
  io.UOLD = s.in("m_u", UOLD, G.n3);
  io.VOLD = s.in("m_v", VOLD, G.n3);
  io.SMF = s.in("m_smf", SMF, G.n2 * 2);
  io.UNEW = s.inout("m_o1", VDUK, G.n2, false);
  io.VNEW = s.inout("m_o2", VDVK, G.n2, false);
  return s.finish(s.ok ? momentum_column(MO_VDIFFU, k, io) : POP_FAIL);
}

extern "C" int pop_grad(int k, double* GRADX, double* GRADY, const double* F, const pop_block* blk) {
  POP_TRY(check_ready("grad", blk, true));
  POP_REQUIRE(GRADX && GRADY && F, "grad: null argument");
  Stage s;
  const double* f = s.in("g_f", F, G.n2);
  double* gx = s.inout("g_x", GRADX, G.n2, false);
  double* gy = s.inout("g_y", GRADY, G.n2, false);
  return s.finish(s.ok ? grad_dev(k, gx, gy, f) : POP_FAIL);
}
extern "C" int pop_div(int k, double* DIV_OUT, const double* UX, const double* UY, const pop_block* blk) {
  POP_TRY(check_ready("div", blk, true));
  POP_REQUIRE(DIV_OUT && UX && UY, "div: null argument");
  Stage s;
  const double* ux = s.in("g_f", UX, G.n2);
  const double* uy = s.in("g_f2", UY, G.n2);
  double* d = s.inout("g_x", DIV_OUT, G.n2, false);
  return s.finish(s.ok ? div_dev(k, d, ux, uy) : POP_FAIL);
}

extern "C" int pop_impvmixt(double* TNEW, const double* TOLD, const double* PSFC, int nfirst, int nlast,
                            const pop_block* blk) {
  POP_TRY(check_ready("impvmixt", blk, true));
  POP_REQUIRE(TNEW && TOLD && PSFC, "impvmixt: null argument");
  Stage s;
  double* tn = s.inout("i_tn", TNEW, G.n3 * G.nt, true);
  const double* to = s.in("i_to", TOLD, G.n3 * G.nt);
  const double* ps = s.in("i_ps", PSFC, G.n2);
  return s.finish(s.ok ? impvmixt_dev(tn, to, ps, nullptr, nfirst, nlast, 0) : POP_FAIL);
}
extern "C" int pop_impvmixt_correct(double* TNEW, const double* PSFC, const double* RHS, int nfirst,
                                    int nlast, const pop_block* blk) {
  POP_TRY(check_ready("impvmixt_correct", blk, true));
  POP_REQUIRE(TNEW && PSFC && RHS, "impvmixt_correct: null argument");
  Stage s;
  double* tn = s.inout("i_tn", TNEW, G.n3 * G.nt, true);
  const double* ps = s.in("i_ps", PSFC, G.n2);
  const double* rhs = s.in("i_rhs", RHS, G.n2 * G.nt);
  return s.finish(s.ok ? impvmixt_dev(tn, nullptr, ps, rhs, nfirst, nlast, 1) : POP_FAIL);
}
extern "C" int pop_impvmixu(double* UNEW, double* VNEW, const pop_block* blk) {
  POP_TRY(check_ready("impvmixu", blk, true));
  POP_REQUIRE(UNEW && VNEW, "impvmixu: null argument");
  Stage s;
  double* u = s.inout("i_u", UNEW, G.n3, true);
  double* v = s.inout("i_v", VNEW, G.n3, true);
  return s.finish(s.ok ? impvmixu_dev(u, v) : POP_FAIL);
}
extern "C" int pop_vmix_coeffs(int k, const double* TMIX, const double* UMIX, const double* VMIX,
                               const double* RHOMIX, const pop_block* blk) {
  POP_TRY(check_ready("vmix_coeffs", blk, true));
  Stage s;
  const double* t = s.in("c_t", TMIX, G.n3 * G.nt);
  const double* u = s.in("c_u", UMIX, G.n3);
  const double* v = s.in("c_v", VMIX, G.n3);
  const double* r = s.in("c_r", RHOMIX, G.n3);
  return s.finish(s.ok ? vmix_coeffs_dev(k, k, t, u, v, r) : POP_FAIL);
}

// ------------------------------------------------------------------ solver
extern "C" int pop_solvers_run(double* sfcPressure, const double* rhsClinic) {
  POP_TRY(check_ready("POP_SolversRun", nullptr, false));
  POP_REQUIRE(sfcPressure && rhsClinic, "POP_SolversRun: null argument");
  Stage s;
  double* x = s.inout("s_x", sfcPressure, G.n2, true);
  const double* b = s.in("s_b", rhsClinic, G.n2);
  return s.finish(s.ok ? solvers_run_dev(x, b) : POP_FAIL);
}
extern "C" int pop_solvers_diagonal(const double* diagonalCorrection, int blockIndx) {
  POP_TRY(check_ready("POP_SolversDiagonal", nullptr, false));
  POP_REQUIRE(diagonalCorrection != nullptr, "POP_SolversDiagonal: null argument");
  POP_REQUIRE(blockIndx == 1, "POP_SolversDiagonal: blockIndx=%d, this rank owns exactly one block", blockIndx);
  Stage s;
  const double* d = s.in("s_d", diagonalCorrection, G.n2);
  return s.finish(s.ok ? solvers_diagonal_dev(d) : POP_FAIL);
}
extern "C" int pop_solvers_get_diagnostics(int* iterationCount, double* residual) {
  if (iterationCount) *iterationCount = G.numIterations;
  if (residual) *residual = G.rmsResidual;
  return POP_SUCCESS;
}
extern "C" int pop_btrop_operator(double* AX, const double* X, int bid) {
  POP_TRY(check_ready("btropOperator", nullptr, false));
  POP_REQUIRE(AX && X && bid == 1, "btropOperator: bad argument");
  Stage s;
  const double* x = s.in("s_b", X, G.n2);
  double* ax = s.inout("s_x", AX, G.n2, false);
  int rc = POP_FAIL;
  if (s.ok) {
    POP_CHECK_CUDA(cudaMemcpyAsync(fld("btropWgtCenter"), fld("centerWgtClinic"), sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
    rc = btrop_operator_dev(ax, x);
  }
  return s.finish(rc);
}
extern "C" int pop_solvers_prep(void) {
  POP_TRY(check_ready("POP_SolversPrep", nullptr, false));
  return solvers_prep_dev();
}
extern "C" int pop_solvers_get_evp_diagnostics(int* subBlocks, int* landSubBlocks, double* maxInverseError) {
  return solvers_evp_diagnostics(subBlocks, landSubBlocks, maxInverseError);
}
extern "C" int pop_solvers_get_eigs(double* mineig, double* maxeig) {
  if (mineig) *mineig = G.pcsiMinEigs;
  if (maxeig) *maxeig = G.pcsiMaxEigs;
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ communication
extern "C" int pop_halo_update_2d_r8(double* array, int fieldLoc, int fieldKind, double fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  Stage s;
  double* a = s.inout("hl_a", array, G.n2, true);
  return s.finish(s.ok ? halo_update(a, 1, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_halo_update_3d_r8(double* array, int nz, int fieldLoc, int fieldKind, double fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  POP_REQUIRE(nz >= 1, "POP_HaloUpdate: nz=%d", nz);
  Stage s;
  double* a = s.inout("hl_a", array, G.n2 * nz, true);
  return s.finish(s.ok ? halo_update(a, nz, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_halo_update_4d_r8(double* array, int nz, int nt, int fieldLoc, int fieldKind,
                                     double fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  POP_REQUIRE(nz >= 1 && nt >= 1, "POP_HaloUpdate: nz=%d nt=%d", nz, nt);
  Stage s;
  double* a = s.inout("hl_a", array, G.n2 * nz * nt, true);
  return s.finish(s.ok ? halo_update(a, nz * nt, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_halo_update_2d_i4(int* array, int fieldLoc, int fieldKind, int fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  Stage s;
  int* a = s.inout("hl_a", array, G.n2, true);
  return s.finish(s.ok ? halo_update_i4(a, 1, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
// the remaining members of the POP_HaloUpdate generic interface (mpi/POP_HaloMod.F90:79-89): 3-d/4-d integer and
// 2-d/3-d/4-d single-precision arrays; the same kernels instantiated for the element type
extern "C" int pop_halo_update_3d_i4(int* array, int nz, int fieldLoc, int fieldKind, int fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  POP_REQUIRE(nz >= 1, "POP_HaloUpdate: nz=%d", nz);
  Stage s;
  int* a = s.inout("hl_a", array, G.n2 * nz, true);
  return s.finish(s.ok ? halo_update_i4(a, nz, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_halo_update_4d_i4(int* array, int nz, int nt, int fieldLoc, int fieldKind, int fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  POP_REQUIRE(nz >= 1 && nt >= 1, "POP_HaloUpdate: nz=%d nt=%d", nz, nt);
  Stage s;
  int* a = s.inout("hl_a", array, G.n2 * nz * nt, true);
  return s.finish(s.ok ? halo_update_i4(a, nz * nt, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_halo_update_2d_r4(float* array, int fieldLoc, int fieldKind, float fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  Stage s;
  float* a = s.inout("hl_a", array, G.n2, true);
  return s.finish(s.ok ? halo_update_r4(a, 1, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_halo_update_3d_r4(float* array, int nz, int fieldLoc, int fieldKind, float fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  POP_REQUIRE(nz >= 1, "POP_HaloUpdate: nz=%d", nz);
  Stage s;
  float* a = s.inout("hl_a", array, G.n2 * nz, true);
  return s.finish(s.ok ? halo_update_r4(a, nz, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_halo_update_4d_r4(float* array, int nz, int nt, int fieldLoc, int fieldKind, float fillValue) {
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: not initialized");
  POP_REQUIRE(nz >= 1 && nt >= 1, "POP_HaloUpdate: nz=%d nt=%d", nz, nt);
  Stage s;
  float* a = s.inout("hl_a", array, G.n2 * nz * nt, true);
  return s.finish(s.ok ? halo_update_r4(a, nz * nt, fieldLoc, fieldKind, fillValue) : POP_FAIL);
}
extern "C" int pop_global_sum_nfields_2d_r8(const double* array, int nfields, int fieldLoc,
                                            const double* mMask, double* sums) {
  POP_REQUIRE(G.initialized, "POP_GlobalSum: not initialized");
  POP_REQUIRE(array && sums, "POP_GlobalSum: null argument");
  POP_REQUIRE(nfields >= 1 && nfields <= POP_RED_NF, "POP_GlobalSum: nfields=%d (max %d)", nfields, POP_RED_NF);
  Stage s;
  const double* a = s.in("gs_a", array, G.n2 * nfields);
  const double* m = s.in("gs_m", mMask, G.n2);
  return s.finish(s.ok ? global_sum_dev(a, nfields, G.n2, fieldLoc, m, sums) : POP_FAIL);
}
extern "C" int pop_global_sum_2d_r8(const double* array, int fieldLoc, const double* mMask, double* sum) {
  return pop_global_sum_nfields_2d_r8(array, 1, fieldLoc, mMask, sum);
}

// ------------------------------------------------------------------ fused drivers
extern "C" int pop_dhdt(void) {
  POP_TRY(check_ready("dhdt", nullptr, false));
  return dhdt_dev();
}
// comp_flux_vel_ghost(DH, errorCode), advection.F90:1014: the flux velocities of UVEL/VVEL(curtime) two cells outside the
// physical domain, which the lw_lim branch of advt needs.  The fused drivers call it themselves; a host that drives
// pop_advt level by level calls it once per step before the first level, as baroclinic_driver does (baroclinic.F90:667).
extern "C" int pop_comp_flux_vel_ghost(const double* DH) {
  POP_TRY(check_ready("comp_flux_vel_ghost", nullptr, false));
  if (!G.use_lw_lim) return POP_SUCCESS;  // :1057
  POP_REQUIRE(DH, "comp_flux_vel_ghost: null DH");
  double* dh = fld("DH");  // the library's own DH (what pop_dhdt fills) is what the lw_lim kernel reads at the surface
  if (DH != dh) POP_CHECK_CUDA(cudaMemcpyAsync(dh, DH, sizeof(double) * G.n2, cudaMemcpyDefault, G.stream));
  return with_p2p_check(lw_flux_prepare_dev(fld_t("UVEL", G.curtime), fld_t("VVEL", G.curtime), dh));
}
extern "C" int pop_baroclinic_driver(void) {
  POP_TRY(check_ready("baroclinic_driver", nullptr, false));
  return with_p2p_check(baroclinic_driver_dev());
}
extern "C" int pop_barotropic_driver(void) {
  POP_TRY(check_ready("barotropic_driver", nullptr, false));
  return with_p2p_check(barotropic_driver_dev());
}
extern "C" int pop_baroclinic_correct_adjust(void) {
  POP_TRY(check_ready("baroclinic_correct_adjust", nullptr, false));
  return with_p2p_check(baroclinic_correct_adjust_dev());
}
extern "C" int pop_step(int ts_type) {
  POP_TRY(check_ready("step", nullptr, false));
  return step_dev(ts_type);
}

// One coupled step from HOST buffers: surface forcing in, surface state out (physical strips).
// ---- coupled step: the surface forcing comes in from host buffers, the new surface state goes out to a host
// buffer, the 3-d state never leaves HBM. STF (read by the first kernel of the step) is copied on the main stream;
// SMF, SHF_QSW and FW follow on the copy stream while the tracer kernel runs, and the step waits for them just
// before their first reader. PSURF leaves as soon as the barotropic solve has produced it, SST/SSS as soon as the
// corrector has; only U1/V1 (final after the barotropic mean is added back, at the very end) are copied after the
// step. Averaging and Robert-filtered steps rewrite the output level at the end, so everything leaves after the step.
int coupled_wait_forcing() {
  if (G.cio.forcing_pending) {
    POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_cp_in, 0));
    G.cio.forcing_pending = false;
  }
  if (G.cio.fw_halo_pending) {
    // the coupler hands over physical cells; the reference fills the ghost cells of the fluxes before they are used
    // (update_ghost_cells_coupler_fluxes, forcing_coupled.F90).  On this path only FW is read off its own cell: DH = ... -
    // FW_OLD is averaged to U points (tgrid_to_ugrid) -- without this the answer depends on where the strips are cut.
    G.cio.fw_halo_pending = false;
    POP_TRY(halo_update(fld("FW"), 1, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
  }
  return POP_SUCCESS;
}
int coupled_after_barotropic() {
  if (!G.cio.out || !G.cio.early) return POP_SUCCESS;
  DevField* f;
  const size_t strip = (size_t)G.nxg * G.ny_local;
  POP_CHECK_CUDA(cudaEventRecord(G.ev_cp_a, G.stream));
  POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream_cp, G.ev_cp_a, 0));
  POP_TRY(find_field("pop_step_coupled", "PSURF", POP_TIME_NEW, &f));
  POP_TRY(strip_copy(f, G.cio.out + 2 * strip, false, 0, 1, G.stream_cp));
  G.cio.psurf_sent = true;
  return POP_SUCCESS;
}
int coupled_after_corrector() {
  if (!G.cio.out || !G.cio.early) return POP_SUCCESS;
  DevField* f;
  const size_t strip = (size_t)G.nxg * G.ny_local;
  POP_CHECK_CUDA(cudaEventRecord(G.ev_cp_b, G.stream));
  POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream_cp, G.ev_cp_b, 0));
  POP_TRY(find_field("pop_step_coupled", "TRACER", POP_TIME_NEW, &f));
  POP_TRY(strip_copy(f, G.cio.out, false, 0, 1, G.stream_cp));              // SST
  POP_TRY(strip_copy(f, G.cio.out + strip, false, G.km, 1, G.stream_cp));   // SSS
  G.cio.ts_sent = true;
  return POP_SUCCESS;
}

extern "C" int pop_step_coupled(int ts_type, const double* STF, const double* SMF, const double* SHF_QSW,
                                const double* FW, double* sfc_out) {
  POP_TRY(check_ready("pop_step_coupled", nullptr, false));
  DevField* f;
  const bool side = !G.no_overlap && (SMF || FW || STF);
  cudaStream_t st_in = side ? G.stream_cp : G.stream;
  if (side) {
    // the copy stream must not overwrite STF / SMF / FW while kernels of a previous (unsynchronised) pop_step still read them
    POP_CHECK_CUDA(cudaEventRecord(G.ev_cp_a, G.stream));
    POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream_cp, G.ev_cp_a, 0));
  }
  // STF is read by the first kernel of the step.  Pinned host buffer: that kernel reads it from the host directly (one value
  // per column) and the copy into the device field joins the others on the copy stream; otherwise it is copied up front.
  G.cio.stf_dev = nullptr;
  if (STF && side && !(getenv("POP_B200_NO_DIRECT_SFC") && getenv("POP_B200_NO_DIRECT_SFC")[0] == '1')) {
#ifndef POP_EMUL
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, STF) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
      G.cio.stf_dev = (const double*)at.devicePointer;
    else
      cudaGetLastError();
#else
    G.cio.stf_dev = STF;
#endif
  }
  if (STF) {
    POP_TRY(find_field("pop_step_coupled", "STF", 0, &f));
    POP_TRY(strip_copy(f, (void*)STF, true, 0, G.nt, G.cio.stf_dev ? st_in : G.stream));
  }
  // SHF_QSW is part of the coupler's forcing set but nothing on this path consumes it (penetrative short-wave
  // absorption, add_sw_absorb, is outside SURVEY section 8): it is accepted and ignored -- not copied, not counted
  (void)SHF_QSW;
  if (SMF) { POP_TRY(find_field("pop_step_coupled", "SMF", 0, &f)); POP_TRY(strip_copy(f, (void*)SMF, true, 0, 2, st_in)); }
  if (FW) { POP_TRY(find_field("pop_step_coupled", "FW", 0, &f)); POP_TRY(strip_copy(f, (void*)FW, true, 0, 1, st_in)); }
  if (side) {
    POP_CHECK_CUDA(cudaEventRecord(G.ev_cp_in, G.stream_cp));
    G.cio.forcing_pending = true;
  }
  G.cio.fw_halo_pending = (FW != nullptr);
  G.cio.out = sfc_out;
  G.cio.early = sfc_out && !G.no_overlap && (ts_type == POP_TS_LEAPFROG || ts_type == POP_TS_EULER);
  G.cio.psurf_sent = G.cio.ts_sent = G.cio.uv_sent = false;
  G.cio.uv_dev = nullptr;
  if (G.cio.early && G.finish_mode == 0 && !(getenv("POP_B200_NO_DIRECT_SFC") && getenv("POP_B200_NO_DIRECT_SFC")[0] == '1')) {
    // a pinned host buffer is device-accessible: the velocity-finish kernel can store U1/V1 into it directly
    const size_t strip = (size_t)G.nxg * G.ny_local;
#ifndef POP_EMUL
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, sfc_out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
      G.cio.uv_dev = (double*)at.devicePointer + 3 * strip;
    else
      cudaGetLastError();
#else
    G.cio.uv_dev = sfc_out + 3 * strip;
#endif
  }
  const int rc = step_dev(ts_type);
  const bool psurf_sent = G.cio.psurf_sent, ts_sent = G.cio.ts_sent, uv_sent = G.cio.uv_sent;
  G.cio.out = nullptr;
  G.cio.early = false;
  G.cio.uv_dev = nullptr;
  G.cio.stf_dev = nullptr;
  if (rc != POP_SUCCESS) {
    if (G.cio.forcing_pending) { cudaStreamSynchronize(G.stream_cp); G.cio.forcing_pending = false; }
    cudaStreamSynchronize(G.stream_cp);
    return rc;
  }
  POP_TRY(coupled_wait_forcing());  // a step that never reached the wait (cannot happen today) must not leave copies behind
  if (sfc_out) {
    // after the step the freshly computed level is `cur` (or the averaged `cur` on an averaging step)
    const size_t strip = (size_t)G.nxg * G.ny_local;
    POP_TRY(find_field("pop_step_coupled", "TRACER", POP_TIME_CUR, &f));
    if (!ts_sent) {
      POP_TRY(strip_copy(f, sfc_out, false, 0, 1));               // SST
      POP_TRY(strip_copy(f, sfc_out + strip, false, G.km, 1));    // SSS
    }
    if (!psurf_sent) {
      POP_TRY(find_field("pop_step_coupled", "PSURF", POP_TIME_CUR, &f));
      POP_TRY(strip_copy(f, sfc_out + 2 * strip, false, 0, 1));
    }
    if (!uv_sent) {
      POP_TRY(find_field("pop_step_coupled", "UVEL", POP_TIME_CUR, &f));
      POP_TRY(strip_copy(f, sfc_out + 3 * strip, false, 0, 1));
      POP_TRY(find_field("pop_step_coupled", "VVEL", POP_TIME_CUR, &f));
      POP_TRY(strip_copy(f, sfc_out + 4 * strip, false, 0, 1));
    } else if (G.cfg.ns_boundary_type == POP_BNDY_TRIPOLE && G.rank == G.nranks - 1) {
      // the kernel left out the tripole seam row (final only after the halo updates and its own barotropic add)
      const size_t src = (size_t)(G.je - 1) * G.nxb + (G.ib - 1), dst = (size_t)(G.ny_local - 1) * G.nxg;
      POP_CHECK_CUDA(cudaMemcpyAsync(sfc_out + 3 * strip + dst, fld_t("UVEL", G.curtime) + src, sizeof(double) * G.nxg,
                                     cudaMemcpyDeviceToHost, G.stream));
      POP_CHECK_CUDA(cudaMemcpyAsync(sfc_out + 4 * strip + dst, fld_t("VVEL", G.curtime) + src, sizeof(double) * G.nxg,
                                     cudaMemcpyDeviceToHost, G.stream));
    }
  }
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream_cp));
  return POP_SUCCESS;
}
