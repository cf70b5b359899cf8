// pop_barotropic.cu -- barotropic driver and the elliptic solver.
//
//   POP_SolversMod.F90: btropOperator :2376-2431; pcg :1200-1503; PCSI :1510-1835; ChronGear
//   :1841-2266; PcsiLanczos :2699-2990 + ratqr :3122-3222; POP_SolversDiagonal :1110-1151;
//   POP_SolversRun :327-495 (clinic == tropic distribution, so POP_RedistributeBlocks is a copy).
//   barotropic.F90: init_barotropic :170-252 (pop_grid.cu), barotropic_driver :267-735.
//
// Every iteration is a fixed sequence of fused stencil-plus-update kernels; dot products are written
// as per-CTA double-double partials by the kernel that produces the operands and finished by
// reduce_finish() (block order, then rank order).  The ChronGear / PCG recurrence scalars live on
// the device (SolverScalars), so an iteration needs no host round trip; the host reads the residual
// norm only on the convergence checks every `convergenceCheckFreq` iterations, as the reference.
#include <cmath>
#include <vector>
#include "pop_dev.cuh"

struct BtView {
  int nxb, nyb, ib, jb, nxp, nyp;  // ib,jb 1-based; physical extent nxp x nyp
  const double *C, *N, *E, *NE, *mask;
};
static BtView bt_view() {
  BtView v;
  v.nxb = G.nxb; v.nyb = G.nyb; v.ib = G.ib; v.jb = G.jb;
  v.nxp = G.ie - G.ib + 1; v.nyp = G.je - G.jb + 1;
  v.C = fld("btropWgtCenter"); v.N = fld("btropWgtNorth"); v.E = fld("btropWgtEast");
  v.NE = fld("btropWgtNE"); v.mask = fld("mMaskTropic");
  return v;
}

// 9-point operator at array point q=(i,j) (0-based); zero outside 1..nxb-2 x 1..nyb-2
__device__ __forceinline__ double btrop_point(const BtView& v, const double* __restrict__ X, int i, int j) {
  if (i < 1 || i > v.nxb - 2 || j < 1 || j > v.nyb - 2) return 0.0;
  const int nxb = v.nxb;
  const size_t q = (size_t)j * nxb + i;
  return v.C[q] * X[q] + v.N[q] * X[q + nxb] + v.N[q - nxb] * X[q - nxb] + v.E[q] * X[q + 1] +
         v.E[q - 1] * X[q - 1] + v.NE[q] * X[q + nxb + 1] + v.NE[q - nxb] * X[q - nxb + 1] +
         v.NE[q - 1] * X[q + nxb - 1] + v.NE[q - nxb - 1] * X[q - nxb - 1];
}
__device__ __forceinline__ bool bt_physical(const BtView& v, int i, int j) {
  return i >= v.ib - 1 && i < v.ib - 1 + v.nxp && j >= v.jb - 1 && j < v.jb - 1 + v.nyp;
}
__device__ __forceinline__ void bt_store_partials(dd* acc, int nf, double* partials, int red_blocks) {
  for (int f = 0; f < nf; f++) {
    dd r = block_reduce_dd(acc[f]);
    if (threadIdx.x == 0) {
      partials[((size_t)f * red_blocks + blockIdx.x) * 2] = r.hi;
      partials[((size_t)f * red_blocks + blockIdx.x) * 2 + 1] = r.lo;
    }
  }
}

// Generic fused solver kernel: grid-stride over all points of the padded block, fixed grid
// (red_blocks CTAs) so that the partial sums are reproducible.
enum {
  BT_APPLY = 0,     // V0 = A V1
  BT_RESID,         // V0 = V2 - A V1                                    [sum0 = V0^2]
  BT_RESID_A0R,     // r = V2 - A V1 ; sum0 += r^2 ; V0 = r * V3(A0R)
  BT_CG_ZS,         // Z(V0) = R(V1)*A0R(V2) ; S(V3) = Z
  BT_CG_QSUM,       // Q(V0) = A S(V1) ; sums: R(V2)*Z(V3), S*Q
  BT_CG_AZSUM,      // AZ(V0) = A Z(V1) ; sums: R(V2)*Z, AZ*Z
  BT_PCG_AS,        // Q(V0) = A S(V1) ; sum: Q*S
  BT_LZ_AP          // R(V0) = A P(V1) - s0*Q1(V2) ; sum: P*R
};
template <int OP>
__global__ void __launch_bounds__(POP_EW_THREADS)
bt_stencil_kernel(BtView v, double* __restrict__ V0, const double* __restrict__ V1,
                  const double* __restrict__ V2, double* __restrict__ V3, double s0, int want_sums,
                  double* __restrict__ partials, int red_blocks) {
  dd acc[2] = {dd{0.0, 0.0}, dd{0.0, 0.0}};
  const size_t n = (size_t)v.nxb * v.nyb;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n;
       q += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)((unsigned)q / (unsigned)v.nxb), i = (int)((unsigned)q - (unsigned)j * (unsigned)v.nxb);  // n < 2^32
    const bool phys = bt_physical(v, i, j);
    const double m = phys ? v.mask[q] : 0.0;
    if (OP == BT_APPLY) {
      V0[q] = btrop_point(v, V1, i, j);
    } else if (OP == BT_RESID) {
      const double r = V2[q] - btrop_point(v, V1, i, j);
      V0[q] = r;
      if (want_sums && phys) acc[0] = dd_add_d(acc[0], (r * r) * m);
    } else if (OP == BT_RESID_A0R) {
      const double r = V2[q] - btrop_point(v, V1, i, j);
      if (want_sums && phys) acc[0] = dd_add_d(acc[0], (r * r) * m);
      V0[q] = r * V3[q];
    } else if (OP == BT_CG_ZS) {
      const double z = V1[q] * V2[q];
      V0[q] = z;
      V3[q] = z;
    } else if (OP == BT_CG_QSUM) {
      const double qv = btrop_point(v, V1, i, j);
      V0[q] = qv;
      if (phys) {
        acc[0] = dd_add_d(acc[0], (V2[q] * V3[q]) * m);
        acc[1] = dd_add_d(acc[1], (V1[q] * qv) * m);
      }
    } else if (OP == BT_CG_AZSUM) {
      const double az = btrop_point(v, V1, i, j);
      V0[q] = az;
      if (phys) {
        acc[0] = dd_add_d(acc[0], (V2[q] * V1[q]) * m);
        acc[1] = dd_add_d(acc[1], (az * V1[q]) * m);
      }
    } else if (OP == BT_PCG_AS) {
      const double qv = btrop_point(v, V1, i, j);
      V0[q] = qv;
      if (phys) acc[0] = dd_add_d(acc[0], (qv * V1[q]) * m);
    } else if (OP == BT_LZ_AP) {
      const double r = btrop_point(v, V1, i, j) - s0 * V2[q];
      V0[q] = r;
      if (phys) acc[0] = dd_add_d(acc[0], (V1[q] * r) * m);
    }
  }
  if (OP == BT_CG_QSUM || OP == BT_CG_AZSUM) bt_store_partials(acc, 2, partials, red_blocks);
  else if (OP == BT_PCG_AS || OP == BT_LZ_AP || ((OP == BT_RESID || OP == BT_RESID_A0R) && want_sums))
    bt_store_partials(acc, 1, partials, red_blocks);
}

// element-wise updates (all points of the padded block, like the reference's whole-array statements)
enum {
  EW_A0R = 0,     // V0 = C != 0 ? 1/C : 0
  EW_CG_UPDATE0,  // X += alpha*S ; R -= alpha*Q
  EW_CG_UPDATE,   // S = Z + beta*S ; Q = AZ + beta*Q ; X += alpha*S ; R -= alpha*Q ; [Z = R*A0R]
  EW_MUL,         // V0 = V1 * V2
  EW_PCSI_INIT,   // R = R*A0R ; Q = (1/csy)*R
  EW_PCSI_QX,     // Q = om*R + (csy*om-1)*Q ; X += Q
  EW_ADD,         // V0 = V0 + V1
  EW_PCG_W1,      // W1(V0) = C!=0 ? R/C : 0 ; sum R*W1
  EW_PCG_S,       // S = W1 + S*(eta1/eta0)
  EW_PCG_UPDATE,  // X += eta1*S ; R -= eta1*Q
  EW_LZ_SR,       // S(V0) = R(V1)*A0R(V2) ; sum S*R
  EW_SCALE,       // V0 = s0 * V1
  EW_AXPY,        // V0 = V0 - s0*V1
  EW_SET,         // V0 = s0
  EW_LZ_SHIFT,    // Q1(V0) = Q(V1) ; Q(V1) = s0*R(V2)
  EW_DOT          // sum V0*V1 (masked, physical cells)
};
struct EwArgs {
  double *V0, *V1, *V2, *V3, *V4, *V5, *V6;
  double s0, s1;
  int flag;
};
template <int OP>
__global__ void __launch_bounds__(POP_EW_THREADS)
bt_ew_kernel(BtView v, EwArgs a, const SolverScalars* __restrict__ sc, double* __restrict__ partials,
             int red_blocks) {
  dd acc[1] = {dd{0.0, 0.0}};
  const size_t n = (size_t)v.nxb * v.nyb;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n;
       q += (size_t)gridDim.x * blockDim.x) {
    if (OP == EW_A0R) {
      const double c = a.V1[q];
      a.V0[q] = (c != 0.0) ? 1.0 / c : 0.0;
    } else if (OP == EW_CG_UPDATE0) {  // V0=X V1=R V2=S V3=Q
      const double al = sc->cgAlpha;
      a.V0[q] = a.V0[q] + al * a.V2[q];
      a.V1[q] = a.V1[q] - al * a.V3[q];
    } else if (OP == EW_CG_UPDATE) {  // V0=X V1=R V2=S V3=Q V4=Z V5=AZ V6=A0R
      const double al = sc->cgAlpha, be = sc->cgBeta;
      const double s = a.V4[q] + be * a.V2[q];
      const double qq = a.V5[q] + be * a.V3[q];
      a.V2[q] = s;
      a.V3[q] = qq;
      a.V0[q] = a.V0[q] + al * s;
      const double r = a.V1[q] - al * qq;
      a.V1[q] = r;
      if (a.flag) a.V4[q] = r * a.V6[q];
    } else if (OP == EW_MUL) {
      a.V0[q] = a.V1[q] * a.V2[q];
    } else if (OP == EW_PCSI_INIT) {  // V0=R V1=Q V2=A0R
      const double r = a.V0[q] * a.V2[q];
      a.V0[q] = r;
      a.V1[q] = a.s0 * r;
    } else if (OP == EW_PCSI_QX) {  // V0=Q V1=X V2=R ; s0 = csomga, s1 = csy*csomga-1
      const double qv = a.s0 * a.V2[q] + a.s1 * a.V0[q];
      a.V0[q] = qv;
      a.V1[q] = a.V1[q] + qv;
    } else if (OP == EW_ADD) {
      a.V0[q] = a.V0[q] + a.V1[q];
    } else if (OP == EW_PCG_W1) {  // V0=W1 V1=R V2=C
      const double c = a.V2[q], r = a.V1[q];
      const double w = (c != 0.0) ? r / c : 0.0;
      a.V0[q] = w;
      const int i = (int)(q % v.nxb), j = (int)(q / v.nxb);
      if (bt_physical(v, i, j)) acc[0] = dd_add_d(acc[0], (r * w) * v.mask[q]);
    } else if (OP == EW_PCG_S) {  // V0=S V1=W1
      a.V0[q] = a.V1[q] + a.V0[q] * (sc->eta1 / sc->eta0);
    } else if (OP == EW_PCG_UPDATE) {  // V0=X V1=R V2=S V3=Q
      const double e = sc->eta1;
      a.V0[q] = a.V0[q] + e * a.V2[q];
      a.V1[q] = a.V1[q] - e * a.V3[q];
    } else if (OP == EW_LZ_SR) {  // V0=S V1=R V2=A0R
      const double s = a.V1[q] * a.V2[q];
      a.V0[q] = s;
      const int i = (int)(q % v.nxb), j = (int)(q / v.nxb);
      if (bt_physical(v, i, j)) acc[0] = dd_add_d(acc[0], (s * a.V1[q]) * v.mask[q]);
    } else if (OP == EW_SCALE) {
      a.V0[q] = a.s0 * a.V1[q];
    } else if (OP == EW_AXPY) {
      a.V0[q] = a.V0[q] - a.s0 * a.V1[q];
    } else if (OP == EW_SET) {
      a.V0[q] = a.s0;
    } else if (OP == EW_LZ_SHIFT) {
      a.V0[q] = a.V1[q];
      a.V1[q] = a.s0 * a.V2[q];
    } else if (OP == EW_DOT) {
      const int i = (int)(q % v.nxb), j = (int)(q / v.nxb);
      if (bt_physical(v, i, j)) acc[0] = dd_add_d(acc[0], (a.V0[q] * a.V1[q]) * v.mask[q]);
    }
  }
  if (OP == EW_PCG_W1 || OP == EW_LZ_SR || OP == EW_DOT) bt_store_partials(acc, 1, partials, red_blocks);
}

#define BT_ST(OP, V0, V1, V2, V3, s0, want)                                                            \
  do {                                                                                                 \
    auto kf_ = bt_stencil_kernel<OP>;                                                                  \
    POP_LAUNCH(kf_, G.red_blocks, POP_EW_THREADS, 0, bt_view(), V0, V1, V2, V3, s0, want, G.d_partials, \
               G.red_blocks);                                                                          \
  } while (0)
template <int OP>
static void bt_ew(double* V0, double* V1 = nullptr, double* V2 = nullptr, double* V3 = nullptr,
                  double* V4 = nullptr, double* V5 = nullptr, double* V6 = nullptr, double s0 = 0.0,
                  double s1 = 0.0, int flag = 0) {
  EwArgs a{V0, V1, V2, V3, V4, V5, V6, s0, s1, flag};
  auto kf = bt_ew_kernel<OP>;
  POP_LAUNCH(kf, G.red_blocks, POP_EW_THREADS, 0, bt_view(), a, (const SolverScalars*)G.d_scal,
             G.d_partials, G.red_blocks);
}
static int bt_halo(double* v) { return halo_update(v, 1, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0); }

// ------------------------------------------------------------------ init / diagonal
int solvers_init_dev() {
  // residualNorm and convergenceCriterion: POP_SolversMod.F90:895-906
  double* w = fld("W2A");
  bt_ew<EW_MUL>(w, fld("TAREA"), fld("TAREA"));
  double s = 0.0;
  POP_TRY(global_sum_dev(w, 1, G.n2, POP_LOC_CENTER, fld("mMaskTropic"), &s));
  G.residualNorm = 1.0 / s;
  G.convergenceCriterion = (G.cfg.convergence_criterion * G.cfg.convergence_criterion) / G.residualNorm;
  return POP_SUCCESS;
}

__global__ void diag_kernel(double* __restrict__ cwc, const double* __restrict__ indep,
                            const double* __restrict__ dc, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) cwc[q] = indep[q] - dc[q];
}
int solvers_diagonal_dev(const double* diagCorr) {
  POP_LAUNCH(diag_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("centerWgtClinic"),
             fld("centerWgtClinicIndep"), diagCorr, G.n2);
  return pop_post_launch("solvers_diagonal");
}

int btrop_operator_dev(double* AX, const double* X) {
  BT_ST(BT_APPLY, AX, X, (const double*)nullptr, (double*)nullptr, 0.0, 0);
  return pop_post_launch("btropOperator");
}

static int refresh_center() {  // POP_SolversRun :390-391 / POP_SolversPrep: btropWgtCenter <- centerWgtClinic
  POP_CHECK_CUDA(cudaMemcpyAsync(fld("btropWgtCenter"), fld("centerWgtClinic"), sizeof(double) * G.n2,
                                 cudaMemcpyDeviceToDevice, G.stream));
  bt_ew<EW_A0R>(fld("BT_A0R"), fld("btropWgtCenter"));
  return POP_SUCCESS;
}


// ------------------------------------------------------------------ EVP block preconditioner
// preconditionerChoice = 'evp' (POP_SolversMod.F90: EvpBlockPartition :2992-3040, EvpPre :2434-2506,
// ExplicitBlockEvpPre :2508-2616, inverse :3042-3120, ExplicitEvp :2618-2696, preconditioner :2273-2369, the EVP
// part of POP_SolversPrep :252-292).  The block is cut into sub-blocks of at most 8 x 8 cells; a sub-block without
// land inverts the five-point operator (centre + corner weights) exactly by error-vector propagation: march
// north-east from a zero south/west edge, read the error arriving at the north/east frame, correct the edge with
// the influence matrix, march again; sub-blocks with land use the diagonal of prep time.
//   Set-up (once per pop_solvers_prep, host): the influence matrices are 15 x 15 inverses per sub-block, sequential
// arithmetic of ~10 k operations each -- built on the host from the weights as they are at prep time, uploaded
// TRANSPOSED (element-major, sub-block fastest) so that the apply kernel's loads coalesce.
//   Apply (every iteration, device): one thread per sub-block; the full 10 x 10 case is unrolled with the marching
// rows in registers, ragged edge sub-blocks take the generic path.  Traffic per application: X in, PX out, and per
// 64 cells 3 x 100 coefficient values + 225 influence values = 66 B/cell.
#define EVP_BS 8
#define EVP_LD (EVP_BS + 2)
#define EVP_L (2 * EVP_BS - 1)
struct EvpDev {
  bool ready = false;
  int xnb = 0, ynb = 0, nsub = 0;
  std::vector<int> xidx, yidx;        // EvpXbidx(1:xnb+1), EvpYbidx(1:ynb+1) ([0] unused)
  int *d_xidx = nullptr, *d_yidx = nullptr, *d_land = nullptr;
  double *d_cc = nullptr, *d_ne = nullptr, *d_ine = nullptr, *d_icc = nullptr, *d_rinv = nullptr;  // [elem][nsub]
  double selfcheck = 0.0;
  int nland = 0;
};
static EvpDev EV;
static bool use_evp() { return G.cfg.preconditioner_choice == POP_PRECOND_EVP; }

static void evp_partition(int m, int mm, int* mb_out, std::vector<int>* mdi) {
  const int mb = (m - 3) / mm + 1;
  mdi->assign(mb + 2, 0);
  (*mdi)[1] = 2;
  if (mb == 1) {
    (*mdi)[mb + 1] = m;
  } else {
    for (int i = 1; i <= mb - 2; i++) (*mdi)[i + 1] = 2 + i * mm;
    (*mdi)[mb] = ((*mdi)[mb - 1] + m) / 2;
    (*mdi)[mb + 1] = m;
  }
  *mb_out = mb;
}
#define HF2(a, i, j) (a)[((j)-1) * EVP_LD + ((i)-1)]
// inverse :3042-3120 (Doolittle LU without pivoting); a is destroyed; cinv(i,k) at [(k-1)*n + (i-1)]
static void evp_inverse_host(std::vector<double>& a, std::vector<double>& cinv, int n) {
  std::vector<double> Lm((size_t)n * n, 0.0), Um((size_t)n * n, 0.0), bv(n + 1, 0.0), d(n + 1, 0.0), x(n + 1, 0.0);
  auto A = [&](int i, int j) -> double& { return a[(size_t)(j - 1) * n + (i - 1)]; };
  auto Lx = [&](int i, int j) -> double& { return Lm[(size_t)(j - 1) * n + (i - 1)]; };
  auto Ux = [&](int i, int j) -> double& { return Um[(size_t)(j - 1) * n + (i - 1)]; };
  for (int k = 1; k <= n - 1; k++)
    for (int i = k + 1; i <= n; i++) {
      const double coeff = A(i, k) / A(k, k);
      Lx(i, k) = coeff;
      for (int j = k + 1; j <= n; j++) A(i, j) = A(i, j) - coeff * A(k, j);
    }
  for (int i = 1; i <= n; i++) Lx(i, i) = 1.0;
  for (int j = 1; j <= n; j++)
    for (int i = 1; i <= j; i++) Ux(i, j) = A(i, j);
  for (int k = 1; k <= n; k++) {
    bv[k] = 1.0;
    d[1] = bv[1];
    for (int i = 2; i <= n; i++) {
      d[i] = bv[i];
      for (int j = 1; j <= i - 1; j++) d[i] = d[i] - Lx(i, j) * d[j];
    }
    x[n] = d[n] / Ux(n, n);
    for (int i = n - 1; i >= 1; i--) {
      x[i] = d[i];
      for (int j = n; j >= i + 1; j--) x[i] = x[i] - Ux(i, j) * x[j];
      x[i] = x[i] / Ux(i, i);
    }
    for (int i = 1; i <= n; i++) cinv[(size_t)(k - 1) * n + (i - 1)] = x[i];
    bv[k] = 0.0;
  }
}
// ExplicitBlockEvpPre :2508-2616; rinv(k,j) at [(j-1)*EVP_L + (k-1)]; returns max|rinv*rin - I|
static double evp_block_pre_host(const double* cc, const double* ne, double* rinv, int n, int m) {
  const int nm = n + m - 5;
  double y[EVP_LD * EVP_LD];
  std::vector<double> rin((size_t)nm * nm, 0.0), work, rtmp((size_t)nm * nm, 0.0);
  auto RIN = [&](int i, int j) -> double& { return rin[(size_t)(j - 1) * nm + (i - 1)]; };
  memset(y, 0, sizeof(y));
  for (int pass = 0; pass < 2; pass++) {
    const int cnt = pass == 0 ? m - 2 : n - 3;
    for (int ii = 1; ii <= cnt; ii++) {
      if (pass == 0) HF2(y, 2, m - ii) = 1.0;
      else HF2(y, ii + 2, 2) = 1.0;
      for (int j = 2; j <= m - 1; j++)
        for (int i = 2; i <= n - 1; i++)
          HF2(y, i + 1, j + 1) = (-HF2(cc, i, j) * HF2(y, i, j) - HF2(ne, i, j - 1) * HF2(y, i + 1, j - 1) -
                                  HF2(ne, i - 1, j) * HF2(y, i - 1, j + 1) - HF2(ne, i - 1, j - 1) * HF2(y, i - 1, j - 1)) /
                                 HF2(ne, i, j);
      const int row = pass == 0 ? ii : m - 2 + ii;
      for (int i = 1; i <= n - 2; i++) RIN(row, i) = -HF2(y, i + 2, m);
      for (int j = 1; j <= m - 3; j++) RIN(row, n - 2 + j) = -HF2(y, n, m - j);
      if (pass == 0) HF2(y, 2, m - ii) = 0.0;
      else HF2(y, ii + 2, 2) = 0.0;
    }
  }
  work = rin;
  evp_inverse_host(work, rtmp, nm);
  double maxvalr = 0.0;
  for (int j = 1; j <= nm; j++)
    for (int i = 1; i <= nm; i++) {
      double w = 0.0;
      for (int k = 1; k <= nm; k++) w = w + rtmp[(size_t)(k - 1) * nm + (i - 1)] * RIN(k, j);
      if (i == j) w = w - 1.0;
      if (fabs(w) > maxvalr) maxvalr = fabs(w);
    }
  for (int j = 1; j <= nm; j++)
    for (int k = 1; k <= nm; k++) rinv[(size_t)(j - 1) * EVP_L + (k - 1)] = rtmp[(size_t)(j - 1) * nm + (k - 1)];
  return maxvalr;
}

static int evp_prep_dev() {
  const int nxb = G.nxb, nyb = G.nyb;
  if (!EV.ready) {
    evp_partition(nxb - 2, EVP_BS, &EV.xnb, &EV.xidx);
    evp_partition(nyb - 2, EVP_BS, &EV.ynb, &EV.yidx);
    EV.nsub = EV.xnb * EV.ynb;
    const size_t ns = (size_t)EV.nsub;
    POP_CHECK_CUDA(cudaMalloc(&EV.d_xidx, sizeof(int) * EV.xidx.size()));
    POP_CHECK_CUDA(cudaMalloc(&EV.d_yidx, sizeof(int) * EV.yidx.size()));
    POP_CHECK_CUDA(cudaMalloc(&EV.d_land, sizeof(int) * ns));
    POP_CHECK_CUDA(cudaMalloc(&EV.d_cc, sizeof(double) * ns * EVP_LD * EVP_LD));
    POP_CHECK_CUDA(cudaMalloc(&EV.d_ne, sizeof(double) * ns * EVP_LD * EVP_LD));
    POP_CHECK_CUDA(cudaMalloc(&EV.d_ine, sizeof(double) * ns * EVP_LD * EVP_LD));
    POP_CHECK_CUDA(cudaMalloc(&EV.d_icc, sizeof(double) * ns * EVP_LD * EVP_LD));
    POP_CHECK_CUDA(cudaMalloc(&EV.d_rinv, sizeof(double) * ns * EVP_L * EVP_L));
    POP_CHECK_CUDA(cudaMemcpy(EV.d_xidx, EV.xidx.data(), sizeof(int) * EV.xidx.size(), cudaMemcpyHostToDevice));
    POP_CHECK_CUDA(cudaMemcpy(EV.d_yidx, EV.yidx.data(), sizeof(int) * EV.yidx.size(), cudaMemcpyHostToDevice));
    EV.ready = true;
  }
  const size_t ns = (size_t)EV.nsub, n2 = G.n2;
  std::vector<double> C(n2), NE(n2);
  POP_CHECK_CUDA(cudaMemcpyAsync(C.data(), fld("btropWgtCenter"), sizeof(double) * n2, cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaMemcpyAsync(NE.data(), fld("btropWgtNE"), sizeof(double) * n2, cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  // element-major copies: a[e * nsub + s]
  std::vector<double> cc(ns * EVP_LD * EVP_LD, 0.0), ne(ns * EVP_LD * EVP_LD, 0.0), ine(ns * EVP_LD * EVP_LD, 0.0),
      icc(ns * EVP_LD * EVP_LD, 0.0), rinv(ns * EVP_L * EVP_L, 0.0);
  std::vector<int> land(ns, 0);
  double worst = 0.0;
  int bad = 0;
  EV.nland = 0;
  double ecc[EVP_LD * EVP_LD], ene[EVP_LD * EVP_LD], er[EVP_L * EVP_L];
  for (int j = 1; j <= EV.ynb; j++) {
    const int js = EV.yidx[j] - 1, je = EV.yidx[j + 1], lm = je - js + 1;
    for (int i = 1; i <= EV.xnb; i++) {
      const int is = EV.xidx[i] - 1, ie = EV.xidx[i + 1], ln = ie - is + 1;
      const size_t sb = (size_t)(j - 1) * EV.xnb + (i - 1);
      memset(ecc, 0, sizeof(ecc));
      memset(ene, 0, sizeof(ene));
      // cc = btropWgtCenter(2:nx1, 2:ny1): element (ii,jj) of the sub-block is 0-based array cell (is+ii-1, js+jj-1)
      for (int jj = 1; jj <= lm; jj++)
        for (int ii = 1; ii <= ln; ii++) {
          const size_t q = (size_t)(js + jj - 1) * nxb + (is + ii - 1);
          HF2(ecc, ii, jj) = C[q];
          HF2(ene, ii, jj) = NE[q];
        }
      int iland = 0;
      double mn = fabs(HF2(ene, 2, 2));
      for (int jj = 2; jj <= lm - 1; jj++)
        for (int ii = 2; ii <= ln - 1; ii++) mn = fabs(HF2(ene, ii, jj)) < mn ? fabs(HF2(ene, ii, jj)) : mn;
      if (mn == 0.0) iland = 1;
      if (is + 2 < G.ib || ie - 1 > G.ie) iland = 2;
      if (js + 2 < G.jb || je - 1 > G.je) iland = 3;
      memset(er, 0, sizeof(er));
      if (iland > 0) {
        land[sb] = 1;
        EV.nland++;
      } else {
        const double chk = evp_block_pre_host(ecc, ene, er, ln, lm);
        worst = chk > worst ? chk : worst;
        if (chk > 1.0e-8) bad++;
      }
      for (int e = 0; e < EVP_LD * EVP_LD; e++) {
        cc[(size_t)e * ns + sb] = ecc[e];
        ne[(size_t)e * ns + sb] = ene[e];
        if (ecc[e] != 0.0) icc[(size_t)e * ns + sb] = 1.0 / ecc[e];
        if (ene[e] != 0.0) ine[(size_t)e * ns + sb] = 1.0 / ene[e];
      }
      for (int e = 0; e < EVP_L * EVP_L; e++) rinv[(size_t)e * ns + sb] = er[e];
    }
  }
  EV.selfcheck = worst;
  POP_REQUIRE(bad == 0, "POP_EXPLICITPRE: error in computing the inverse, error > 1.0e-8 (%g) in %d sub-blocks; check EVP "
              "sub-block size", worst, bad);  // POP_SolversMod.F90:2606-2614
  POP_CHECK_CUDA(cudaMemcpy(EV.d_land, land.data(), sizeof(int) * ns, cudaMemcpyHostToDevice));
  POP_CHECK_CUDA(cudaMemcpy(EV.d_cc, cc.data(), sizeof(double) * cc.size(), cudaMemcpyHostToDevice));
  POP_CHECK_CUDA(cudaMemcpy(EV.d_ne, ne.data(), sizeof(double) * ne.size(), cudaMemcpyHostToDevice));
  POP_CHECK_CUDA(cudaMemcpy(EV.d_ine, ine.data(), sizeof(double) * ine.size(), cudaMemcpyHostToDevice));
  POP_CHECK_CUDA(cudaMemcpy(EV.d_icc, icc.data(), sizeof(double) * icc.size(), cudaMemcpyHostToDevice));
  POP_CHECK_CUDA(cudaMemcpy(EV.d_rinv, rinv.data(), sizeof(double) * rinv.size(), cudaMemcpyHostToDevice));
  return POP_SUCCESS;
}
void evp_release() {
  cudaFree(EV.d_xidx); cudaFree(EV.d_yidx); cudaFree(EV.d_land); cudaFree(EV.d_cc); cudaFree(EV.d_ne);
  cudaFree(EV.d_ine); cudaFree(EV.d_icc); cudaFree(EV.d_rinv);
  EV = EvpDev();
}

struct EvpArgs {
  int nxb, xnb, ynb, nsub;
  const int *xidx, *yidx, *land;
  const double *cc, *ne, *ine, *icc, *rinv;
  const double* X;
  double* PX;
};
// ExplicitEvp (:2618-2696) for one sub-block.  N, M > 0: compile-time extents (everything unrolled, y in registers);
// N = M = 0: run-time extents n, m.  Element (i,j) (1-based) of a coefficient array is a[((j-1)*EVP_LD + (i-1)) * ns].
template <int N, int M>
__device__ __forceinline__ void evp_solve(const EvpArgs& a, int sb, int is, int js, int n_rt, int m_rt) {
  const int n = N ? N : n_rt, m = M ? M : m_rt;
  const int nm = n + m - 5;
  const size_t ns = (size_t)a.nsub;
  const double* cc = a.cc + sb;
  const double* ne = a.ne + sb;
  const double* ine = a.ine + sb;
  const double* rinv = a.rinv + sb;
#define CF(p, i, j) ldg((p) + (size_t)(((j)-1) * EVP_LD + ((i)-1)) * ns)
#define YY(i, j) y[((j)-1) * EVP_LD + ((i)-1)]
  const double* X = a.X + (size_t)(js - 1) * a.nxb + (is - 1);  // X(is,js): sub-block element (1,1), 1-based array indices
  double y[EVP_LD * EVP_LD];
#pragma unroll
  for (int e = 0; e < EVP_LD * EVP_LD; e++) y[e] = 0.0;
#pragma unroll
  for (int j = 2; j <= (M ? M : EVP_LD) - 1; j++)
    if (j <= m - 1) {
#pragma unroll
      for (int i = 2; i <= (N ? N : EVP_LD) - 1; i++)
        if (i <= n - 1)
          YY(i + 1, j + 1) = (ldg(X + (size_t)(j - 1) * a.nxb + (i - 1)) - CF(cc, i, j) * YY(i, j) -
                              CF(ne, i, j - 1) * YY(i + 1, j - 1) - CF(ne, i - 1, j) * YY(i - 1, j + 1) -
                              CF(ne, i - 1, j - 1) * YY(i - 1, j - 1)) * CF(ine, i, j);
    }
  double r[EVP_L];
#pragma unroll
  for (int q = 1; q <= EVP_L; q++) {
    double v = 0.0;
    if (q <= n - 2) v = YY((N ? q + 2 : q + 2), m);                       // r(1:n-2) = y(3:n, m)
    else if (q <= nm) v = YY(n, m - 1 - (q - (n - 1)));                  // r(n-1:nm) = y(n, m-1:3:-1)
    r[q - 1] = v;
  }
#pragma unroll
  for (int j = 1; j <= EVP_LD - 2; j++)
    if (j <= m - 2) {
      double acc = YY(2, m - j);
#pragma unroll
      for (int k = 1; k <= EVP_L; k++)
        if (k <= nm) acc = acc + ldg(rinv + (size_t)((j - 1) * EVP_L + (k - 1)) * ns) * r[k - 1];
      YY(2, m - j) = acc;
    }
#pragma unroll
  for (int i = 1; i <= EVP_LD - 3; i++)
    if (i <= n - 3) {
      double acc = YY(i + 2, 2);
#pragma unroll
      for (int k = 1; k <= EVP_L; k++)
        if (k <= nm) acc = acc + ldg(rinv + (size_t)((m - 2 + i - 1) * EVP_L + (k - 1)) * ns) * r[k - 1];
      YY(i + 2, 2) = acc;
    }
#pragma unroll
  for (int j = 2; j <= (M ? M : EVP_LD) - 2; j++)
    if (j <= m - 2) {
#pragma unroll
      for (int i = 2; i <= (N ? N : EVP_LD) - 2; i++)
        if (i <= n - 2)
          YY(i + 1, j + 1) = (ldg(X + (size_t)(j - 1) * a.nxb + (i - 1)) - CF(cc, i, j) * YY(i, j) -
                              CF(ne, i, j - 1) * YY(i + 1, j - 1) - CF(ne, i - 1, j) * YY(i - 1, j + 1) -
                              CF(ne, i - 1, j - 1) * YY(i - 1, j - 1)) * CF(ine, i, j);
    }
  double* PX = a.PX + (size_t)(js - 1) * a.nxb + (is - 1);
#pragma unroll
  for (int j = 2; j <= (M ? M : EVP_LD) - 1; j++)
    if (j <= m - 1) {
#pragma unroll
      for (int i = 2; i <= (N ? N : EVP_LD) - 1; i++)
        if (i <= n - 1) PX[(size_t)(j - 1) * a.nxb + (i - 1)] = YY(i, j);
    }
#undef CF
#undef YY
}
__global__ void __launch_bounds__(128)
evp_apply_kernel(const EvpArgs a) {
  const int sb = blockIdx.x * blockDim.x + threadIdx.x;
  if (sb >= a.nsub) return;
  const int bi = sb % a.xnb + 1, bj = sb / a.xnb + 1;
  // preconditioner :2333-2342: the frame of the sub-block in array indices (1-based)
  const int is = a.xidx[bi], ie = a.xidx[bi + 1] + 1, js = a.yidx[bj], je = a.yidx[bj + 1] + 1;
  const int ln = ie - is + 1, lm = je - js + 1;
  if (a.land[sb]) {
    const size_t ns = (size_t)a.nsub;
    for (int jj = 2; jj <= lm - 1; jj++)
      for (int ii = 2; ii <= ln - 1; ii++) {
        const size_t q = (size_t)(js + jj - 2) * a.nxb + (is + ii - 2);
        a.PX[q] = ldg(a.X + q) * ldg(a.icc + (size_t)((jj - 1) * EVP_LD + (ii - 1)) * ns + sb);
      }
    return;
  }
  if (ln == EVP_LD && lm == EVP_LD) evp_solve<EVP_LD, EVP_LD>(a, sb, is, js, ln, lm);
  else evp_solve<0, 0>(a, sb, is, js, ln, lm);
}
// PX = (PC) X on the sub-block interiors, 0 elsewhere (PX(:,:,bid) = 0 first, :2309)
static int precond_apply(double* PX, const double* X) {
  POP_REQUIRE(EV.ready, "POP_Solvers: the EVP preconditioner needs pop_solvers_prep (POP_SolversPrep) first");
  POP_CHECK_CUDA(cudaMemsetAsync(PX, 0, sizeof(double) * G.n2, G.stream));
  EvpArgs a;
  a.nxb = G.nxb; a.xnb = EV.xnb; a.ynb = EV.ynb; a.nsub = EV.nsub;
  a.xidx = EV.d_xidx; a.yidx = EV.d_yidx; a.land = EV.d_land;
  a.cc = EV.d_cc; a.ne = EV.d_ne; a.ine = EV.d_ine; a.icc = EV.d_icc; a.rinv = EV.d_rinv;
  a.X = X; a.PX = PX;
  POP_LAUNCH(evp_apply_kernel, (unsigned)((EV.nsub + 127) / 128), 128, 0, a);
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ ChronGear :1841-2266
static int chrongear(double* X, const double* B) {
  const int maxIt = G.cfg.max_iterations, freq = G.cfg.convergence_check_freq;
  double *R = fld("BT_R"), *S = fld("BT_S"), *Q = fld("BT_Q"), *Z = fld("BT_Z"), *AZ = fld("BT_AZ"),
         *A0R = fld("BT_A0R");
  double rr = 0.0;
  G.numIterations = maxIt;
  const bool evp = use_evp();
  BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 0);
  POP_TRY(bt_halo(R));
  if (evp) {  // :2009-2033: Z = (PC) r, halo update, S = Z
    POP_TRY(precond_apply(Z, R));
    POP_TRY(bt_halo(Z));
    POP_CHECK_CUDA(cudaMemcpyAsync(S, Z, sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
  } else {
    BT_ST(BT_CG_ZS, Z, R, A0R, S, 0.0, 0);
  }
  BT_ST(BT_CG_QSUM, Q, S, R, Z, 0.0, 1);
  POP_TRY(bt_halo(Q));
  POP_TRY(reduce_finish(2, RED_POST_CG_INIT, nullptr));
  bt_ew<EW_CG_UPDATE0>(X, R, S, Q);
  if (evp) POP_TRY(precond_apply(Z, R));
  else bt_ew<EW_MUL>(Z, R, A0R);
  for (int m = 1; m <= maxIt; m++) {
    POP_TRY(bt_halo(Z));
    BT_ST(BT_CG_AZSUM, AZ, Z, R, (double*)nullptr, 0.0, 1);
    POP_TRY(reduce_finish(2, RED_POST_CG_ITER, nullptr));
    const bool check = (m % freq == 0);
    bt_ew<EW_CG_UPDATE>(X, R, S, Q, Z, AZ, A0R, 0.0, 0.0, (check || evp) ? 0 : 1);
    if (check) {
      BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 1);
      POP_TRY(bt_halo(R));
      POP_TRY(reduce_finish(1, RED_POST_RR, &rr));
      if (rr < G.convergenceCriterion) {
        G.numIterations = m;
        break;
      }
      if (!evp) bt_ew<EW_MUL>(Z, R, A0R);
    }
    if (evp) POP_TRY(precond_apply(Z, R));
  }
  G.rmsResidual = sqrt(rr * G.residualNorm);
  POP_TRY(pop_post_launch("ChronGear"));
  POP_REQUIRE(!(G.numIterations == maxIt && G.convergenceCriterion != 0.0),
              "POP_SolversRun: ChronGear solver did not converge in %d iterations (rms residual %g)",
              maxIt, G.rmsResidual);
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ PCSI :1510-1835
// One fused kernel per iteration.  The reference iteration m is
//     halo(R) ; Q = om_m R + (csy om_m - 1) Q ; X = X + Q ; R = (B - A X) A0R        [check: sum r^2]
// and here the same arithmetic is cut one half-iteration later: kernel K(m) takes X_m and produces
//     r = B - A X_m  [check: sum r^2] ; R = r A0R ; Q_{m+1} = om_{m+1} R + (csy om_{m+1} - 1) Q_m ;
//     X_{m+1} = X_m + Q_{m+1}
// so R never touches memory and every operand is read once: X (stencil, through L1), B, Q, the four
// stored weights (A0R = 1/C is recomputed from the centre weight, bit-identical to EW_A0R) in, Q and
// the NEXT X out = 72 B per point instead of 13 array passes.  X is double-buffered because a
// neighbouring CTA still reads X_m while this one writes X_{m+1}.  Ghost cells: the reference
// refreshes R's ghost cells before the update, which makes Q and X ghost cells bitwise copies of
// their source points; here the kernel updates every point from the locally computed R (exactly what
// the reference does on ghost rows no message targets, i.e. closed boundaries) and one fused halo
// pass then overwrites the ghost cells of Q_{m+1} and X_{m+1} with those copies.
#define PC_TX 64
#define PC_TY 4
struct PcsiArgs {
  BtView v;
  const double* X;   // X_m
  double* Xn;        // X_{m+1}
  const double* Q;   // Q_m
  double* Qn;        // Q_{m+1} (other buffer)
  const double* B;
  double om, c1;     // om_{m+1}, csy*om_{m+1} - 1
  int advance;       // 0: last iteration (residual only)
  double* partials;  // [gridDim.x*gridDim.y][2] dd block partials of sum r^2 (SUM only)
  // single rank: ghost cells are filled in the same pass by evaluating the update at the cell the
  // halo update would copy from (east-west wrap, tripole fold), so no halo launch follows
  int map_ghost, do_ew, do_tripole, je0, nxg;
  const int *iglob, *jglob;
};
template <bool SUM>
__global__ void __launch_bounds__(PC_TX * PC_TY, 8)
pcsi_iter_kernel(const PcsiArgs a) {
  pdl_wait();
  pdl_trigger();
  const BtView& v = a.v;
  const int tx = threadIdx.x % PC_TX, ty = threadIdx.x / PC_TX;
  const int i = blockIdx.x * PC_TX + tx, j = blockIdx.y * PC_TY + ty;
  dd acc{0.0, 0.0};
  if (i < v.nxb && j < v.nyb) {
    const int nxb = v.nxb;
    const size_t qd = (size_t)j * nxb + i;  // cell written by this thread
    int is = i, js = j;                     // cell whose update it evaluates
    if (a.map_ghost) {
      if (a.do_tripole && j > a.je0) {
        int ig = a.nxg - a.iglob[i] + 1;
        if (ig == 0) ig = a.nxg;
        if (ig >= 1 && ig <= a.nxg) {
          is = POP_NGHOST - 1 + ig;
          js = 2 * a.je0 + 1 - j;
        }
      } else if (a.do_ew && (i < POP_NGHOST || i >= nxb - POP_NGHOST) && a.jglob[j] > 0) {
        is = (i < POP_NGHOST) ? i + (nxb - 2 * POP_NGHOST) : i - (nxb - 2 * POP_NGHOST);
      }
    }
    const bool ghost_copy = (is != i) || (js != j);
    {
    const int i = is, j = js;
    const size_t q = (size_t)j * nxb + i;
    const double c = ldg(v.C + q);
    double ax = 0.0;
    if (i >= 1 && i <= nxb - 2 && j >= 1 && j <= v.nyb - 2) {
      const double* X = a.X;  // written by the previous pass: L2 loads (ldcg), see pop_dev.cuh
      ax = c * ldcg(X + q) + ldg(v.N + q) * ldcg(X + q + nxb) + ldg(v.N + q - nxb) * ldcg(X + q - nxb) +
           ldg(v.E + q) * ldcg(X + q + 1) + ldg(v.E + q - 1) * ldcg(X + q - 1) +
           ldg(v.NE + q) * ldcg(X + q + nxb + 1) + ldg(v.NE + q - nxb) * ldcg(X + q - nxb + 1) +
           ldg(v.NE + q - 1) * ldcg(X + q + nxb - 1) + ldg(v.NE + q - nxb - 1) * ldcg(X + q - nxb - 1);
    }
    const double r = ldg(a.B + q) - ax;
    if (SUM && !ghost_copy && bt_physical(v, i, j)) acc = dd_add_d(acc, (r * r) * ldg(v.mask + q));
    if (a.advance) {
      const double a0r = (c != 0.0) ? 1.0 / c : 0.0;
      const double R = r * a0r;
      const double qv = a.om * R + a.c1 * ldcg(a.Q + q);
      a.Qn[qd] = qv;
      a.Xn[qd] = ldcg(a.X + q) + qv;
    }
    }
  }
  if (SUM) {
    dd rsum = block_reduce_dd(acc);
    if (threadIdx.x == 0) {
      const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      a.partials[b * 2] = rsum.hi;
      a.partials[b * 2 + 1] = rsum.lo;
    }
  }
}

// ---- two iterations per pass (temporal blocking) -----------------------------------------------
// K2(m): a CTA loads X_m on its tile plus a 2-cell ring, evaluates iteration m on the tile plus a 1-cell
// ring (X_{m+1} stays in shared memory), then iteration m+1 on the tile, and writes only Q_{m+2} and
// X_{m+2}: the seven input arrays and the two outputs cross HBM once per TWO iterations.  The values are
// the same bits as two K passes: a ring cell that is a ghost cell of the block is evaluated at the
// cell a halo update would copy from (east-west wrap / tripole mirror, read from global memory), and
// ghost rows owned by a neighbouring rank hold that rank's bits two cells deep.  One 2-level halo
// update of (X,Q) follows every pass (P > 1: one strip exchange per two iterations).
#define P2_TX 64
#ifndef P2_TY
#define P2_TY 10
#endif
#ifndef P2_MINB
#define P2_MINB 4
#endif
#define P2_NT 256
// every staged tile starts two cells west of the tile (an even column: TMA needs the first element of a
// box on a 16-byte boundary) and is P2_XW wide; X, N, NE start two rows south, the others one row south
#define P2_XW (P2_TX + 4)
#define P2_PAD16(n) (((n) + 15) / 16 * 16)  // TMA destinations start on 128-byte boundaries
#define P2_N_X P2_PAD16((P2_TY + 4) * P2_XW)
#define P2_N_1 P2_PAD16((P2_TY + 2) * P2_XW)
#define P2_N_2 P2_PAD16((P2_TY + 3) * P2_XW)
#define P2_SMEM_DOUBLES (P2_N_X + 5 * P2_N_1 + 2 * P2_N_2 + 16)
struct Pcsi2Args {
  BtView v;
  const double *X, *Q, *B;  // X_m, Q_m
  double *Xn, *Qn;          // X_{m+2}, Q_{m+2}
  double om1, c11, om2, c12;
  double* partials;
  int do_ew, do_tripole, je0, nxg;
  const int *iglob, *jglob;
  int use_tma;
  // tile rows of this launch: rowsel 0 -> blockIdx.y + row0; rowsel 1 -> the first and the last tile row (the rows a
  // neighbouring strip needs), launched ahead of the interior so that the exchange overlaps the interior tiles
  int rowsel, row0, nty;
  // deep strips (see pcsi()): rows above jw_max are not written by the tile pass; with `deep` the ghost cells that have
  // a source cell in this strip (east-west wrap, tripole fold) are written together with their source
  int deep, jw_max;
  int early_const;  // see the kernel head
  PopTmap tmX, tmC, tmB, tmQ, tmN, tmE, tmNE;  // 2-d tensor maps with the box of each staged tile
};
// cooperative asynchronous staging of a w x h window of a 2-d field (origin gi0,gj0; zero outside)
__device__ __forceinline__ void p2_stage(double* dst, const double* __restrict__ src, int w, int h, int gi0,
                                         int gj0, int nxb, int nyb, int tid) {
  for (int p = tid; p < w * h; p += P2_NT) {
    const int gi = gi0 + p % w, gj = gj0 + p / w;
    const bool in = (gi >= 0 && gi < nxb && gj >= 0 && gj < nyb);
    cp_async8(dst + p, src + (in ? (size_t)gj * nxb + gi : 0), in);
  }
}
template <bool SUM>
__global__ void __launch_bounds__(P2_NT, P2_MINB)
pcsi_iter2_kernel(const POP_GRID_CONSTANT Pcsi2Args a) {
  // Programmatic dependent launch: the CTAs of this pass may start while the last wave of the previous pass drains.
  // What that pass writes (X, Q) must not be touched before pdl_wait(); the five arrays that are constant during the
  // solve (B and the operator weights) are requested before it (a.early_const: not for the first pass of a solve, whose
  // predecessors are the kernels that set those arrays up).
  const bool early = a.use_tma && a.early_const;
  if (!early) pdl_wait();
  pdl_trigger();
  POP_DYN_SMEM(smem_raw);
  double* sX = (double*)smem_raw;  // X_m              rows j0-2 .. j0+TY+1
  double* sX1 = sX + P2_N_X;       // X_{m+1}          rows j0-1 .. j0+TY
  double* sC = sX1 + P2_N_1;       // centre weight    rows j0-1 ..
  double* sB = sC + P2_N_1;        // right-hand side
  double* sQ = sB + P2_N_1;        // Q_m, then Q_{m+1}
  double* sE = sQ + P2_N_1;        // east weight
  double* sN = sE + P2_N_1;        // north weight     rows j0-2 .. j0+TY
  double* sNE = sN + P2_N_2;       // NE weight        rows j0-2 .. j0+TY
  uint64_t* s_bar = (uint64_t*)(sNE + P2_N_2);
  const BtView& v = a.v;
  const int nxb = v.nxb, nyb = v.nyb;
  // rowsel = E > 0: the E first and the E last tile rows (gridDim.y = 2E)
  const int trow = a.rowsel ? ((int)blockIdx.y < a.rowsel ? (int)blockIdx.y : a.nty - 2 * a.rowsel + (int)blockIdx.y)
                            : (int)blockIdx.y + a.row0;
  const int i0 = POP_NGHOST + blockIdx.x * P2_TX, j0 = POP_NGHOST + trow * P2_TY;  // 0-based tile origin
  const int tid = threadIdx.x;
  // ---- every operand of both iterations is requested up front (one exposed memory latency per CTA):
  // seven TMA box copies issued by one thread (no per-element instructions), or per-thread cp.async
  // check passes: the mask values of this thread's cells are requested before the wait for the tiles
  double msk[(P2_TY + 2 + 3) / 4 + 1];
  if (SUM) {
    const int r0p = -1 + ((tid / P2_TX) * (P2_TY + 2)) / (P2_NT / P2_TX);
#pragma unroll
    for (int t = 0; t < (P2_TY + 2 + 3) / 4 + 1; t++) {
      const int gjm = j0 + r0p + t, gim = i0 + tid % P2_TX;
      msk[t] = (r0p + t >= 0 && r0p + t < P2_TY && bt_physical(v, gim, gjm)) ? ldg(v.mask + (size_t)gjm * nxb + gim) : 0.0;
    }
  }
  if (a.use_tma) {
    if (tid == 0) {
      mbar_init(s_bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
      constexpr uint32_t bytes = 8u * P2_XW * ((P2_TY + 4) + 4 * (P2_TY + 2) + 2 * (P2_TY + 3));
      mbar_expect_tx(s_bar, bytes);
      tma_load_box2d(sC, &a.tmC, i0 - 2, j0 - 1, s_bar);
      tma_load_box2d(sB, &a.tmB, i0 - 2, j0 - 1, s_bar);
      tma_load_box2d(sE, &a.tmE, i0 - 2, j0 - 1, s_bar);
      tma_load_box2d(sN, &a.tmN, i0 - 2, j0 - 2, s_bar);
      tma_load_box2d(sNE, &a.tmNE, i0 - 2, j0 - 2, s_bar);
      if (early) pdl_wait();
      tma_load_box2d(sX, &a.tmX, i0 - 2, j0 - 2, s_bar);
      tma_load_box2d(sQ, &a.tmQ, i0 - 2, j0 - 1, s_bar);
    }
    mbar_wait(s_bar, 0u);
    if (early) pdl_wait();  // every thread: the edge tiles read X, Q with plain loads as well, and all write X, Q
  } else {
    p2_stage(sX, a.X, P2_XW, P2_TY + 4, i0 - 2, j0 - 2, nxb, nyb, tid);
    p2_stage(sC, v.C, P2_XW, P2_TY + 2, i0 - 2, j0 - 1, nxb, nyb, tid);
    p2_stage(sB, a.B, P2_XW, P2_TY + 2, i0 - 2, j0 - 1, nxb, nyb, tid);
    p2_stage(sQ, a.Q, P2_XW, P2_TY + 2, i0 - 2, j0 - 1, nxb, nyb, tid);
    p2_stage(sE, v.E, P2_XW, P2_TY + 2, i0 - 2, j0 - 1, nxb, nyb, tid);
    p2_stage(sN, v.N, P2_XW, P2_TY + 3, i0 - 2, j0 - 2, nxb, nyb, tid);
    p2_stage(sNE, v.NE, P2_XW, P2_TY + 3, i0 - 2, j0 - 2, nxb, nyb, tid);
    cp_async_wait_all();
    __syncthreads();
  }
  // Thread <-> column ii of a strip of rows; it marches north keeping the 3x3 X window and the three
  // "southern" weights in registers, so an evaluation costs 11 shared-memory loads instead of 20.
  dd acc{0.0, 0.0};
  constexpr int NSTRIP = P2_NT / P2_TX;  // 4 strips of rows
  const int ii = tid % P2_TX, strip = tid / P2_TX;
  const int gi = i0 + ii;
  // where the first ghost ring of the block has a source cell the update is evaluated there (global memory)
  auto eval_at_source = [&](int gi_, int gj_, bool& mapped, double& x1) {
    int si = gi_, sj = gj_;
    if (a.do_tripole && gj_ > a.je0) {
      int ig = a.nxg - a.iglob[gi_] + 1;
      if (ig == 0) ig = a.nxg;
      if (ig >= 1 && ig <= a.nxg) { si = POP_NGHOST - 1 + ig; sj = 2 * a.je0 + 1 - gj_; }
    } else if (a.do_ew && (gi_ < POP_NGHOST || gi_ >= nxb - POP_NGHOST) && a.jglob[gj_] > 0) {
      si = (gi_ < POP_NGHOST) ? gi_ + (nxb - 2 * POP_NGHOST) : gi_ - (nxb - 2 * POP_NGHOST);
    }
    mapped = (si != gi_ || sj != gj_);
    if (!mapped) return;
    const size_t q = (size_t)sj * nxb + si;
    const double* X = a.X;  // written by the previous pass: L2 loads (ldcg), see pop_dev.cuh
    const double c = ldg(v.C + q);
    const double ax = c * ldcg(X + q) + ldg(v.N + q) * ldcg(X + q + nxb) + ldg(v.N + q - nxb) * ldcg(X + q - nxb) +
                      ldg(v.E + q) * ldcg(X + q + 1) + ldg(v.E + q - 1) * ldcg(X + q - 1) +
                      ldg(v.NE + q) * ldcg(X + q + nxb + 1) + ldg(v.NE + q - nxb) * ldcg(X + q - nxb + 1) +
                      ldg(v.NE + q - 1) * ldcg(X + q + nxb - 1) + ldg(v.NE + q - nxb - 1) * ldcg(X + q - nxb - 1);
    const double r = ldg(a.B + q) - ax;
    const double a0r = (c != 0.0) ? 1.0 / c : 0.0;
    x1 = ldcg(X + q) + (a.om1 * (r * a0r) + a.c11 * ldcg(a.Q + q));
  };
  const bool edge_cta = (i0 + P2_TX + 1 >= nxb - POP_NGHOST) || (i0 - 1 < POP_NGHOST) || (j0 - 1 < POP_NGHOST) ||
                        (j0 + P2_TY >= nyb - POP_NGHOST) ||  // the tile + ring touches ghost cells of the block
                        (a.do_tripole && j0 + P2_TY > a.je0);  // ... or the fold rows of a deep strip
  // ---- iteration m on the tile + ring: rows -1 .. TY split over the strips
  {
    constexpr int NR1 = P2_TY + 2;
    const int r0 = -1 + (strip * NR1) / NSTRIP, r1 = -1 + ((strip + 1) * NR1) / NSTRIP;  // [r0, r1)
    // window rows (r0-1, r0) of X_m at columns ii-1..ii+1; tile row jj is row jj+2 of sX
    const double* xr = sX + (r0 + 1) * P2_XW + (ii + 2);
    double xm0 = xr[-1], xm1 = xr[0], xm2 = xr[1];
    double xc0 = xr[P2_XW - 1], xc1 = xr[P2_XW], xc2 = xr[P2_XW + 1];
    double n_s = sN[(r0 + 1) * P2_XW + ii + 2];      // N(ii, r0-1): sN row jj+2
    double ne_s = sNE[(r0 + 1) * P2_XW + ii + 2];    // NE(ii, r0-1)
    double ne_sw = sNE[(r0 + 1) * P2_XW + ii + 1];   // NE(ii-1, r0-1)
    for (int jj = r0; jj < r1; jj++) {
      const int o1 = (jj + 1) * P2_XW + ii + 2;       // cell in the tiles whose first row is j0-1, first col i0-1
      const int oX = (jj + 1) * P2_XW + ii + 2;       // same for first col i0-2
      const double* xn = sX + (jj + 3) * P2_XW + (ii + 2);
      const double xp0 = xn[-1], xp1 = xn[0], xp2 = xn[1];
      const double c = sC[o1], n = sN[o1 + P2_XW], e = sE[oX], ew = sE[oX - 1];
      const double ne = sNE[oX + P2_XW], nw = sNE[oX + P2_XW - 1];
      const int gj = j0 + jj;
      double x1 = 0.0, qv = 0.0;
      if (gi >= 1 && gi <= nxb - 2 && gj >= 1 && gj <= nyb - 2) {
        bool mapped = false;
        if (edge_cta) eval_at_source(gi, gj, mapped, x1);
        if (!mapped) {
          const double ax = c * xc1 + n * xp1 + n_s * xm1 + e * xc2 + ew * xc0 + ne * xp2 + ne_s * xm2 + nw * xp0 +
                            ne_sw * xm0;
          const double r = sB[o1] - ax;
          if (SUM && jj >= 0 && jj < P2_TY && bt_physical(v, gi, gj)) acc = dd_add_d(acc, (r * r) * msk[jj - r0]);
          const double a0r = (c != 0.0) ? 1.0 / c : 0.0;
          qv = a.om1 * (r * a0r) + a.c11 * sQ[o1];
          x1 = xc1 + qv;
        }
      }
      sX1[o1] = x1;
      sQ[o1] = qv;  // Q_{m+1} of the cell (only this thread touches it before the barrier)
      xm0 = xc0; xm1 = xc1; xm2 = xc2;
      xc0 = xp0; xc1 = xp1; xc2 = xp2;
      n_s = n; ne_s = ne; ne_sw = nw;
    }
    // the two ring columns ii = -1 and ii = TX: one cell per thread, straight from the tiles
    if (tid < 2 * NR1) {
      const int jj = tid % NR1 - 1, ic = (tid < NR1) ? -1 : P2_TX;
      const int gic = i0 + ic, gj = j0 + jj;
      const int o1 = (jj + 1) * P2_XW + ic + 2, oX = (jj + 1) * P2_XW + ic + 2;
      double x1 = 0.0;
      if (gic >= 1 && gic <= nxb - 2 && gj >= 1 && gj <= nyb - 2) {
        bool mapped = false;
        if (edge_cta) eval_at_source(gic, gj, mapped, x1);
        if (!mapped) {
          const double* x = sX + (jj + 2) * P2_XW + (ic + 2);
          const double c = sC[o1];
          const double ax = c * x[0] + sN[o1 + P2_XW] * x[P2_XW] + sN[o1] * x[-P2_XW] + sE[oX] * x[1] +
                            sE[oX - 1] * x[-1] + sNE[oX + P2_XW] * x[P2_XW + 1] + sNE[oX] * x[-P2_XW + 1] +
                            sNE[oX + P2_XW - 1] * x[P2_XW - 1] + sNE[oX - 1] * x[-P2_XW - 1];
          const double r = sB[o1] - ax;
          const double a0r = (c != 0.0) ? 1.0 / c : 0.0;
          x1 = x[0] + (a.om1 * (r * a0r) + a.c11 * sQ[o1]);
        }
      }
      sX1[o1] = x1;
    }
  }
  __syncthreads();
  // ---- iteration m+1 on the tile (physical cells only): rows 0 .. TY-1 split over the strips
  {
    const int r0 = (strip * P2_TY) / NSTRIP, r1 = ((strip + 1) * P2_TY) / NSTRIP;
    // tile row jj is row jj+1 of sX1
    const double* xr = sX1 + r0 * P2_XW + (ii + 2);
    double xm0 = xr[-1], xm1 = xr[0], xm2 = xr[1];
    double xc0 = xr[P2_XW - 1], xc1 = xr[P2_XW], xc2 = xr[P2_XW + 1];
    double n_s = sN[(r0 + 1) * P2_XW + ii + 2];
    double ne_s = sNE[(r0 + 1) * P2_XW + ii + 2];
    double ne_sw = sNE[(r0 + 1) * P2_XW + ii + 1];
    for (int jj = r0; jj < r1; jj++) {
      const int o1 = (jj + 1) * P2_XW + ii + 2, oX = (jj + 1) * P2_XW + ii + 2;
      const double* xn = sX1 + (jj + 2) * P2_XW + (ii + 2);
      const double xp0 = xn[-1], xp1 = xn[0], xp2 = xn[1];
      const double c = sC[o1], n = sN[o1 + P2_XW], e = sE[oX], ew = sE[oX - 1];
      const double ne = sNE[oX + P2_XW], nw = sNE[oX + P2_XW - 1];
      const int gj = j0 + jj;
      if (gi <= nxb - POP_NGHOST - 1 && gj <= a.jw_max) {
        const double ax = c * xc1 + n * xp1 + n_s * xm1 + e * xc2 + ew * xc0 + ne * xp2 + ne_s * xm2 + nw * xp0 +
                          ne_sw * xm0;
        const double r = sB[o1] - ax;
        const double a0r = (c != 0.0) ? 1.0 / c : 0.0;
        const double qv = a.om2 * (r * a0r) + a.c12 * sQ[o1];
        const double xv = xc1 + qv;
        const size_t q = (size_t)gj * nxb + gi;
        a.Qn[q] = qv;
        a.Xn[q] = xv;
        if (a.deep && edge_cta) {  // the copies a halo update would make of this cell (edge tiles only)
          const int nxg = nxb - 2 * POP_NGHOST;
          if (a.do_ew && a.jglob[gj] > 0) {
            if (gi < 2 * POP_NGHOST) { a.Qn[q + nxg] = qv; a.Xn[q + nxg] = xv; }
            else if (gi >= nxg) { a.Qn[q - nxg] = qv; a.Xn[q - nxg] = xv; }
          }
          if (a.do_tripole && gj >= a.je0 - (POP_NGHOST - 1)) {
            const int t = a.nxg - a.iglob[gi] + 1;  // global column of the fold row cell this one is the source of
            const size_t qr = (size_t)(2 * a.je0 + 1 - gj) * nxb;
            a.Qn[qr + POP_NGHOST - 1 + t] = qv;
            a.Xn[qr + POP_NGHOST - 1 + t] = xv;
            if (a.do_ew) {
              if (t <= POP_NGHOST) { a.Qn[qr + nxb - POP_NGHOST - 1 + t] = qv; a.Xn[qr + nxb - POP_NGHOST - 1 + t] = xv; }
              else if (t > a.nxg - POP_NGHOST) { a.Qn[qr + t - (a.nxg - POP_NGHOST) - 1] = qv; a.Xn[qr + t - (a.nxg - POP_NGHOST) - 1] = xv; }
            }
          }
        }
      }
      xm0 = xc0; xm1 = xc1; xm2 = xc2;
      xc0 = xp0; xc1 = xp1; xc2 = xp2;
      n_s = n; ne_s = ne; ne_sw = nw;
    }
  }
  if (SUM) {
    dd rsum = block_reduce_dd(acc);
    if (threadIdx.x == 0) {
      const size_t b = (size_t)trow * gridDim.x + blockIdx.x;  // global tile index: the sum order does not depend on the split
      a.partials[b * 2] = rsum.hi;
      a.partials[b * 2 + 1] = rsum.lo;
    }
  }
}

// PCSI with usePreconditioner (:1651-1659, :1734-1742): the preconditioner couples the cells of a sub-block, so the
// residual goes through memory: evp_apply -> halo -> (Q, X) update -> residual, four launches per iteration.
static int pcsi_evp(double* X, const double* B) {
  const int maxIt = G.cfg.max_iterations, freq = G.cfg.convergence_check_freq,
            start = G.cfg.convergence_check_start;
  double *R = fld("BT_R"), *PR = fld("BT_S"), *Q = fld("BT_Q");
  double rr = 0.0;
  const double csalpha = 2.0 / (G.pcsiMaxEigs - G.pcsiMinEigs);
  const double csbeta = (G.pcsiMaxEigs + G.pcsiMinEigs) / (G.pcsiMaxEigs - G.pcsiMinEigs);
  const double csy = csbeta / csalpha;
  double csomga = 2.0 / csy;
  BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 0);
  POP_TRY(precond_apply(PR, R));
  bt_ew<EW_SCALE>(Q, PR, nullptr, nullptr, nullptr, nullptr, nullptr, 1.0 / csy);
  POP_TRY(bt_halo(Q));
  bt_ew<EW_ADD>(X, Q);
  BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 0);
  POP_TRY(bt_halo(R));
  G.numIterations = maxIt;
  for (int m = 1; m <= maxIt; m++) {
    csomga = 1.0 / (csy - csomga / (4.0 * csalpha * csalpha));
    POP_TRY(precond_apply(PR, R));
    POP_TRY(bt_halo(PR));
    bt_ew<EW_PCSI_QX>(Q, X, PR, nullptr, nullptr, nullptr, nullptr, csomga, csy * csomga - 1.0);
    const bool check = (m % freq == 0) && (m >= start);
    BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, check ? 1 : 0);
    if (check) {
      POP_TRY(reduce_finish(1, RED_POST_RR, &rr));
      if (rr < G.convergenceCriterion) {
        G.numIterations = m;
        break;
      }
    }
  }
  G.rmsResidual = sqrt(rr * G.residualNorm);
  return pop_post_launch("PCSI/EVP");
}

// ---- deep strips (P > 1) ------------------------------------------------------------------------------------
// On P strips every pass used to be followed by a strip exchange (a peer-memory kernel with a flag round trip): at
// 1/8 of the tx0.1 grid the exchange cost more than the pass.  The iteration therefore runs on private copies of
// its ten arrays whose strips carry `gd` ghost rows on either side instead of two: after an exchange all gd rows
// hold the owner's bits, and every iteration makes one more ghost row stale (the pass recomputes the ghost rows
// redundantly, bit for bit what their owner computes), so ONE exchange serves gd iterations.  Ghost cells with a
// source in the same strip (east-west wrap, tripole fold) are written by the pass itself.  The residual sum is a
// double-double sum over tiles, i.e. the correctly rounded exact sum whatever the tiling: shifting the tiles by the
// ghost depth changes neither a bit of the answer nor an iteration count (the tests run depths 8, 12 and 22).
struct DeepBt {
  double* buf = nullptr;  // levels: C N E NE mask B X0 Q0 X1 Q1 X2 Q2
  int* jglob = nullptr;
  int gd = 0, nyd = 0;
  size_t n2d = 0;
};
static DeepBt DB;
enum { DL_C = 0, DL_N, DL_E, DL_NE, DL_MASK, DL_B, DL_X0, DL_Q0, DL_X1, DL_Q1, DL_X2, DL_Q2, DL_COUNT };
void deep_release() {
  cudaFree(DB.buf);
  cudaFree(DB.jglob);
  DB = DeepBt();
}
static int deep_prepare(int gd) {
  if (DB.buf && DB.gd == gd && DB.nyd == G.ny_local + 2 * gd) return POP_SUCCESS;
  deep_release();
  DB.gd = gd;
  DB.nyd = G.ny_local + 2 * gd;
  DB.n2d = (size_t)G.nxb * DB.nyd;
  POP_CHECK_CUDA(cudaMalloc(&DB.buf, sizeof(double) * DB.n2d * DL_COUNT));
  POP_CHECK_CUDA(cudaMemset(DB.buf, 0, sizeof(double) * DB.n2d * DL_COUNT));
  std::vector<int> jg(DB.nyd);
  for (int j = 0; j < DB.nyd; j++) {  // as j_glob of the strip (pop_core.cu), gd rows deep
    int v = G.j0 + j - gd;
    if (v < 1) v = 0;
    else if (v > G.nyg) v = (G.cfg.ns_boundary_type == POP_BNDY_TRIPOLE) ? -v : 0;
    jg[j] = v;
  }
  POP_CHECK_CUDA(cudaMalloc(&DB.jglob, sizeof(int) * DB.nyd));
  POP_CHECK_CUDA(cudaMemcpy(DB.jglob, jg.data(), sizeof(int) * DB.nyd, cudaMemcpyHostToDevice));
  return POP_SUCCESS;
}
// plain strip (nyb rows) -> rows gd-2 .. of a deep level, and back
static int deep_load(int level, const double* src) {
  POP_CHECK_CUDA(cudaMemcpyAsync(DB.buf + level * DB.n2d + (size_t)(DB.gd - POP_NGHOST) * G.nxb, src, sizeof(double) * G.n2,
                                 cudaMemcpyDeviceToDevice, G.stream));
  return POP_SUCCESS;
}
static int deep_store(double* dst, const double* deep_level) {
  POP_CHECK_CUDA(cudaMemcpyAsync(dst, deep_level + (size_t)(DB.gd - POP_NGHOST) * G.nxb, sizeof(double) * G.n2,
                                 cudaMemcpyDeviceToDevice, G.stream));
  return POP_SUCCESS;
}

static int pcsi(double* X, const double* B) {
  if (use_evp()) return pcsi_evp(X, B);
  const int maxIt = G.cfg.max_iterations, freq = G.cfg.convergence_check_freq,
            start = G.cfg.convergence_check_start;
  double *R = fld("BT_R"), *A0R = fld("BT_A0R");
  double* W = fld("BT_PCSI");  // [X0, Q0, X1, Q1, X2, Q2]: (X,Q) pairs are 2-level fields for the halo pass
  double *Xb[3] = {W, W + 2 * G.n2, W + 4 * G.n2}, *Qb[3] = {W + G.n2, W + 3 * G.n2, W + 5 * G.n2};
  double rr = 0.0;
  const double csalpha = 2.0 / (G.pcsiMaxEigs - G.pcsiMinEigs);
  const double csbeta = (G.pcsiMaxEigs + G.pcsiMinEigs) / (G.pcsiMaxEigs - G.pcsiMinEigs);
  const double csy = csbeta / csalpha;
  double csomga = 2.0 / csy;
  // ---- prologue, as the reference (:1640-1700)
  double* Q = Qb[0];
  BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 0);
  bt_ew<EW_PCSI_INIT>(R, Q, A0R, nullptr, nullptr, nullptr, nullptr, 1.0 / csy);
  POP_TRY(bt_halo(Q));
  bt_ew<EW_ADD>(X, Q);
  BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 0);
  POP_TRY(bt_halo(R));
  bt_ew<EW_MUL>(R, R, A0R);
  POP_TRY(bt_halo(R));
  // ---- first half-iteration: Q_1, X_1 from the halo-updated preconditioned residual
  csomga = 1.0 / (csy - csomga / (4.0 * csalpha * csalpha));
  POP_CHECK_CUDA(cudaMemcpyAsync(Xb[0], X, sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
  bt_ew<EW_PCSI_QX>(Q, Xb[0], R, nullptr, nullptr, nullptr, nullptr, csomga, csy * csomga - 1.0);
  // the other pairs start as copies so that cells no pass writes (closed-boundary ghost rows) agree
  POP_CHECK_CUDA(cudaMemcpyAsync(Xb[1], Xb[0], sizeof(double) * 2 * G.n2, cudaMemcpyDeviceToDevice, G.stream));
  POP_CHECK_CUDA(cudaMemcpyAsync(Xb[2], Xb[0], sizeof(double) * 2 * G.n2, cudaMemcpyDeviceToDevice, G.stream));
  G.numIterations = maxIt;
  const bool blocking = !G.no_pcsi_blocking;
  // deep strips: see DeepBt.  Needs the peer-memory exchange, a closed or tripole north-south boundary and strips
  // at least gd rows high.
  // Depth: deep_halo rows (even: a pass consumes two).  Opt-in POP_B200_DEEP_HALO_WAVES=1: among the depths >= 6 the
  // largest one whose pass fits the fewest waves of CTAs (a 3600 x 300 strip with 12 ghost rows is 1824 tiles on 592
  // resident CTAs = 3.08 waves, with 6 rows 1767 = 2.98) -- measured slower: twice the exchanges cost more than the
  // tail of the fourth wave (profiles/r2_ncu_summary.md section 6).
  int gd = (G.deep_halo / 2) * 2;
  if (gd < 4) gd = 0;
  if (gd > 6 && getenv("POP_B200_DEEP_HALO_WAVES") && getenv("POP_B200_DEEP_HALO_WAVES")[0] == '1') {
    const long slots = (long)P2_MINB * G.sm_count, tx = (G.nxg + P2_TX - 1) / P2_TX;
    auto waves = [&](int d) { return ((long)((G.ny_local + 2 * d - 2 * POP_NGHOST + P2_TY - 1) / P2_TY) * tx + slots - 1) / slots; };
    int best = gd;
    for (int d = gd - 2; d >= 6; d -= 2)
      if (waves(d) < waves(best)) best = d;
    gd = best;
  }
  const bool deep = blocking && gd > 0 && G.ny_local >= gd && G.cfg.ns_boundary_type != POP_BNDY_CYCLIC &&
                    ((G.nranks > 1 && G.p2p_on && (size_t)8 * gd * G.nxg <= G.p2p_cap) || (G.nranks == 1 && G.deep_force));
  BtView view = bt_view();
  const int* jglob = G.d_jglob;
  int nyv = G.nyb;        // rows of the arrays the passes work on
  int je0 = G.je - 1;     // 0-based last physical row
  int valid = 0;          // deep: ghost rows that still hold the owner's bits
  if (deep) {
    if (getenv("POP_B200_TRACE")) fprintf(stderr, "[pop_b200] rank %d: P-CSI on deep strips, %d ghost rows\n", G.rank, gd);
    POP_TRY(deep_prepare(gd));
    nyv = DB.nyd; je0 = gd + G.ny_local - 1; jglob = DB.jglob;
    const double* srcs[8] = {view.C, view.N, view.E, view.NE, view.mask, B, Xb[0], Qb[0]};
    for (int l = 0; l < 8; l++) POP_TRY(deep_load(l, srcs[l]));
    POP_TRY(halo_exchange_deep(DB.buf, 8, gd, DB.n2d));
    for (int l = DL_X1; l <= DL_X2; l += 2)
      POP_CHECK_CUDA(cudaMemcpyAsync(DB.buf + l * DB.n2d, DB.buf + DL_X0 * DB.n2d, sizeof(double) * 2 * DB.n2d,
                                     cudaMemcpyDeviceToDevice, G.stream));
    view.nyb = DB.nyd; view.jb = gd + 1;
    view.C = DB.buf + DL_C * DB.n2d; view.N = DB.buf + DL_N * DB.n2d; view.E = DB.buf + DL_E * DB.n2d;
    view.NE = DB.buf + DL_NE * DB.n2d; view.mask = DB.buf + DL_MASK * DB.n2d;
    B = DB.buf + DL_B * DB.n2d;
    Xb[0] = DB.buf + DL_X0 * DB.n2d; Qb[0] = DB.buf + DL_Q0 * DB.n2d;
    Xb[1] = DB.buf + DL_X1 * DB.n2d; Qb[1] = DB.buf + DL_Q1 * DB.n2d;
    Xb[2] = DB.buf + DL_X2 * DB.n2d; Qb[2] = DB.buf + DL_Q2 * DB.n2d;
    valid = gd;
  }
  // One strip without a cyclic north-south boundary: every ghost cell of (X,Q) has its source in the strip, so the tile
  // pass writes the ghost cells itself (as on deep strips) and no halo kernel runs between two passes.
  const bool self_ghost = deep || (G.nranks == 1 && G.cfg.ns_boundary_type != POP_BNDY_CYCLIC &&
                                   !(getenv("POP_B200_PCSI_HALO_KERNEL") && getenv("POP_B200_PCSI_HALO_KERNEL")[0] == '1'));
  // rows the tile pass may write: on the last strip the rows above the top physical row belong to the fold (or are
  // closed-boundary zeros)
  const int jw_max = (self_ghost && G.rank == G.nranks - 1) ? je0 : nyv - POP_NGHOST - 1;
  const dim3 grid1((unsigned)((G.nxb + PC_TX - 1) / PC_TX), (unsigned)((nyv + PC_TY - 1) / PC_TY), 1);
  const dim3 grid2((unsigned)((G.nxg + P2_TX - 1) / P2_TX), (unsigned)((nyv - 2 * POP_NGHOST + P2_TY - 1) / P2_TY), 1);
  const int nblk1 = (int)(grid1.x * grid1.y), nblk2 = (int)(grid2.x * grid2.y);
  POP_TRY(reduce_reserve_partials(nblk1 > nblk2 ? nblk1 : nblk2));
  const int do_ew = (G.cfg.ew_boundary_type == POP_BNDY_CYCLIC) ? 1 : 0;
  const int do_tp = (G.cfg.ns_boundary_type == POP_BNDY_TRIPOLE && G.rank == G.nranks - 1) ? 1 : 0;
  // opt-in (POP_B200_OVERLAP_EXCHANGE=1): boundary tile rows first, strip exchange concurrent with the interior tiles
  // deep strips, opt-in POP_B200_DEEP_OVERLAP=1: interior tile rows of the pass after an exchange run beside the exchange
  // (bitwise the same; measured no faster: with one exchange per six passes the two extra launches and events cost
  // what the overlap hides)
  const bool deep_overlap = deep && (G.nranks > 1 || G.deep_force) && grid2.y >= 6 && G.stream_x != nullptr &&
                            getenv("POP_B200_DEEP_OVERLAP") && getenv("POP_B200_DEEP_OVERLAP")[0] == '1';
  const bool xover = !deep && blocking && G.overlap_exchange && G.nranks > 1 && G.p2p_on && grid2.y >= 3 &&
                     G.cfg.ew_boundary_type == POP_BNDY_CYCLIC;
  // Convergence checks.  One rank: the host reads rr right away.  P > 1 ranks: a check costs an all-gather
  // and a host round trip on every rank, so its verdict is read one check period LATER, when it has long
  // arrived; X_m of the check is kept (its buffer pair sits out of the rotation), and when a check turns out to have
  // converged the answer is that X_m and numIterations is that m -- the same bits as stopping immediately, a few
  // passes late.
  // (POP_B200_SYNC_CHECKS=1: immediate verdicts on P ranks too; POP_B200_LAGGED_CHECKS=1: lagged on one rank, a test aid)
  const bool lagged = ((G.nranks > 1) && !(getenv("POP_B200_SYNC_CHECKS") && getenv("POP_B200_SYNC_CHECKS")[0] == '1')) ||
                      (getenv("POP_B200_LAGGED_CHECKS") && getenv("POP_B200_LAGGED_CHECKS")[0] == '1');
  struct { bool valid; int m, slot, buf; } pend = {false, 0, 0, -1};
  const double* result = nullptr;
  if (lagged && !G.ev_chk[0]) {
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_chk[0], cudaEventDisableTiming));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_chk[1], cudaEventDisableTiming));
    POP_CHECK_CUDA(cudaEventCreateWithFlags(&G.ev_chk_go, cudaEventDisableTiming));
    int lo = 0, hi = 0;
    POP_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    POP_CHECK_CUDA(cudaStreamCreateWithPriority(&G.stream_chk, cudaStreamNonBlocking, hi));
  }
  // tensor maps of the staged tiles (X and Q for both buffers); without TMA the kernel stages with cp.async
  PopTmap tmXb[3], tmQb[3], tmC, tmB, tmN, tmE, tmNE;
  bool use_tma = blocking && !G.no_tma;
  if (use_tma) {
    const BtView& bv = view;
    use_tma = make_tmap_2d(&tmXb[0], Xb[0], P2_XW, P2_TY + 4, nyv) && make_tmap_2d(&tmXb[1], Xb[1], P2_XW, P2_TY + 4, nyv) &&
              make_tmap_2d(&tmXb[2], Xb[2], P2_XW, P2_TY + 4, nyv) && make_tmap_2d(&tmQb[2], Qb[2], P2_XW, P2_TY + 2, nyv) &&
              make_tmap_2d(&tmQb[0], Qb[0], P2_XW, P2_TY + 2, nyv) && make_tmap_2d(&tmQb[1], Qb[1], P2_XW, P2_TY + 2, nyv) &&
              make_tmap_2d(&tmC, bv.C, P2_XW, P2_TY + 2, nyv) && make_tmap_2d(&tmB, B, P2_XW, P2_TY + 2, nyv) &&
              make_tmap_2d(&tmN, bv.N, P2_XW, P2_TY + 3, nyv) && make_tmap_2d(&tmE, bv.E, P2_XW, P2_TY + 2, nyv) &&
              make_tmap_2d(&tmNE, bv.NE, P2_XW, P2_TY + 3, nyv);
  }
#ifndef POP_EMUL
  if (blocking) {
    const int smem2 = (int)(sizeof(double) * P2_SMEM_DOUBLES);
    POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)pcsi_iter2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)pcsi_iter2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  }
#endif
  // (X_m, Q_m) live in pair `cur`, a pass writes pair `nxt`.  With lagged checks X_m of a check must outlive its
  // verdict: the pair is simply retired from the rotation until then (pend.buf: one at a time, the next check frees it),
  // so three pairs suffice and nothing is copied.
  int cur = 0, nxt = 1;
  int m = 1, npass = 0, npass2 = 0;
  while (m <= maxIt) {
    const bool check = (m % freq == 0) && (m >= start);
    const bool next_is_check = ((m + 1) % freq == 0) && (m + 1 >= start);
    int adv;  // iterations this pass advances X by
    int nblk;
    if (blocking && m + 2 <= maxIt && !next_is_check) {
      if (check && lagged && pend.valid) POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_chk[pend.slot], 0));
      // Deep strips: the pass after an exchange is split.  Only the two outermost tile rows on either side read ghost rows
      // the exchange rewrites (a tile reads two rows beyond itself), so the interior tile rows run beside the exchange
      // (on its own high-priority stream, launched first) and the edge rows follow it.
      bool split = false;
      if (deep && valid < 2) {  // the next pass consumes two ghost rows
        if (deep_overlap) {
          POP_CHECK_CUDA(cudaEventRecord(G.ev_xb, G.stream));
          POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream_x, G.ev_xb, 0));
          std::swap(G.stream, G.stream_x);
          const int rcx = halo_exchange_deep(Xb[cur], 2, gd, DB.n2d);
          std::swap(G.stream, G.stream_x);
          POP_TRY(rcx);
          POP_CHECK_CUDA(cudaEventRecord(G.ev_xx, G.stream_x));
          split = true;
        } else {
          POP_TRY(halo_exchange_deep(Xb[cur], 2, gd, DB.n2d));
        }
        valid = gd;
      }
      Pcsi2Args a;
      a.v = view;
      a.deep = self_ghost ? 1 : 0; a.jw_max = jw_max;
      a.early_const = (npass2++ > 0 && !G.no_pdl) ? 1 : 0;
      a.X = Xb[cur]; a.Q = Qb[cur]; a.B = B; a.Xn = Xb[nxt]; a.Qn = Qb[nxt];
      csomga = 1.0 / (csy - csomga / (4.0 * csalpha * csalpha));  // om_{m+1}
      a.om1 = csomga; a.c11 = csy * csomga - 1.0;
      csomga = 1.0 / (csy - csomga / (4.0 * csalpha * csalpha));  // om_{m+2}
      a.om2 = csomga; a.c12 = csy * csomga - 1.0;
      a.partials = G.d_partials_big;
      a.do_ew = do_ew; a.do_tripole = do_tp; a.je0 = je0; a.nxg = G.nxg;
      a.iglob = G.d_iglob; a.jglob = jglob;
      a.use_tma = use_tma ? 1 : 0;
      a.rowsel = 0; a.row0 = 0; a.nty = (int)grid2.y;
      if (use_tma) {
        a.tmX = tmXb[cur]; a.tmQ = tmQb[cur]; a.tmC = tmC; a.tmB = tmB; a.tmN = tmN; a.tmE = tmE; a.tmNE = tmNE;
      }
      const size_t smem2 = sizeof(double) * P2_SMEM_DOUBLES;
      if (xover) {
        // boundary tile rows first; their rows go to the neighbours on the exchange stream while the interior tiles run
        a.rowsel = 1;
        if (check) POP_LAUNCH(pcsi_iter2_kernel<true>, dim3(grid2.x, 2, 1), P2_NT, smem2, a);
        else POP_LAUNCH(pcsi_iter2_kernel<false>, dim3(grid2.x, 2, 1), P2_NT, smem2, a);
        POP_CHECK_CUDA(cudaEventRecord(G.ev_xb, G.stream));
        a.rowsel = 0; a.row0 = 1;
        if (check) POP_LAUNCH(pcsi_iter2_kernel<true>, dim3(grid2.x, grid2.y - 2, 1), P2_NT, smem2, a);
        else POP_LAUNCH(pcsi_iter2_kernel<false>, dim3(grid2.x, grid2.y - 2, 1), P2_NT, smem2, a);
        POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream_x, G.ev_xb, 0));
        std::swap(G.stream, G.stream_x);
        const int rcx = halo_exchange_rows(Xb[nxt], 2);
        std::swap(G.stream, G.stream_x);
        POP_TRY(rcx);
        POP_CHECK_CUDA(cudaEventRecord(G.ev_xx, G.stream_x));
        POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_xx, 0));
        POP_TRY(halo_ew_own_rows(Xb[nxt], 2));  // east-west ghost columns of the rows the interior tiles wrote
        adv = 2;
        nblk = nblk2;
      } else if (split) {
        constexpr int E = 2;
        a.rowsel = 0; a.row0 = E;
        if (check) POP_LAUNCH(pcsi_iter2_kernel<true>, dim3(grid2.x, grid2.y - 2 * E, 1), P2_NT, smem2, a);
        else POP_LAUNCH(pcsi_iter2_kernel<false>, dim3(grid2.x, grid2.y - 2 * E, 1), P2_NT, smem2, a);
        POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_xx, 0));
        a.rowsel = E; a.row0 = 0;
        if (check) POP_LAUNCH(pcsi_iter2_kernel<true>, dim3(grid2.x, 2 * E, 1), P2_NT, smem2, a);
        else POP_LAUNCH(pcsi_iter2_kernel<false>, dim3(grid2.x, 2 * E, 1), P2_NT, smem2, a);
        valid -= 2;
        adv = 2;
        nblk = nblk2;
      } else {
      {
        // CUDA-event sample of the pass kernel alone (every 16th pass, uniformly over the solve; the solver
        // timers are otherwise off)
        const bool sample = (!check && (npass++ % 16) == 8);
        if (sample) G.timer_suppress--;
        {
          ScopedTimer tk("PCSI_PASS2_KERNEL");
          if (check) POP_LAUNCH_PDL(pcsi_iter2_kernel<true>, grid2, P2_NT, smem2, a);
          else POP_LAUNCH_PDL(pcsi_iter2_kernel<false>, grid2, P2_NT, smem2, a);
        }
        if (sample) G.timer_suppress++;
      }
      if (deep) valid -= 2;
      else if (!self_ghost) POP_TRY(halo_update(Xb[nxt], 2, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
      adv = 2;
      nblk = nblk2;
      }
    } else {
      csomga = 1.0 / (csy - csomga / (4.0 * csalpha * csalpha));  // om_{m+1}
      if (check && lagged && pend.valid) POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_chk[pend.slot], 0));
      if (deep && valid < 1) {
        POP_TRY(halo_exchange_deep(Xb[cur], 2, gd, DB.n2d));
        valid = gd;
      }
      PcsiArgs a;
      a.v = view;
      a.X = Xb[cur]; a.Xn = Xb[nxt]; a.Q = Qb[cur]; a.Qn = Qb[nxt]; a.B = B;
      a.om = csomga; a.c1 = csy * csomga - 1.0;
      a.advance = (m < maxIt) ? 1 : 0;
      a.partials = G.d_partials_big;
      // one rank: the pass itself fills the ghost cells it can map to a source cell (east-west wrap, tripole
      // fold); north-south cyclic rows come from the halo update like the rows of a neighbouring strip
      a.map_ghost = (deep || (G.nranks == 1 && G.cfg.ns_boundary_type != POP_BNDY_CYCLIC)) ? 1 : 0;
      a.do_ew = do_ew; a.do_tripole = do_tp;
      a.je0 = je0; a.nxg = G.nxg; a.iglob = G.d_iglob; a.jglob = jglob;
      if (check) POP_LAUNCH_PDL(pcsi_iter_kernel<true>, grid1, PC_TX * PC_TY, 0, a);
      else POP_LAUNCH_PDL(pcsi_iter_kernel<false>, grid1, PC_TX * PC_TY, 0, a);
      if (deep) valid -= a.advance;
      else if (a.advance && !a.map_ghost)
        POP_TRY(halo_update(Xb[nxt], 2, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
      adv = a.advance;
      nblk = nblk1;
    }
    if (check && !lagged) {
      POP_TRY(reduce_finish_n(1, RED_POST_RR, &rr, G.d_partials_big, nblk));
      if (rr < G.convergenceCriterion) {
        G.numIterations = m;
        break;
      }
    } else if (check) {
      if (pend.valid) {  // verdict of the previous check
        POP_CHECK_CUDA(cudaEventSynchronize(G.ev_chk[pend.slot]));
        rr = G.h_sums[2 + pend.slot];
        if (rr < G.convergenceCriterion) {
          G.numIterations = pend.m;
          result = Xb[pend.buf];
          break;
        }
      }
      // the sum over tiles and ranks, the all-gather and the copy of the verdict to the host run on their own
      // stream beside the following passes; X_m stays where it is (its pair leaves the rotation)
      const int slot = pend.valid ? (pend.slot ^ 1) : 0;
      POP_CHECK_CUDA(cudaEventRecord(G.ev_chk_go, G.stream));
      POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream_chk, G.ev_chk_go, 0));
      std::swap(G.stream, G.stream_chk);
      int rcc = reduce_finish_n(1, RED_POST_RR, nullptr, G.d_partials_big, nblk);
      if (rcc == POP_SUCCESS &&
          (cudaMemcpyAsync(G.h_sums + 2 + slot, G.d_sums, sizeof(double), cudaMemcpyDeviceToHost, G.stream) != cudaSuccess ||
           cudaEventRecord(G.ev_chk[slot], G.stream) != cudaSuccess))
        rcc = POP_FAIL;
      std::swap(G.stream, G.stream_chk);
      POP_REQUIRE(rcc == POP_SUCCESS, "PCSI: convergence check could not be queued");
      pend.valid = true; pend.m = m; pend.slot = slot; pend.buf = cur;
    }
    if (adv) {  // rotate: the pair just read is free again unless a pending check holds it
      const int old = cur;
      cur = nxt;
      nxt = (pend.valid && pend.buf == old) ? 3 - old - cur : old;
    }
    m += (adv > 0) ? adv : 1;
  }
  if (lagged && !result && pend.valid) {  // the loop ran out: the last check may still have converged
    POP_CHECK_CUDA(cudaEventSynchronize(G.ev_chk[pend.slot]));
    rr = G.h_sums[2 + pend.slot];
    if (rr < G.convergenceCriterion) {
      G.numIterations = pend.m;
      result = Xb[pend.buf];
    }
  }
  // the answer is X_m of the converged (or last) residual evaluation
  if (deep) {  // physical rows back to the caller's strip; its ghost cells as a halo update leaves them
    POP_TRY(deep_store(X, result ? result : Xb[cur]));
    POP_TRY(bt_halo(X));
  } else {
    POP_CHECK_CUDA(cudaMemcpyAsync(X, result ? result : Xb[cur], sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
  }
  G.rmsResidual = sqrt(rr * G.residualNorm);
  return pop_post_launch("PCSI");  // PCSI returns silently when not converged (:1828-1830)
}

// ------------------------------------------------------------------ pcg :1200-1503
static int pcg(double* X, const double* B) {
  const int maxIt = G.cfg.max_iterations, freq = G.cfg.convergence_check_freq;
  double *R = fld("BT_R"), *S = fld("BT_S"), *Q = fld("BT_Q"), *W1 = fld("BT_Z");
  double rr = 0.0;
  BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 0);
  bt_ew<EW_SET>(S, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0);
  POP_TRY(bt_halo(R));
  SolverScalars sc0;
  memset(&sc0, 0, sizeof(sc0));
  sc0.eta0 = 1.0;
  POP_CHECK_CUDA(cudaMemcpyAsync(G.d_scal, &sc0, sizeof(sc0), cudaMemcpyHostToDevice, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  G.numIterations = maxIt;
  for (int m = 1; m <= maxIt; m++) {
    if (use_evp()) {  // :1310-1345
      POP_TRY(precond_apply(W1, R));
      bt_ew<EW_DOT>(R, W1);
      POP_TRY(bt_halo(W1));
    } else {
      bt_ew<EW_PCG_W1>(W1, R, fld("btropWgtCenter"));
    }
    POP_TRY(reduce_finish(1, RED_POST_PCG_ETA1, nullptr));
    bt_ew<EW_PCG_S>(S, W1);
    BT_ST(BT_PCG_AS, Q, S, (const double*)nullptr, (double*)nullptr, 0.0, 1);
    POP_TRY(bt_halo(Q));
    POP_TRY(reduce_finish(1, RED_POST_PCG_ETA2, nullptr));
    bt_ew<EW_PCG_UPDATE>(X, R, S, Q);
    if (m % freq == 0) {
      BT_ST(BT_RESID, R, X, B, (double*)nullptr, 0.0, 1);
      POP_TRY(bt_halo(R));
      POP_TRY(reduce_finish(1, RED_POST_RR, &rr));
      if (rr < G.convergenceCriterion) {
        G.numIterations = m;
        break;
      }
    }
  }
  G.rmsResidual = sqrt(rr * G.residualNorm);
  POP_TRY(pop_post_launch("pcg"));
  POP_REQUIRE(!(G.numIterations == maxIt && G.convergenceCriterion != 0.0),
              "POP_SolversRun: pcg solver did not converge in %d iterations (rms residual %g)", maxIt,
              G.rmsResidual);
  return POP_SUCCESS;
}

// ------------------------------------------------------------------ Lanczos :2699-2990, ratqr :3122-3222
static int ratqr(int n, double eps1, const double* d, const double* e, double* mineig) {
  std::vector<double> bd(n + 1, 0.0), w(n + 1, 0.0);
  double f, ep, delta = 0.0, p, q = 0.0, qp, r, s = 0.0, tot;
  for (int i = 1; i <= n; i++) w[i] = d[i - 1];
  tot = w[1];
  for (int i = 1; i <= n; i++) {
    p = q;
    bd[i] = e[i - 1] * e[i - 1];
    q = 0.0;
    if (i != n) q = fabs(e[i]);
    const double c = w[i] - p - q;
    tot = (c < tot) ? c : tot;
  }
  bd[1] = 0.0;
  if (tot < 0.0) tot = 0.0;
  else
    for (int i = 1; i <= n; i++) w[i] = w[i] - tot;
  for (;;) {
    tot = tot + s;
    delta = w[n] - s;
    if (delta <= eps1) break;
    f = bd[n] / delta;
    qp = delta + f;
    p = 1.0;
    for (int ii = 1; ii <= n - 1; ii++) {
      const int i = n - ii;
      q = w[i] - s - f;
      r = q / qp;
      p = p * r + 1.0;
      ep = f * r;
      w[i + 1] = qp + ep;
      delta = q - ep;
      if (delta <= eps1) break;
      f = bd[i] / q;
      qp = delta + f;
      bd[i + 1] = qp * ep;
    }
    if (delta <= eps1) break;
    w[1] = qp;
    s = qp / p;
    if (tot + s <= tot) return 1;
  }
  *mineig = tot;
  return 0;
}

static int pcsi_lanczos() {
  const int maxstep = G.cfg.max_lanczos_step;
  double *R = fld("BT_R"), *S = fld("BT_S"), *Q = fld("BT_Q"), *Q1 = fld("BT_Z"), *P = fld("BT_AZ"),
         *A0R = fld("BT_A0R");
  std::vector<double> vcsa(maxstep + 1, 0.0), vcsb(maxstep + 1, 0.0);
  double csa, csb, csc, mineig, u = 0.0, v = 0.0, sum = 0.0;
  bt_ew<EW_SET>(R, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1.0);
  bt_ew<EW_SET>(Q, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0);
  bt_ew<EW_SET>(Q1, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0);
  const bool evp = use_evp();
  if (evp) {
    POP_TRY(precond_apply(S, R));
    bt_ew<EW_DOT>(S, R);
  } else {
    bt_ew<EW_LZ_SR>(S, R, A0R);
  }
  POP_TRY(reduce_finish(1, RED_POST_NONE, &sum));
  csc = -sum;
  POP_REQUIRE(csc > 0.0, "PcsiLanczos: preconditioned operator is not negative definite (csc=%g)", csc);
  bt_ew<EW_SCALE>(Q, R, nullptr, nullptr, nullptr, nullptr, nullptr, 1 / sqrt(csc));
  POP_TRY(bt_halo(Q));
  csb = 0.0;
  mineig = 1.0;
  G.lanczosSteps = 0;
  for (int m = 1; m <= maxstep; m++) {
    G.lanczosSteps = m;
    if (evp) POP_TRY(precond_apply(P, Q));
    else bt_ew<EW_MUL>(P, Q, A0R);
    POP_TRY(bt_halo(P));
    BT_ST(BT_LZ_AP, R, P, Q1, (double*)nullptr, csb, 1);
    POP_TRY(reduce_finish(1, RED_POST_NONE, &sum));
    csa = -sum;
    bt_ew<EW_AXPY>(R, Q, nullptr, nullptr, nullptr, nullptr, nullptr, csa);
    POP_TRY(bt_halo(R));
    if (evp) {
      POP_TRY(precond_apply(S, R));
      bt_ew<EW_DOT>(S, R);
    } else {
      bt_ew<EW_LZ_SR>(S, R, A0R);
    }
    POP_TRY(reduce_finish(1, RED_POST_NONE, &sum));
    csc = -sum;
    csb = sqrt(csc);
    vcsa[m] = csa;
    vcsb[m] = csb;
    if (m == 1) u = vcsa[1] + vcsb[1];
    else {
      const double c = vcsa[m] + vcsb[m] + vcsb[m - 1];
      u = (u > c) ? u : c;
    }
    POP_REQUIRE(csb != 0.0, "PcsiLanczos: breakdown (csb = 0) at step %d", m);
    bt_ew<EW_LZ_SHIFT>(Q1, Q, R, nullptr, nullptr, nullptr, nullptr, 1 / csb);
    if (m % 10 == 0 || m == maxstep) {
      std::vector<double> mcsa(m + 1, 0.0), mcsb(m + 1, 0.0);
      for (int i = 1; i <= m - 1; i++) { mcsa[i - 1] = vcsa[i]; mcsb[i] = vcsb[i]; }
      mcsa[m - 1] = vcsa[m];
      mcsb[0] = 0.0;
      POP_REQUIRE(ratqr(m, 1.0e-8, mcsa.data(), mcsb.data(), &v) == 0, "PcsiLanczos: ratqr failed");
      if (fabs(1 - v / mineig) < G.cfg.lanczos_convergence_criterion) break;
      mineig = v;
    }
  }
  G.pcsiMaxEigs = u;
  G.pcsiMinEigs = v;
  return pop_post_launch("PcsiLanczos");
}

int solvers_prep_dev() {
  POP_TRY(refresh_center());
  if (use_evp()) POP_TRY(evp_prep_dev());  // POP_SolversPrep :252-292, before the Lanczos estimate that uses it
  if (G.cfg.solver_choice == POP_SOLVER_PCSI) return pcsi_lanczos();
  return POP_SUCCESS;
}
int solvers_evp_diagnostics(int* nsub, int* nland, double* selfcheck) {
  POP_REQUIRE(EV.ready, "POP_SolversGetEvpDiagnostics: the EVP preconditioner has not been prepared");
  if (nsub) *nsub = EV.nsub;
  if (nland) *nland = EV.nland;
  if (selfcheck) *selfcheck = EV.selfcheck;
  return POP_SUCCESS;
}

int solvers_run_dev(double* X, const double* B) {
  ScopedTimer tm("SOLVER");
  POP_TRY(refresh_center());
  G.timer_suppress++;
  int rc;
  switch (G.cfg.solver_choice) {
    case POP_SOLVER_PCG: rc = pcg(X, B); break;
    case POP_SOLVER_PCSI: rc = pcsi(X, B); break;
    default: rc = chrongear(X, B); break;
  }
  G.timer_suppress--;
  return rc;
}

// ------------------------------------------------------------------ barotropic_driver :267-735
struct BtDrv {
  const double *ZX, *ZY, *GXc, *GXo, *GYc, *GYo, *FCOR, *HU, *UBo, *VBo, *RCALCT, *TAREA, *Pc, *FW, *PGUESS,
      *indep;
  const int *CHECKER, *CONSTNT;
  double *UH, *VH, *W3, *W4, *RHS, *dc, *cwc, *Pn, *GXn, *GYn, *UBn, *VBn;
  double c2dtp, beta, gamma, dtp, rcheck, rconst;
  int leapfrog, impcor, sfc;
  size_t n;
};
__global__ void bt_rhs1_kernel(BtDrv d) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= d.n) return;
  double w3, w4;
  if (d.leapfrog) {
    w3 = d.c2dtp * (d.ZX[q] - d.gamma * d.GXc[q] - (1.0 - d.gamma) * d.GXo[q]);
    w4 = d.c2dtp * (d.ZY[q] - d.gamma * d.GYc[q] - (1.0 - d.gamma) * d.GYo[q]);
  } else {
    w3 = d.c2dtp * (d.ZX[q] - d.GXc[q]);
    w4 = d.c2dtp * (d.ZY[q] - d.GYc[q]);
  }
  double uh, vh;
  if (d.impcor) {
    const double w1 = d.c2dtp * d.beta * d.FCOR[q];
    const double w2 = 1.0 / (1.0 + w1 * w1);
    uh = w2 * (w3 + w1 * w4) + d.UBo[q];
    vh = w2 * (w4 - w1 * w3) + d.VBo[q];
  } else {
    uh = w3 + d.UBo[q];
    vh = w4 + d.VBo[q];
  }
  d.UH[q] = uh;
  d.VH[q] = vh;
  const double gx = d.leapfrog ? d.GXo[q] : d.GXc[q], gy = d.leapfrog ? d.GYo[q] : d.GYc[q];
  d.W3[q] = d.HU[q] * (uh + d.beta * d.c2dtp * gx);
  d.W4[q] = d.HU[q] * (vh + d.beta * d.c2dtp * gy);
}
__global__ void bt_rhs2_kernel(BtDrv d) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= d.n) return;
  double rhs = d.RHS[q] / (d.beta * d.c2dtp);
  double dc = 0.0;
  if (d.sfc == POP_SFC_VARTHICK) {
    dc = (d.RCALCT[q] != 0.0) ? d.TAREA[q] / (d.beta * d.c2dtp * d.dtp * POP_GRAV) : 0.0;
    rhs = rhs - dc * d.Pc[q] - d.FW[q] * d.TAREA[q] / (d.beta * d.c2dtp);
  } else if (d.sfc != POP_SFC_RIGID) {
    dc = (d.RCALCT[q] != 0.0) ? d.TAREA[q] / (d.beta * d.c2dtp * d.dtp * POP_GRAV) : 0.0;
    rhs = rhs - dc * d.Pc[q];
  }
  d.RHS[q] = rhs;
  d.cwc[q] = d.indep[q] - dc;  // POP_SolversDiagonal
  d.Pn[q] = d.PGUESS[q];
}
__global__ void bt_pcheck_kernel(double* __restrict__ out, const double* __restrict__ P,
                                 const int* __restrict__ CHECKER, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = P[q] * CHECKER[q];
}
__global__ void bt_fin1_kernel(BtDrv d, const double* __restrict__ sums) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= d.n) return;
  const double xcheck = sums[0];
  d.Pn[q] = d.Pn[q] + d.CONSTNT[q] * d.rcheck * xcheck - d.CHECKER[q] * d.rconst * xcheck;
}
__global__ void bt_fin2_kernel(BtDrv d) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= d.n) return;
  const double gxr = d.leapfrog ? d.GXo[q] : d.GXc[q], gyr = d.leapfrog ? d.GYo[q] : d.GYc[q];
  d.UBn[q] = d.UH[q] - d.beta * d.c2dtp * (d.GXn[q] - gxr);
  d.VBn[q] = d.VH[q] - d.beta * d.c2dtp * (d.GYn[q] - gyr);
}

int barotropic_driver_dev() {
  ScopedTimer tm("BAROTROPIC");
  const int o = G.oldtime, c = G.curtime, n_ = G.newtime;
  BtDrv d;
  memset(&d, 0, sizeof(d));
  d.ZX = fld("ZX"); d.ZY = fld("ZY");
  d.GXc = fld_t("GRADPX", c); d.GXo = fld_t("GRADPX", o); d.GYc = fld_t("GRADPY", c); d.GYo = fld_t("GRADPY", o);
  d.FCOR = fld("FCOR"); d.HU = fld("HU"); d.UBo = fld_t("UBTROP", o); d.VBo = fld_t("VBTROP", o);
  d.RCALCT = fld("RCALCT"); d.TAREA = fld("TAREA"); d.Pc = fld_t("PSURF", c); d.FW = fld("FW");
  d.PGUESS = fld("PGUESS"); d.indep = fld("centerWgtClinicIndep");
  d.CHECKER = fldi("CHECKER"); d.CONSTNT = fldi("CONSTNT");
  d.UH = fld("UH_BT"); d.VH = fld("VH_BT"); d.W3 = fld("W2A"); d.W4 = fld("W2B"); d.RHS = fld("RHS_BT");
  d.dc = fld("DIAGC"); d.cwc = fld("centerWgtClinic"); d.Pn = fld_t("PSURF", n_);
  d.GXn = fld_t("GRADPX", n_); d.GYn = fld_t("GRADPY", n_); d.UBn = fld_t("UBTROP", n_); d.VBn = fld_t("VBTROP", n_);
  d.c2dtp = G.c2dtp; d.beta = G.beta; d.gamma = G.gamma; d.dtp = G.dtp; d.rcheck = G.rcheck; d.rconst = G.rconst;
  d.leapfrog = G.leapfrogts; d.impcor = G.cfg.impcor; d.sfc = G.cfg.sfc_layer_type;
  d.n = G.n2;
  const unsigned grid = ew_grid(G.n2);
  POP_LAUNCH(bt_rhs1_kernel, grid, POP_EW_THREADS, 0, d);
  POP_TRY(div_dev(1, d.RHS, d.W3, d.W4));
  POP_LAUNCH(bt_rhs2_kernel, grid, POP_EW_THREADS, 0, d);
  POP_TRY(halo_update(d.RHS, 1, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
  POP_TRY(solvers_run_dev(d.Pn, d.RHS));
  if (G.cfg.sfc_layer_type == POP_SFC_VARTHICK) {
    // remove the checkerboard null space: barotropic.F90:606-640
    POP_LAUNCH(bt_pcheck_kernel, grid, POP_EW_THREADS, 0, d.W3, d.Pn, d.CHECKER, G.n2);
    POP_TRY(global_sum_dev(d.W3, 1, G.n2, POP_LOC_CENTER, nullptr, nullptr));
    POP_LAUNCH(bt_fin1_kernel, grid, POP_EW_THREADS, 0, d, (const double*)G.d_sums);
  }
  POP_TRY(grad_dev(1, d.GXn, d.GYn, d.Pn));
  POP_LAUNCH(bt_fin2_kernel, grid, POP_EW_THREADS, 0, d);
  POP_TRY(halo_update(d.Pn, 1, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
  POP_TRY(halo_update(d.GXn, 1, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));
  POP_TRY(halo_update(d.GYn, 1, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));
  return pop_post_launch("barotropic_driver");
}
