// pop_lwlim.cu -- tracer advection with one-dimensional flux limiters (tadvect_ctype = 'lw_lim').
//
//   advection.F90: init (grid coefficients) :566-694; comp_flux_vel_ghost :1014-1120; the lw_lim branches of advt
//   :1667-1708 and comp_flux_vel :2036, :2086-2092, :2116-2121; advt_lw_lim :2684-2825; lw_lim :2832-3280.
//
// The scheme limits in z, then in x on the z-updated field, then in y on the z- and x-updated field, so the tendency
// of a cell needs the flux velocities two cells away -- one more than a block with two ghost cells can derive from its
// own velocities.  The reference therefore computes the flux velocities of every level once per step, halo-updates
// them, and keeps the outermost rows / columns (comp_flux_vel_ghost).  Here:
//   lw_flux_kernel (pass 1)  UTE, WTKB of every level from U, V, DH            -> halo update (3-d) -> lw_strips_kernel
//   lw_flux_kernel (pass 2)  the same kernel again, now with this step's rows / columns: UTE, VTN, WTKB of every level
//                            exactly as comp_flux_vel leaves them for lw_lim (local values one ghost cell out, the
//                            neighbour's two cells out), as 3-d fields
//   lw_lim_kernel            L(T) of every level for up to NTC tracers: a CTA owns a tile of columns plus a two-cell
//                            ring (thread <-> cell of the extended tile), marches k with the four levels of the
//                            vertical limiter and the flux through the top face in registers, and passes the z- and
//                            x-updated field between the three sweeps through shared memory
// The column kernel of pop_tracer.cu then takes L(T) of these tracers from memory instead of computing it.
#include <cmath>
#include "pop_dev.cuh"

#ifndef LW_TX
#define LW_TX 28
#endif
#ifndef LW_TY
#define LW_TY 12
#endif
#ifndef LW_MINB
#define LW_MINB 1   // resident CTAs per SM the register allocation aims for
#endif
#define LW_EX (LW_TX + 4)
#define LW_EY (LW_TY + 4)
#define LW_NT (LW_EX * LW_EY)
#define LW_NTC 2  // tracers per pass (the chunk of the column kernel)

struct LwDev {
  double *UTE_jbm2 = nullptr, *WTKB_jbm2 = nullptr, *WTKB_jep2 = nullptr;  // [km][nxb]
  double *WTKB_ibm2 = nullptr, *WTKB_iep2 = nullptr;                       // [km][nyb]
  int km = 0, nxb = 0, nyb = 0;
  bool coef_ready = false;
};
static LwDev LW;

void lw_release() {
  cudaFree(LW.UTE_jbm2);
  LW = LwDev();
}

// ---- k-invariant grid coefficients (:602-688), full-cell form; the partial-cell factor is applied per level
__global__ void lw_coef_kernel(int nxb, int nyb, const double* __restrict__ DXT, const double* __restrict__ DYT,
                               const double* __restrict__ HTE, const double* __restrict__ HTN, double* __restrict__ PX,
                               double* __restrict__ PY, double* __restrict__ E2U, double* __restrict__ N2V) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)nxb * nyb) return;
  const int i = (int)(q % nxb), j = (int)(q / nxb);
  PX[q] = (i + 1 < nxb) ? 1.0 / (DXT[q] + DXT[q + 1]) : 0.0;
  PY[q] = (j + 1 < nyb) ? 1.0 / (DYT[q] + DYT[q + nxb]) : 0.0;
  E2U[q] = 1.0 / HTE[q];
  N2V[q] = 1.0 / HTN[q];
}

// ---- flux velocities of every level (comp_flux_vel, :1970-2132, as called for lw_lim) -----------------------------
struct LwFluxArgs {
  int nxb, nyb, km;
  size_t n2;
  const double *U, *V, *DH, *DYU, *DXU, *TAREA_R, *DZU;
  const int* KMT;
  double *UTE3, *VTN3, *WTKB3;
  const double *UTE_jbm2, *WTKB_jbm2, *WTKB_jep2, *WTKB_ibm2, *WTKB_iep2;
};
template <bool PBC>
__global__ void __launch_bounds__(256) lw_flux_kernel(const LwFluxArgs a) {
  const int i = blockIdx.x * 32 + threadIdx.x % 32, j = blockIdx.y * 8 + threadIdx.x / 32;
  const int nxb = a.nxb, nyb = a.nyb, km = a.km;
  if (i >= nxb || j >= nyb) return;
  const size_t q = (size_t)j * nxb + i, n2 = a.n2;
  // the loops of comp_flux_vel run over i = ib-1..ie+1, j = jb-1..je+2 (0-based 1..nxb-2, 1..nyb-1)
  const bool in = (i >= 1 && i <= nxb - 2 && j >= 1);
  // U*DYU, V*DXU at (i,j), (i-1,j), (i,j-1), (i-1,j-1), guarded at the array edge
  const bool hw = (i >= 1), hs = (j >= 1);
  const double dyu00 = a.DYU[q], dyuw0 = hw ? a.DYU[q - 1] : 0.0, dyu0s = hs ? a.DYU[q - nxb] : 0.0,
               dyuws = (hw && hs) ? a.DYU[q - nxb - 1] : 0.0;
  const double dxu00 = a.DXU[q], dxuw0 = hw ? a.DXU[q - 1] : 0.0, dxu0s = hs ? a.DXU[q - nxb] : 0.0,
               dxuws = (hw && hs) ? a.DXU[q - nxb - 1] : 0.0;
  const double tarea_r = a.TAREA_R[q];
  const int kmt = a.KMT[q];
  double wtk = a.DH[q];
  for (int k = 1; k <= km; k++) {
    const size_t l = (size_t)(k - 1) * n2 + q;
    double u00 = a.U[l], uw0 = hw ? a.U[l - 1] : 0.0, u0s = hs ? a.U[l - nxb] : 0.0, uws = (hw && hs) ? a.U[l - nxb - 1] : 0.0;
    double v00 = a.V[l], vw0 = hw ? a.V[l - 1] : 0.0, v0s = hs ? a.V[l - nxb] : 0.0, vws = (hw && hs) ? a.V[l - nxb - 1] : 0.0;
    u00 = u00 * dyu00; uw0 = uw0 * dyuw0; u0s = u0s * dyu0s; uws = uws * dyuws;
    v00 = v00 * dxu00; vw0 = vw0 * dxuw0; v0s = v0s * dxu0s; vws = vws * dxuws;
    if (PBC) {  // :2040-2066
      const size_t z = (size_t)k * n2 + q;
      const double z00 = a.DZU[z], zw0 = hw ? a.DZU[z - 1] : 0.0, z0s = hs ? a.DZU[z - nxb] : 0.0,
                   zws = (hw && hs) ? a.DZU[z - nxb - 1] : 0.0;
      u00 = u00 * z00; uw0 = uw0 * zw0; u0s = u0s * z0s; uws = uws * zws;
      v00 = v00 * z00; vw0 = vw0 * zw0; v0s = v0s * z0s; vws = vws * zws;
    }
    double ute = in ? 0.5 * (u00 + u0s) : 0.0, utw = in ? 0.5 * (uw0 + uws) : 0.0;
    double vtn = in ? 0.5 * (v00 + vw0) : 0.0, vts = in ? 0.5 * (v0s + vws) : 0.0;
    {  // :2086-2092
      const double* s = a.UTE_jbm2 + (size_t)(k - 1) * nxb;
      if (j == 0) {
        ute = s[i];
        if (i >= 1) utw = s[i - 1];
      }
      // UTE(ib-2,:) = UTW(ib-1,:), VTN(:,jb-2) = VTS(:,jb-1): the formulas of the loops one cell further out
      if (i == 0) ute = (j == 0) ? s[0] : 0.5 * (u00 + u0s);
      if (j == 0) vtn = (i >= 1 && i <= nxb - 2) ? 0.5 * (v00 + vw0) : 0.0;
    }
    double wtkb = 0.0;
    if (k < km) {
      const double FC = (vtn - vts + ute - utw) * tarea_r;
      if (PBC) wtkb = (k < kmt) ? wtk + FC : 0.0;
      else wtkb = (k < kmt) ? wtk + c_vc.dz[k] * FC : 0.0;
      {  // :2116-2121
        if (j == 0) wtkb = a.WTKB_jbm2[(size_t)(k - 1) * nxb + i];
        if (j == nyb - 1) wtkb = a.WTKB_jep2[(size_t)(k - 1) * nxb + i];
        if (i == 0) wtkb = a.WTKB_ibm2[(size_t)(k - 1) * nyb + j];
        if (i == nxb - 1) wtkb = a.WTKB_iep2[(size_t)(k - 1) * nyb + j];
      }
    }
    a.UTE3[l] = ute;
    a.VTN3[l] = vtn;
    a.WTKB3[l] = wtkb;
    wtk = wtkb;
  }
}

// the rows / columns comp_flux_vel_ghost keeps (:1107-1113)
__global__ void lw_strips_kernel(int nxb, int nyb, int km, size_t n2, const double* __restrict__ UTE3,
                                 const double* __restrict__ WTKB3, double* __restrict__ UTE_jbm2,
                                 double* __restrict__ WTKB_jbm2, double* __restrict__ WTKB_jep2,
                                 double* __restrict__ WTKB_ibm2, double* __restrict__ WTKB_iep2) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nx = (size_t)km * nxb, ny = (size_t)km * nyb;
  if (p < nx) {
    const int i = (int)(p % nxb), k = (int)(p / nxb);
    const double* W = WTKB3 + (size_t)k * n2;
    UTE_jbm2[p] = UTE3[(size_t)k * n2 + i];
    WTKB_jbm2[p] = W[i];
    WTKB_jep2[p] = W[(size_t)(nyb - 1) * nxb + i];
  } else if (p < nx + ny) {
    const size_t t = p - nx;
    const int j = (int)(t % nyb), k = (int)(t / nyb);
    const double* W = WTKB3 + (size_t)k * n2;
    WTKB_ibm2[t] = W[(size_t)j * nxb];
    WTKB_iep2[t] = W[(size_t)j * nxb + nxb - 1];
  }
}

int lw_flux_prepare_dev(const double* U, const double* V, const double* DH) {
  POP_REQUIRE(G.use_lw_lim, "comp_flux_vel_ghost: no tracer uses lw_lim");
  ScopedTimer tm("LW_FLUX_VEL");
  const int nxb = G.nxb, nyb = G.nyb, km = G.km;
  if (!LW.UTE_jbm2 || LW.km != km || LW.nxb != nxb || LW.nyb != nyb) {
    lw_release();
    const size_t n = (size_t)km * (3 * (size_t)nxb + 2 * (size_t)nyb);
    POP_CHECK_CUDA(cudaMalloc(&LW.UTE_jbm2, sizeof(double) * n));
    POP_CHECK_CUDA(cudaMemsetAsync(LW.UTE_jbm2, 0, sizeof(double) * n, G.stream));
    LW.WTKB_jbm2 = LW.UTE_jbm2 + (size_t)km * nxb;
    LW.WTKB_jep2 = LW.WTKB_jbm2 + (size_t)km * nxb;
    LW.WTKB_ibm2 = LW.WTKB_jep2 + (size_t)km * nxb;
    LW.WTKB_iep2 = LW.WTKB_ibm2 + (size_t)km * nyb;
    LW.km = km; LW.nxb = nxb; LW.nyb = nyb;
  }
  if (!LW.coef_ready || G.lw_coef_dirty) {
    POP_TRY(alloc_field("LW_PX", 1, false)); POP_TRY(alloc_field("LW_PY", 1, false));
    POP_TRY(alloc_field("LW_E2U", 1, false)); POP_TRY(alloc_field("LW_N2V", 1, false));
    POP_LAUNCH(lw_coef_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, nxb, nyb, fld("DXT"), fld("DYT"), fld("HTE"), fld("HTN"),
               fld("LW_PX"), fld("LW_PY"), fld("LW_E2U"), fld("LW_N2V"));
    LW.coef_ready = true;
    G.lw_coef_dirty = false;
  }
  POP_TRY(alloc_field("LW_UTE", km, false)); POP_TRY(alloc_field("LW_VTN", km, false));
  POP_TRY(alloc_field("LW_WTKB", km, false));
  LwFluxArgs a;
  a.nxb = nxb; a.nyb = nyb; a.km = km; a.n2 = G.n2;
  a.U = U; a.V = V; a.DH = DH; a.DYU = fld("DYU"); a.DXU = fld("DXU"); a.TAREA_R = fld("TAREA_R");
  a.DZU = G.cfg.partial_bottom_cells ? fld("DZU") : nullptr;
  a.KMT = fldi("KMT");
  a.UTE3 = fld("LW_UTE"); a.VTN3 = fld("LW_VTN"); a.WTKB3 = fld("LW_WTKB");
  a.UTE_jbm2 = LW.UTE_jbm2; a.WTKB_jbm2 = LW.WTKB_jbm2; a.WTKB_jep2 = LW.WTKB_jep2;
  a.WTKB_ibm2 = LW.WTKB_ibm2; a.WTKB_iep2 = LW.WTKB_iep2;
  const dim3 grid((unsigned)((nxb + 31) / 32), (unsigned)((nyb + 7) / 8), 1);
  // pass 1 (comp_flux_vel_ghost): the rows / columns kept from the previous step are in place, as in the reference, so
  // ghost cells no halo update reaches (closed boundaries) hold what they hold there
  if (a.DZU) POP_LAUNCH(lw_flux_kernel<true>, grid, 256, 0, a);
  else POP_LAUNCH(lw_flux_kernel<false>, grid, 256, 0, a);
  POP_TRY(halo_update(a.UTE3, km, POP_LOC_EFACE, POP_KIND_VECTOR, 0.0));
  POP_TRY(halo_update(a.WTKB3, km, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
  const size_t ns = (size_t)km * ((size_t)nxb + nyb);
  POP_LAUNCH(lw_strips_kernel, ew_grid(ns), POP_EW_THREADS, 0, nxb, nyb, km, G.n2, a.UTE3, a.WTKB3, LW.UTE_jbm2,
             LW.WTKB_jbm2, LW.WTKB_jep2, LW.WTKB_ibm2, LW.WTKB_iep2);
  // pass 2: what comp_flux_vel returns during the tracer update of this step
  if (a.DZU) POP_LAUNCH(lw_flux_kernel<true>, grid, 256, 0, a);
  else POP_LAUNCH(lw_flux_kernel<false>, grid, 256, 0, a);
  G.lw_flux_ready = true;
  return pop_post_launch("comp_flux_vel_ghost");
}

// ---- the limited scheme ----------------------------------------------------------------------------------------------
struct LwArgs {
  int nxb, nyb, km, k0, k1;
  size_t n2;
  const double* X;      // TMIX (nxb,nyb,km,nt)
  // L(T) of tracer n, level k goes to LTK[n * tstride + (k-1) * lstride + q], ocean cells only: the caller's output
  // array itself (TRACER(new) of the fused driver, the LTK slab of pop_advt), where the column kernel picks it up
  // before it overwrites the cell with its own result -- no scratch copy of a 4-d field
  double* LTK;
  size_t tstride, lstride;
  int n[LW_NTC];        // 0-based tracer index of each slot, -1: slot unused
  const double *UTE3, *VTN3, *WTKB3, *DH, *TAREA_R, *DXT, *DYT, *PX, *PY, *E2U, *N2V, *DZT;
  const int* KMT;
  int varthick;
};

// limited face value: `a` = LW * dTR, `b` = MU * (difference on the upwind side)
__device__ __forceinline__ double lw_psi(double dTR, double dTRu, double lw, double mu) {
  if (dTR > 0.0 && dTRu > 0.0) return fmin(lw * dTR, mu * dTRu);
  if (dTR < 0.0 && dTRu < 0.0) return fmax(lw * dTR, mu * dTRu);
  return 0.0;
}

template <bool PBC>
__global__ void __launch_bounds__(LW_NT, LW_MINB) lw_lim_kernel(const LwArgs a) {
  __shared__ double s_ue[LW_NT], s_vn[LW_NT], s_lwx[LW_NT], s_lwy[LW_NT], s_ute[LW_NT], s_vtn[LW_NT];
  __shared__ double s_xs[LW_NTC][LW_NT], s_te[LW_NTC][LW_NT];
  double (*s_tn)[LW_NT] = s_te;  // the north-face values reuse the tile of the east-face values (a barrier apart)
  __shared__ int s_kmt[LW_NT];
  __shared__ double s_rdt[POP_KMAX + 2], s_p5[POP_KMAX + 2];  // 1 / c2dtt(k); p5_dz_ph_r(k) (:599-600), 0 at k = 0
  const int c = threadIdx.x, ei = c % LW_EX, ej = c / LW_EX;
  const int nxb = a.nxb, nyb = a.nyb, km = a.km;
  for (int k = c; k <= km + 1; k += LW_NT) {
    s_rdt[k] = (k >= 1 && k <= km) ? 1.0 / c_vc.c2dtt[k] : 0.0;
    s_p5[k] = (k >= 1 && k < km) ? 1.0 / (c_vc.dz[k] + c_vc.dz[k + 1]) : ((k == km) ? 0.5 / c_vc.dz[km] : 0.0);
  }
  const int gi = POP_NGHOST + blockIdx.x * LW_TX - 2 + ei, gj = POP_NGHOST + blockIdx.y * LW_TY - 2 + ej;  // 0-based
  const bool valid = (gi < nxb && gj < nyb);  // gi, gj >= 0 always
  const size_t q = valid ? (size_t)gj * nxb + gi : 0, n2 = a.n2;
  const int kmt = valid ? a.KMT[q] : 0;
  s_kmt[c] = kmt;
  // cells whose east / north face, x update, y update this thread evaluates (and whose operands exist in the tile)
  const bool he = valid && ei >= 1 && ei <= LW_TX + 1 && gi + 2 < nxb;                 // faces i = ib-1 .. ie
  const bool ux = valid && ei >= 2 && ei <= LW_TX + 1 && gi <= nxb - POP_NGHOST - 1;   // i = ib .. ie, every row
  const bool hn = ux && ej >= 1 && ej <= LW_TY + 1 && gj >= 1 && gj + 2 < nyb;          // faces j = jb-1 .. je
  const bool uy = ux && ej >= 2 && ej <= LW_TY + 1 && gj <= nyb - POP_NGHOST - 1;       // physical cells of the tile
  const double tarea_r = valid ? a.TAREA_R[q] : 0.0;
  const double dxt = valid ? a.DXT[q] : 0.0, dxte = (valid && gi + 1 < nxb) ? a.DXT[q + 1] : 0.0;
  const double dyt = valid ? a.DYT[q] : 0.0, dytn = (valid && gj + 1 < nyb) ? a.DYT[q + nxb] : 0.0;
  const double px = valid ? a.PX[q] : 0.0, py = valid ? a.PY[q] : 0.0;
  const double e2u0 = valid ? a.E2U[q] : 0.0, n2v0 = valid ? a.N2V[q] : 0.0;
  // column state: four levels of every tracer, the flux through the top face, the vertical velocities
  double Xm[LW_NTC], Xk[LW_NTC], Xp[LW_NTC], Xpp[LW_NTC], aux[LW_NTC];
  const double* Xn[LW_NTC];
#pragma unroll
  for (int m = 0; m < LW_NTC; m++) {
    Xn[m] = (a.n[m] >= 0 && valid) ? a.X + (size_t)a.n[m] * km * n2 + q : nullptr;
    Xm[m] = Xk[m] = Xp[m] = Xpp[m] = aux[m] = 0.0;
  }
  const int k0 = a.k0;
  auto level = [&](const double* x, int k) { return (x && k >= 1 && k <= km) ? x[(size_t)(k - 1) * n2] : 0.0; };
  auto wlev = [&](int k) { return (valid && k >= 1 && k <= km) ? a.WTKB3[(size_t)(k - 1) * n2 + q] : 0.0; };
  // vertical velocity at the top of level k as lw_lim sees it (:2780-2784)
  auto wtop = [&](int k) { return (k == 1) ? ((a.varthick || !valid) ? 0.0 : a.DH[q]) : wlev(k - 1); };
  auto dzt_at = [&](int k) { return (PBC && valid) ? a.DZT[(size_t)k * n2 + q] : 0.0; };
  // LW_z, MU_z of the bottom face of level k (:2911-2975): tracer independent
  auto zcoef = [&](int k, double wt, double wb, double wbp1, double& lwz, double& muz) {
    lwz = 0.0; muz = 0.0;
    if (!(k + 1 <= kmt)) return;
    const double adv_dt = c_vc.c2dtt[k], adv_dt_r = s_rdt[k];
    if (PBC) {
      const double zk = dzt_at(k), zp = dzt_at(k + 1);
      if (wb > 0.0) {
        lwz = (zp - adv_dt * wb) / (zk + zp);
        if (wbp1 > 0.0) muz = (zp * adv_dt_r - wbp1) / wb;
        else if (wbp1 < 0.0) muz = -wbp1 / wb * (zp + adv_dt * wbp1) / (zp + dzt_at(k + 2));
      } else if (wb < 0.0) {
        lwz = (zk + adv_dt * wb) / (zk + zp);
        if (wt < 0.0) muz = -(zk * adv_dt_r + wt) / wb;
        else if (wt > 0.0) muz = -wt / wb * (zk - adv_dt * wt) / (dzt_at(k - 1) + zk);
      }
    } else {
      const double p5k = s_p5[k];
      if (wb > 0.0) {
        lwz = c_vc.dz[k + 1] * p5k - (adv_dt * p5k) * wb;
        if (wbp1 > 0.0) muz = (c_vc.dz[k + 1] * adv_dt_r - wbp1) / wb;
        else if (wbp1 < 0.0) muz = -wbp1 / wb * (c_vc.dz[k + 1] + adv_dt * wbp1) * s_p5[k + 1];
      } else if (wb < 0.0) {
        lwz = c_vc.dz[k] * p5k + (adv_dt * p5k) * wb;
        if (wt < 0.0) muz = -(c_vc.dz[k] * adv_dt_r + wt) / wb;
        else if (wt > 0.0) muz = -wt / wb * (c_vc.dz[k] - adv_dt * wt) * s_p5[k - 1];
      }
    }
  };
  // the limited flux through that face (:3096-3137)
  auto zface = [&](int k, double wb, double lwz, double muz, double xm, double xk, double xp, double xpp) {
    if (!(k + 1 <= kmt)) return 0.0;
    const double dTR = xp - xk;
    if (wb > 0.0) {
      double f = wb * xp;
      if (k + 2 <= kmt) {
        const double dTRp1 = xpp - xp;
        if ((dTR > 0.0 && dTRp1 > 0.0) || (dTR < 0.0 && dTRp1 < 0.0)) f = wb * (xp - lw_psi(dTR, dTRp1, lwz, muz));
      }
      return f;
    }
    if (wb < 0.0) {
      double f = wb * xk;
      if (k > 1) {
        const double dTRm1 = xk - xm;
        if ((dTR > 0.0 && dTRm1 > 0.0) || (dTR < 0.0 && dTRm1 < 0.0)) f = wb * (xk + lw_psi(dTR, dTRm1, lwz, muz));
      }
      return f;
    }
    return 0.0;
  };
  double wtk = wtop(k0), wtkb = wlev(k0), wtkbp1 = wlev(k0 + 1);
  __syncthreads();  // s_rdt, s_p5, s_kmt
  {
    double lwz0 = 0.0, muz0 = 0.0;
    if (k0 > 1) zcoef(k0 - 1, wtop(k0 - 1), wtk, wtkb, lwz0, muz0);
#pragma unroll
    for (int m = 0; m < LW_NTC; m++) {
      Xm[m] = level(Xn[m], k0 - 1); Xk[m] = level(Xn[m], k0); Xp[m] = level(Xn[m], k0 + 1); Xpp[m] = level(Xn[m], k0 + 2);
      // flux through the top face of the first level: WTK_EFF * X at the surface (:2795), else the bottom flux of k0-1
      if (k0 == 1) aux[m] = wtk * Xk[m];
      else aux[m] = zface(k0 - 1, wtk, lwz0, muz0, level(Xn[m], k0 - 2), Xm[m], Xk[m], Xp[m]);
    }
  }
  // running pointers of the level loop (one add per level and array instead of a 64-bit product per access); the flux
  // velocities of a level are requested one level ahead
  const int n2i = (int)n2;
  const double* pu = a.UTE3 + (size_t)(k0 - 1) * n2 + q;
  const double* pv = a.VTN3 + (size_t)(k0 - 1) * n2 + q;
  const double* pw = a.WTKB3 + (size_t)(k0 + 1) * n2 + q;  // WTKB of level k + 2
  const double* pxn[LW_NTC];
  double* pl[LW_NTC];
#pragma unroll
  for (int m = 0; m < LW_NTC; m++) {
    pxn[m] = Xn[m] ? Xn[m] + (size_t)(k0 + 2) * n2 : nullptr;  // X of level k + 3
    pl[m] = (a.n[m] >= 0) ? a.LTK + (size_t)a.n[m] * a.tstride + (size_t)(k0 - 1) * a.lstride + q : nullptr;
  }
  double ute_n = valid ? *pu : 0.0, vtn_n = valid ? *pv : 0.0;
  for (int k = k0; k <= a.k1; k++) {
    const double adv_dt = c_vc.c2dtt[k];
    // ---- tracer-independent part of the level
    const double ute = ute_n, vtn = vtn_n;
    pu += n2i; pv += n2i;
    if (valid && k < a.k1) { ute_n = *pu; vtn_n = *pv; }
    double e2u = e2u0, n2v = n2v0, wgt = tarea_r, dzk = 0.0;
    if (PBC && valid) {  // :609-617, :659-667, :2761-2768
      dzk = a.DZT[(size_t)k * n2 + q];
      if (gi + 1 < nxb) e2u = e2u0 / fmin(dzk, a.DZT[(size_t)k * n2 + q + 1]);
      if (gj + 1 < nyb) n2v = n2v0 / fmin(dzk, a.DZT[(size_t)k * n2 + q + nxb]);
      wgt = tarea_r / dzk;
    }
    const double ue = adv_dt * ute * e2u, vn = adv_dt * vtn * n2v;
    s_ute[c] = ute; s_vtn[c] = vtn; s_ue[c] = ue; s_vn[c] = vn;
    // LW_x, LW_y of the east / north face of this cell (:2977-3061): one formula, chosen by the sign of the velocity
    s_lwx[c] = (ue > 0.0) ? (dxt - ue) * px : ((ue < 0.0) ? (dxte + ue) * px : dxt * px);
    s_lwy[c] = (vn > 0.0) ? (dyt - vn) * py : ((vn < 0.0) ? (dytn + vn) * py : dyt * py);
    // ---- z sweep (:3096-3149)
    double xout[LW_NTC], xs[LW_NTC];
    double lwz, muz;
    zcoef(k, wtk, wtkb, wtkbp1, lwz, muz);
#pragma unroll
    for (int m = 0; m < LW_NTC; m++) {
      const double auxb = zface(k, wtkb, lwz, muz, Xm[m], Xk[m], Xp[m], Xpp[m]);
      if (PBC) xout[m] = (aux[m] - auxb - (wtk - wtkb) * Xk[m]) / dzk;
      else xout[m] = (aux[m] - auxb - (wtk - wtkb) * Xk[m]) * c_vc.dzr[k];
      if (!valid) xout[m] = 0.0;
      xs[m] = Xk[m] - adv_dt * xout[m];
      s_xs[m][c] = xs[m];
      aux[m] = auxb;
    }
    __syncthreads();
    // ---- x sweep: MU_x of this face (:2977-3015), limited value on the face (:3151-3188)
    const double CE = ute * wgt;
    if (he) {
      const double uw = s_ue[c - 1], uee = s_ue[c + 1];
      double mux = 0.0;
      if (ue > 0.0) {
        if (uw > 0.0) mux = (dxt - uw) / ue;
        else if (uw < 0.0) mux = -uw / ue * s_lwx[c - 1];
      } else if (ue < 0.0) {
        if (uee < 0.0) mux = -(dxte + uee) / ue;
        else if (uee > 0.0) mux = -uee / ue * s_lwx[c + 1];
      }
      const double lwx = s_lwx[c];
      const int kE = s_kmt[c + 1];
      const double me = (k <= kmt && k <= kE) ? 1.0 : 0.0;
      const double mw = (k <= s_kmt[c - 1] && k <= kmt) ? 1.0 : 0.0;            // KMASKE(i-1)
      const double mee = (k <= kE && k <= s_kmt[c + 2]) ? 1.0 : 0.0;            // KMASKE(i+1)
#pragma unroll
      for (int m = 0; m < LW_NTC; m++) {
        const double x0 = s_xs[m][c], x1 = s_xs[m][c + 1];
        const double dTR = (x1 - x0) * me;
        double te;
        if (CE > 0.0) te = x0 + lw_psi(dTR, (x0 - s_xs[m][c - 1]) * mw, lwx, mux);
        else if (CE < 0.0) te = x1 - lw_psi(dTR, (s_xs[m][c + 2] - x1) * mee, lwx, mux);
        else te = x0 + lwx * dTR;
        s_te[m][c] = te;
      }
    }
    __syncthreads();
    if (ux) {  // :3205-3213
      const double CW = -s_ute[c - 1] * wgt;
#pragma unroll
      for (int m = 0; m < LW_NTC; m++) {
        const double work1 = CE * s_te[m][c] + CW * s_te[m][c - 1] - (CE + CW) * Xk[m];
        xout[m] = xout[m] + work1;
        xs[m] = xs[m] - adv_dt * work1;
        s_xs[m][c] = xs[m];
      }
    }
    __syncthreads();
    // ---- y sweep (:3017-3061, :3220-3256)
    const double CN = vtn * wgt;
    if (hn) {
      const double vs = s_vn[c - LW_EX], vnn = s_vn[c + LW_EX];
      double muy = 0.0;
      if (vn > 0.0) {
        if (vs > 0.0) muy = (dyt - vs) / vn;
        else if (vs < 0.0) muy = -vs / vn * s_lwy[c - LW_EX];
      } else if (vn < 0.0) {
        if (vnn < 0.0) muy = -(dytn + vnn) / vn;
        else if (vnn > 0.0) muy = -vnn / vn * s_lwy[c + LW_EX];
      }
      const double lwy = s_lwy[c];
      const int kN = s_kmt[c + LW_EX];
      const double mn = (k <= kmt && k <= kN) ? 1.0 : 0.0;
      const double ms = (k <= s_kmt[c - LW_EX] && k <= kmt) ? 1.0 : 0.0;        // KMASKN(j-1)
      const double mnn = (k <= kN && k <= s_kmt[c + 2 * LW_EX]) ? 1.0 : 0.0;    // KMASKN(j+1)
#pragma unroll
      for (int m = 0; m < LW_NTC; m++) {
        const double x0 = s_xs[m][c], x1 = s_xs[m][c + LW_EX];
        const double dTR = (x1 - x0) * mn;
        double tn;
        if (CN > 0.0) tn = x0 + lw_psi(dTR, (x0 - s_xs[m][c - LW_EX]) * ms, lwy, muy);
        else if (CN < 0.0) tn = x1 - lw_psi(dTR, (s_xs[m][c + 2 * LW_EX] - x1) * mnn, lwy, muy);
        else tn = x0 + lwy * dTR;
        s_tn[m][c] = tn;
      }
    }
    __syncthreads();
    if (uy) {  // :3063-3067, :3263-3272
      const double CW = -s_ute[c - 1] * wgt, CS = -s_vtn[c - LW_EX] * wgt;
      double DIV;
      if (PBC) DIV = (wtk - wtkb) / dzk + CE + CW + CN + CS;
      else DIV = (wtk - wtkb) * c_vc.dzr[k] + CE + CW + CN + CS;
#pragma unroll
      for (int m = 0; m < LW_NTC; m++) {
        if (a.n[m] < 0 || k > kmt) continue;
        const double L = xout[m] + CN * s_tn[m][c] + CS * s_tn[m][c - LW_EX] - (CN + CS - DIV) * Xk[m];
        *pl[m] = L;
      }
    }
    // ---- next level
    wtk = wtkb; wtkb = wtkbp1; wtkbp1 = (valid && k + 2 <= km) ? *pw : 0.0;
    pw += n2i;
#pragma unroll
    for (int m = 0; m < LW_NTC; m++) {
      Xm[m] = Xk[m]; Xk[m] = Xp[m]; Xp[m] = Xpp[m]; Xpp[m] = (pxn[m] && k + 3 <= km) ? *pxn[m] : 0.0;
      if (pxn[m]) pxn[m] += n2i;
      if (pl[m]) pl[m] += a.lstride;
    }
    __syncthreads();
  }
}

// L(T) of levels k0..k1 for the tracers `slots` (0-based indices, -1: unused) into `out` (see LwArgs::LTK)
int lw_lim_dev(const int* slots, const double* TMIX, int k0, int k1, double* out, size_t tstride, size_t lstride) {
  POP_REQUIRE(G.lw_flux_ready, "advt (lw_lim): comp_flux_vel_ghost has not been called for this step");
  ScopedTimer tm("ADVT_LW_LIM");
  LwArgs a;
  a.nxb = G.nxb; a.nyb = G.nyb; a.km = G.km; a.k0 = k0; a.k1 = k1; a.n2 = G.n2;
  a.X = TMIX; a.LTK = out; a.tstride = tstride; a.lstride = lstride;
  for (int m = 0; m < LW_NTC; m++) a.n[m] = slots[m];
  a.UTE3 = fld("LW_UTE"); a.VTN3 = fld("LW_VTN"); a.WTKB3 = fld("LW_WTKB"); a.DH = fld("DH");
  a.TAREA_R = fld("TAREA_R"); a.DXT = fld("DXT"); a.DYT = fld("DYT");
  a.PX = fld("LW_PX"); a.PY = fld("LW_PY"); a.E2U = fld("LW_E2U"); a.N2V = fld("LW_N2V");
  a.DZT = G.cfg.partial_bottom_cells ? fld("DZT") : nullptr;
  a.KMT = fldi("KMT");
  a.varthick = (G.cfg.sfc_layer_type == POP_SFC_VARTHICK) ? 1 : 0;
  const dim3 grid((unsigned)((G.nxg + LW_TX - 1) / LW_TX), (unsigned)((G.ny_local + LW_TY - 1) / LW_TY), 1);
  if (a.DZT) POP_LAUNCH(lw_lim_kernel<true>, grid, LW_NT, 0, a);
  else POP_LAUNCH(lw_lim_kernel<false>, grid, LW_NT, 0, a);
  return pop_post_launch("advt_lw_lim");
}
