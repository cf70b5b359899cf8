// pop_gm.cu -- Gent-McWilliams / Redi isopycnal tracer mixing:
//   tracer_diffs_and_isopyc_slopes  source/hmix_gm_submeso_share.F90:149-434
//   init_gm / hdifft_gm             source/hmix_gm.F90:283-1095, 1102-2219
// for constant kappa_isop/kappa_thic, slope_control 'notanh', constant surface background diffusivity,
// no transition layer, no KPP boundary layer depth (BL_DEPTH = zw(1)); both the 'cancellation' short cut
// (ah == ah_bolus, slm_r == slm_b) and the general skew-flux form.
//
// The reference keeps TX/TY/TZ, RX/RY, SLX/SLY, SF_SLX/SF_SLY, KAPPA_ISOP/KAPPA_THIC/HOR_DIFF as block arrays
// (about 26*km + 3*km*nt 2-d fields) that every k pass re-reads. Here two kernels do the work:
//   gm_column_kernel  one thread per (column, level) over the physical cells + the first ghost ring: density
//                     gradients, the quarter-cell slopes either side of the interface below the level, tapers and
//                     diffusivities, and the VDC side effect (VDC += VDC_GM); only slopes (8/level) and
//                     diffusivities (4-6/level) reach HBM
//   gm_flux_kernel    one thread per (physical column, level), tracers looped inside: tracer differences are rebuilt from TMIX,
//                     face fluxes FX/FY and the vertical fluxes through the top and bottom faces are formed and
//                     differenced in registers; only the tendency GTK reaches HBM
// Nothing is carried between levels (FZTOP is re-evaluated where it is needed), so all levels run in parallel.
// Every expression keeps the reference's operand order (-fmad=false), so the tendency is bit-identical to the
// oracle's slab-by-slab restatement.
#include <cmath>
#include <cstring>
#include <vector>

#include "pop_ctx.h"
#include "pop_dev.cuh"
#include "pop_state.cuh"

namespace {
constexpr double GM_EPS = 1.0e-10, GM_EPS2 = 1.0e-20;  // pop_constants.F90:55-56
enum { XE = 0, XW = 1, YN = 2, YS = 3 };               // SLX(ieast), SLX(iwest), SLY(jnorth), SLY(jsouth)
enum { KTP = 0, KBT = 1 };

struct GmArgs {
  GridView g;
  const double *TMIX, *HYX, *HXY, *RBR, *DXT, *DYT;
  double *SL;   // (n2, 4 directions, 2 halves, km)
  double *KI, *HD, *KT;  // (n2, 2 halves, km)
  double* VDC;
  double* HDT;  // (n2, km, nt) tendency
  StateOpt so;
  double ah, ah_bolus, ah_bkg_srfbl, slm_r, slm_b;
  int diff_tapering, cancel;
};

__device__ __forceinline__ size_t sl_ix(const GmArgs& a, int d, int h, int kk) {  // kk 1-based
  return ((size_t)((kk - 1) * 2 + h) * 4 + d) * a.g.n2;
}
__device__ __forceinline__ size_t k2_ix(const GmArgs& a, int h, int kk) { return (size_t)((kk - 1) * 2 + h) * a.g.n2; }

// taper of slope_control_notanh (hmix_gm.F90:1503-1538)
__device__ __forceinline__ double gm_taper(double SLA, double slm) {
  if (SLA > 0.2 * slm && SLA < 0.6 * slm)
    return 0.5 * (1.0 - (2.5 * SLA / slm - 1.0) * (4.0 - fabs(10.0 * SLA / slm - 4.0)));
  if (SLA >= 0.6 * slm) return 0.0;
  return 1.0;
}

// diffusivities of one half cell from its four slopes (hmix_gm.F90:1414-1662); returns KAPPA_ISOP
__device__ __forceinline__ double gm_half(const GmArgs& a, size_t q, int h, int kk, int kmt, const double* s,
                                          double rbr, double dxt, double dyt) {
  const int kid = kk + h - 1;  // = kk + kk_sub - 2 with kk_sub = h + 1
  const double SLA = c_vc.dzw[kid] * sqrt(0.5 * ((s[XE] * s[XE] + s[XW] * s[XW]) / (dxt * dxt) +
                                                 (s[YN] * s[YN] + s[YS] * s[YS]) / (dyt * dyt))) + GM_EPS;
  const double dz_bottom = (kk == 1) ? 0.0 : c_vc.zt[kk - 1];
  const bool in_bl = (dz_bottom <= c_vc.zw[1]);  // BL_DEPTH = zw(1) without KPP
  double W1 = c_vc.zt[kk] * rbr / SLA;
  W1 = (1.0 < W1) ? 1.0 : W1;
  double TAPER1 = (0.5 + 2.0 * (W1 - 0.5) * (1.0 - fabs(W1 - 0.5)));
  if (!in_bl) TAPER1 = 1.0;
  const double TAPER2 = gm_taper(SLA, a.slm_r);
  const double TAPER3 = a.diff_tapering ? gm_taper(SLA, a.slm_b) : TAPER2;
  double hd;
  if (kk == 1 && h == KTP) hd = a.ah_bkg_srfbl;
  else hd = in_bl ? a.ah_bkg_srfbl * (1.0 - TAPER1 * TAPER2) * 1.0 : 0.0;
  double ki = TAPER1 * TAPER2 * a.ah;
  double kt = TAPER1 * TAPER3 * a.ah_bolus;
  if ((h == KBT && kk == kmt) || (h == KTP && kk == 1)) { ki = 0.0; kt = 0.0; }  // bottom / top B.C.
  const size_t o = k2_ix(a, h, kk) + q;
  a.KI[o] = ki;
  a.HD[o] = hd;
  if (!a.cancel) a.KT[o] = kt;
#pragma unroll
  for (int d = 0; d < 4; d++) a.SL[sl_ix(a, d, h, kk) + q] = s[d];
  return ki;
}

struct GmLevel {
  double drdt, drds, rx[4], temp, salt;
};
// horizontal density differences of level kk at column q (hmix_gm_submeso_share.F90:221-290, 318-386)
__device__ __forceinline__ void gm_level(const GmArgs& a, size_t q, int kk, int kmt, int kmte, int kmtw, int kmtn,
                                         int kmts, GmLevel* L) {
  const size_t n2 = a.g.n2, nxb = a.g.nxb;
  const double* T = a.TMIX + (size_t)(kk - 1) * n2;
  const double* S = T + (size_t)a.g.km * n2;
  const double t0 = T[q], te = T[q + 1], tw = T[q - 1], tn = T[q + nxb], ts = T[q - nxb];
  const double s0 = S[q], se = S[q + 1], sw = S[q - 1], sn = S[q + nxb], ss = S[q - nxb];
#define CLIP(x) ((-2.0 > (x)) ? -2.0 : (x))
  const double c0 = CLIP(t0), ce = CLIP(te), cw = CLIP(tw), cn = CLIP(tn), cs = CLIP(ts);
#undef CLIP
  const double me = (kk <= kmt && kk <= kmte) ? 1.0 : 0.0, mw = (kk <= kmtw && kk <= kmt) ? 1.0 : 0.0;
  const double mn = (kk <= kmt && kk <= kmtn) ? 1.0 : 0.0, ms = (kk <= kmts && kk <= kmt) ? 1.0 : 0.0;
  const double txpe = me * (ce - c0), txpw = mw * (c0 - cw), typn = mn * (cn - c0), typs = ms * (c0 - cs);
  const double sxe = me * (se - s0), sxw = mw * (s0 - sw), syn = mn * (sn - s0), sys = ms * (s0 - ss);
  state_cell(a.so, kk, t0, s0, nullptr, nullptr, &L->drdt, &L->drds);
  L->rx[XE] = L->drdt * txpe + L->drds * sxe;
  L->rx[XW] = L->drdt * txpw + L->drds * sxw;
  L->rx[YN] = L->drdt * typn + L->drds * syn;
  L->rx[YS] = L->drdt * typs + L->drds * sys;
  L->temp = c0;
  L->salt = s0;
}

// One thread per (column, level): nothing is carried from level to level (each interface needs the density
// differences of its two levels only), so the km levels of a column run in parallel -- a gx1v7-sized strip has
// only 1.2e5 columns, far too few to fill 148 SMs with one thread per column. Thread (q, kk) owns the interface
// below level kk: the bottom half of level kk, the top half of level kk+1 and VDC_GM(kk); it evaluates the level
// differences of kk and kk+1 (the EOS derivatives of a level are thus evaluated twice, by threads kk-1 and kk).
__global__ void __launch_bounds__(128, 8) gm_column_kernel(const GmArgs a) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;  // 0-based, first ghost ring included
  const int j = 1 + blockIdx.y;
  const int kk = 1 + blockIdx.z;
  const GridView& g = a.g;
  if (i > g.nxb - 2 || j > g.nyb - 2) return;
  const size_t n2 = g.n2, nxb = g.nxb, q = (size_t)j * nxb + i;
  const int km = g.km;
  const int kmt = g.KMT[q], kmte = g.KMT[q + 1], kmtw = g.KMT[q - 1], kmtn = g.KMT[q + nxb], kmts = g.KMT[q - nxb];
  const double rbr = a.RBR[q], dxt = a.DXT[q], dyt = a.DYT[q];
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  if (kk == 1) gm_half(a, q, KTP, 1, kmt, s, rbr, dxt, dyt);  // the top half of level 1 has no slopes
  if (kk == km) {                                             // nor has the bottom half of level km
    gm_half(a, q, KBT, km, kmt, s, rbr, dxt, dyt);
    return;
  }
  const double hyx = a.HYX[q], hyxw = a.HYX[q - 1], hxy = a.HXY[q], hxys = a.HXY[q - nxb];
  GmLevel L;
  gm_level(a, q, kk, kmt, kmte, kmtw, kmtn, kmts, &L);
  const double KMASK = (kk < kmt) ? 1.0 : 0.0;
  const double tdn = a.TMIX[(size_t)kk * n2 + q];
  const double tempn = (-2.0 > tdn) ? -2.0 : tdn;
  const double tz2 = L.salt - a.TMIX[((size_t)km + kk) * n2 + q];
  const double tzp = L.temp - tempn;
  double rz = L.drdt * tzp + L.drds * tz2;  // :302-303
  rz = (rz < -GM_EPS2) ? rz : -GM_EPS2;
#pragma unroll
  for (int d = 0; d < 4; d++) s[d] = KMASK * L.rx[d] / rz;
  const double kib = gm_half(a, q, KBT, kk, kmt, s, rbr, dxt, dyt);
  const double vb = hyx * (s[XE] * s[XE]) + hyxw * (s[XW] * s[XW]) + hxy * (s[YN] * s[YN]) + hxys * (s[YS] * s[YS]);
  gm_level(a, q, kk + 1, kmt, kmte, kmtw, kmtn, kmts, &L);
  double rz1 = L.drdt * tzp + L.drds * tz2;  // :388-390
  rz1 = (rz1 < -GM_EPS2) ? rz1 : -GM_EPS2;
#pragma unroll
  for (int d = 0; d < 4; d++) s[d] = (kk + 1 <= kmt) ? L.rx[d] / rz1 : 0.0;
  const double kit = gm_half(a, q, KTP, kk + 1, kmt, s, rbr, dxt, dyt);
  const double vt = hyx * (s[XE] * s[XE]) + hyxw * (s[XW] * s[XW]) + hxy * (s[YN] * s[YN]) + hxys * (s[YS] * s[YS]);
  // effective vertical diffusion coefficient, VDC += VDC_GM (:1725-1743)
  const double W = c_vc.dzw[kk] * KMASK * g.TAREA_R[q] *
                   (c_vc.dz[kk] * 0.25 * kib * vb + c_vc.dz[kk + 1] * 0.25 * kit * vt);
  for (int n = 0; n < g.vdc_nd; n++) {
    double* v = a.VDC + ((size_t)n * g.vdc_nk + (kk - g.vdc_k0)) * n2 + q;
    *v = *v + W;
  }
}

// ---- face fluxes FX / FY of hdifft_gm (:1765-1896). The diffusivity sums and the skew coefficients of a face
// do not depend on the tracer: they are formed once per (column, level) and applied to every tracer.
struct GmFace {
  double C, m, W;          // CX or CY, its 0/1 mask, the eight-term diffusivity sum WORK3 / WORK4
  double W1, W2, W3, W4;   // general form only: KAPPA_ISOP*slope*dz - SF_slope at (qa,top) (qa,bottom) (qb,top) (qb,bottom)
};
// face between cell qa and its east (qb = qa+1) / north (qb = qa+nxb) neighbour; da/db: slope direction stored
// at qa (east/north) and at qb (west/south)
template <bool CANCEL>
__device__ __forceinline__ GmFace gm_face_coef(const GmArgs& a, size_t qa, size_t qb, int da, int db, double H, int k) {
  GmFace f;
  const int ka = a.g.KMT[qa], kb = a.g.KMT[qb];
  const bool open = (k <= ka && k <= kb);
  f.C = open ? H * 0.25 : 0.0;
  f.m = open ? 1.0 : 0.0;
  const size_t ot = k2_ix(a, KTP, k), ob = k2_ix(a, KBT, k);
  const double kiat = a.KI[ot + qa], kiab = a.KI[ob + qa], kibt = a.KI[ot + qb], kibb = a.KI[ob + qb];
  f.W = kiat + a.HD[ot + qa] + kiab + a.HD[ob + qa] + kibt + a.HD[ot + qb] + kibb + a.HD[ob + qb];
  f.W1 = f.W2 = f.W3 = f.W4 = 0.0;
  if (!CANCEL) {
    const double dzk = c_vc.dz[k];
    const double sat = a.SL[sl_ix(a, da, KTP, k) + qa], sab = a.SL[sl_ix(a, da, KBT, k) + qa];
    const double sbt = a.SL[sl_ix(a, db, KTP, k) + qb], sbb = a.SL[sl_ix(a, db, KBT, k) + qb];
    const double sfat = (k <= ka) ? a.KT[ot + qa] * sat * dzk : 0.0;  // SF_SLX / SF_SLY (:1682-1700)
    const double sfab = (k <= ka) ? a.KT[ob + qa] * sab * dzk : 0.0;
    const double sfbt = (k <= kb) ? a.KT[ot + qb] * sbt * dzk : 0.0;
    const double sfbb = (k <= kb) ? a.KT[ob + qb] * sbb * dzk : 0.0;
    f.W1 = kiat * sat * dzk - sfat;
    f.W2 = kiab * sab * dzk - sfab;
    f.W3 = kibt * sbt * dzk - sfbt;
    f.W4 = kibb * sbb * dzk - sfbb;
  }
  return f;
}
template <bool CANCEL>
__device__ __forceinline__ double gm_face_flux(const GmArgs& a, const GmFace& f, const double* Tn, size_t qa, size_t qb,
                                               int k, int kp1) {
  const size_t n2 = a.g.n2;
  const double* Tk = Tn + (size_t)(k - 1) * n2;
  const double tx = f.m * (Tk[qb] - Tk[qa]);
  double F = c_vc.dz[k] * f.C * tx * f.W;
  if (!CANCEL) {
    // TZ(:,:,k) = TMIX(k-1) - TMIX(k), zero at k = 1 (hmix_gm_submeso_share.F90:296-297, hmix_gm.F90:1849-1850)
    const double* Tm = Tk - n2;
    const double* Tp = Tn + (size_t)(kp1 - 1) * n2;
    const double tza = (k >= 2) ? Tm[qa] - Tk[qa] : 0.0, tzb = (k >= 2) ? Tm[qb] - Tk[qb] : 0.0;
    const double tzpa = (kp1 >= 2) ? (Tp - n2)[qa] - Tp[qa] : 0.0, tzpb = (kp1 >= 2) ? (Tp - n2)[qb] - Tp[qb] : 0.0;
    F = F - f.C * (f.W1 * tza + f.W2 * tzpa + f.W3 * tzb + f.W4 * tzpb);
  }
  return F;
}

// ---- vertical flux through the bottom face of level k < km (:1911-2040): tracer-independent part ...
struct GmFz {
  double KMASK, kib, kit1, sb[4], st[4], fb[4], ft[4];
};
template <bool CANCEL>
__device__ __forceinline__ GmFz gm_fz_coef(const GmArgs& a, size_t q, int k, int kmt) {
  GmFz c;
  const int kp1 = k + 1;
  c.KMASK = (k < kmt) ? 1.0 : 0.0;
  const size_t ob = k2_ix(a, KBT, k) + q, ot1 = k2_ix(a, KTP, kp1) + q;
  c.kib = a.KI[ob];
  c.kit1 = a.KI[ot1];
#pragma unroll
  for (int d = 0; d < 4; d++) {
    c.sb[d] = a.SL[sl_ix(a, d, KBT, k) + q];
    c.st[d] = a.SL[sl_ix(a, d, KTP, kp1) + q];
    c.fb[d] = c.ft[d] = 0.0;
  }
  if (!CANCEL) {
    const double ktb = a.KT[ob], ktt1 = a.KT[ot1];
#pragma unroll
    for (int d = 0; d < 4; d++) {
      c.fb[d] = (k <= kmt) ? ktb * c.sb[d] * c_vc.dz[k] : 0.0;
      c.ft[d] = (kp1 <= kmt) ? ktt1 * c.st[d] * c_vc.dz[kp1] : 0.0;
    }
  }
  return c;
}
// ... and its application to the tracer differences t0 (level k) and t1 (level k+1), ordered TX(i), TX(i-1), TY(j), TY(j-1)
template <bool CANCEL>
__device__ __forceinline__ double gm_fz_eval(const GmArgs& a, const GmFz& c, int k, const double* t0, const double* t1,
                                             double hyx, double hyxw, double hxy, double hxys) {
  const double dz_bottom = c_vc.dz[k + 1];
  if (!CANCEL) {  // :1917-1986
    double w3 = 0.0;
    w3 = w3 + (c_vc.dz[k] * c.kib *
               (c.sb[XE] * hyx * t0[XE] + c.sb[YN] * hxy * t0[YN] + c.sb[XW] * hyxw * t0[XW] + c.sb[YS] * hxys * t0[YS]));
    w3 = w3 + (c.fb[XE] * hyx * t0[XE] + c.fb[YN] * hxy * t0[YN] + c.fb[XW] * hyxw * t0[XW] + c.fb[YS] * hxys * t0[YS]);
    w3 = w3 + (dz_bottom * c.kit1 *
               (c.st[XE] * hyx * t1[XE] + c.st[YN] * hxy * t1[YN] + c.st[XW] * hyxw * t1[XW] + c.st[YS] * hxys * t1[YS]));
    w3 = w3 + (1.0 * (c.ft[XE] * hyx * t1[XE] + c.ft[YN] * hxy * t1[YN] + c.ft[XW] * hyxw * t1[XW] +
                      c.ft[YS] * hxys * t1[YS]));
    return -c.KMASK * 0.25 * w3;
  }
  // cancellation_occurs :1990-2033
  double w3 = (c_vc.dz[k] * c.kib *
               (c.sb[XE] * hyx * t0[XE] + c.sb[YN] * hxy * t0[YN] + c.sb[XW] * hyxw * t0[XW] + c.sb[YS] * hxys * t0[YS]));
  w3 = w3 + (dz_bottom * c.kit1 *
             (c.st[XE] * hyx * t1[XE] + c.st[YN] * hxy * t1[YN] + c.st[XW] * hyxw * t1[XW] + c.st[YS] * hxys * t1[YS]));
  return -c.KMASK * 0.5 * w3;
}

// One thread per (physical column, level); the tracers are looped inside so that the diffusivities and slopes
// are read once. FZTOP of the reference (the flux through the top face, carried from the k-1 call) is
// re-evaluated by the thread that needs it: same expression, same bits.
template <bool CANCEL>
__global__ void __launch_bounds__(128, CANCEL ? 4 : 3) gm_flux_kernel(const GmArgs a) {
  const GridView& g = a.g;
  const int i = (g.ib - 1) + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (g.jb - 1) + blockIdx.y;
  const int km = g.km, k = 1 + (int)blockIdx.z;
  if (i > g.ie - 1 || j > g.je - 1) return;
  const size_t n2 = g.n2, nxb = g.nxb, q = (size_t)j * nxb + i, qw = q - 1, qs = q - nxb;
  const int kmt = g.KMT[q], kmte = g.KMT[q + 1], kmtw = g.KMT[qw], kmtn = g.KMT[q + nxb], kmts = g.KMT[qs];
  const double hyx = a.HYX[q], hyxw = a.HYX[qw], hxy = a.HXY[q], hxys = a.HXY[qs];
  const double scale_r = g.TAREA_R[q];
  const int kp1 = (k == km) ? k : k + 1;
  const GmFace fe = gm_face_coef<CANCEL>(a, q, q + 1, XE, XW, hyx, k), fw = gm_face_coef<CANCEL>(a, qw, q, XE, XW, hyxw, k);
  const GmFace fn = gm_face_coef<CANCEL>(a, q, q + nxb, YN, YS, hxy, k), fs = gm_face_coef<CANCEL>(a, qs, q, YN, YS, hxys, k);
  GmFz ct, cb;
  if (k > 1) ct = gm_fz_coef<CANCEL>(a, q, k - 1, kmt);
  if (k < km) cb = gm_fz_coef<CANCEL>(a, q, k, kmt);
  for (int n = 0; n < g.nt; n++) {
    const double* Tn = a.TMIX + (size_t)n * km * n2;
    const double fxe = gm_face_flux<CANCEL>(a, fe, Tn, q, q + 1, k, kp1), fxw = gm_face_flux<CANCEL>(a, fw, Tn, qw, q, k, kp1);
    const double fyn = gm_face_flux<CANCEL>(a, fn, Tn, q, q + nxb, k, kp1), fys = gm_face_flux<CANCEL>(a, fs, Tn, qs, q, k, kp1);
    // masked tracer differences of levels k-1, k, k+1: TX(i), TX(i-1), TY(j), TY(j-1)
    double t[3][4];
#pragma unroll
    for (int l = 0; l < 3; l++) {
      const int kk = k - 1 + l;
      if (kk < 1 || kk > km) continue;
      const double* T = Tn + (size_t)(kk - 1) * n2;
      const double me = (kk <= kmt && kk <= kmte) ? 1.0 : 0.0, mw = (kk <= kmtw && kk <= kmt) ? 1.0 : 0.0;
      const double mn = (kk <= kmt && kk <= kmtn) ? 1.0 : 0.0, ms = (kk <= kmts && kk <= kmt) ? 1.0 : 0.0;
      const double t0 = T[q];
      t[l][XE] = me * (T[q + 1] - t0);
      t[l][XW] = mw * (t0 - T[qw]);
      t[l][YN] = mn * (T[q + nxb] - t0);
      t[l][YS] = ms * (t0 - T[qs]);
    }
    // zero flux B.C. at the surface (:1664); FZTOP = fz of the level above otherwise (:2036, :2065)
    const double fztop = (k == 1) ? 0.0 : gm_fz_eval<CANCEL>(a, ct, k - 1, t[0], t[1], hyx, hyxw, hxy, hxys);
    double G;
    if (k < km) {
      const double fz = gm_fz_eval<CANCEL>(a, cb, k, t[1], t[2], hyx, hyxw, hxy, hxys);
      G = (fxe - fxw + fyn - fys + fztop - fz) * c_vc.dzr[k] * scale_r;
    } else {
      G = (fxe - fxw + fyn - fys + fztop) * c_vc.dzr[k] * scale_r;
    }
    a.HDT[((size_t)n * km + (k - 1)) * n2 + q] = G;
  }
}

__global__ void gm_metric_kernel(const double* __restrict__ HTE, const double* __restrict__ HUS,
                                 const double* __restrict__ HTN, const double* __restrict__ HUW,
                                 double* __restrict__ HYX, double* __restrict__ HXY, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  HYX[q] = HTE[q] / HUS[q];  // hmix_gm_submeso_share.F90:127-131
  HXY[q] = HTN[q] / HUW[q];
}
}  // namespace

int gm_alloc_fields() {
  const int km = G.km;
  POP_TRY(alloc_field("TLAT", 1, false));
  for (const char* f : {"GM_HYX", "GM_HXY", "GM_RBR"}) POP_TRY(alloc_field(f, 1, false));
  POP_TRY(alloc_field("GM_SL", 8 * km, false));
  POP_TRY(alloc_field("GM_KI", 2 * km, false));
  POP_TRY(alloc_field("GM_HD", 2 * km, false));
  POP_TRY(alloc_field("GM_KT", 2 * km, false));
  POP_TRY(alloc_field("GM_HDT", km * G.nt, false));
  if (G.cfg.vmix_itype == POP_VMIX_GIVEN) POP_TRY(alloc_field("GM_VDC_BASE", G.vdc_nk * G.vdc_nd, false));
  G.gm_dirty = true;
  return POP_SUCCESS;
}

// init_meso_mixing + init_gm: metric ratios on the device; the inverse Rossby radius on the host, whose libm
// sin() is the one the reference's FCORT = 2*omega*sin(TLAT) (grid.F90:1159) and the oracle use
static int gm_init_dev() {
  POP_LAUNCH(gm_metric_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("HTE"), fld("HUS"), fld("HTN"), fld("HUW"),
             fld("GM_HYX"), fld("GM_HXY"), G.n2);
  std::vector<double> tlat(G.n2), rbr(G.n2);
  POP_CHECK_CUDA(cudaMemcpyAsync(tlat.data(), fld("TLAT"), sizeof(double) * G.n2, cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  for (size_t q = 0; q < G.n2; q++) {  // hmix_gm.F90:881-886
    const double fcort = 2.0 * POP_OMEGA * sin(tlat[q]);
    double r = fabs(fcort) / 200.0;
    r = (r < 1.0 / 1.5e+6) ? r : 1.0 / 1.5e+6;
    r = (r > 1.e-7) ? r : 1.e-7;
    rbr[q] = r;
  }
  POP_CHECK_CUDA(cudaMemcpyAsync(fld("GM_RBR"), rbr.data(), sizeof(double) * G.n2, cudaMemcpyHostToDevice, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}

// vmix GIVEN stands in for KPP, which rebuilds VDC every step before GM adds to it: each step starts from the
// coefficients the caller supplied last
int gm_begin_step() {
  if (G.cfg.hmix_tracer_itype != POP_HMIX_GM || G.cfg.vmix_itype != POP_VMIX_GIVEN) return POP_SUCCESS;
  const size_t bytes = sizeof(double) * G.n2 * G.vdc_nk * G.vdc_nd;
  if (G.gm_vdc_dirty) {
    POP_CHECK_CUDA(cudaMemcpyAsync(fld("GM_VDC_BASE"), fld("VDC"), bytes, cudaMemcpyDeviceToDevice, G.stream));
    G.gm_vdc_dirty = false;
  } else {
    POP_CHECK_CUDA(cudaMemcpyAsync(fld("VDC"), fld("GM_VDC_BASE"), bytes, cudaMemcpyDeviceToDevice, G.stream));
  }
  return POP_SUCCESS;
}

// all levels of hdifft_gm for the mixing-time tracers: fills GM_HDT and adds VDC_GM to VDC
int gm_tendency_dev(const double* TMIX) {
  ScopedTimer tm("HMIX_TRACER_GM");
  if (G.gm_dirty) {
    POP_TRY(gm_init_dev());
    G.gm_dirty = false;
  }
  GmArgs a;
  memset(&a, 0, sizeof(a));
  a.g = grid_view();
  a.TMIX = TMIX;
  a.HYX = fld("GM_HYX"); a.HXY = fld("GM_HXY"); a.RBR = fld("GM_RBR"); a.DXT = fld("DXT"); a.DYT = fld("DYT");
  a.SL = fld("GM_SL"); a.KI = fld("GM_KI"); a.HD = fld("GM_HD"); a.KT = fld("GM_KT");
  a.VDC = fld("VDC"); a.HDT = fld("GM_HDT");
  a.so = StateOpt{G.cfg.state_itype, G.cfg.state_range_iopt};
  const pop_config& c = G.cfg;
  a.ah = c.ah_gm; a.ah_bolus = c.ah_bolus; a.ah_bkg_srfbl = c.ah_bkg_srfbl; a.slm_r = c.slm_r; a.slm_b = c.slm_b;
  a.diff_tapering = (c.slm_r != c.slm_b);                       // hmix_gm.F90:964-968
  a.cancel = !(a.diff_tapering || c.ah_gm != c.ah_bolus);       // :970-983
  if (G.gm_force_general) a.cancel = 0;
  const int nt = 128;
  dim3 g1((unsigned)((G.nxb - 2 + nt - 1) / nt), (unsigned)(G.nyb - 2), (unsigned)G.km);
  POP_LAUNCH(gm_column_kernel, g1, nt, 0, a);
  dim3 g2((unsigned)((G.ie - G.ib + 1 + nt - 1) / nt), (unsigned)(G.je - G.jb + 1), (unsigned)G.km);
  if (a.cancel) POP_LAUNCH(gm_flux_kernel<true>, g2, nt, 0, a);
  else POP_LAUNCH(gm_flux_kernel<false>, g2, nt, 0, a);
  return pop_post_launch("hdifft_gm");
}
