/* o_gm.c -- ORACLE (test infrastructure, never linked into the product): Gent-McWilliams / Redi
 * isopycnal tracer mixing restated slab by slab from
 *   tracer_diffs_and_isopyc_slopes   source/hmix_gm_submeso_share.F90:149-434
 *   init_gm                          source/hmix_gm.F90:283-1095
 *   hdifft_gm                        source/hmix_gm.F90:1102-2219
 * for the option set pop_config exposes: kappa_isop_type = kappa_thic_type = 'constant', kappa_freq = 'never',
 * slope_control = 'notanh', use_const_ah_bkg_srfbl = .true., ah_bkg_bottom = 0, transition_layer_on = .false.,
 * no KPP boundary layer (BL_DEPTH = zw(1)), diagnostics (bolus velocities, tavg, cfl) omitted.
 * Both the cancellation_occurs branch (ah == ah_bolus and slm_r == slm_b) and the general branch are here.
 * Parity unpinned: the reference tree holds no golden data for GM; the restatement is pinned by its own
 * invariants (tests/test_oracle_gm.py). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pop_oracle.h"

#define NB M.nblocks
enum { IEAST = 1, IWEST = 2, JNORTH = 1, JSOUTH = 2, KTP = 1, KBT = 2 };
static const double eps = 1.0e-10, eps2 = 1.0e-20; /* pop_constants.F90:55-56 */

typedef struct {
  int ready, diff_tapering, cancellation_occurs, vdc_saved;
  double *HYX, *HXY, *HYXW, *HXYS, *RBR, *BL_DEPTH; /* [NB][n2] */
  double *SLX, *SLY, *SF_SLX, *SF_SLY;              /* [NB][km][2 halves][2 directions][n2] */
  double *TX, *TY, *TZ;                             /* [NB][nt][km][n2] */
  double *RX, *RY;                                  /* [NB][km][2][n2] */
  double *KAPPA_ISOP, *KAPPA_THIC, *HOR_DIFF;       /* [NB][km][2][n2] */
  double *FZTOP;                                    /* [NB][nt][n2] */
  double* VDC_BASE;                                 /* vmix GIVEN: the caller's coefficients */
} ogm;
static ogm Gm;
static int force_general = 0; /* tests: take the general (no cancellation) branch even when the terms cancel */
void o_gm_force_general(int on) { force_general = on; }

/* (i,j) plane of a (nxb,nyb,2,2,km) array: direction d, half h, level kk (all 1-based) */
#define S4(a, b, d, h, kk) ((a) + ((size_t)(b)*M.km * 4 + ((size_t)((kk)-1) * 2 + ((h)-1)) * 2 + ((d)-1)) * M.n2)
/* (i,j) plane of a (nxb,nyb,2,km) array */
#define S2(a, b, h, kk) ((a) + ((size_t)(b)*M.km * 2 + (size_t)((kk)-1) * 2 + ((h)-1)) * M.n2)
#define T4(a, b, kk, n) ((a) + ((size_t)(b)*M.nt * M.km + (size_t)((n)-1) * M.km + ((kk)-1)) * M.n2)

static double* zalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }

void o_gm_reset(void) {
  double** p[] = {&Gm.HYX, &Gm.HXY, &Gm.HYXW, &Gm.HXYS, &Gm.RBR, &Gm.BL_DEPTH, &Gm.SLX, &Gm.SLY, &Gm.SF_SLX,
                  &Gm.SF_SLY, &Gm.TX, &Gm.TY, &Gm.TZ, &Gm.RX, &Gm.RY, &Gm.KAPPA_ISOP, &Gm.KAPPA_THIC,
                  &Gm.HOR_DIFF, &Gm.FZTOP, &Gm.VDC_BASE};
  for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) free(*p[i]);
  memset(&Gm, 0, sizeof(Gm));
}

/* init_meso_mixing (hmix_gm_submeso_share.F90:65-141) + init_gm (hmix_gm.F90:770-990) */
static void gm_init(void) {
  const size_t n2t = M.n2 * (size_t)NB;
  const int km = M.km, nt = M.nt;
  const pop_config* c = &M.cfg;
  Gm.HYX = zalloc(n2t); Gm.HXY = zalloc(n2t); Gm.HYXW = zalloc(n2t); Gm.HXYS = zalloc(n2t);
  Gm.RBR = zalloc(n2t); Gm.BL_DEPTH = zalloc(n2t);
  Gm.SLX = zalloc(n2t * 4 * km); Gm.SLY = zalloc(n2t * 4 * km);
  Gm.SF_SLX = zalloc(n2t * 4 * km); Gm.SF_SLY = zalloc(n2t * 4 * km);
  Gm.TX = zalloc(n2t * km * nt); Gm.TY = zalloc(n2t * km * nt); Gm.TZ = zalloc(n2t * km * nt);
  Gm.RX = zalloc(n2t * 2 * km); Gm.RY = zalloc(n2t * 2 * km);
  Gm.KAPPA_ISOP = zalloc(n2t * 2 * km); Gm.KAPPA_THIC = zalloc(n2t * 2 * km); Gm.HOR_DIFF = zalloc(n2t * 2 * km);
  Gm.FZTOP = zalloc(n2t * nt);
  for (size_t q = 0; q < n2t * 2 * km; q++) { /* :856-858 */
    Gm.KAPPA_ISOP[q] = c->ah_gm;
    Gm.KAPPA_THIC[q] = c->ah_bolus;
    Gm.HOR_DIFF[q] = c->ah_gm;
  }
  for (int b = 0; b < NB; b++) {
    double *HYX = B2(Gm.HYX, b), *HXY = B2(Gm.HXY, b), *HYXW = B2(Gm.HYXW, b), *HXYS = B2(Gm.HXYS, b);
    double* RBR = B2(Gm.RBR, b);
    for (size_t q = 0; q < M.n2; q++) {
      HYX[q] = B2(M.HTE, b)[q] / B2(M.HUS, b)[q];
      HXY[q] = B2(M.HTN, b)[q] / B2(M.HUW, b)[q];
    }
    for (int j = 1; j <= M.nyb; j++) /* eoshift(dim=1, shift=-1) / eoshift(dim=2, shift=-1): :870-871 */
      for (int i = 1; i <= M.nxb; i++) {
        HYXW[IX2(i, j)] = (i > 1) ? HYX[IX2(i - 1, j)] : 0.0;
        HXYS[IX2(i, j)] = (j > 1) ? HXY[IX2(i, j - 1)] : 0.0;
      }
    for (size_t q = 0; q < M.n2; q++) { /* inverse Rossby radius, :881-886 */
      const double fcort = 2.0 * O_OMEGA * sin(B2(M.TLAT, b)[q]); /* grid.F90:1159 */
      double r = fabs(fcort) / 200.0;
      r = (r < 1.0 / 1.5e+6) ? r : 1.0 / 1.5e+6;
      r = (r > 1.e-7) ? r : 1.e-7;
      RBR[q] = r;
    }
  }
  Gm.diff_tapering = (c->slm_r != c->slm_b); /* :964-968 */
  Gm.cancellation_occurs = !(Gm.diff_tapering || c->ah_gm != c->ah_bolus); /* :970-983, const/const kappa */
  if (force_general) Gm.cancellation_occurs = 0;
  Gm.ready = 1;
}

/* vmix GIVEN stands in for KPP, whose coefficients are rebuilt every step before GM adds to them
   (vmix_kpp.F90 vmix_coeffs_kpp at k == 1): every step starts from the caller's coefficients. */
void o_gm_begin_step(void) {
  if (M.cfg.hmix_tracer_itype != POP_HMIX_GM || M.cfg.vmix_itype != POP_VMIX_GIVEN) return;
  const size_t n = (size_t)NB * M.vdc_nd * M.vdc_nk * M.n2;
  if (!Gm.vdc_saved) {
    free(Gm.VDC_BASE);
    Gm.VDC_BASE = (double*)malloc(sizeof(double) * n);
    memcpy(Gm.VDC_BASE, M.VDC, sizeof(double) * n);
    Gm.vdc_saved = 1;
  } else {
    memcpy(M.VDC, Gm.VDC_BASE, sizeof(double) * n);
  }
}

/* tracer_diffs_and_isopyc_slopes: hmix_gm_submeso_share.F90:149-434 */
static void gm_slopes(const double* TMIX, int b) {
  const int km = M.km, nt = M.nt, nxb = M.nxb, nyb = M.nyb;
  const size_t n2 = M.n2;
  const int *KMT = M.KMT + (size_t)b * n2, *KMTE = M.KMTE + (size_t)b * n2, *KMTN = M.KMTN + (size_t)b * n2;
  double *KMASK = zalloc(n2), *KMASKE = zalloc(n2), *KMASKN = zalloc(n2), *DRDT = zalloc(n2), *DRDS = zalloc(n2);
  double *RZ = zalloc(n2), *TXP[2], *TYP[2], *TZP[2], *TEMP[2];
  for (int s = 0; s < 2; s++) { TXP[s] = zalloc(n2); TYP[s] = zalloc(n2); TZP[s] = zalloc(n2); TEMP[s] = zalloc(n2); }
  int kn = 0, ks = 1;
  /* horizontal differences and RX/RY at level kk from TEMP/TXP/TYP slot s (:221-290 and :329-386) */
#define LEVEL_DIFFS(kk, s)                                                                                  \
  do {                                                                                                      \
    for (int j = 1; j <= nyb; j++)                                                                          \
      for (int i = 1; i <= nxb - 1; i++)                                                                    \
        TXP[s][IX2(i, j)] = KMASKE[IX2(i, j)] * (TEMP[s][IX2(i + 1, j)] - TEMP[s][IX2(i, j)]);              \
    for (int j = 1; j <= nyb - 1; j++)                                                                      \
      for (int i = 1; i <= nxb; i++)                                                                        \
        TYP[s][IX2(i, j)] = KMASKN[IX2(i, j)] * (TEMP[s][IX2(i, j + 1)] - TEMP[s][IX2(i, j)]);              \
    for (int n = 1; n <= nt; n++) {                                                                         \
      const double* T = KN4(TMIX, kk, n);                                                                   \
      double *tx = T4(Gm.TX, b, kk, n), *ty = T4(Gm.TY, b, kk, n);                                          \
      for (int j = 1; j <= nyb; j++)                                                                        \
        for (int i = 1; i <= nxb - 1; i++)                                                                  \
          tx[IX2(i, j)] = KMASKE[IX2(i, j)] * (T[IX2(i + 1, j)] - T[IX2(i, j)]);                            \
      for (int j = 1; j <= nyb - 1; j++)                                                                    \
        for (int i = 1; i <= nxb; i++)                                                                      \
          ty[IX2(i, j)] = KMASKN[IX2(i, j)] * (T[IX2(i, j + 1)] - T[IX2(i, j)]);                            \
    }                                                                                                       \
    o_state(kk, kk, KN4(TMIX, kk, 1), KN4(TMIX, kk, 2), b, NULL, NULL, DRDT, DRDS);                         \
    {                                                                                                       \
      const double *tx2 = T4(Gm.TX, b, kk, 2), *ty2 = T4(Gm.TY, b, kk, 2);                                  \
      double *rxe = S2(Gm.RX, b, IEAST, kk), *rxw = S2(Gm.RX, b, IWEST, kk);                                \
      double *ryn = S2(Gm.RY, b, JNORTH, kk), *rys = S2(Gm.RY, b, JSOUTH, kk);                              \
      for (size_t q = 0; q < n2; q++) {                                                                     \
        rxe[q] = DRDT[q] * TXP[s][q] + DRDS[q] * tx2[q];                                                    \
        ryn[q] = DRDT[q] * TYP[s][q] + DRDS[q] * ty2[q];                                                    \
      }                                                                                                     \
      for (int j = 1; j <= nyb; j++)                                                                        \
        for (int i = 2; i <= nxb; i++)                                                                      \
          rxw[IX2(i, j)] = DRDT[IX2(i, j)] * TXP[s][IX2(i - 1, j)] + DRDS[IX2(i, j)] * tx2[IX2(i - 1, j)];  \
      for (int j = 2; j <= nyb; j++)                                                                        \
        for (int i = 1; i <= nxb; i++)                                                                      \
          rys[IX2(i, j)] = DRDT[IX2(i, j)] * TYP[s][IX2(i, j - 1)] + DRDS[IX2(i, j)] * ty2[IX2(i, j - 1)];  \
    }                                                                                                       \
  } while (0)

  for (int kk = 1; kk <= km; kk++) {
    for (size_t q = 0; q < n2; q++) KMASK[q] = (kk < KMT[q]) ? 1.0 : 0.0;
    if (kk == 1) {
      const double* T1 = KN4(TMIX, kk, 1);
      for (size_t q = 0; q < n2; q++) {
        KMASKE[q] = (kk <= KMT[q] && kk <= KMTE[q]) ? 1.0 : 0.0;
        KMASKN[q] = (kk <= KMT[q] && kk <= KMTN[q]) ? 1.0 : 0.0;
        TEMP[kn][q] = (-2.0 > T1[q]) ? -2.0 : T1[q];
      }
      LEVEL_DIFFS(kk, kn);
    }
    if (kk < km) {
      const double *Tk = KN4(TMIX, kk, 1), *Tk1 = KN4(TMIX, kk + 1, 1);
      const double *Sk = KN4(TMIX, kk, 2), *Sk1 = KN4(TMIX, kk + 1, 2);
      double *tz1 = T4(Gm.TZ, b, kk + 1, 1), *tz2 = T4(Gm.TZ, b, kk + 1, 2);
      for (size_t q = 0; q < n2; q++) {
        TEMP[ks][q] = (-2.0 > Tk1[q]) ? -2.0 : Tk1[q];
        tz1[q] = Tk[q] - Tk1[q];
        tz2[q] = Sk[q] - Sk1[q];
        TZP[ks][q] = TEMP[kn][q] - TEMP[ks][q];
        double rz = DRDT[q] * TZP[ks][q] + DRDS[q] * tz2[q]; /* :302-303 */
        RZ[q] = (rz < -eps2) ? rz : -eps2;
      }
      {
        const double *rxe = S2(Gm.RX, b, IEAST, kk), *rxw = S2(Gm.RX, b, IWEST, kk);
        const double *ryn = S2(Gm.RY, b, JNORTH, kk), *rys = S2(Gm.RY, b, JSOUTH, kk);
        double *sxe = S4(Gm.SLX, b, IEAST, KBT, kk), *sxw = S4(Gm.SLX, b, IWEST, KBT, kk);
        double *syn = S4(Gm.SLY, b, JNORTH, KBT, kk), *sys = S4(Gm.SLY, b, JSOUTH, KBT, kk);
        for (size_t q = 0; q < n2; q++) { /* :305-310 */
          sxe[q] = KMASK[q] * rxe[q] / RZ[q];
          sxw[q] = KMASK[q] * rxw[q] / RZ[q];
          syn[q] = KMASK[q] * ryn[q] / RZ[q];
          sys[q] = KMASK[q] * rys[q] / RZ[q];
        }
      }
      for (size_t q = 0; q < n2; q++) { /* :318-321 */
        KMASKE[q] = (kk + 1 <= KMT[q] && kk + 1 <= KMTE[q]) ? 1.0 : 0.0;
        KMASKN[q] = (kk + 1 <= KMT[q] && kk + 1 <= KMTN[q]) ? 1.0 : 0.0;
      }
      LEVEL_DIFFS(kk + 1, ks);
      {
        const double *rxe = S2(Gm.RX, b, IEAST, kk + 1), *rxw = S2(Gm.RX, b, IWEST, kk + 1);
        const double *ryn = S2(Gm.RY, b, JNORTH, kk + 1), *rys = S2(Gm.RY, b, JSOUTH, kk + 1);
        double *sxe = S4(Gm.SLX, b, IEAST, KTP, kk + 1), *sxw = S4(Gm.SLX, b, IWEST, KTP, kk + 1);
        double *syn = S4(Gm.SLY, b, JNORTH, KTP, kk + 1), *sys = S4(Gm.SLY, b, JSOUTH, KTP, kk + 1);
        for (size_t q = 0; q < n2; q++) { /* :388-405 */
          double rz = DRDT[q] * TZP[ks][q] + DRDS[q] * tz2[q];
          rz = (rz < -eps2) ? rz : -eps2;
          if (kk + 1 <= KMT[q]) {
            sxe[q] = rxe[q] / rz;
            sxw[q] = rxw[q] / rz;
            syn[q] = ryn[q] / rz;
            sys[q] = rys[q] / rz;
          }
        }
      }
    }
    int t = kn; kn = ks; ks = t;
  }
#undef LEVEL_DIFFS
  free(KMASK); free(KMASKE); free(KMASKN); free(DRDT); free(DRDS); free(RZ);
  for (int s = 0; s < 2; s++) { free(TXP[s]); free(TYP[s]); free(TZP[s]); free(TEMP[s]); }
}

/* the k == 1 block of hdifft_gm: diffusivities, tapers, boundary conditions, SF_SLX/SF_SLY (:1190-1706) */
static void gm_coefficients(int b) {
  const int km = M.km;
  const size_t n2 = M.n2;
  const pop_config* c = &M.cfg;
  const int* KMT = M.KMT + (size_t)b * n2;
  const double *DXT = B2(M.DXT, b), *DYT = B2(M.DYT, b), *RBR = B2(Gm.RBR, b);
  double* BL = B2(Gm.BL_DEPTH, b);
  double* hd = S2(Gm.HOR_DIFF, b, KTP, 1);
  for (size_t q = 0; q < n2; q++) { hd[q] = c->ah_bkg_srfbl; BL[q] = M.zw[1]; } /* :1207-1209 */
  for (size_t q = 0; q < n2 * 2 * km; q++) { /* :1345-1346, :1367-1368 */
    (Gm.KAPPA_ISOP + (size_t)b * 2 * km * n2)[q] = c->ah_gm;
    (Gm.KAPPA_THIC + (size_t)b * 2 * km * n2)[q] = c->ah_bolus;
  }
  for (int kk = 1; kk <= km; kk++) {
    const double dz_bottom = (kk == 1) ? 0.0 : M.zt[kk - 1];
    for (int sub = KTP; sub <= KBT; sub++) {
      const int kid = kk + sub - 2;
      const double *sx1 = S4(Gm.SLX, b, 1, sub, kk), *sx2 = S4(Gm.SLX, b, 2, sub, kk);
      const double *sy1 = S4(Gm.SLY, b, 1, sub, kk), *sy2 = S4(Gm.SLY, b, 2, sub, kk);
      double *KI = S2(Gm.KAPPA_ISOP, b, sub, kk), *KT = S2(Gm.KAPPA_THIC, b, sub, kk), *HD = S2(Gm.HOR_DIFF, b, sub, kk);
      for (size_t q = 0; q < n2; q++) {
        const double SLA = M.dzw[kid] * sqrt(0.5 * ((sx1[q] * sx1[q] + sx2[q] * sx2[q]) / (DXT[q] * DXT[q]) +
                                                    (sy1[q] * sy1[q] + sy2[q] * sy2[q]) / (DYT[q] * DYT[q]))) + eps;
        /* Large et al. (1997) taper near the surface, notanh form (:1462-1470) */
        double W1 = M.zt[kk] * RBR[q] / SLA;
        W1 = (1.0 < W1) ? 1.0 : W1;
        double TAPER1 = (0.5 + 2.0 * (W1 - 0.5) * (1.0 - fabs(W1 - 0.5)));
        if (!(dz_bottom <= BL[q])) TAPER1 = 1.0;
        double TAPER2 = 1.0, TAPER3 = 1.0; /* slope_control_notanh :1503-1538 */
        if (SLA > 0.2 * c->slm_r && SLA < 0.6 * c->slm_r)
          TAPER2 = 0.5 * (1.0 - (2.5 * SLA / c->slm_r - 1.0) * (4.0 - fabs(10.0 * SLA / c->slm_r - 4.0)));
        else if (SLA >= 0.6 * c->slm_r)
          TAPER2 = 0.0;
        if (Gm.diff_tapering) {
          if (SLA > 0.2 * c->slm_b && SLA < 0.6 * c->slm_b)
            TAPER3 = 0.5 * (1.0 - (2.5 * SLA / c->slm_b - 1.0) * (4.0 - fabs(10.0 * SLA / c->slm_b - 4.0)));
          else if (SLA >= 0.6 * c->slm_b)
            TAPER3 = 0.0;
        } else {
          TAPER3 = TAPER2;
        }
        if (!(kk == 1 && sub == KTP)) /* :1617-1624, KAPPA_VERTICAL = 1 */
          HD[q] = (dz_bottom <= BL[q]) ? c->ah_bkg_srfbl * (1.0 - TAPER1 * TAPER2) * 1.0 : 0.0;
        KI[q] = TAPER1 * TAPER2 * KI[q];
        KT[q] = TAPER1 * TAPER3 * KT[q];
      }
    }
    {
      double *KI = S2(Gm.KAPPA_ISOP, b, KBT, kk), *KT = S2(Gm.KAPPA_THIC, b, KBT, kk);
      for (size_t q = 0; q < n2; q++)
        if (kk == KMT[q]) { KI[q] = 0.0; KT[q] = 0.0; } /* bottom B.C. :1652-1655 */
    }
  }
  memset(S2(Gm.KAPPA_ISOP, b, KTP, 1), 0, sizeof(double) * n2); /* top B.C. :1661-1662 */
  memset(S2(Gm.KAPPA_THIC, b, KTP, 1), 0, sizeof(double) * n2);
  memset(Gm.FZTOP + (size_t)b * M.nt * n2, 0, sizeof(double) * n2 * M.nt); /* :1664 */
  for (int kk = 1; kk <= km; kk++)
    for (int sub = KTP; sub <= KBT; sub++) {
      const double* KT = S2(Gm.KAPPA_THIC, b, sub, kk);
      for (int d = 1; d <= 2; d++) {
        const double *sx = S4(Gm.SLX, b, d, sub, kk), *sy = S4(Gm.SLY, b, d, sub, kk);
        double *fx = S4(Gm.SF_SLX, b, d, sub, kk), *fy = S4(Gm.SF_SLY, b, d, sub, kk);
        for (size_t q = 0; q < n2; q++)
          if (kk <= KMT[q]) { /* :1682-1700 */
            fx[q] = KT[q] * sx[q] * M.dz[kk];
            fy[q] = KT[q] * sy[q] * M.dz[kk];
          }
      }
    }
}

/* hdifft_gm: hmix_gm.F90:1102-2219. HDTK is (nxb,nyb,nt), zeroed by the dispatcher. */
void o_hdifft_gm(int k, double* GTK, const double* TMIX, const double* UMIX, const double* VMIX, int b) {
  (void)UMIX; (void)VMIX;
  const int km = M.km, nt = M.nt, nxb = M.nxb, nyb = M.nyb;
  const size_t n2 = M.n2;
#pragma omp critical(o_gm_init)
  if (!Gm.ready) gm_init();
  const int *KMT = M.KMT + (size_t)b * n2, *KMTE = M.KMTE + (size_t)b * n2, *KMTN = M.KMTN + (size_t)b * n2;
  const double *HYX = B2(Gm.HYX, b), *HXY = B2(Gm.HXY, b), *HYXW = B2(Gm.HYXW, b), *HXYS = B2(Gm.HXYS, b);
  const double* TAREA_R = B2(M.TAREA_R, b);
  if (k == 1) {
    gm_slopes(TMIX, b); /* horizontal_mix.F90:551-553 */
    gm_coefficients(b);
  }
  double *CX = zalloc(n2), *CY = zalloc(n2), *KMASK = zalloc(n2);
  double *WORK1 = zalloc(n2), *WORK2 = zalloc(n2), *WORK3 = zalloc(n2), *WORK4 = zalloc(n2);
  double *FX = zalloc(n2 * nt), *FY = zalloc(n2 * nt);
  for (size_t q = 0; q < n2; q++) { /* :1214-1217 */
    CX[q] = (k <= KMT[q] && k <= KMTE[q]) ? HYX[q] * 0.25 : 0.0;
    CY[q] = (k <= KMT[q] && k <= KMTN[q]) ? HXY[q] * 0.25 : 0.0;
    KMASK[q] = (k < KMT[q]) ? 1.0 : 0.0;
  }
#define SQ(x) ((x) * (x))
  if (k < km) { /* effective vertical diffusion coefficient :1725-1743 */
    const double *KIb = S2(Gm.KAPPA_ISOP, b, KBT, k), *KIt = S2(Gm.KAPPA_ISOP, b, KTP, k + 1);
    const double *xe = S4(Gm.SLX, b, IEAST, KBT, k), *xw = S4(Gm.SLX, b, IWEST, KBT, k);
    const double *yn = S4(Gm.SLY, b, JNORTH, KBT, k), *ys = S4(Gm.SLY, b, JSOUTH, KBT, k);
    const double *xe1 = S4(Gm.SLX, b, IEAST, KTP, k + 1), *xw1 = S4(Gm.SLX, b, IWEST, KTP, k + 1);
    const double *yn1 = S4(Gm.SLY, b, JNORTH, KTP, k + 1), *ys1 = S4(Gm.SLY, b, JSOUTH, KTP, k + 1);
    for (size_t q = 0; q < n2; q++)
      WORK1[q] = M.dzw[k] * KMASK[q] * TAREA_R[q] *
                 (M.dz[k] * 0.25 * KIb[q] *
                      (HYX[q] * SQ(xe[q]) + HYXW[q] * SQ(xw[q]) + HXY[q] * SQ(yn[q]) + HXYS[q] * SQ(ys[q])) +
                  M.dz[k + 1] * 0.25 * KIt[q] *
                      (HYX[q] * SQ(xe1[q]) + HYXW[q] * SQ(xw1[q]) + HXY[q] * SQ(yn1[q]) + HXYS[q] * SQ(ys1[q])));
    for (int n = 0; n < M.vdc_nd; n++) {
      double* VDC = M.VDC + (((size_t)b * M.vdc_nd + n) * M.vdc_nk + (k - M.vdc_k0)) * n2;
      for (size_t q = 0; q < n2; q++) VDC[q] = VDC[q] + WORK1[q];
    }
  }
  { /* combined isopycnal and horizontal diffusion coefficients :1765-1791 */
    const double *KIt = S2(Gm.KAPPA_ISOP, b, KTP, k), *KIb = S2(Gm.KAPPA_ISOP, b, KBT, k);
    const double *HDt = S2(Gm.HOR_DIFF, b, KTP, k), *HDb = S2(Gm.HOR_DIFF, b, KBT, k);
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb - 1; i++) {
        const size_t q = IX2(i, j), e = IX2(i + 1, j);
        WORK3[q] = KIt[q] + HDt[q] + KIb[q] + HDb[q] + KIt[e] + HDt[e] + KIb[e] + HDb[e];
      }
    for (int j = 1; j <= nyb - 1; j++)
      for (int i = 1; i <= nxb; i++) {
        const size_t q = IX2(i, j), e = IX2(i, j + 1);
        WORK4[q] = KIt[q] + HDt[q] + KIb[q] + HDb[q] + KIt[e] + HDt[e] + KIb[e] + HDb[e];
      }
  }
  const int kp1 = (k == km) ? k : k + 1;
  const double dz_bottom = (k < km) ? M.dz[kp1] : 0.0, factor = (k < km) ? 1.0 : 0.0;
  for (int n = 1; n <= nt; n++) { /* :1825-1826 */
    const double *tx = T4(Gm.TX, b, k, n), *ty = T4(Gm.TY, b, k, n);
    double *fx = FX + (size_t)(n - 1) * n2, *fy = FY + (size_t)(n - 1) * n2;
    for (size_t q = 0; q < n2; q++) {
      fx[q] = M.dz[k] * CX[q] * tx[q] * WORK3[q];
      fy[q] = M.dz[k] * CY[q] * ty[q] * WORK4[q];
    }
  }
  if (!Gm.cancellation_occurs) { /* :1830-1896 */
    const double *KIt = S2(Gm.KAPPA_ISOP, b, KTP, k), *KIb = S2(Gm.KAPPA_ISOP, b, KBT, k);
    {
      const double *set = S4(Gm.SLX, b, IEAST, KTP, k), *seb = S4(Gm.SLX, b, IEAST, KBT, k);
      const double *swt = S4(Gm.SLX, b, IWEST, KTP, k), *swb = S4(Gm.SLX, b, IWEST, KBT, k);
      const double *fet = S4(Gm.SF_SLX, b, IEAST, KTP, k), *feb = S4(Gm.SF_SLX, b, IEAST, KBT, k);
      const double *fwt = S4(Gm.SF_SLX, b, IWEST, KTP, k), *fwb = S4(Gm.SF_SLX, b, IWEST, KBT, k);
      for (int j = 1; j <= nyb; j++)
        for (int i = 1; i <= nxb - 1; i++) {
          const size_t q = IX2(i, j), e = IX2(i + 1, j);
          WORK1[q] = KIt[q] * set[q] * M.dz[k] - fet[q];
          WORK2[q] = KIb[q] * seb[q] * M.dz[k] - feb[q];
          WORK3[q] = KIt[e] * swt[e] * M.dz[k] - fwt[e];
          WORK4[q] = KIb[e] * swb[e] * M.dz[k] - fwb[e];
        }
    }
    for (int n = 1; n <= nt; n++) {
      if (n > 2 && k < km) {
        double* tz = T4(Gm.TZ, b, k + 1, n);
        const double *Tk = KN4(TMIX, k, n), *Tk1 = KN4(TMIX, k + 1, n);
        for (size_t q = 0; q < n2; q++) tz[q] = Tk[q] - Tk1[q];
      }
      const double *tzk = T4(Gm.TZ, b, k, n), *tzp = T4(Gm.TZ, b, kp1, n);
      double* fx = FX + (size_t)(n - 1) * n2;
      for (int j = 1; j <= nyb; j++)
        for (int i = 1; i <= nxb - 1; i++) {
          const size_t q = IX2(i, j), e = IX2(i + 1, j);
          fx[q] = fx[q] - CX[q] * (WORK1[q] * tzk[q] + WORK2[q] * tzp[q] + WORK3[q] * tzk[e] + WORK4[q] * tzp[e]);
        }
    }
    {
      const double *snt = S4(Gm.SLY, b, JNORTH, KTP, k), *snb = S4(Gm.SLY, b, JNORTH, KBT, k);
      const double *sst = S4(Gm.SLY, b, JSOUTH, KTP, k), *ssb = S4(Gm.SLY, b, JSOUTH, KBT, k);
      const double *fnt = S4(Gm.SF_SLY, b, JNORTH, KTP, k), *fnb = S4(Gm.SF_SLY, b, JNORTH, KBT, k);
      const double *fst = S4(Gm.SF_SLY, b, JSOUTH, KTP, k), *fsb = S4(Gm.SF_SLY, b, JSOUTH, KBT, k);
      for (int j = 1; j <= nyb - 1; j++)
        for (int i = 1; i <= nxb; i++) {
          const size_t q = IX2(i, j), e = IX2(i, j + 1);
          WORK1[q] = KIt[q] * snt[q] * M.dz[k] - fnt[q];
          WORK2[q] = KIb[q] * snb[q] * M.dz[k] - fnb[q];
          WORK3[q] = KIt[e] * sst[e] * M.dz[k] - fst[e];
          WORK4[q] = KIb[e] * ssb[e] * M.dz[k] - fsb[e];
        }
    }
    for (int n = 1; n <= nt; n++) {
      const double *tzk = T4(Gm.TZ, b, k, n), *tzp = T4(Gm.TZ, b, kp1, n);
      double* fy = FY + (size_t)(n - 1) * n2;
      for (int j = 1; j <= nyb - 1; j++)
        for (int i = 1; i <= nxb; i++) {
          const size_t q = IX2(i, j), e = IX2(i, j + 1);
          fy[q] = fy[q] - CY[q] * (WORK1[q] * tzk[q] + WORK2[q] * tzp[q] + WORK3[q] * tzk[e] + WORK4[q] * tzp[e]);
        }
    }
  }
  /* vertical fluxes and the flux divergence :1899-2076 */
  const double* KIb = S2(Gm.KAPPA_ISOP, b, KBT, k);
  const double* KIt1 = S2(Gm.KAPPA_ISOP, b, KTP, kp1);
  const double *xe = S4(Gm.SLX, b, IEAST, KBT, k), *xw = S4(Gm.SLX, b, IWEST, KBT, k);
  const double *yn = S4(Gm.SLY, b, JNORTH, KBT, k), *ys = S4(Gm.SLY, b, JSOUTH, KBT, k);
  const double *xe1 = S4(Gm.SLX, b, IEAST, KTP, kp1), *xw1 = S4(Gm.SLX, b, IWEST, KTP, kp1);
  const double *yn1 = S4(Gm.SLY, b, JNORTH, KTP, kp1), *ys1 = S4(Gm.SLY, b, JSOUTH, KTP, kp1);
  const double *fxe = S4(Gm.SF_SLX, b, IEAST, KBT, k), *fxw = S4(Gm.SF_SLX, b, IWEST, KBT, k);
  const double *fyn = S4(Gm.SF_SLY, b, JNORTH, KBT, k), *fys = S4(Gm.SF_SLY, b, JSOUTH, KBT, k);
  const double *fxe1 = S4(Gm.SF_SLX, b, IEAST, KTP, kp1), *fxw1 = S4(Gm.SF_SLX, b, IWEST, KTP, kp1);
  const double *fyn1 = S4(Gm.SF_SLY, b, JNORTH, KTP, kp1), *fys1 = S4(Gm.SF_SLY, b, JSOUTH, KTP, kp1);
  const int ib = M.ib[b], ie = M.ie[b], jb = M.jb[b], je = M.je[b];
  for (int n = 1; n <= nt; n++) {
    double* G = GTK + (size_t)(n - 1) * n2;
    const double *fx = FX + (size_t)(n - 1) * n2, *fy = FY + (size_t)(n - 1) * n2;
    double* FZTOP = Gm.FZTOP + ((size_t)b * nt + (n - 1)) * n2;
    const double *tx = T4(Gm.TX, b, k, n), *ty = T4(Gm.TY, b, k, n);
    const double *tx1 = T4(Gm.TX, b, kp1, n), *ty1 = T4(Gm.TY, b, kp1, n);
    memset(G, 0, sizeof(double) * n2);
    for (int j = jb; j <= je; j++)
      for (int i = ib; i <= ie; i++) {
        const size_t q = IX2(i, j), w = IX2(i - 1, j), s = IX2(i, j - 1);
        if (k < km) {
          double w3, fz;
          if (!Gm.cancellation_occurs) {
            w3 = 0.0;
            w3 = w3 + (M.dz[k] * KIb[q] *
                       (xe[q] * HYX[q] * tx[q] + yn[q] * HXY[q] * ty[q] + xw[q] * HYX[w] * tx[w] + ys[q] * HXY[s] * ty[s]));
            w3 = w3 + (fxe[q] * HYX[q] * tx[q] + fyn[q] * HXY[q] * ty[q] + fxw[q] * HYX[w] * tx[w] +
                       fys[q] * HXY[s] * ty[s]);
            w3 = w3 + (dz_bottom * KIt1[q] *
                       (xe1[q] * HYX[q] * tx1[q] + yn1[q] * HXY[q] * ty1[q] + xw1[q] * HYX[w] * tx1[w] +
                        ys1[q] * HXY[s] * ty1[s]));
            w3 = w3 + (factor * (fxe1[q] * HYX[q] * tx1[q] + fyn1[q] * HXY[q] * ty1[q] + fxw1[q] * HYX[w] * tx1[w] +
                                 fys1[q] * HXY[s] * ty1[s]));
            fz = -KMASK[q] * 0.25 * w3;
          } else {
            w3 = (M.dz[k] * KIb[q] *
                  (xe[q] * HYX[q] * tx[q] + yn[q] * HXY[q] * ty[q] + xw[q] * HYX[w] * tx[w] + ys[q] * HXY[s] * ty[s]));
            w3 = w3 + (dz_bottom * KIt1[q] *
                       (xe1[q] * HYX[q] * tx1[q] + yn1[q] * HXY[q] * ty1[q] + xw1[q] * HYX[w] * tx1[w] +
                        ys1[q] * HXY[s] * ty1[s]));
            fz = -KMASK[q] * 0.5 * w3;
          }
          G[q] = (fx[q] - fx[w] + fy[q] - fy[s] + FZTOP[q] - fz) * M.dzr[k] * TAREA_R[q];
          FZTOP[q] = fz;
        } else {
          G[q] = (fx[q] - fx[w] + fy[q] - fy[s] + FZTOP[q]) * M.dzr[k] * TAREA_R[q];
          FZTOP[q] = 0.0;
        }
      }
  }
#undef SQ
  free(CX); free(CY); free(KMASK); free(WORK1); free(WORK2); free(WORK3); free(WORK4); free(FX); free(FY);
}

/* harness access to the GM work arrays (tests only) */
void* o_gm_field(const char* name) {
  if (!Gm.ready) return NULL;
#define F(n, p) if (!strcmp(name, n)) return (void*)(p)
  F("SLX", Gm.SLX); F("SLY", Gm.SLY); F("KAPPA_ISOP", Gm.KAPPA_ISOP); F("KAPPA_THIC", Gm.KAPPA_THIC);
  F("HOR_DIFF", Gm.HOR_DIFF); F("SF_SLX", Gm.SF_SLX); F("SF_SLY", Gm.SF_SLY); F("RBR", Gm.RBR);
  F("HYX", Gm.HYX); F("HXY", Gm.HXY); F("TX", Gm.TX); F("TY", Gm.TY); F("TZ", Gm.TZ);
#undef F
  return NULL;
}
int o_gm_cancellation(void) { return Gm.ready ? Gm.cancellation_occurs : -1; }
