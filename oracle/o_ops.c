/*
 * o_ops.c -- ORACLE (test infrastructure): the slab operators of the hot path, one (block, level)
 * at a time with the reference's loop bounds and operation order.  Full-cell branch only
 * (partial_bottom_cells = .false.).  References cited per function.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pop_oracle.h"

#define NXB (M.nxb)
#define NYB (M.nyb)
#define ALL2 for (int j = 1; j <= NYB; j++) for (int i = 1; i <= NXB; i++)
#define PHYS2(b) for (int j = M.jb[b]; j <= M.je[b]; j++) for (int i = M.ib[b]; i <= M.ie[b]; i++)
static double* tmp2(void) { return (double*)calloc(M.n2, sizeof(double)); }
static double* tmp2n(int n) { return (double*)calloc(M.n2 * n, sizeof(double)); }

/* ------------------------------------------------------------------ state_mod.F90:258-683 */
/* MWJF coefficients: McDougall, Wright, Jackett & Feistel (2003), as tabulated at
   state_mod.F90:136-165 */
static const double p001 = 0.001;
#define mwjfnp0s0t0 (9.99843699e+2 * p001)
#define mwjfnp0s0t1 (7.35212840e+0 * p001)
#define mwjfnp0s0t2 (-5.45928211e-2 * p001)
#define mwjfnp0s0t3 (3.98476704e-4 * p001)
#define mwjfnp0s1t0 (2.96938239e+0 * p001)
#define mwjfnp0s1t1 (-7.23268813e-3 * p001)
#define mwjfnp0s2t0 (2.12382341e-3 * p001)
#define mwjfnp1s0t0 (1.04004591e-2 * p001)
#define mwjfnp1s0t2 (1.03970529e-7 * p001)
#define mwjfnp1s1t0 (5.18761880e-6 * p001)
#define mwjfnp2s0t0 (-3.24041825e-8 * p001)
#define mwjfnp2s0t2 (-1.23869360e-11 * p001)
#define mwjfdp0s0t0 1.0e+0
#define mwjfdp0s0t1 7.28606739e-3
#define mwjfdp0s0t2 (-4.60835542e-5)
#define mwjfdp0s0t3 3.68390573e-7
#define mwjfdp0s0t4 1.80809186e-10
#define mwjfdp0s1t0 2.14691708e-3
#define mwjfdp0s1t1 (-9.27062484e-6)
#define mwjfdp0s1t3 (-1.78343643e-10)
#define mwjfdp0sqt0 4.76534122e-6
#define mwjfdp0sqt2 1.63410736e-9
#define mwjfdp1s0t0 5.30848875e-6
#define mwjfdp2s0t3 (-3.03175128e-16)
#define mwjfdp3s0t1 (-1.27934137e-17)
/* linear EOS: state_mod.F90:177-182 */
#define T_leos_ref 19.0
#define S_leos_ref 0.035
#define rho_leos_ref 1.025022
#define alf 2.55e-4
#define bet 7.64e-1

void o_state(int k, int kk, const double* TEMPK, const double* SALTK, int b, double* RHOOUT,
             double* RHOFULL, double* DRHODT, double* DRHODS) {
  (void)k; (void)b;
  double t0 = o_now();
  if (M.cfg.state_itype == POP_STATE_LINEAR) { /* :664-672 */
    for (size_t q = 0; q < M.n2; q++) {
      if (RHOOUT) RHOOUT[q] = bet * (SALTK[q] - S_leos_ref) - alf * (TEMPK[q] - T_leos_ref);
      if (RHOFULL)
        RHOFULL[q] = rho_leos_ref + bet * (SALTK[q] - S_leos_ref) - alf * (TEMPK[q] - T_leos_ref);
      if (DRHODT) DRHODT[q] = -alf;
      if (DRHODS) DRHODS[q] = bet;
    }
    M.timer[OT_STATE] += o_now() - t0;
    return;
  }
  /* MWJF :420-500 */
  double p = 10.0 * M.pressz[kk];
  double mwjfnums0t0 = mwjfnp0s0t0 + p * (mwjfnp1s0t0 + p * mwjfnp2s0t0);
  double mwjfnums0t1 = mwjfnp0s0t1;
  double mwjfnums0t2 = mwjfnp0s0t2 + p * (mwjfnp1s0t2 + p * mwjfnp2s0t2);
  double mwjfnums0t3 = mwjfnp0s0t3;
  double mwjfnums1t0 = mwjfnp0s1t0 + p * mwjfnp1s1t0;
  double mwjfnums1t1 = mwjfnp0s1t1;
  double mwjfnums2t0 = mwjfnp0s2t0;
  double mwjfdens0t0 = mwjfdp0s0t0 + p * mwjfdp1s0t0;
  double mwjfdens0t1 = mwjfdp0s0t1 + (p * p * p) * mwjfdp3s0t1;
  double mwjfdens0t2 = mwjfdp0s0t2;
  double mwjfdens0t3 = mwjfdp0s0t3 + (p * p) * mwjfdp2s0t3;
  double mwjfdens0t4 = mwjfdp0s0t4;
  double mwjfdens1t0 = mwjfdp0s1t0;
  double mwjfdens1t1 = mwjfdp0s1t1;
  double mwjfdens1t3 = mwjfdp0s1t3;
  double mwjfdensqt0 = mwjfdp0sqt0;
  double mwjfdensqt2 = mwjfdp0sqt2;
  for (size_t q = 0; q < M.n2; q++) {
    double TQ, SQ;
    if (M.cfg.state_range_iopt == POP_STATE_RANGE_ENFORCE) { /* :394-398 */
      TQ = fmin(TEMPK[q], M.tmax[kk]);
      TQ = fmax(TQ, M.tmin[kk]);
      SQ = fmin(SALTK[q], M.smax[kk]);
      SQ = fmax(SQ, M.smin[kk]);
    } else { /* ignore :355-358 -- SQ = max(SALTK,0) discards the preceding min */
      TQ = fmin(TEMPK[q], 1000.0);
      TQ = fmax(TQ, -1000.0);
      SQ = fmin(SALTK[q], 1000.0);
      SQ = fmax(SALTK[q], 0.0);
    }
    SQ = 1000.0 * SQ;
    double SQR = sqrt(SQ);
    double WORK1 = mwjfnums0t0 + TQ * (mwjfnums0t1 + TQ * (mwjfnums0t2 + mwjfnums0t3 * TQ)) +
                   SQ * (mwjfnums1t0 + mwjfnums1t1 * TQ + mwjfnums2t0 * SQ);
    double WORK2 = mwjfdens0t0 +
                   TQ * (mwjfdens0t1 + TQ * (mwjfdens0t2 + TQ * (mwjfdens0t3 + mwjfdens0t4 * TQ))) +
                   SQ * (mwjfdens1t0 + TQ * (mwjfdens1t1 + TQ * TQ * mwjfdens1t3) +
                         SQR * (mwjfdensqt0 + TQ * TQ * mwjfdensqt2));
    double DENOMK = 1.0 / WORK2;
    if (RHOOUT) RHOOUT[q] = WORK1 * DENOMK;
    if (RHOFULL) RHOFULL[q] = WORK1 * DENOMK;
    if (DRHODT) {
      double WORK3 = mwjfnums0t1 + TQ * (2.0 * mwjfnums0t2 + 3.0 * mwjfnums0t3 * TQ) + mwjfnums1t1 * SQ;
      double WORK4 = mwjfdens0t1 + SQ * mwjfdens1t1 +
                     TQ * (2.0 * (mwjfdens0t2 + SQ * SQR * mwjfdensqt2) +
                           TQ * (3.0 * (mwjfdens0t3 + SQ * mwjfdens1t3) + TQ * 4.0 * mwjfdens0t4));
      DRHODT[q] = (WORK3 - WORK1 * DENOMK * WORK4) * DENOMK;
    }
    if (DRHODS) {
      double WORK3 = mwjfnums1t0 + mwjfnums1t1 * TQ + 2.0 * mwjfnums2t0 * SQ;
      double WORK4 = mwjfdens1t0 + TQ * (mwjfdens1t1 + TQ * TQ * mwjfdens1t3) +
                     1.5 * SQR * (mwjfdensqt0 + TQ * TQ * mwjfdensqt2);
      DRHODS[q] = (WORK3 - WORK1 * DENOMK * WORK4) * DENOMK * 1000.0;
    }
  }
  M.timer[OT_STATE] += o_now() - t0;
}

/* ------------------------------------------------------------ grid.F90:3297-3420 */
void o_tgrid_to_ugrid(double* AU, const double* AT, int b) {
  const double *AU0 = B2(M.AU0, b), *AUN = B2(M.AUN, b), *AUE = B2(M.AUE, b), *AUNE = B2(M.AUNE, b);
  for (int j = 1; j <= NYB - 1; j++)
    for (int i = 1; i <= NXB - 1; i++)
      AU[IX2(i, j)] = AU0[IX2(i, j)] * AT[IX2(i, j)] + AUN[IX2(i, j)] * AT[IX2(i, j + 1)] +
                      AUE[IX2(i, j)] * AT[IX2(i + 1, j)] + AUNE[IX2(i, j)] * AT[IX2(i + 1, j + 1)];
  for (int i = 1; i <= NXB; i++) AU[IX2(i, NYB)] = 0.0;
  for (int j = 1; j <= NYB; j++) AU[IX2(NXB, j)] = 0.0;
}
void o_ugrid_to_tgrid(double* AT, const double* AU, int b) {
  const double *AT0 = B2(M.AT0, b), *ATS = B2(M.ATS, b), *ATW = B2(M.ATW, b), *ATSW = B2(M.ATSW, b);
  for (int j = 2; j <= NYB; j++)
    for (int i = 2; i <= NXB; i++)
      AT[IX2(i, j)] = AT0[IX2(i, j)] * AU[IX2(i, j)] + ATS[IX2(i, j)] * AU[IX2(i, j - 1)] +
                      ATW[IX2(i, j)] * AU[IX2(i - 1, j)] + ATSW[IX2(i, j)] * AU[IX2(i - 1, j - 1)];
  for (int i = 1; i <= NXB; i++) AT[IX2(i, 1)] = 0.0;
  for (int j = 1; j <= NYB; j++) AT[IX2(1, j)] = 0.0;
}

/* ------------------------------------------------------------ operators.F90:49-192 */
void o_grad(int k, double* GX, double* GY, const double* F, int b) {
  const int* KMU = M.KMU + (size_t)b * M.n2;
  const double *DXUR = B2(M.DXUR, b), *DYUR = B2(M.DYUR, b);
  memset(GX, 0, sizeof(double) * M.n2);
  memset(GY, 0, sizeof(double) * M.n2);
  for (int j = 1; j <= NYB - 1; j++)
    for (int i = 1; i <= NXB - 1; i++)
      if (k <= KMU[IX2(i, j)]) {
        GX[IX2(i, j)] = DXUR[IX2(i, j)] * 0.5 *
                        (F[IX2(i + 1, j + 1)] - F[IX2(i, j)] - F[IX2(i, j + 1)] + F[IX2(i + 1, j)]);
        GY[IX2(i, j)] = DYUR[IX2(i, j)] * 0.5 *
                        (F[IX2(i + 1, j + 1)] - F[IX2(i, j)] + F[IX2(i, j + 1)] - F[IX2(i + 1, j)]);
      }
}
void o_div(int k, double* D, const double* UX, const double* UY, int b) {
  const int* KMT = M.KMT + (size_t)b * M.n2;
  const double *DYU = B2(M.DYU, b), *DXU = B2(M.DXU, b);
  memset(D, 0, sizeof(double) * M.n2);
  for (int j = 2; j <= NYB; j++)
    for (int i = 2; i <= NXB; i++)
      if (k <= KMT[IX2(i, j)])
        D[IX2(i, j)] =
            0.5 * (UX[IX2(i, j)] * DYU[IX2(i, j)] + UX[IX2(i, j - 1)] * DYU[IX2(i, j - 1)] -
                   UX[IX2(i - 1, j)] * DYU[IX2(i - 1, j)] - UX[IX2(i - 1, j - 1)] * DYU[IX2(i - 1, j - 1)] +
                   UY[IX2(i, j)] * DXU[IX2(i, j)] + UY[IX2(i - 1, j)] * DXU[IX2(i - 1, j)] -
                   UY[IX2(i, j - 1)] * DXU[IX2(i, j - 1)] - UY[IX2(i - 1, j - 1)] * DXU[IX2(i - 1, j - 1)]);
}

/* ------------------------------------------------------------ pressure_grad.F90:187-306 */
void o_gradp(int k, double* PKX, double* PKY, const double* RO, const double* RC,
             const double* RN, int b) {
  double *RHOAVG = tmp2(), *RHOKX = tmp2(), *RHOKY = tmp2();
  double *SUMX = B2(M.SUMX, b), *SUMY = B2(M.SUMY, b), *RMX = B2(M.RHOKMX, b),
         *RMY = B2(M.RHOKMY, b);
  if (M.cfg.lpressure_avg && M.leapfrogts) {
    for (size_t q = 0; q < M.n2; q++) RHOAVG[q] = 0.25 * (RN[q] + 2.0 * RC[q] + RO[q]) * M.bouss[k];
  } else {
    for (size_t q = 0; q < M.n2; q++) RHOAVG[q] = RC[q] * M.bouss[k];
  }
  o_grad(k, RHOKX, RHOKY, RHOAVG, b);
  if (k == 1)
    for (size_t q = 0; q < M.n2; q++) {
      RMX[q] = RHOKX[q]; RMY[q] = RHOKY[q]; SUMX[q] = 0.0; SUMY[q] = 0.0;
    }
  double factor = M.dzw[k - 1] * O_GRAV * 0.5;
  for (size_t q = 0; q < M.n2; q++) {
    SUMX[q] = SUMX[q] + factor * (RHOKX[q] + RMX[q]);
    SUMY[q] = SUMY[q] + factor * (RHOKY[q] + RMY[q]);
    PKX[q] = SUMX[q];
    PKY[q] = SUMY[q];
    RMX[q] = RHOKX[q];
    RMY[q] = RHOKY[q];
  }
  free(RHOAVG); free(RHOKX); free(RHOKY);
}

/* ------------------------------------------------------------ advection.F90:1970-2132 */
static void zero_ghost_cells(int b, double* a) {
  ALL2 if (i < M.ib[b] || i > M.ie[b] || j < M.jb[b] || j > M.je[b]) a[IX2(i, j)] = 0.0;
}
void o_comp_flux_vel(int k, const double* UUU, const double* VVV, const double* WTK, double* UTE,
                     double* UTW, double* VTN, double* VTS, double* WTKB, int b) {
  zero_ghost_cells(b, UTE); zero_ghost_cells(b, UTW); zero_ghost_cells(b, VTN);
  zero_ghost_cells(b, VTS);
  if (k > M.km) {
    memset(WTKB, 0, sizeof(double) * M.n2);
    return;
  }
  const double *U = K3(UUU, k), *V = K3(VVV, k), *DYU = B2(M.DYU, b), *DXU = B2(M.DXU, b);
  const int jend = M.use_lw_lim ? M.je[b] + 2 : M.je[b] + 1; /* :2036 */
  if (PBC) { /* :2040-2066: the flux velocities carry the thickness of the U cells */
    const double* DZU = DZU3(b, k);
#define UD_(i, j) (U[IX2(i, j)] * DYU[IX2(i, j)] * DZU[IX2(i, j)])
#define VD_(i, j) (V[IX2(i, j)] * DXU[IX2(i, j)] * DZU[IX2(i, j)])
    for (int j = M.jb[b] - 1; j <= jend; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
        UTE[IX2(i, j)] = 0.5 * (UD_(i, j) + UD_(i, j - 1));
        UTW[IX2(i, j)] = 0.5 * (UD_(i - 1, j) + UD_(i - 1, j - 1));
        VTN[IX2(i, j)] = 0.5 * (VD_(i, j) + VD_(i - 1, j));
        VTS[IX2(i, j)] = 0.5 * (VD_(i, j - 1) + VD_(i - 1, j - 1));
      }
#undef UD_
#undef VD_
  } else
  for (int j = M.jb[b] - 1; j <= jend; j++)
    for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
      UTE[IX2(i, j)] = 0.5 * (U[IX2(i, j)] * DYU[IX2(i, j)] + U[IX2(i, j - 1)] * DYU[IX2(i, j - 1)]);
      UTW[IX2(i, j)] = 0.5 * (U[IX2(i - 1, j)] * DYU[IX2(i - 1, j)] + U[IX2(i - 1, j - 1)] * DYU[IX2(i - 1, j - 1)]);
      VTN[IX2(i, j)] = 0.5 * (V[IX2(i, j)] * DXU[IX2(i, j)] + V[IX2(i - 1, j)] * DXU[IX2(i - 1, j)]);
      VTS[IX2(i, j)] = 0.5 * (V[IX2(i, j - 1)] * DXU[IX2(i, j - 1)] + V[IX2(i - 1, j - 1)] * DXU[IX2(i - 1, j - 1)]);
    }
  if (M.use_lw_lim) { /* :2086-2092: the outermost ghost cells come from comp_flux_vel_ghost */
    const int jb2 = M.jb[b] - 2, ib2 = M.ib[b] - 2;
    const double* ute_s = M.UTE_jbm2 + ((size_t)b * M.km + (k - 1)) * NXB;
    for (int i = 1; i <= NXB; i++) UTE[IX2(i, jb2)] = ute_s[i - 1];
    for (int i = 2; i <= NXB; i++) UTW[IX2(i, jb2)] = UTE[IX2(i - 1, jb2)];
    for (int j = 1; j <= NYB; j++) UTE[IX2(ib2, j)] = UTW[IX2(ib2 + 1, j)];
    for (int i = 1; i <= NXB; i++) VTN[IX2(i, jb2)] = VTS[IX2(i, jb2 + 1)];
  }
  if (k < M.km) {
    const int* KMT = M.KMT + (size_t)b * M.n2;
    const double* TR = B2(M.TAREA_R, b);
    for (size_t q = 0; q < M.n2; q++) {
      double FC = (VTN[q] - VTS[q] + UTE[q] - UTW[q]) * TR[q];
      if (PBC) WTKB[q] = (k < KMT[q]) ? WTK[q] + FC : 0.0; /* :2110-2111 */
      else WTKB[q] = (k < KMT[q]) ? WTK[q] + M.dz[k] * FC : 0.0;
    }
    if (M.use_lw_lim) { /* :2116-2121 */
      const size_t sx = ((size_t)b * M.km + (k - 1)) * NXB, sy = ((size_t)b * M.km + (k - 1)) * NYB;
      for (int i = 1; i <= NXB; i++) WTKB[IX2(i, M.jb[b] - 2)] = M.WTKB_jbm2[sx + i - 1];
      for (int i = 1; i <= NXB; i++) WTKB[IX2(i, M.je[b] + 2)] = M.WTKB_jep2[sx + i - 1];
      for (int j = 1; j <= NYB; j++) WTKB[IX2(M.ib[b] - 2, j)] = M.WTKB_ibm2[sy + j - 1];
      for (int j = 1; j <= NYB; j++) WTKB[IX2(M.ie[b] + 2, j)] = M.WTKB_iep2[sy + j - 1];
    }
  } else {
    memset(WTKB, 0, sizeof(double) * M.n2);
  }
}

/* advt_centered: advection.F90:2139-2306 */
static void advt_centered(int k, double* LTK, const double* TRCR, const double* WTK,
                          const double* WTKB, const double* UTE, const double* VTN, int b) {
  const double* TR = B2(M.TAREA_R, b);
  for (int n = 1; n <= M.nt; n++) {
    if (M.cfg.tadvect_itype[n - 1] != POP_TADVECT_CENTERED) continue;
    const double* T = KN4(TRCR, k, n);
    double* L = LTK + (size_t)(n - 1) * M.n2;
    PHYS2(b)
    L[IX2(i, j)] = 0.5 *
                   ((VTN[IX2(i, j)] - VTN[IX2(i, j - 1)] + UTE[IX2(i, j)] - UTE[IX2(i - 1, j)]) * T[IX2(i, j)] +
                    VTN[IX2(i, j)] * T[IX2(i, j + 1)] - VTN[IX2(i, j - 1)] * T[IX2(i, j - 1)] +
                    UTE[IX2(i, j)] * T[IX2(i + 1, j)] - UTE[IX2(i - 1, j)] * T[IX2(i - 1, j)]) *
                   TR[IX2(i, j)];
    const double* DZT = PBC ? DZT3(b, k) : NULL;
    if (PBC) PHYS2(b) L[IX2(i, j)] = L[IX2(i, j)] / DZT[IX2(i, j)]; /* :2223-2238 */
    if (k == 1) {
      if (M.cfg.sfc_layer_type != POP_SFC_VARTHICK)
        for (size_t q = 0; q < M.n2; q++) L[q] = L[q] + M.dzr[k] * WTK[q] * T[q];
    } else {
      const double* Tm = KN4(TRCR, k - 1, n);
      if (PBC) for (size_t q = 0; q < M.n2; q++) L[q] = L[q] + 0.5 / DZT[q] * WTK[q] * (Tm[q] + T[q]); /* :2278-2280 */
      else for (size_t q = 0; q < M.n2; q++) L[q] = L[q] + M.dz2r[k] * WTK[q] * (Tm[q] + T[q]);
    }
    if (k < M.km) {
      const double* Tp = KN4(TRCR, k + 1, n);
      if (PBC) for (size_t q = 0; q < M.n2; q++) L[q] = L[q] - 0.5 / DZT[q] * WTKB[q] * (T[q] + Tp[q]);
      else for (size_t q = 0; q < M.n2; q++) L[q] = L[q] - M.dz2r[k] * WTKB[q] * (T[q] + Tp[q]);
    }
  }
}

/* hupw3: advection.F90:2488-2677 */
static void hupw3(int k, double* XOUT, const double* X, const double* CN, const double* CS,
                  const double* CE, const double* CW, double* TRACER_E, double* TRACER_N, int b) {
  const int *KMTE = M.KMTE + (size_t)b * M.n2, *KMTW = M.KMTW + (size_t)b * M.n2,
            *KMTEE = M.KMTEE + (size_t)b * M.n2, *KMTN = M.KMTN + (size_t)b * M.n2,
            *KMTS = M.KMTS + (size_t)b * M.n2, *KMTNN = M.KMTNN + (size_t)b * M.n2;
  for (int n = 1; n <= M.nt; n++) {
    if (M.cfg.tadvect_itype[n - 1] != POP_TADVECT_UPWIND3) continue;
    const double* Xk = KN4(X, k, n);
    double *TE = TRACER_E + (size_t)(n - 1) * M.n2, *TN = TRACER_N + (size_t)(n - 1) * M.n2,
           *XO = XOUT + (size_t)(n - 1) * M.n2;
    zero_ghost_cells(b, TE);
    zero_ghost_cells(b, TN);
    for (int j = M.jb[b]; j <= M.je[b]; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b]; i++) {
        size_t q = IX2(i, j);
        double work, ap, bp, gp, am, bm, dm;
        if (k <= KMTE[q]) { work = B2(M.TBETXP, b)[q]; ap = B2(M.TALFXP, b)[q]; }
        else { work = B2(M.TBETXP, b)[q] + B2(M.TALFXP, b)[q]; ap = 0.0; }
        if (k <= KMTW[q]) { bp = work; gp = B2(M.TGAMXP, b)[q]; }
        else { bp = work + B2(M.TGAMXP, b)[q]; gp = 0.0; }
        if (k <= KMTEE[q]) { am = B2(M.TALFXM, b)[q]; dm = B2(M.TDELXM, b)[q]; }
        else { am = B2(M.TALFXM, b)[q] + B2(M.TDELXM, b)[q]; dm = 0.0; }
        bm = B2(M.TBETXM, b)[q];
        if (CE[q] > 0.0)
          TE[q] = ap * Xk[IX2(i + 1, j)] + bp * Xk[IX2(i, j)] + gp * Xk[IX2(i - 1, j)];
        else
          TE[q] = am * Xk[IX2(i + 1, j)] + bm * Xk[IX2(i, j)] + dm * Xk[IX2(i + 2, j)];
      }
    PHYS2(b) XO[IX2(i, j)] = CE[IX2(i, j)] * TE[IX2(i, j)] + CW[IX2(i, j)] * TE[IX2(i - 1, j)];
    for (int j = M.jb[b] - 1; j <= M.je[b]; j++)
      for (int i = M.ib[b]; i <= M.ie[b]; i++) {
        size_t q = IX2(i, j);
        double work, ap, bp, gp, am, bm, dm;
        if (k <= KMTN[q]) { work = B2(M.TBETYP, b)[q]; ap = B2(M.TALFYP, b)[q]; }
        else { work = B2(M.TBETYP, b)[q] + B2(M.TALFYP, b)[q]; ap = 0.0; }
        if (k <= KMTS[q]) { bp = work; gp = B2(M.TGAMYP, b)[q]; }
        else { bp = work + B2(M.TGAMYP, b)[q]; gp = 0.0; }
        if (k <= KMTNN[q]) { am = B2(M.TALFYM, b)[q]; dm = B2(M.TDELYM, b)[q]; }
        else { am = B2(M.TALFYM, b)[q] + B2(M.TDELYM, b)[q]; dm = 0.0; }
        bm = B2(M.TBETYM, b)[q];
        if (CN[q] > 0.0)
          TN[q] = ap * Xk[IX2(i, j + 1)] + bp * Xk[IX2(i, j)] + gp * Xk[IX2(i, j - 1)];
        else
          TN[q] = am * Xk[IX2(i, j + 1)] + bm * Xk[IX2(i, j)] + dm * Xk[IX2(i, j + 2)];
      }
    for (int j = M.jb[b] - 1; j <= M.je[b]; j++)
      for (int i = M.ib[b]; i <= M.ie[b]; i++)
        XO[IX2(i, j)] = XO[IX2(i, j)] + CN[IX2(i, j)] * TN[IX2(i, j)] + CS[IX2(i, j)] * TN[IX2(i, j - 1)];
  }
}

/* advt_upwind3: advection.F90:2313-2481 */
static void advt_upwind3(int k, double* LTK, const double* TRCR, const double* WTK,
                         const double* WTKB, const double* UTE, const double* UTW,
                         const double* VTN, const double* VTS, int b) {
  const int km = M.km;
  const int* KMT = M.KMT + (size_t)b * M.n2;
  const double* TR = B2(M.TAREA_R, b);
  double *FVN = tmp2(), *FVS = tmp2(), *FUE = tmp2(), *FUW = tmp2(), *AZM = tmp2(), *DZM = tmp2();
  double *TE = tmp2n(M.nt), *TN = tmp2n(M.nt), *AUXB = tmp2();
  for (size_t q = 0; q < M.n2; q++) {
    FVN[q] = VTN[q] * TR[q];
    FVS[q] = -VTS[q] * TR[q];
    FUE[q] = UTE[q] * TR[q];
    FUW[q] = -UTW[q] * TR[q];
  }
  hupw3(k, LTK, TRCR, FVN, FVS, FUE, FUW, TE, TN, b);
  for (size_t q = 0; q < M.n2; q++) {
    if (k < KMT[q] - 1) { AZM[q] = M.talfzm[k]; DZM[q] = M.tdelzm[k]; }
    else { AZM[q] = M.talfzm[k] + M.tdelzm[k]; DZM[q] = 0.0; }
  }
  for (int n = 1; n <= M.nt; n++) {
    if (M.cfg.tadvect_itype[n - 1] != POP_TADVECT_UPWIND3) continue;
    double* L = LTK + (size_t)(n - 1) * M.n2;
    double* AUX = M.AUX + ((size_t)b * M.nt + (n - 1)) * M.n2;
    const double* T = KN4(TRCR, k, n);
    const double* Tm = (k > 1) ? KN4(TRCR, k - 1, n) : NULL;
    const double* Tp = (k < km) ? KN4(TRCR, k + 1, n) : NULL;
    const double* Tpp = (k + 2 <= km) ? KN4(TRCR, k + 2, n) : NULL;
    if (k < km - 1 && k > 1) {
      for (size_t q = 0; q < M.n2; q++) {
        double TPLUS = M.talfzp[k] * Tp[q] + M.tbetzp[k] * T[q] + M.tgamzp[k] * Tm[q];
        double TMINUS = AZM[q] * Tp[q] + M.tbetzm[k] * T[q] + DZM[q] * Tpp[q];
        AUXB[q] = (WTKB[q] - fabs(WTKB[q])) * TPLUS + (WTKB[q] + fabs(WTKB[q])) * TMINUS;
      }
    } else if (k == 1) {
      for (size_t q = 0; q < M.n2; q++) {
        double TPLUS = M.talfzp[k] * Tp[q] + M.tbetzp[k] * T[q];
        double TMINUS = AZM[q] * Tp[q] + M.tbetzm[k] * T[q] + DZM[q] * Tpp[q];
        AUXB[q] = (WTKB[q] - fabs(WTKB[q])) * TPLUS + (WTKB[q] + fabs(WTKB[q])) * TMINUS;
      }
    } else if (k == km - 1) {
      for (size_t q = 0; q < M.n2; q++) {
        double TPLUS = M.talfzp[k] * Tp[q] + M.tbetzp[k] * T[q] + M.tgamzp[k] * Tm[q];
        double TMINUS = AZM[q] * Tp[q] + M.tbetzm[k] * T[q];
        AUXB[q] = (WTKB[q] - fabs(WTKB[q])) * TPLUS + (WTKB[q] + fabs(WTKB[q])) * TMINUS;
      }
    } else {
      for (size_t q = 0; q < M.n2; q++) AUXB[q] = 0.0;
    }
    if (k == 1) {
      if (M.cfg.sfc_layer_type != POP_SFC_VARTHICK) {
        for (size_t q = 0; q < M.n2; q++) {
          double FLUX_T = M.dzr[k] * WTK[q] * T[q];
          L[q] = L[q] + FLUX_T - M.dz2r[k] * AUXB[q];
        }
      } else {
        for (size_t q = 0; q < M.n2; q++) L[q] = L[q] - M.dz2r[k] * AUXB[q];
      }
    } else {
      for (size_t q = 0; q < M.n2; q++) L[q] = L[q] + M.dz2r[k] * (AUX[q] - AUXB[q]);
    }
    for (size_t q = 0; q < M.n2; q++) AUX[q] = AUXB[q];
  }
  free(FVN); free(FVS); free(FUE); free(FUW); free(AZM); free(DZM); free(TE); free(TN); free(AUXB);
}

/* comp_flux_vel_ghost: advection.F90:1014-1120.  Flux velocities of every level with halo-updated ghost cells; the
   rows / columns two cells outside the physical domain are kept for comp_flux_vel. */
void o_comp_flux_vel_ghost(void) {
  if (!M.use_lw_lim) return;
  const int nb = M.nblocks, km = M.km, c = M.curtime;
  const size_t nall = M.n2 * nb;
  double *UTE = (double*)calloc(nall, sizeof(double)), *UTW = (double*)calloc(nall, sizeof(double)),
         *VTN = (double*)calloc(nall, sizeof(double)), *VTS = (double*)calloc(nall, sizeof(double)),
         *WTK = (double*)calloc(nall, sizeof(double)), *WTKB = (double*)calloc(nall, sizeof(double));
  for (int k = 1; k <= km; k++) {
    for (int b = 0; b < nb; b++) {
      if (!M.active[b]) continue;
      if (k == 1) memcpy(B2(WTK, b), B2(M.DH, b), sizeof(double) * M.n2);
      else memcpy(B2(WTK, b), B2(WTKB, b), sizeof(double) * M.n2);
      o_comp_flux_vel(k, B3(M.UVEL[c], b), B3(M.VVEL[c], b), B2(WTK, b), B2(UTE, b), B2(UTW, b), B2(VTN, b),
                      B2(VTS, b), B2(WTKB, b), b);
    }
    oracle_halo_2d(UTE, POP_LOC_EFACE, POP_KIND_VECTOR, 0.0);
    oracle_halo_2d(WTKB, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0);
    for (int b = 0; b < nb; b++) {
      if (!M.active[b]) continue;
      const size_t sx = ((size_t)b * km + (k - 1)) * NXB, sy = ((size_t)b * km + (k - 1)) * NYB;
      const double *E = B2(UTE, b), *W = B2(WTKB, b);
      for (int i = 1; i <= NXB; i++) {
        M.UTE_jbm2[sx + i - 1] = E[IX2(i, M.jb[b] - 2)];
        M.WTKB_jbm2[sx + i - 1] = W[IX2(i, M.jb[b] - 2)];
        M.WTKB_jep2[sx + i - 1] = W[IX2(i, M.je[b] + 2)];
      }
      for (int j = 1; j <= NYB; j++) {
        M.WTKB_ibm2[sy + j - 1] = W[IX2(M.ib[b] - 2, j)];
        M.WTKB_iep2[sy + j - 1] = W[IX2(M.ie[b] + 2, j)];
      }
    }
  }
  free(UTE); free(UTW); free(VTN); free(VTS); free(WTK); free(WTKB);
}

/* lw_lim: advection.F90:2832-3280.  Second-order forward-in-time scheme with one-dimensional flux limiters:
   z, then x (on the z-updated field), then y (on the z- and x-updated field). */
static void lw_lim(int k, double adv_dt, const double* UVEL_E_dt, const double* VVEL_N_dt, double* XOUT,
                   const double* X, const double* WTK, const double* WTKB, const double* WTKBp1, double* AUXB,
                   const double* CN, const double* CS, const double* CE, const double* CW, const double* KMASKE,
                   const double* KMASKN, double* TRACER_E, double* TRACER_N, int b) {
  const int km = M.km, ib = M.ib[b], ie = M.ie[b], jb = M.jb[b], je = M.je[b];
  const int* KMT = M.KMT + (size_t)b * M.n2;
  const double *DXT = B2(M.DXT, b), *DYT = B2(M.DYT, b), *PX = B2(M.p5_DXT_ph_R, b), *PY = B2(M.p5_DYT_ph_R, b);
  const double adv_dt_r = 1.0 / adv_dt;
  double *DIV = tmp2(), *XSTAR = tmp2(), *MU_z = tmp2(), *LW_z = tmp2(), *MU_x = tmp2(), *LW_x = tmp2(),
         *MU_y = tmp2(), *LW_y = tmp2();
  /* ---- tracer-independent coefficients, :2911-2975 */
  if (PBC) {
    const double *Zk = DZT3(b, k), *Zp = DZT3(b, k + 1), *Zm = DZT3(b, k - 1);
    const double* Zpp = (k + 2 <= km + 1) ? DZT3(b, k + 2) : NULL;
    for (int j = jb - 2; j <= je + 2; j++)
      for (int i = ib - 2; i <= ie + 2; i++) {
        const size_t q = IX2(i, j);
        if (WTKB[q] > 0.0) {
          LW_z[q] = (Zp[q] - adv_dt * WTKB[q]) / (Zk[q] + Zp[q]);
          if (WTKBp1[q] > 0.0) MU_z[q] = (Zp[q] * adv_dt_r - WTKBp1[q]) / WTKB[q];
          else if (WTKBp1[q] < 0.0)
            MU_z[q] = -WTKBp1[q] / WTKB[q] * (Zp[q] + adv_dt * WTKBp1[q]) / (Zp[q] + Zpp[q]);
          else MU_z[q] = 0.0;
        } else if (WTKB[q] < 0.0) {
          LW_z[q] = (Zk[q] + adv_dt * WTKB[q]) / (Zk[q] + Zp[q]);
          if (WTK[q] < 0.0) MU_z[q] = -(Zk[q] * adv_dt_r + WTK[q]) / WTKB[q];
          else if (WTK[q] > 0.0) MU_z[q] = -WTK[q] / WTKB[q] * (Zk[q] - adv_dt * WTK[q]) / (Zm[q] + Zk[q]);
          else MU_z[q] = 0.0;
        }
      }
  } else {
    const double dzk_div_dt = M.dz[k] * adv_dt_r;
    const double dzkp1_div_dt = (k < km) ? M.dz[k + 1] * adv_dt_r : 0.0;
    const double work1 = M.dz[k] * M.p5_dz_ph_r[k];
    const double work2 = (k < km) ? M.dz[k + 1] * M.p5_dz_ph_r[k] : 0.0;
    const double work3 = adv_dt * M.p5_dz_ph_r[k];
    for (int j = jb - 2; j <= je + 2; j++)
      for (int i = ib - 2; i <= ie + 2; i++) {
        const size_t q = IX2(i, j);
        if (WTKB[q] > 0.0) {
          LW_z[q] = work2 - work3 * WTKB[q];
          if (WTKBp1[q] > 0.0) MU_z[q] = (dzkp1_div_dt - WTKBp1[q]) / WTKB[q];
          else if (WTKBp1[q] < 0.0)
            MU_z[q] = -WTKBp1[q] / WTKB[q] * (M.dz[k + 1] + adv_dt * WTKBp1[q]) * M.p5_dz_ph_r[k + 1];
          else MU_z[q] = 0.0;
        } else if (WTKB[q] < 0.0) {
          LW_z[q] = work1 + work3 * WTKB[q];
          if (WTK[q] < 0.0) MU_z[q] = -(dzk_div_dt + WTK[q]) / WTKB[q];
          else if (WTK[q] > 0.0) MU_z[q] = -WTK[q] / WTKB[q] * (M.dz[k] - adv_dt * WTK[q]) * M.p5_dz_ph_r[k - 1];
          else MU_z[q] = 0.0;
        }
      }
  }
  /* :2977-3015 */
  for (int j = jb - 2; j <= je + 2; j++) {
    int i = ib - 2;
    if (UVEL_E_dt[IX2(i, j)] < 0.0) LW_x[IX2(i, j)] = (DXT[IX2(i + 1, j)] + UVEL_E_dt[IX2(i, j)]) * PX[IX2(i, j)];
    for (i = ib - 1; i <= ie; i++) {
      const size_t q = IX2(i, j);
      if (UVEL_E_dt[q] > 0.0) {
        LW_x[q] = (DXT[q] - UVEL_E_dt[q]) * PX[q];
        if (UVEL_E_dt[IX2(i - 1, j)] > 0.0) MU_x[q] = (DXT[q] - UVEL_E_dt[IX2(i - 1, j)]) / UVEL_E_dt[q];
        else MU_x[q] = 0.0;
      } else if (UVEL_E_dt[q] < 0.0) {
        LW_x[q] = (DXT[IX2(i + 1, j)] + UVEL_E_dt[q]) * PX[q];
        if (UVEL_E_dt[IX2(i + 1, j)] < 0.0) MU_x[q] = -(DXT[IX2(i + 1, j)] + UVEL_E_dt[IX2(i + 1, j)]) / UVEL_E_dt[q];
        else MU_x[q] = 0.0;
      } else {
        LW_x[q] = DXT[q] * PX[q];
      }
    }
    i = ie + 1;
    if (UVEL_E_dt[IX2(i, j)] > 0.0) LW_x[IX2(i, j)] = (DXT[IX2(i, j)] - UVEL_E_dt[IX2(i, j)]) * PX[IX2(i, j)];
    for (i = ib - 1; i <= ie; i++) {
      const size_t q = IX2(i, j);
      if (UVEL_E_dt[q] > 0.0 && UVEL_E_dt[IX2(i - 1, j)] < 0.0)
        MU_x[q] = -UVEL_E_dt[IX2(i - 1, j)] / UVEL_E_dt[q] * LW_x[IX2(i - 1, j)];
      if (UVEL_E_dt[q] < 0.0 && UVEL_E_dt[IX2(i + 1, j)] > 0.0)
        MU_x[q] = -UVEL_E_dt[IX2(i + 1, j)] / UVEL_E_dt[q] * LW_x[IX2(i + 1, j)];
    }
  }
  /* :3017-3061 */
  for (int i = ib; i <= ie; i++)
    if (VVEL_N_dt[IX2(i, jb - 2)] < 0.0)
      LW_y[IX2(i, jb - 2)] = (DYT[IX2(i, jb - 1)] + VVEL_N_dt[IX2(i, jb - 2)]) * PY[IX2(i, jb - 2)];
  for (int j = jb - 1; j <= je; j++)
    for (int i = ib; i <= ie; i++) {
      const size_t q = IX2(i, j);
      if (VVEL_N_dt[q] > 0.0) {
        LW_y[q] = (DYT[q] - VVEL_N_dt[q]) * PY[q];
        if (VVEL_N_dt[IX2(i, j - 1)] > 0.0) MU_y[q] = (DYT[q] - VVEL_N_dt[IX2(i, j - 1)]) / VVEL_N_dt[q];
        else MU_y[q] = 0.0;
      } else if (VVEL_N_dt[q] < 0.0) {
        LW_y[q] = (DYT[IX2(i, j + 1)] + VVEL_N_dt[q]) * PY[q];
        if (VVEL_N_dt[IX2(i, j + 1)] < 0.0) MU_y[q] = -(DYT[IX2(i, j + 1)] + VVEL_N_dt[IX2(i, j + 1)]) / VVEL_N_dt[q];
        else MU_y[q] = 0.0;
      } else {
        LW_y[q] = DYT[q] * PY[q];
      }
    }
  for (int i = ib; i <= ie; i++)
    if (VVEL_N_dt[IX2(i, je + 1)] > 0.0)
      LW_y[IX2(i, je + 1)] = (DYT[IX2(i, je + 1)] - VVEL_N_dt[IX2(i, je + 1)]) * PY[IX2(i, je + 1)];
  for (int j = jb - 1; j <= je; j++)
    for (int i = ib; i <= ie; i++) {
      const size_t q = IX2(i, j);
      if (VVEL_N_dt[q] > 0.0 && VVEL_N_dt[IX2(i, j - 1)] < 0.0)
        MU_y[q] = -VVEL_N_dt[IX2(i, j - 1)] / VVEL_N_dt[q] * LW_y[IX2(i, j - 1)];
      if (VVEL_N_dt[q] < 0.0 && VVEL_N_dt[IX2(i, j + 1)] > 0.0)
        MU_y[q] = -VVEL_N_dt[IX2(i, j + 1)] / VVEL_N_dt[q] * LW_y[IX2(i, j + 1)];
    }
  /* :3063-3067 */
  for (size_t q = 0; q < M.n2; q++) {
    if (PBC) DIV[q] = (WTK[q] - WTKB[q]) / DZT3(b, k)[q] + CE[q] + CW[q] + CN[q] + CS[q];
    else DIV[q] = (WTK[q] - WTKB[q]) * M.dzr[k] + CE[q] + CW[q] + CN[q] + CS[q];
  }
  /* ---- tracers, :3073-3274 */
  for (int n = 1; n <= M.nt; n++) {
    if (M.cfg.tadvect_itype[n - 1] != POP_TADVECT_LW_LIM) continue;
    double *TE = TRACER_E + (size_t)(n - 1) * M.n2, *TN = TRACER_N + (size_t)(n - 1) * M.n2,
           *XO = XOUT + (size_t)(n - 1) * M.n2, *AB = AUXB + (size_t)(n - 1) * M.n2;
    const double* AUX = M.AUX + ((size_t)b * M.nt + (n - 1)) * M.n2;
    const double* Xk = KN4(X, k, n);
    const double* Xp = (k + 1 <= km) ? KN4(X, k + 1, n) : NULL;
    const double* Xpp = (k + 2 <= km) ? KN4(X, k + 2, n) : NULL;
    const double* Xm = (k > 1) ? KN4(X, k - 1, n) : NULL;
    zero_ghost_cells(b, TE); zero_ghost_cells(b, TN); zero_ghost_cells(b, AB);
    for (int j = jb - 2; j <= je + 2; j++) {
      for (int i = ib - 2; i <= ie + 2; i++) { /* z, :3096-3149 */
        const size_t q = IX2(i, j);
        if (k + 1 <= KMT[q]) {
          const double dTR = Xp[q] - Xk[q];
          if (WTKB[q] > 0.0) {
            AB[q] = WTKB[q] * Xp[q];
            if (k + 2 <= KMT[q]) {
              const double dTRp1 = Xpp[q] - Xp[q];
              if (dTR > 0.0 && dTRp1 > 0.0) AB[q] = WTKB[q] * (Xp[q] - fmin(LW_z[q] * dTR, MU_z[q] * dTRp1));
              else if (dTR < 0.0 && dTRp1 < 0.0) AB[q] = WTKB[q] * (Xp[q] - fmax(LW_z[q] * dTR, MU_z[q] * dTRp1));
            }
          } else if (WTKB[q] < 0.0) {
            AB[q] = WTKB[q] * Xk[q];
            if (k > 1) {
              const double dTRm1 = Xk[q] - Xm[q];
              if (dTR > 0.0 && dTRm1 > 0.0) AB[q] = WTKB[q] * (Xk[q] + fmin(LW_z[q] * dTR, MU_z[q] * dTRm1));
              else if (dTR < 0.0 && dTRm1 < 0.0) AB[q] = WTKB[q] * (Xk[q] + fmax(LW_z[q] * dTR, MU_z[q] * dTRm1));
            }
          } else {
            AB[q] = 0.0;
          }
        } else {
          AB[q] = 0.0;
        }
        if (PBC) XO[q] = (AUX[q] - AB[q] - (WTK[q] - WTKB[q]) * Xk[q]) / DZT3(b, k)[q];
        else XO[q] = (AUX[q] - AB[q] - (WTK[q] - WTKB[q]) * Xk[q]) * M.dzr[k];
        XSTAR[q] = Xk[q] - adv_dt * XO[q];
      }
      for (int i = ib - 1; i <= ie; i++) { /* x faces, :3151-3188 */
        const size_t q = IX2(i, j);
        const double dTR = (XSTAR[IX2(i + 1, j)] - XSTAR[q]) * KMASKE[q];
        if (CE[q] > 0.0) {
          const double dTRm1 = (XSTAR[q] - XSTAR[IX2(i - 1, j)]) * KMASKE[IX2(i - 1, j)];
          if (dTR > 0.0 && dTRm1 > 0.0) TE[q] = XSTAR[q] + fmin(LW_x[q] * dTR, MU_x[q] * dTRm1);
          else if (dTR < 0.0 && dTRm1 < 0.0) TE[q] = XSTAR[q] + fmax(LW_x[q] * dTR, MU_x[q] * dTRm1);
          else TE[q] = XSTAR[q];
        } else if (CE[q] < 0.0) {
          const double dTRp1 = (XSTAR[IX2(i + 2, j)] - XSTAR[IX2(i + 1, j)]) * KMASKE[IX2(i + 1, j)];
          if (dTR > 0.0 && dTRp1 > 0.0) TE[q] = XSTAR[IX2(i + 1, j)] - fmin(LW_x[q] * dTR, MU_x[q] * dTRp1);
          else if (dTR < 0.0 && dTRp1 < 0.0) TE[q] = XSTAR[IX2(i + 1, j)] - fmax(LW_x[q] * dTR, MU_x[q] * dTRp1);
          else TE[q] = XSTAR[IX2(i + 1, j)];
        } else {
          TE[q] = XSTAR[q] + LW_x[q] * dTR;
        }
      }
    }
    for (int j = jb - 2; j <= je + 2; j++) /* x update, :3205-3213 */
      for (int i = ib; i <= ie; i++) {
        const size_t q = IX2(i, j);
        const double work1 = CE[q] * TE[q] + CW[q] * TE[IX2(i - 1, j)] - (CE[q] + CW[q]) * Xk[q];
        XO[q] = XO[q] + work1;
        XSTAR[q] = XSTAR[q] - adv_dt * work1;
      }
    for (int j = jb - 1; j <= je; j++) /* y faces, :3220-3256 */
      for (int i = ib; i <= ie; i++) {
        const size_t q = IX2(i, j);
        const double dTR = (XSTAR[IX2(i, j + 1)] - XSTAR[q]) * KMASKN[q];
        if (CN[q] > 0.0) {
          const double dTRm1 = (XSTAR[q] - XSTAR[IX2(i, j - 1)]) * KMASKN[IX2(i, j - 1)];
          if (dTR > 0.0 && dTRm1 > 0.0) TN[q] = XSTAR[q] + fmin(LW_y[q] * dTR, MU_y[q] * dTRm1);
          else if (dTR < 0.0 && dTRm1 < 0.0) TN[q] = XSTAR[q] + fmax(LW_y[q] * dTR, MU_y[q] * dTRm1);
          else TN[q] = XSTAR[q];
        } else if (CN[q] < 0.0) {
          const double dTRp1 = (XSTAR[IX2(i, j + 2)] - XSTAR[IX2(i, j + 1)]) * KMASKN[IX2(i, j + 1)];
          if (dTR > 0.0 && dTRp1 > 0.0) TN[q] = XSTAR[IX2(i, j + 1)] - fmin(LW_y[q] * dTR, MU_y[q] * dTRp1);
          else if (dTR < 0.0 && dTRp1 < 0.0) TN[q] = XSTAR[IX2(i, j + 1)] - fmax(LW_y[q] * dTR, MU_y[q] * dTRp1);
          else TN[q] = XSTAR[IX2(i, j + 1)];
        } else {
          TN[q] = XSTAR[q] + LW_y[q] * dTR;
        }
      }
    for (int j = jb - 1; j <= je; j++) /* y update + divergence term, :3263-3272 */
      for (int i = ib; i <= ie; i++) {
        const size_t q = IX2(i, j);
        XO[q] = XO[q] + CN[q] * TN[q] + CS[q] * TN[IX2(i, j - 1)] - (CN[q] + CS[q] - DIV[q]) * Xk[q];
      }
  }
  free(DIV); free(XSTAR); free(MU_z); free(LW_z); free(MU_x); free(LW_x); free(MU_y); free(LW_y);
}

/* advt_lw_lim: advection.F90:2684-2825 */
static void advt_lw_lim(int k, double* LTK, const double* TMIX, const double* WTK, const double* WTKB,
                        const double* WTKBp1, const double* UTE, const double* UTW, const double* VTN,
                        const double* VTS, int b) {
  const int kk = PBC ? k : 1;
  const double adv_dt = M.c2dtt[k];
  const int *KMT = M.KMT + (size_t)b * M.n2, *KMTE = M.KMTE + (size_t)b * M.n2, *KMTN = M.KMTN + (size_t)b * M.n2;
  const double* TR = B2(M.TAREA_R, b);
  const double* E2U = M.UTE_to_UVEL_E + ((size_t)b * (PBC ? M.km : 1) + (kk - 1)) * M.n2;
  const double* N2V = M.VTN_to_VVEL_N + ((size_t)b * (PBC ? M.km : 1) + (kk - 1)) * M.n2;
  double *UE = tmp2(), *VN = tmp2(), *FVN = tmp2(), *FVS = tmp2(), *FUE = tmp2(), *FUW = tmp2(), *KE = tmp2(),
         *KN = tmp2(), *WEFF = tmp2();
  double *TE = tmp2n(M.nt), *TN = tmp2n(M.nt), *AUXB = tmp2n(M.nt);
  for (size_t q = 0; q < M.n2; q++) {
    UE[q] = adv_dt * UTE[q] * E2U[q];
    VN[q] = adv_dt * VTN[q] * N2V[q];
    const double w = PBC ? TR[q] / DZT3(b, k)[q] : TR[q];
    FVN[q] = VTN[q] * w;
    FVS[q] = -VTS[q] * w;
    FUE[q] = UTE[q] * w;
    FUW[q] = -UTW[q] * w;
    KE[q] = (k <= KMT[q] && k <= KMTE[q]) ? 1.0 : 0.0;
    KN[q] = (k <= KMT[q] && k <= KMTN[q]) ? 1.0 : 0.0;
    WEFF[q] = (k == 1 && M.cfg.sfc_layer_type == POP_SFC_VARTHICK) ? 0.0 : WTK[q];
  }
  if (k == 1)
    for (int n = 1; n <= M.nt; n++) {
      if (M.cfg.tadvect_itype[n - 1] != POP_TADVECT_LW_LIM) continue;
      double* AUX = M.AUX + ((size_t)b * M.nt + (n - 1)) * M.n2;
      const double* T = KN4(TMIX, 1, n);
      for (size_t q = 0; q < M.n2; q++) AUX[q] = WEFF[q] * T[q];
    }
  lw_lim(k, adv_dt, UE, VN, LTK, TMIX, WEFF, WTKB, WTKBp1, AUXB, FVN, FVS, FUE, FUW, KE, KN, TE, TN, b);
  for (int n = 1; n <= M.nt; n++) {
    if (M.cfg.tadvect_itype[n - 1] != POP_TADVECT_LW_LIM) continue;
    memcpy(M.AUX + ((size_t)b * M.nt + (n - 1)) * M.n2, AUXB + (size_t)(n - 1) * M.n2, sizeof(double) * M.n2);
  }
  free(UE); free(VN); free(FVN); free(FVS); free(FUE); free(FUW); free(KE); free(KN); free(WEFF);
  free(TE); free(TN); free(AUXB);
}

/* advt: advection.F90:1577-1963 */
void o_advt(int k, double* LTK, double* WTK, const double* TMIX, const double* TRCR,
            const double* UUU, const double* VVV, int b) {
  double t0 = o_now();
  double *UTE = tmp2(), *UTW = tmp2(), *VTN = tmp2(), *VTS = tmp2(), *WTKB = tmp2();
  double* FP = M.use_lw_lim ? M.FLUX_VEL_prev + (size_t)b * 5 * M.n2 : NULL;
  if (k == 1 || !M.use_lw_lim) { /* :1667-1676 */
    o_comp_flux_vel(k, UUU, VVV, WTK, UTE, UTW, VTN, VTS, WTKB, b);
  } else {
    memcpy(UTE, FP, sizeof(double) * M.n2); memcpy(UTW, FP + M.n2, sizeof(double) * M.n2);
    memcpy(VTN, FP + 2 * M.n2, sizeof(double) * M.n2); memcpy(VTS, FP + 3 * M.n2, sizeof(double) * M.n2);
    memcpy(WTKB, FP + 4 * M.n2, sizeof(double) * M.n2);
  }
  memset(LTK, 0, sizeof(double) * M.n2 * M.nt);
  if (M.use_lw_lim) { /* :1688-1708: flux velocities of level k+1 are kept for the next call */
    o_comp_flux_vel(k + 1, UUU, VVV, WTKB, FP, FP + M.n2, FP + 2 * M.n2, FP + 3 * M.n2, FP + 4 * M.n2, b);
    advt_lw_lim(k, LTK, TMIX, WTK, WTKB, FP + 4 * M.n2, UTE, UTW, VTN, VTS, b);
  }
  int up3 = 0, cen = 0;
  for (int n = 0; n < M.nt; n++) {
    if (M.cfg.tadvect_itype[n] == POP_TADVECT_UPWIND3) up3 = 1;
    if (M.cfg.tadvect_itype[n] == POP_TADVECT_CENTERED) cen = 1;
  }
  if (up3) advt_upwind3(k, LTK, TRCR, WTK, WTKB, UTE, UTW, VTN, VTS, b);
  if (cen) advt_centered(k, LTK, TRCR, WTK, WTKB, UTE, VTN, b);
  memcpy(WTK, WTKB, sizeof(double) * M.n2); /* advection.F90:1960 WTK = WTKB */
  free(UTE); free(UTW); free(VTN); free(VTS); free(WTKB);
  M.timer[OT_ADVT] += o_now() - t0;
}

/* advu: advection.F90:1127-1570 */
void o_advu(int k, double* LUK, double* LVK, double* WUK, const double* UUU, const double* VVV,
            int b) {
  double t0 = o_now();
  const double *U = K3(UUU, k), *V = K3(VVV, k), *DYU = B2(M.DYU, b), *DXU = B2(M.DXU, b),
               *UR = B2(M.UAREA_R, b), *KXU = B2(M.KXU, b), *KYU = B2(M.KYU, b);
  const int* KMU = M.KMU + (size_t)b * M.n2;
  double *UUE = tmp2(), *UUW = tmp2(), *VUN = tmp2(), *VUS = tmp2(), *WUKB = tmp2();
  const double* DZU = PBC ? DZU3(b, k) : NULL;
  if (PBC) { /* :1245-1303 */
#define UD_(i, j) (U[IX2(i, j)] * DYU[IX2(i, j)] * DZU[IX2(i, j)])
#define VD_(i, j) (V[IX2(i, j)] * DXU[IX2(i, j)] * DZU[IX2(i, j)])
    for (int j = M.jb[b] - 1; j <= M.je[b] + 1; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
        UUW[IX2(i, j)] = 0.25 * (UD_(i, j) + UD_(i - 1, j)) +
                         0.125 * (UD_(i, j - 1) + UD_(i - 1, j - 1) + UD_(i, j + 1) + UD_(i - 1, j + 1));
        UUE[IX2(i, j)] = 0.25 * (UD_(i + 1, j) + UD_(i, j)) +
                         0.125 * (UD_(i + 1, j - 1) + UD_(i, j - 1) + UD_(i + 1, j + 1) + UD_(i, j + 1));
        VUS[IX2(i, j)] = 0.25 * (VD_(i, j) + VD_(i, j - 1)) +
                         0.125 * (VD_(i - 1, j) + VD_(i - 1, j - 1) + VD_(i + 1, j) + VD_(i + 1, j - 1));
        VUN[IX2(i, j)] = 0.25 * (VD_(i, j + 1) + VD_(i, j)) +
                         0.125 * (VD_(i - 1, j + 1) + VD_(i - 1, j) + VD_(i + 1, j + 1) + VD_(i + 1, j));
      }
#undef UD_
#undef VD_
  } else
  for (int j = M.jb[b] - 1; j <= M.je[b] + 1; j++)
    for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
      UUW[IX2(i, j)] = 0.25 * (U[IX2(i, j)] * DYU[IX2(i, j)] + U[IX2(i - 1, j)] * DYU[IX2(i - 1, j)]) +
                       0.125 * (U[IX2(i, j - 1)] * DYU[IX2(i, j - 1)] + U[IX2(i - 1, j - 1)] * DYU[IX2(i - 1, j - 1)] +
                                U[IX2(i, j + 1)] * DYU[IX2(i, j + 1)] + U[IX2(i - 1, j + 1)] * DYU[IX2(i - 1, j + 1)]);
      UUE[IX2(i, j)] = 0.25 * (U[IX2(i + 1, j)] * DYU[IX2(i + 1, j)] + U[IX2(i, j)] * DYU[IX2(i, j)]) +
                       0.125 * (U[IX2(i + 1, j - 1)] * DYU[IX2(i + 1, j - 1)] + U[IX2(i, j - 1)] * DYU[IX2(i, j - 1)] +
                                U[IX2(i + 1, j + 1)] * DYU[IX2(i + 1, j + 1)] + U[IX2(i, j + 1)] * DYU[IX2(i, j + 1)]);
      VUS[IX2(i, j)] = 0.25 * (V[IX2(i, j)] * DXU[IX2(i, j)] + V[IX2(i, j - 1)] * DXU[IX2(i, j - 1)]) +
                       0.125 * (V[IX2(i - 1, j)] * DXU[IX2(i - 1, j)] + V[IX2(i - 1, j - 1)] * DXU[IX2(i - 1, j - 1)] +
                                V[IX2(i + 1, j)] * DXU[IX2(i + 1, j)] + V[IX2(i + 1, j - 1)] * DXU[IX2(i + 1, j - 1)]);
      VUN[IX2(i, j)] = 0.25 * (V[IX2(i, j + 1)] * DXU[IX2(i, j + 1)] + V[IX2(i, j)] * DXU[IX2(i, j)]) +
                       0.125 * (V[IX2(i - 1, j + 1)] * DXU[IX2(i - 1, j + 1)] + V[IX2(i - 1, j)] * DXU[IX2(i - 1, j)] +
                                V[IX2(i + 1, j + 1)] * DXU[IX2(i + 1, j + 1)] + V[IX2(i + 1, j)] * DXU[IX2(i + 1, j)]);
    }
  for (size_t q = 0; q < M.n2; q++) {
    if (PBC) WUKB[q] = WUK[q] + (VUN[q] - VUS[q] + UUE[q] - UUW[q]) * UR[q]; /* :1352-1353 */
    else WUKB[q] = WUK[q] + M.c2dz[k] * 0.5 * (VUN[q] - VUS[q] + UUE[q] - UUW[q]) * UR[q];
  }
  memset(LUK, 0, sizeof(double) * M.n2);
  memset(LVK, 0, sizeof(double) * M.n2);
  PHYS2(b) {
    double cc = VUS[IX2(i, j + 1)] - VUS[IX2(i, j)] + UUW[IX2(i + 1, j)] - UUW[IX2(i, j)];
    LUK[IX2(i, j)] = 0.5 * (cc * U[IX2(i, j)] + VUS[IX2(i, j + 1)] * U[IX2(i, j + 1)] -
                            VUS[IX2(i, j)] * U[IX2(i, j - 1)] + UUW[IX2(i + 1, j)] * U[IX2(i + 1, j)] -
                            UUW[IX2(i, j)] * U[IX2(i - 1, j)]) * UR[IX2(i, j)];
    LVK[IX2(i, j)] = 0.5 * (cc * V[IX2(i, j)] + VUS[IX2(i, j + 1)] * V[IX2(i, j + 1)] -
                            VUS[IX2(i, j)] * V[IX2(i, j - 1)] + UUW[IX2(i + 1, j)] * V[IX2(i + 1, j)] -
                            UUW[IX2(i, j)] * V[IX2(i - 1, j)]) * UR[IX2(i, j)];
    if (PBC) { /* :1381-1405: ... * UAREA_R / DZU */
      LUK[IX2(i, j)] = LUK[IX2(i, j)] / DZU[IX2(i, j)];
      LVK[IX2(i, j)] = LVK[IX2(i, j)] / DZU[IX2(i, j)];
    }
  }
  if (k == 1) {
    for (size_t q = 0; q < M.n2; q++) {
      LUK[q] = LUK[q] + M.dzr[k] * WUK[q] * U[q];
      LVK[q] = LVK[q] + M.dzr[k] * WUK[q] * V[q];
    }
  } else {
    const double *Um = K3(UUU, k - 1), *Vm = K3(VVV, k - 1);
    for (size_t q = 0; q < M.n2; q++) {
      if (PBC) { /* :1443-1447 */
        LUK[q] = LUK[q] + 0.5 / DZU[q] * WUK[q] * (Um[q] + U[q]);
        LVK[q] = LVK[q] + 0.5 / DZU[q] * WUK[q] * (Vm[q] + V[q]);
      } else {
        LUK[q] = LUK[q] + M.dz2r[k] * WUK[q] * (Um[q] + U[q]);
        LVK[q] = LVK[q] + M.dz2r[k] * WUK[q] * (Vm[q] + V[q]);
      }
    }
  }
  if (k < M.km) {
    const double *Up = K3(UUU, k + 1), *Vp = K3(VVV, k + 1);
    for (size_t q = 0; q < M.n2; q++) {
      if (PBC) { /* :1462-1466 */
        LUK[q] = LUK[q] - 0.5 / DZU[q] * WUKB[q] * (U[q] + Up[q]);
        LVK[q] = LVK[q] - 0.5 / DZU[q] * WUKB[q] * (V[q] + Vp[q]);
      } else {
        LUK[q] = LUK[q] - M.dz2r[k] * WUKB[q] * (U[q] + Up[q]);
        LVK[q] = LVK[q] - M.dz2r[k] * WUKB[q] * (V[q] + Vp[q]);
      }
    }
  }
  PHYS2(b) {
    size_t q = IX2(i, j);
    if (k <= KMU[q]) {
      LUK[q] = LUK[q] + U[q] * V[q] * KYU[q] - (V[q] * V[q]) * KXU[q];
      LVK[q] = LVK[q] + U[q] * V[q] * KXU[q] - (U[q] * U[q]) * KYU[q];
    } else {
      LUK[q] = 0.0;
      LVK[q] = 0.0;
    }
  }
  memcpy(WUK, WUKB, sizeof(double) * M.n2);
  free(UUE); free(UUW); free(VUN); free(VUS); free(WUKB);
  M.timer[OT_ADVU] += o_now() - t0;
}

/* ------------------------------------------------------------ horizontal mixing */
/* 5-pt tracer coefficients: hmix_del2.F90:1064-1078 / hmix_del4.F90:1014-1023 */
static void tracer_coeffs(int k, int b, double* CC, double* CN, double* CS, double* CE, double* CW) {
  const int *KMT = M.KMT + (size_t)b * M.n2, *KMTN = M.KMTN + (size_t)b * M.n2,
            *KMTS = M.KMTS + (size_t)b * M.n2, *KMTE = M.KMTE + (size_t)b * M.n2,
            *KMTW = M.KMTW + (size_t)b * M.n2;
  if (PBC) { /* hmix_del2.F90:1034-1062 / hmix_del4.F90:964-988: the face is as thick as the thinner of its two cells */
    const double* DZT = DZT3(b, k);
    memset(CN, 0, sizeof(double) * M.n2); memset(CS, 0, sizeof(double) * M.n2);
    memset(CE, 0, sizeof(double) * M.n2); memset(CW, 0, sizeof(double) * M.n2);
#define MIN_(a, c) ((a) < (c) ? (a) : (c))
    for (int j = M.jb[b] - 1; j <= M.je[b] + 1; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
        size_t q = IX2(i, j);
        CN[q] = B2(M.DTN, b)[q] * MIN_(DZT[q], DZT[IX2(i, j + 1)]) / DZT[q];
        CS[q] = B2(M.DTS, b)[q] * MIN_(DZT[q], DZT[IX2(i, j - 1)]) / DZT[q];
        CE[q] = B2(M.DTE, b)[q] * MIN_(DZT[q], DZT[IX2(i + 1, j)]) / DZT[q];
        CW[q] = B2(M.DTW, b)[q] * MIN_(DZT[q], DZT[IX2(i - 1, j)]) / DZT[q];
      }
#undef MIN_
    for (size_t q = 0; q < M.n2; q++) {
      if (!(k <= KMTN[q] && k <= KMT[q])) CN[q] = 0.0;
      if (!(k <= KMTS[q] && k <= KMT[q])) CS[q] = 0.0;
      if (!(k <= KMTE[q] && k <= KMT[q])) CE[q] = 0.0;
      if (!(k <= KMTW[q] && k <= KMT[q])) CW[q] = 0.0;
      CC[q] = -(CN[q] + CS[q] + CE[q] + CW[q]);
    }
    return;
  }
  for (size_t q = 0; q < M.n2; q++) {
    CN[q] = (k <= KMTN[q] && k <= KMT[q]) ? B2(M.DTN, b)[q] : 0.0;
    CS[q] = (k <= KMTS[q] && k <= KMT[q]) ? B2(M.DTS, b)[q] : 0.0;
    CE[q] = (k <= KMTE[q] && k <= KMT[q]) ? B2(M.DTE, b)[q] : 0.0;
    CW[q] = (k <= KMTW[q] && k <= KMT[q]) ? B2(M.DTW, b)[q] : 0.0;
    CC[q] = -(CN[q] + CS[q] + CE[q] + CW[q]);
  }
}
/* hdifft_del2: hmix_del2.F90:970-1144 */
static void hdifft_del2(int k, double* HDTK, const double* TMIX, int b) {
  double *CC = tmp2(), *CN = tmp2(), *CS = tmp2(), *CE = tmp2(), *CW = tmp2();
  tracer_coeffs(k, b, CC, CN, CS, CE, CW);
  memset(HDTK, 0, sizeof(double) * M.n2 * M.nt);
  for (int n = 1; n <= M.nt; n++) {
    const double* T = KN4(TMIX, k, n);
    double* H = HDTK + (size_t)(n - 1) * M.n2;
    PHYS2(b)
    H[IX2(i, j)] = M.ah * (CC[IX2(i, j)] * T[IX2(i, j)] + CN[IX2(i, j)] * T[IX2(i, j + 1)] +
                           CS[IX2(i, j)] * T[IX2(i, j - 1)] + CE[IX2(i, j)] * T[IX2(i + 1, j)] +
                           CW[IX2(i, j)] * T[IX2(i - 1, j)]);
  }
  free(CC); free(CN); free(CS); free(CE); free(CW);
}
/* hdifft_del4: hmix_del4.F90:889-1106 */
static void hdifft_del4(int k, double* HDTK, const double* TMIX, int b) {
  double *CC = tmp2(), *CN = tmp2(), *CS = tmp2(), *CE = tmp2(), *CW = tmp2(), *D2 = tmp2();
  const double* AHF = B2(M.AHF, b);
  tracer_coeffs(k, b, CC, CN, CS, CE, CW);
  for (int n = 1; n <= M.nt; n++) {
    const double* T = KN4(TMIX, k, n);
    double* H = HDTK + (size_t)(n - 1) * M.n2;
    for (int j = M.jb[b] - 1; j <= M.je[b] + 1; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
        double v = CC[IX2(i, j)] * T[IX2(i, j)] + CN[IX2(i, j)] * T[IX2(i, j + 1)] +
                   CS[IX2(i, j)] * T[IX2(i, j - 1)] + CE[IX2(i, j)] * T[IX2(i + 1, j)] +
                   CW[IX2(i, j)] * T[IX2(i - 1, j)];
        D2[IX2(i, j)] = M.cfg.lvariable_hmixt ? AHF[IX2(i, j)] * v : v;
      }
    memset(H, 0, sizeof(double) * M.n2);
    PHYS2(b)
    H[IX2(i, j)] = M.ah * (CC[IX2(i, j)] * D2[IX2(i, j)] + CN[IX2(i, j)] * D2[IX2(i, j + 1)] +
                           CS[IX2(i, j)] * D2[IX2(i, j - 1)] + CE[IX2(i, j)] * D2[IX2(i + 1, j)] +
                           CW[IX2(i, j)] * D2[IX2(i - 1, j)]);
  }
  free(CC); free(CN); free(CS); free(CE); free(CW); free(D2);
}
void o_hdifft_gm(int k, double* HDTK, const double* TMIX, const double* UMIX, const double* VMIX,
                 int b); /* o_gm.c */
/* hdifft dispatcher: horizontal_mix.F90:486-619 */
void o_hdifft(int k, double* HDTK, const double* TMIX, const double* UMIX, const double* VMIX,
              int b) {
  double t0 = o_now();
  memset(HDTK, 0, sizeof(double) * M.n2 * M.nt);
  switch (M.cfg.hmix_tracer_itype) {
    case POP_HMIX_DEL2: hdifft_del2(k, HDTK, TMIX, b); break;
    case POP_HMIX_DEL4: hdifft_del4(k, HDTK, TMIX, b); break;
    case POP_HMIX_GM: o_hdifft_gm(k, HDTK, TMIX, UMIX, VMIX, b); break;
    default: break;
  }
  M.timer[OT_HDIFFT] += o_now() - t0;
}

/* 5+5-pt momentum stencil used by del2 and (twice) by del4 */
static inline double mom_stencil(const double* CCa, const double* DUN, const double* DUS,
                                 const double* DUE, const double* DUW, const double* DMC,
                                 const double* DMN, const double* DMS, const double* DME,
                                 const double* DMW, const double* A, const double* Bf, int i, int j,
                                 double sgn) {
  double s1 = CCa[IX2(i, j)] * A[IX2(i, j)] + DUN[IX2(i, j)] * A[IX2(i, j + 1)] +
              DUS[IX2(i, j)] * A[IX2(i, j - 1)] + DUE[IX2(i, j)] * A[IX2(i + 1, j)] +
              DUW[IX2(i, j)] * A[IX2(i - 1, j)];
  double s2 = DMC[IX2(i, j)] * Bf[IX2(i, j)] + DMN[IX2(i, j)] * Bf[IX2(i, j + 1)] +
              DMS[IX2(i, j)] * Bf[IX2(i, j - 1)] + DME[IX2(i, j)] * Bf[IX2(i + 1, j)] +
              DMW[IX2(i, j)] * Bf[IX2(i - 1, j)];
  return (sgn > 0) ? (s1 + s2) : (s1 - s2);
}
/* hdiffu_del2: hmix_del2.F90:670-963 ; hdiffu_del4: hmix_del4.F90:597-882 */
void o_hdiffu(int k, double* HDUK, double* HDVK, const double* UMIXK, const double* VMIXK, int b) {
  double t0 = o_now();
  const int* KMU = M.KMU + (size_t)b * M.n2;
  const double *DUC = B2(M.DUC, b), *DUM = B2(M.DUM, b), *DUN = B2(M.DUN, b), *DUS = B2(M.DUS, b),
               *DUE = B2(M.DUE, b), *DUW = B2(M.DUW, b), *DMC = B2(M.DMC, b), *DMN = B2(M.DMN, b),
               *DMS = B2(M.DMS, b), *DME = B2(M.DME, b), *DMW = B2(M.DMW, b), *AMF = B2(M.AMF, b);
  double* CC = tmp2();
  for (size_t q = 0; q < M.n2; q++) CC[q] = DUC[q] + DUM[q];
  memset(HDUK, 0, sizeof(double) * M.n2);
  memset(HDVK, 0, sizeof(double) * M.n2);
  double *PN = NULL, *PS = NULL, *PE = NULL, *PW = NULL;
  if (PBC) { /* hmix_del2.F90:852-863, hmix_del4.F90:683-694: neighbour weights scaled by min(DZU)/DZU */
    const double* DZU = DZU3(b, k);
    PN = tmp2(); PS = tmp2(); PE = tmp2(); PW = tmp2();
#define MIN_(a, c) ((a) < (c) ? (a) : (c))
    for (int j = M.jb[b] - 1; j <= M.je[b] + 1; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
        size_t q = IX2(i, j);
        PN[q] = DUN[q] * MIN_(DZU[IX2(i, j + 1)], DZU[q]) / DZU[q];
        PS[q] = DUS[q] * MIN_(DZU[IX2(i, j - 1)], DZU[q]) / DZU[q];
        PE[q] = DUE[q] * MIN_(DZU[IX2(i + 1, j)], DZU[q]) / DZU[q];
        PW[q] = DUW[q] * MIN_(DZU[IX2(i - 1, j)], DZU[q]) / DZU[q];
      }
#undef MIN_
    DUN = PN; DUS = PS; DUE = PE; DUW = PW;
  }
  if (M.cfg.hmix_momentum_itype == POP_HMIX_DEL2) {
    PHYS2(b) {
      HDUK[IX2(i, j)] = M.am * mom_stencil(CC, DUN, DUS, DUE, DUW, DMC, DMN, DMS, DME, DMW, UMIXK, VMIXK, i, j, +1);
      HDVK[IX2(i, j)] = M.am * mom_stencil(CC, DUN, DUS, DUE, DUW, DMC, DMN, DMS, DME, DMW, VMIXK, UMIXK, i, j, -1);
    }
  } else {
    double *D2U = tmp2(), *D2V = tmp2();
    for (int j = M.jb[b] - 1; j <= M.je[b] + 1; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b] + 1; i++) {
        D2U[IX2(i, j)] = mom_stencil(CC, DUN, DUS, DUE, DUW, DMC, DMN, DMS, DME, DMW, UMIXK, VMIXK, i, j, +1);
        D2V[IX2(i, j)] = mom_stencil(CC, DUN, DUS, DUE, DUW, DMC, DMN, DMS, DME, DMW, VMIXK, UMIXK, i, j, -1);
      }
    for (size_t q = 0; q < M.n2; q++) {
      if (M.cfg.lvariable_hmixu) {
        if (k <= KMU[q]) { D2U[q] = AMF[q] * D2U[q]; D2V[q] = AMF[q] * D2V[q]; }
        else { D2U[q] = 0.0; D2V[q] = 0.0; }
      } else if (k > KMU[q]) { D2U[q] = 0.0; D2V[q] = 0.0; }
    }
    PHYS2(b) {
      HDUK[IX2(i, j)] = M.am * mom_stencil(CC, DUN, DUS, DUE, DUW, DMC, DMN, DMS, DME, DMW, D2U, D2V, i, j, +1);
      HDVK[IX2(i, j)] = M.am * mom_stencil(CC, DUN, DUS, DUE, DUW, DMC, DMN, DMS, DME, DMW, D2V, D2U, i, j, -1);
    }
    free(D2U); free(D2V);
  }
  for (size_t q = 0; q < M.n2; q++)
    if (k > KMU[q]) { HDUK[q] = 0.0; HDVK[q] = 0.0; }
  free(CC);
  if (PBC) { free(PN); free(PS); free(PE); free(PW); }
  M.timer[OT_HDIFFU] += o_now() - t0;
}

/* ------------------------------------------------------------ vertical_mix.F90 */
static inline const double* vdc_level(int b, int k, int mt2) { /* VDC(:,:,k,mt2,bid) */
  int kk = k;
  if (M.vdc_nk == 1) kk = 1;
  return M.VDC + (((size_t)b * M.vdc_nd + (mt2 - 1)) * M.vdc_nk + (kk - M.vdc_k0)) * M.n2;
}
static inline double* vdc_level_w(int b, int k, int mt2) { return (double*)vdc_level(b, k, mt2); }
static inline const double* vvc_level(int b, int k) {
  int kk = (M.vvc_nk == 1) ? 1 : k;
  return M.VVC + ((size_t)b * M.vvc_nk + (kk - 1)) * M.n2;
}
/* vdifft: vertical_mix.F90:691-847 */
void o_vdifft(int k, double* VDTK, const double* TOLD, const double* STF, int b) {
  const int* KMT = M.KMT + (size_t)b * M.n2;
  int kp1 = (k < M.km) ? k + 1 : M.km;
  for (int n = 1; n <= M.nt; n++) {
    int mt2 = (n < M.vdc_nd) ? n : M.vdc_nd;
    double* VTF = M.VTF + ((size_t)b * M.nt + (n - 1)) * M.n2;
    const double *VDC = vdc_level(b, k, mt2), *T = KN4(TOLD, k, n), *Tp = KN4(TOLD, kp1, n);
    const double* S = STF + (size_t)(n - 1) * M.n2;
    double* VD = VDTK + (size_t)(n - 1) * M.n2;
    if (k == 1)
      for (size_t q = 0; q < M.n2; q++) VTF[q] = (KMT[q] >= 1) ? S[q] : 0.0;
    if (PBC) { /* :790-803 */
      const double *DZT = DZT3(b, k), *DZTp = DZT3(b, kp1);
      for (size_t q = 0; q < M.n2; q++) {
        double VTFB = (KMT[q] > k) ? VDC[q] * (T[q] - Tp[q]) / (0.5 * (DZT[q] + DZTp[q])) : 0.0;
        VD[q] = (k <= KMT[q]) ? (VTF[q] - VTFB) / DZT[q] : 0.0;
        VTF[q] = VTFB;
      }
    } else
    for (size_t q = 0; q < M.n2; q++) {
      double VTFB = (KMT[q] > k) ? VDC[q] * (T[q] - Tp[q]) * M.dzwr[k] : 0.0;
      VD[q] = (k <= KMT[q]) ? (VTF[q] - VTFB) * M.dzr[k] : 0.0;
      VTF[q] = VTFB;
    }
  }
}
/* vdiffu: vertical_mix.F90:853-1026 */
void o_vdiffu(int k, double* VDUK, double* VDVK, const double* UOLD, const double* VOLD,
              const double* SMF, int b) {
  const int* KMU = M.KMU + (size_t)b * M.n2;
  int kp1 = (k < M.km) ? k + 1 : M.km;
  double *VUF = B2(M.VUF, b), *VVF = B2(M.VVF, b);
  const double *VVC = vvc_level(b, k), *U = K3(UOLD, k), *Up = K3(UOLD, kp1), *V = K3(VOLD, k),
               *Vp = K3(VOLD, kp1);
  double *VUFB = tmp2(), *VVFB = tmp2();
  if (k == 1)
    for (size_t q = 0; q < M.n2; q++) {
      VUF[q] = (KMU[q] >= 1) ? SMF[q] : 0.0;
      VVF[q] = (KMU[q] >= 1) ? SMF[M.n2 + q] : 0.0;
    }
  if (PBC) { /* :946-958: kp1 = min(k+1, km), so at k = km WORK = p5*DZU(km) */
    const double *DZU = DZU3(b, k), *DZUp = DZU3(b, kp1);
    for (size_t q = 0; q < M.n2; q++) {
      double W = (k < M.km) ? 0.5 * (DZU[q] + DZUp[q]) : 0.5 * DZUp[q];
      VUFB[q] = VVC[q] * (U[q] - Up[q]) / W;
      VVFB[q] = VVC[q] * (V[q] - Vp[q]) / W;
    }
  } else
  for (size_t q = 0; q < M.n2; q++) {
    VUFB[q] = VVC[q] * (U[q] - Up[q]) * M.dzwr[k];
    VVFB[q] = VVC[q] * (V[q] - Vp[q]) * M.dzwr[k];
  }
  PHYS2(b) {
    size_t q = IX2(i, j);
    if (k == KMU[q]) {
      double vmag = M.cfg.bottom_drag * sqrt(U[q] * U[q] + V[q] * V[q]);
      VUFB[q] = vmag * U[q];
      VVFB[q] = vmag * V[q];
    }
  }
  for (size_t q = 0; q < M.n2; q++) {
    if (PBC) { /* :991-995 */
      const double* DZU = DZU3(b, k);
      VDUK[q] = (k <= KMU[q]) ? (VUF[q] - VUFB[q]) / DZU[q] : 0.0;
      VDVK[q] = (k <= KMU[q]) ? (VVF[q] - VVFB[q]) / DZU[q] : 0.0;
    } else {
      VDUK[q] = (k <= KMU[q]) ? (VUF[q] - VUFB[q]) * M.dzr[k] : 0.0;
      VDVK[q] = (k <= KMU[q]) ? (VVF[q] - VVFB[q]) * M.dzr[k] : 0.0;
    }
    VUF[q] = VUFB[q];
    VVF[q] = VVFB[q];
  }
  free(VUFB); free(VVFB);
}

/* impvmixt: vertical_mix.F90:1164-1382 ; impvmixt_correct: :1460-1672 (correct != 0) */
static void impvmixt_common(double* TNEW, const double* TOLD, const double* PSFC,
                            const double* RHS, int nfirst, int nlast, int b, int correct) {
  const int km = M.km;
  if (nfirst > nlast || nfirst > M.nt) return;
  double t0 = o_now();
  const int* KMT = M.KMT + (size_t)b * M.n2;
  double* hfac_t = (double*)malloc(sizeof(double) * (km + 2));
  for (int k = 1; k <= km; k++) hfac_t[k] = M.dz[k] / M.c2dtt[k];
  double *H1 = tmp2(), *A = tmp2(), *B = tmp2(), *C = tmp2(), *D = tmp2();
  double *E = (double*)calloc(M.n3, sizeof(double)), *F = (double*)calloc(M.n3, sizeof(double));
  for (size_t q = 0; q < M.n2; q++)
    H1[q] = (M.cfg.sfc_layer_type == POP_SFC_VARTHICK) ? hfac_t[1] + PSFC[q] / (O_GRAV * M.c2dtt[1])
                                                        : hfac_t[1];
  for (int n = nfirst; n <= nlast; n++) {
    int mt2 = (n < M.vdc_nd) ? n : M.vdc_nd;
    {
      const double* VDC = vdc_level(b, 1, mt2);
      const double* R1 = correct ? RHS + (size_t)(n - 1) * M.n2 : KN4(TNEW, 1, n);
      PHYS2(b) {
        size_t q = IX2(i, j);
        A[q] = M.afac_t[1] * VDC[q];
        D[q] = H1[q] + A[q];
        K3(E, 1)[q] = A[q] / D[q];
        B[q] = H1[q] * K3(E, 1)[q];
        K3(F, 1)[q] = hfac_t[1] * R1[q] / D[q];
      }
    }
    for (int k = 2; k <= km; k++) {
      const double* VDC = vdc_level(b, k, mt2);
      const double* Rk = KN4(TNEW, k, n);
      PHYS2(b) {
        size_t q = IX2(i, j);
        C[q] = A[q];
        double hf = hfac_t[k];
        if (PBC) { /* :1280-1286 / :1577-1582: DZT(k+1) is read at k = km too (DZT is dimensioned 0:km+1) */
          A[q] = M.cfg.aidif * VDC[q] / (0.5 * (DZT3(b, k)[q] + DZT3(b, k + 1)[q]));
          hf = DZT3(b, k)[q] / M.c2dtt[k];
        } else {
          A[q] = M.afac_t[k] * VDC[q];
        }
        if (k > KMT[q]) {
          K3(F, k)[q] = 0.0;
        } else {
          if (k == KMT[q]) D[q] = hf + B[q];
          else D[q] = hf + A[q] + B[q];
          K3(E, k)[q] = A[q] / D[q];
          B[q] = (hf + B[q]) * K3(E, k)[q];
          if (correct) K3(F, k)[q] = C[q] * K3(F, k - 1)[q] / D[q];
          else K3(F, k)[q] = (hf * Rk[q] + C[q] * K3(F, k - 1)[q]) / D[q];
        }
      }
    }
    for (int k = km - 1; k >= 1; k--)
      PHYS2(b) {
        size_t q = IX2(i, j);
        if (k < KMT[q]) K3(F, k)[q] = K3(F, k)[q] + K3(E, k)[q] * K3(F, k + 1)[q];
      }
    for (int k = 1; k <= km; k++) {
      double* Tn = KN4(TNEW, k, n);
      const double* To = correct ? Tn : KN4(TOLD, k, n);
      PHYS2(b) Tn[IX2(i, j)] = To[IX2(i, j)] + K3(F, k)[IX2(i, j)];
    }
  }
  free(hfac_t); free(H1); free(A); free(B); free(C); free(D); free(E); free(F);
  M.timer[OT_VMIXT] += o_now() - t0;
}
void o_impvmixt(double* TNEW, const double* TOLD, const double* PSFC, int nfirst, int nlast,
                int b) {
  impvmixt_common(TNEW, TOLD, PSFC, NULL, nfirst, nlast, b, 0);
}
void o_impvmixt_correct(double* TNEW, const double* PSFC, const double* RHS, int nfirst,
                        int nlast, int b) {
  impvmixt_common(TNEW, NULL, PSFC, RHS, nfirst, nlast, b, 1);
}

/* impvmixu: vertical_mix.F90:1679-1881 */
void o_impvmixu(double* UNEW, double* VNEW, int b) {
  double t0 = o_now();
  const int km = M.km;
  const int* KMU = M.KMU + (size_t)b * M.n2;
  double* hfac_u = (double*)malloc(sizeof(double) * (km + 2));
  for (int k = 1; k <= km; k++) hfac_u[k] = M.dz[k] / M.c2dtu;
  double *A = tmp2(), *B = tmp2(), *C = tmp2(), *D = tmp2();
  double *E = (double*)calloc(M.n3, sizeof(double)), *F1 = (double*)calloc(M.n3, sizeof(double)),
         *F2 = (double*)calloc(M.n3, sizeof(double));
  {
    const double* VVC = vvc_level(b, 1);
    PHYS2(b) {
      size_t q = IX2(i, j);
      A[q] = M.afac_u[1] * VVC[q];
      D[q] = hfac_u[1] + A[q];
      K3(E, 1)[q] = A[q] / D[q];
      B[q] = hfac_u[1] * K3(E, 1)[q];
      K3(F1, 1)[q] = hfac_u[1] * K3(UNEW, 1)[q] / D[q];
      K3(F2, 1)[q] = hfac_u[1] * K3(VNEW, 1)[q] / D[q];
    }
  }
  for (int k = 2; k <= km; k++) {
    const double* VVC = vvc_level(b, k);
    PHYS2(b) {
      size_t q = IX2(i, j);
      C[q] = A[q];
      double hf = hfac_u[k];
      if (PBC) { /* :1777-1784 */
        hf = DZU3(b, k)[q] / M.c2dtu;
        A[q] = M.cfg.aidif * VVC[q] / (0.5 * (DZU3(b, k)[q] + DZU3(b, k + 1)[q]));
      } else {
        A[q] = M.afac_u[k] * VVC[q];
      }
      if (k < KMU[q]) {
        D[q] = hf + A[q] + B[q];
      } else if (k == KMU[q]) {
        D[q] = hf + B[q];
      }
      if (k <= KMU[q]) {
        K3(E, k)[q] = A[q] / D[q];
        B[q] = (hf + B[q]) * K3(E, k)[q];
        K3(F1, k)[q] = (hf * K3(UNEW, k)[q] + C[q] * K3(F1, k - 1)[q]) / D[q];
        K3(F2, k)[q] = (hf * K3(VNEW, k)[q] + C[q] * K3(F2, k - 1)[q]) / D[q];
      } else {
        K3(F1, k)[q] = 0.0;
        K3(F2, k)[q] = 0.0;
      }
    }
  }
  for (int k = km - 1; k >= 1; k--)
    PHYS2(b) {
      size_t q = IX2(i, j);
      if (k < KMU[q]) {
        K3(F1, k)[q] = K3(F1, k)[q] + K3(E, k)[q] * K3(F1, k + 1)[q];
        K3(F2, k)[q] = K3(F2, k)[q] + K3(E, k)[q] * K3(F2, k + 1)[q];
      }
    }
  for (int k = 1; k <= km; k++)
    PHYS2(b) {
      size_t q = IX2(i, j);
      K3(UNEW, k)[q] = K3(F1, k)[q];
      K3(VNEW, k)[q] = K3(F2, k)[q];
    }
  free(hfac_u); free(A); free(B); free(C); free(D); free(E); free(F1); free(F2);
  M.timer[OT_VMIXU] += o_now() - t0;
}

/* vmix_coeffs: vertical_mix.F90:518-670 -> vmix_coeffs_const (vmix_const.F90:144-233),
   vmix_coeffs_rich (vmix_rich.F90:179-414).  GIVEN: VDC/VVC are supplied, nothing to do. */
void o_vmix_coeffs(int k, const double* TMIX, const double* UMIX, const double* VMIX,
                   const double* RHOMIX, int b) {
  const pop_config* c = &M.cfg;
  const int km = M.km;
  const int *KMT = M.KMT + (size_t)b * M.n2, *KMU = M.KMU + (size_t)b * M.n2;
  if (c->vmix_itype == POP_VMIX_GIVEN) return;
  int kp1 = (k + 1 < km) ? k + 1 : km;
  if (c->vmix_itype == POP_VMIX_CONST) {
    if (!c->convection_diff) return;
    double *RHOK = tmp2(), *RHOKP = tmp2();
    double* VVC = (double*)vvc_level(b, k);
    double* VDC = vdc_level_w(b, k, 1);
    for (size_t q = 0; q < M.n2; q++) { VVC[q] = c->const_vvc; VDC[q] = c->const_vdc; }
    o_state(k, kp1, KN4(TMIX, k, 1), KN4(TMIX, k, 2), b, RHOK, NULL, NULL, NULL);
    o_state(kp1, kp1, KN4(TMIX, kp1, 1), KN4(TMIX, kp1, 2), b, RHOKP, NULL, NULL, NULL);
    double vvconv = (c->convect_visc != 0.0) ? c->convect_visc : c->const_vvc;
    for (size_t q = 0; q < M.n2; q++)
      if (RHOK[q] > RHOKP[q] && k < KMT[q]) { VDC[q] = c->convect_diff; VVC[q] = vvconv; }
    free(RHOK); free(RHOKP);
    return;
  }
  /* Richardson-number mixing (full cells) */
  const double eps = 1.0e-10; /* pop_constants eps */
  double *UTK = B2(M.UTK, b), *VTK = B2(M.VTK, b);
  double *UTKP = tmp2(), *VTKP = tmp2(), *RHOK = tmp2(), *RICH = tmp2(), *RICHU = tmp2();
  if (k == 1) {
    o_ugrid_to_tgrid(UTK, K3(UMIX, k), b);
    o_ugrid_to_tgrid(VTK, K3(VMIX, k), b);
  }
  o_ugrid_to_tgrid(UTKP, K3(UMIX, kp1), b);
  o_ugrid_to_tgrid(VTKP, K3(VMIX, kp1), b);
  o_state(k, kp1, KN4(TMIX, k, 1), KN4(TMIX, k, 2), b, RHOK, NULL, NULL, NULL);
  double critnu = c->convection_diff ? c->convect_diff : (0.25 * M.dz[k] * M.dzw[k]) / M.c2dtt[k];
  double* VDC = vdc_level_w(b, k, 1);
  const double* RHOP = K3(RHOMIX, kp1);
  for (size_t q = 0; q < M.n2; q++) {
    if (k < KMT[q]) {
      RICH[q] = -O_GRAV * M.dzw[k] * (RHOK[q] - RHOP[q]) /
                ((UTK[q] - UTKP[q]) * (UTK[q] - UTKP[q]) + (VTK[q] - VTKP[q]) * (VTK[q] - VTKP[q]) + eps);
      double d = 1.0 + 5.0 * RICH[q];
      double v = c->bckgrnd_vdc + (c->bckgrnd_vvc + c->rich_mix / (d * d)) / d;
      VDC[q] = (critnu < v) ? critnu : v;
    } else {
      RICH[q] = 0.0;
      VDC[q] = 0.0;
    }
    if (RICH[q] < 0.0) VDC[q] = critnu;
  }
  o_tgrid_to_ugrid(RICHU, RICH, b);
  if (c->convection_diff) critnu = c->convect_visc;
  double* VVC = (double*)vvc_level(b, k);
  for (size_t q = 0; q < M.n2; q++) {
    if (k < KMU[q]) {
      double d = 1.0 + 5.0 * RICHU[q];
      double v = c->bckgrnd_vvc + c->rich_mix / (d * d);
      VVC[q] = (critnu < v) ? critnu : v;
    } else {
      RICHU[q] = 0.0;
      VVC[q] = 0.0;
    }
    if (RICHU[q] < 0.0) VVC[q] = critnu;
  }
  memcpy(UTK, UTKP, sizeof(double) * M.n2);
  memcpy(VTK, VTKP, sizeof(double) * M.n2);
  free(UTKP); free(VTKP); free(RHOK); free(RICH); free(RICHU);
}
