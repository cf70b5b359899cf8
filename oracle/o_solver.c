/*
 * o_solver.c -- ORACLE (test infrastructure): barotropic elliptic solver.
 * Restates source/POP_SolversMod.F90: POP_SolversInit :783-820,895-906; POP_SolversDiagonal
 * :1110-1151; POP_SolversRun :327-495 (clinic == tropic distribution, redistribution = copy);
 * pcg :1200-1503; PCSI :1510-1835; ChronGear :1841-2266; btropOperator :2376-2431;
 * PcsiLanczos :2699-2990; ratqr :3122-3222; the EVP block preconditioner :2273-2369,2434-2696,2992-3120.
 * Block loops and element-wise loops run as OpenMP teams (the reference threads the same loops over blocks,
 * POP_SolversMod.F90 "!$OMP PARALLEL DO PRIVATE(iblock)"); every global sum is combined in block order.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pop_oracle.h"

void* o_alloc_d(size_t n);
#define NB (M.nblocks)
#define NTOT (M.n2 * (size_t)M.nblocks)
static double* vec(void) { return (double*)calloc(NTOT, sizeof(double)); }

int o_solvers_init(void) {
  size_t n = NTOT;
  M.centerWgtClinicIndep = o_alloc_d(n); M.centerWgtClinic = o_alloc_d(n);
  M.btropWgtCenter = o_alloc_d(n); M.btropWgtNorth = o_alloc_d(n); M.btropWgtEast = o_alloc_d(n);
  M.btropWgtNE = o_alloc_d(n); M.mMaskTropic = o_alloc_d(n);
  double* work0 = vec();
  for (int b = 0; b < NB; b++) {
    const double *HU = B2(M.HU, b), *DXUR = B2(M.DXUR, b), *DYUR = B2(M.DYUR, b),
                 *DXU = B2(M.DXU, b), *DYU = B2(M.DYU, b), *TAREA = B2(M.TAREA, b);
    for (int j = 2; j <= M.nyb; j++)
      for (int i = 2; i <= M.nxb; i++) {
        double xne = 0.25 * HU[IX2(i, j)] * DXUR[IX2(i, j)] * DYU[IX2(i, j)];
        double xse = 0.25 * HU[IX2(i, j - 1)] * DXUR[IX2(i, j - 1)] * DYU[IX2(i, j - 1)];
        double xnw = 0.25 * HU[IX2(i - 1, j)] * DXUR[IX2(i - 1, j)] * DYU[IX2(i - 1, j)];
        double xsw = 0.25 * HU[IX2(i - 1, j - 1)] * DXUR[IX2(i - 1, j - 1)] * DYU[IX2(i - 1, j - 1)];
        double yne = 0.25 * HU[IX2(i, j)] * DYUR[IX2(i, j)] * DXU[IX2(i, j)];
        double yse = 0.25 * HU[IX2(i, j - 1)] * DYUR[IX2(i, j - 1)] * DXU[IX2(i, j - 1)];
        double ynw = 0.25 * HU[IX2(i - 1, j)] * DYUR[IX2(i - 1, j)] * DXU[IX2(i - 1, j)];
        double ysw = 0.25 * HU[IX2(i - 1, j - 1)] * DYUR[IX2(i - 1, j - 1)] * DXU[IX2(i - 1, j - 1)];
        size_t q = IX2(i, j);
        double ne = xne + yne, ase = xse + yse, anw = xnw + ynw, asw = xsw + ysw;
        B2(M.btropWgtNE, b)[q] = ne;
        B2(M.btropWgtEast, b)[q] = xne + xse - yne - yse;
        B2(M.btropWgtNorth, b)[q] = yne + ynw - xne - xnw;
        B2(M.centerWgtClinicIndep, b)[q] = -(ne + ase + anw + asw);
        B2(work0, b)[q] = TAREA[q] * TAREA[q];
        B2(M.mMaskTropic, b)[q] = B2(M.RCALCT, b)[q];
      }
  }
  M.residualNorm = 1.0 / oracle_global_sum(work0, POP_LOC_CENTER, M.mMaskTropic);
  M.convergenceCriterion = (M.cfg.convergence_criterion * M.cfg.convergence_criterion) / M.residualNorm;
  free(work0);
  return 0;
}

void o_solvers_diagonal(const double* diagCorr, int b) {
  for (size_t q = 0; q < M.n2; q++)
    B2(M.centerWgtClinic, b)[q] = B2(M.centerWgtClinicIndep, b)[q] - diagCorr[q];
}

/* btropOperator: AX(:,:,bid) for one block of block-major vectors */
void o_btrop_operator(double* AXv, const double* Xv, int b) {
  double* AX = B2(AXv, b);
  const double* X = B2(Xv, b);
  const double *C = B2(M.btropWgtCenter, b), *N = B2(M.btropWgtNorth, b),
               *E = B2(M.btropWgtEast, b), *NE = B2(M.btropWgtNE, b);
  memset(AX, 0, sizeof(double) * M.n2);
  for (int j = 2; j <= M.nyb - 1; j++)
    for (int i = 2; i <= M.nxb - 1; i++)
      AX[IX2(i, j)] = C[IX2(i, j)] * X[IX2(i, j)] + N[IX2(i, j)] * X[IX2(i, j + 1)] +
                      N[IX2(i, j - 1)] * X[IX2(i, j - 1)] + E[IX2(i, j)] * X[IX2(i + 1, j)] +
                      E[IX2(i - 1, j)] * X[IX2(i - 1, j)] + NE[IX2(i, j)] * X[IX2(i + 1, j + 1)] +
                      NE[IX2(i, j - 1)] * X[IX2(i + 1, j - 1)] + NE[IX2(i - 1, j)] * X[IX2(i - 1, j + 1)] +
                      NE[IX2(i - 1, j - 1)] * X[IX2(i - 1, j - 1)];
}

static void halo(double* v) { oracle_halo_2d(v, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0); }
static double gsum(const double* w) { return oracle_global_sum(w, POP_LOC_CENTER, M.mMaskTropic); }
static void set_a0r(double* A0R) {
  for (size_t q = 0; q < NTOT; q++)
    A0R[q] = (M.btropWgtCenter[q] != 0.0) ? 1.0 / M.btropWgtCenter[q] : 0.0;
}


/* ------------------------------------------------------------------------------------------------
 * EVP block preconditioner (preconditionerChoice = 'evp', the production default,
 * bld/namelist_files/namelist_defaults_pop.xml:262): POP_SolversMod.F90
 *   EvpBlockPartition :2992-3040, EvpPre :2434-2506, ExplicitBlockEvpPre :2508-2616, inverse :3042-3120,
 *   ExplicitEvp :2618-2696, preconditioner :2273-2369, the EVP part of POP_SolversPrep :252-292.
 * Every block is cut into sub-blocks of at most EvpXbs x EvpYbs = 8 x 8 cells; on a sub-block without land the
 * five-point operator (centre + the four corner weights; the small N/S/E/W weights are dropped) is inverted
 * exactly by error-vector propagation (Roache 1995): march north-east from a guessed south/west edge, read the
 * error that arrives at the north/east frame, correct the guess with the pre-computed influence matrix, march
 * again.  Sub-blocks with land use the diagonal.  The coefficient copies are those of PREP time (the reference
 * freezes EvpCenterWgt in POP_SolversPrep and keeps using it when the centre weight changes with the time step).
 * Arrays keep the Fortran shapes: evp arrays (EVP_LD, EVP_LD, nsub) and rinv (EVP_L, EVP_L, nsub), first index
 * fastest, per block.
 * ------------------------------------------------------------------------------------------------ */
#define EVP_XBS 8 /* POP_SolversMod.F90:158-159 */
#define EVP_YBS 8
#define EVP_LD (EVP_XBS + 2)
#define EVP_L (EVP_XBS + EVP_YBS - 1)
static struct {
  int ready, xnb, ynb, nsub;
  int *xidx, *yidx; /* EvpXbidx(1:xnb+1), EvpYbidx(1:ynb+1): [0] unused */
  int* land;        /* landIndx(xnb, ynb, nblocks) */
  double *cc, *icc, *ne, *ine, *rinv;
} E;
#define EV2(a, b, ib) ((a) + ((size_t)(b)*E.nsub + (ib)) * EVP_LD * EVP_LD)   /* (EVP_LD,EVP_LD) slab of sub-block ib (0-based) */
#define EVR(b, ib) (E.rinv + ((size_t)(b)*E.nsub + (ib)) * EVP_L * EVP_L)
#define F2(a, i, j) (a)[((j)-1) * EVP_LD + ((i)-1)] /* 1-based (i,j) of an (EVP_LD, *) array */
#define RV(r, k, j) (r)[((j)-1) * EVP_L + ((k)-1)]  /* rinv(k,j) */

/* EvpBlockPartition :2992-3040 */
static void evp_partition(int m, int mm, int* mb_out, int** mdi_out) {
  int mb = (m - 3) / mm + 1;
  int* mdi = (int*)calloc(mb + 2, sizeof(int));
  mdi[1] = 2;
  if (mb == 1) {
    mdi[mb + 1] = m;
  } else {
    for (int i = 1; i <= mb - 2; i++) mdi[i + 1] = 2 + i * mm;
    mdi[mb] = (mdi[mb - 1] + m) / 2;
    mdi[mb + 1] = m;
  }
  *mb_out = mb;
  *mdi_out = mdi;
}

/* inverse :3042-3120: Doolittle LU without pivoting, column by column; a is destroyed */
static void evp_inverse(double* a, double* cinv, int n) {
#define A_(i, j) a[((j)-1) * n + ((i)-1)]
#define L_(i, j) Lm[((j)-1) * n + ((i)-1)]
#define U_(i, j) Um[((j)-1) * n + ((i)-1)]
  double *Lm = (double*)calloc((size_t)n * n, sizeof(double)), *Um = (double*)calloc((size_t)n * n, sizeof(double));
  double *bv = (double*)calloc(n + 1, sizeof(double)), *d = (double*)calloc(n + 1, sizeof(double)),
         *x = (double*)calloc(n + 1, sizeof(double));
  for (int k = 1; k <= n - 1; k++)
    for (int i = k + 1; i <= n; i++) {
      double coeff = A_(i, k) / A_(k, k);
      L_(i, k) = coeff;
      for (int j = k + 1; j <= n; j++) A_(i, j) = A_(i, j) - coeff * A_(k, j);
    }
  for (int i = 1; i <= n; i++) L_(i, i) = 1.0;
  for (int j = 1; j <= n; j++)
    for (int i = 1; i <= j; i++) U_(i, j) = A_(i, j);
  for (int k = 1; k <= n; k++) {
    bv[k] = 1.0;
    d[1] = bv[1];
    for (int i = 2; i <= n; i++) {
      d[i] = bv[i];
      for (int j = 1; j <= i - 1; j++) d[i] = d[i] - L_(i, j) * d[j];
    }
    x[n] = d[n] / U_(n, n);
    for (int i = n - 1; i >= 1; i--) {
      x[i] = d[i];
      for (int j = n; j >= i + 1; j--) x[i] = x[i] - U_(i, j) * x[j];
      x[i] = x[i] / U_(i, i);
    }
    for (int i = 1; i <= n; i++) cinv[(k - 1) * n + (i - 1)] = x[i];
    bv[k] = 0.0;
  }
  free(Lm); free(Um); free(bv); free(d); free(x);
#undef A_
#undef L_
#undef U_
}

/* ExplicitBlockEvpPre :2508-2616.  cc, ne: (n,m) sub-block copies with leading dimension EVP_LD; rinv: (nm,nm)
   stored with leading dimension EVP_L.  Returns max|rinv*rin - I| (the reference's self-check, bound 1e-8). */
static double evp_block_pre(const double* cc, const double* ne, double* rinv, int n, int m) {
  const int nm = n + m - 5;
  double y[EVP_LD * EVP_LD];
  double *rin = (double*)calloc((size_t)nm * nm, sizeof(double)), *work = (double*)calloc((size_t)nm * nm, sizeof(double)),
         *rtmp = (double*)calloc((size_t)nm * nm, sizeof(double));
#define RIN(i, j) rin[((j)-1) * nm + ((i)-1)]
  memset(y, 0, sizeof(y));
  for (int pass = 0; pass < 2; pass++) {
    const int cnt = pass == 0 ? m - 2 : n - 3;
    for (int ii = 1; ii <= cnt; ii++) {
      if (pass == 0) F2(y, 2, m - ii) = 1.0; /* from the west edge */
      else F2(y, ii + 2, 2) = 1.0;           /* from the south edge */
      for (int j = 2; j <= m - 1; j++)
        for (int i = 2; i <= n - 1; i++)
          F2(y, i + 1, j + 1) = (-F2(cc, i, j) * F2(y, i, j) - F2(ne, i, j - 1) * F2(y, i + 1, j - 1) -
                                 F2(ne, i - 1, j) * F2(y, i - 1, j + 1) - F2(ne, i - 1, j - 1) * F2(y, i - 1, j - 1)) /
                                F2(ne, i, j);
      const int row = pass == 0 ? ii : m - 2 + ii;
      for (int i = 1; i <= n - 2; i++) RIN(row, i) = -F2(y, i + 2, m);
      for (int j = 1; j <= m - 3; j++) RIN(row, n - 2 + j) = -F2(y, n, m - j);
      if (pass == 0) F2(y, 2, m - ii) = 0.0;
      else F2(y, ii + 2, 2) = 0.0;
    }
  }
  memcpy(work, rin, sizeof(double) * nm * nm);
  evp_inverse(work, rtmp, nm);
  double maxvalr = 0.0;
  for (int j = 1; j <= nm; j++)
    for (int i = 1; i <= nm; i++) {
      double w = 0.0;
      for (int k = 1; k <= nm; k++) w = w + rtmp[(k - 1) * nm + (i - 1)] * RIN(k, j);
      if (i == j) w = w - 1.0;
      if (fabs(w) > maxvalr) maxvalr = fabs(w);
    }
  for (int j = 1; j <= nm; j++)
    for (int k = 1; k <= nm; k++) RV(rinv, k, j) = rtmp[(j - 1) * nm + (k - 1)];
  free(rin); free(work); free(rtmp);
#undef RIN
  return maxvalr;
}

static double g_evp_selfcheck = 0.0; /* largest max|rinv*rin - I| over all ocean sub-blocks of the last prep */
double oracle_evp_selfcheck(void) { return g_evp_selfcheck; }
int oracle_evp_counts(int* nsub, int* nland) {
  if (!E.ready) return -1;
  int nl = 0;
  for (int q = 0; q < E.nsub * NB; q++) nl += E.land[q];
  *nsub = E.nsub * NB;
  *nland = nl;
  return 0;
}

/* POP_SolversInit :1056-1081 (allocation) + the EVP part of POP_SolversPrep :252-292 with EvpPre :2434-2506 */
static int evp_prep(void) {
  if (!E.ready) {
    evp_partition(M.nxb - 2, EVP_XBS, &E.xnb, &E.xidx);
    evp_partition(M.nyb - 2, EVP_YBS, &E.ynb, &E.yidx);
    E.nsub = E.xnb * E.ynb;
    size_t ns = (size_t)E.nsub * NB;
    E.land = (int*)calloc(ns, sizeof(int));
    E.cc = (double*)calloc(ns * EVP_LD * EVP_LD, sizeof(double));
    E.icc = (double*)calloc(ns * EVP_LD * EVP_LD, sizeof(double));
    E.ne = (double*)calloc(ns * EVP_LD * EVP_LD, sizeof(double));
    E.ine = (double*)calloc(ns * EVP_LD * EVP_LD, sizeof(double));
    E.rinv = (double*)calloc(ns * EVP_L * EVP_L, sizeof(double));
    E.ready = 1;
  }
  size_t ns = (size_t)E.nsub * NB;
  memset(E.cc, 0, sizeof(double) * ns * EVP_LD * EVP_LD);
  memset(E.icc, 0, sizeof(double) * ns * EVP_LD * EVP_LD);
  memset(E.ne, 0, sizeof(double) * ns * EVP_LD * EVP_LD);
  memset(E.ine, 0, sizeof(double) * ns * EVP_LD * EVP_LD);
  double worst = 0.0;
  int bad = 0;
  _Pragma("omp parallel for schedule(dynamic) reduction(max : worst) reduction(+ : bad)")
  for (int b = 0; b < NB; b++) {
    /* cc = btropWgtCenter(2:nx1, 2:ny1), ne = btropWgtNE(2:nx1, 2:ny1): index i of cc is array index i+1 */
    const double *C = B2(M.btropWgtCenter, b), *NE = B2(M.btropWgtNE, b);
    for (int j = 1; j <= E.ynb; j++) {
      const int js = E.yidx[j] - 1, je = E.yidx[j + 1], lm = je - js + 1;
      for (int i = 1; i <= E.xnb; i++) {
        const int is = E.xidx[i] - 1, ie = E.xidx[i + 1], ln = ie - is + 1;
        const int ib = (j - 1) * E.xnb + i - 1;
        double *ecc = EV2(E.cc, b, ib), *ene = EV2(E.ne, b, ib);
        for (int jj = 1; jj <= lm; jj++)
          for (int ii = 1; ii <= ln; ii++) {
            F2(ecc, ii, jj) = C[IX2(is + ii - 1 + 1, js + jj - 1 + 1)];
            F2(ene, ii, jj) = NE[IX2(is + ii - 1 + 1, js + jj - 1 + 1)];
          }
        int iland = 0;
        double mn = fabs(F2(ene, 2, 2));
        for (int jj = 2; jj <= lm - 1; jj++)
          for (int ii = 2; ii <= ln - 1; ii++)
            if (fabs(F2(ene, ii, jj)) < mn) mn = fabs(F2(ene, ii, jj));
        if (mn == 0.0) iland = 1;
        if (is + 2 < M.ib[b] || ie - 1 > M.ie[b]) iland = 2;
        if (js + 2 < M.jb[b] || je - 1 > M.je[b]) iland = 3;
        double* rinv = EVR(b, ib);
        memset(rinv, 0, sizeof(double) * EVP_L * EVP_L);
        if (iland > 0) {
          E.land[(size_t)b * E.nsub + ib] = 1;
        } else {
          E.land[(size_t)b * E.nsub + ib] = 0;
          double chk = evp_block_pre(ecc, ene, rinv, ln, lm);
          if (chk > worst) worst = chk;
          if (chk > 1.0e-8) bad++; /* 'error in computing the inverse, error > 1.0e-8; Check EVP sub-block size!' */
        }
      }
    }
  }
  g_evp_selfcheck = worst;
  /* save inverses of the coefficients (:277-283) */
  for (size_t q = 0; q < ns * EVP_LD * EVP_LD; q++) {
    if (E.cc[q] != 0.0) E.icc[q] = 1.0 / E.cc[q];
    if (E.ne[q] != 0.0) E.ine[q] = 1.0 / E.ne[q];
  }
  return bad ? -1 : 0;
}

/* ExplicitEvp :2618-2696 on one sub-block: tu (n-2, m-2) solution written into PX; f (n,m) with leading dim EVP_LD */
static void evp_explicit(const double* cc, const double* ne, const double* ine, const double* rinv, const double* f,
                         double* y, int n, int m) {
  const int nm = n + m - 5;
  double r[EVP_L];
  /* y(:,:) = 0 and y(2:n-1,2:m-1) = tu = 0 on entry (PX was zeroed): the caller passes a zeroed y */
  for (int j = 2; j <= m - 1; j++)
    for (int i = 2; i <= n - 1; i++)
      F2(y, i + 1, j + 1) = (F2(f, i, j) - F2(cc, i, j) * F2(y, i, j) - F2(ne, i, j - 1) * F2(y, i + 1, j - 1) -
                             F2(ne, i - 1, j) * F2(y, i - 1, j + 1) - F2(ne, i - 1, j - 1) * F2(y, i - 1, j - 1)) *
                            F2(ine, i, j);
  for (int i = 1; i <= n - 2; i++) r[i - 1] = F2(y, i + 2, m);
  for (int q = n - 1, jj = m - 1; q <= nm; q++, jj--) r[q - 1] = F2(y, n, jj);
  for (int j = 1; j <= m - 2; j++)
    for (int k = 1; k <= nm; k++) F2(y, 2, m - j) = F2(y, 2, m - j) + RV(rinv, k, j) * r[k - 1];
  for (int i = 1; i <= n - 3; i++)
    for (int k = 1; k <= nm; k++) F2(y, i + 2, 2) = F2(y, i + 2, 2) + RV(rinv, k, m - 2 + i) * r[k - 1];
  for (int j = 2; j <= m - 2; j++)
    for (int i = 2; i <= n - 2; i++)
      F2(y, i + 1, j + 1) = (F2(f, i, j) - F2(cc, i, j) * F2(y, i, j) - F2(ne, i, j - 1) * F2(y, i + 1, j - 1) -
                             F2(ne, i - 1, j) * F2(y, i - 1, j + 1) - F2(ne, i - 1, j - 1) * F2(y, i - 1, j - 1)) *
                            F2(ine, i, j);
}

/* preconditioner :2273-2369 (EVP branch): PX(:,:,bid) for one block */
void o_preconditioner(double* PXv, const double* Xv, int b) {
  double* PX = B2(PXv, b);
  const double* X = B2(Xv, b);
  memset(PX, 0, sizeof(double) * M.n2);
  for (int j = 1; j <= E.ynb; j++) {
    const int js = E.yidx[j], je = E.yidx[j + 1] + 1, lm = je - js + 1;
    for (int i = 1; i <= E.xnb; i++) {
      const int is = E.xidx[i], ie = E.xidx[i + 1] + 1, ln = ie - is + 1;
      const int ib = (j - 1) * E.xnb + i - 1;
      double f[EVP_LD * EVP_LD], y[EVP_LD * EVP_LD];
      memset(f, 0, sizeof(f));
      for (int jj = 2; jj <= lm - 1; jj++)
        for (int ii = 2; ii <= ln - 1; ii++) F2(f, ii, jj) = X[IX2(is + ii - 1, js + jj - 1)];
      if (E.land[(size_t)b * E.nsub + ib] == 1) {
        const double* icc = EV2(E.icc, b, ib);
        for (int jj = 2; jj <= lm - 1; jj++)
          for (int ii = 2; ii <= ln - 1; ii++) PX[IX2(is + ii - 1, js + jj - 1)] = F2(f, ii, jj) * F2(icc, ii, jj);
      } else {
        memset(y, 0, sizeof(y));
        evp_explicit(EV2(E.cc, b, ib), EV2(E.ne, b, ib), EV2(E.ine, b, ib), EVR(b, ib), f, y, ln, lm);
        for (int jj = 2; jj <= lm - 1; jj++)
          for (int ii = 2; ii <= ln - 1; ii++) PX[IX2(is + ii - 1, js + jj - 1)] = F2(y, ii, jj);
      }
    }
  }
}
static int use_evp(void) { return M.cfg.preconditioner_choice == POP_PRECOND_EVP; }

/* pcg :1200-1503 */
static int pcg(double* X, const double* B) {
  const int maxIt = M.cfg.max_iterations, freq = M.cfg.convergence_check_freq;
  double *R = vec(), *S = vec(), *Q = vec(), *work0 = vec(), *work1 = vec();
  double eta0, eta1, rr = 0.0;
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (int b = 0; b < NB; b++) {
    o_btrop_operator(S, X, b);
    for (size_t q = 0; q < M.n2; q++) {
      B2(R, b)[q] = B2(B, b)[q] - B2(S, b)[q];
      B2(S, b)[q] = 0.0;
    }
  }
  halo(R);
  eta0 = 1.0;
  M.numIterations = maxIt;
  for (int m = 1; m <= maxIt; m++) {
    if (use_evp()) { /* :1310-1345: work1 = (PC)r, halo-updated before it enters S */
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (int b = 0; b < NB; b++) {
        o_preconditioner(work1, R, b);
        for (size_t q = 0; q < M.n2; q++) B2(work0, b)[q] = B2(R, b)[q] * B2(work1, b)[q];
      }
      halo(work1);
    } else {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (size_t q = 0; q < NTOT; q++) {
        work1[q] = (M.btropWgtCenter[q] != 0.0) ? R[q] / M.btropWgtCenter[q] : 0.0;
        work0[q] = R[q] * work1[q];
      }
    }
    eta1 = gsum(work0);
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) {
      for (size_t q = 0; q < M.n2; q++) B2(S, b)[q] = B2(work1, b)[q] + B2(S, b)[q] * (eta1 / eta0);
      o_btrop_operator(Q, S, b);
      for (size_t q = 0; q < M.n2; q++) B2(work0, b)[q] = B2(Q, b)[q] * B2(S, b)[q];
    }
    halo(Q);
    eta0 = eta1;
    eta1 = eta0 / gsum(work0);
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) {
      for (size_t q = 0; q < M.n2; q++) {
        B2(X, b)[q] = B2(X, b)[q] + eta1 * B2(S, b)[q];
        B2(R, b)[q] = B2(R, b)[q] - eta1 * B2(Q, b)[q];
      }
      if (m % freq == 0) {
        o_btrop_operator(R, X, b);
        for (size_t q = 0; q < M.n2; q++) {
          B2(R, b)[q] = B2(B, b)[q] - B2(R, b)[q];
          B2(work0, b)[q] = B2(R, b)[q] * B2(R, b)[q];
        }
      }
    }
    if (m % freq == 0) {
      halo(R);
      rr = gsum(work0);
      if (rr < M.convergenceCriterion) {
        M.numIterations = m;
        break;
      }
    }
  }
  M.rmsResidual = sqrt(rr * M.residualNorm);
  free(R); free(S); free(Q); free(work0); free(work1);
  if (M.numIterations == maxIt && M.convergenceCriterion != 0.0) return -1;
  return 0;
}

/* ChronGear :1841-2266 */
static int chrongear(double* X, const double* B) {
  const int maxIt = M.cfg.max_iterations, freq = M.cfg.convergence_check_freq;
  double *R = vec(), *S = vec(), *Q = vec(), *Z = vec(), *AZ = vec(), *WORK0 = vec(), *A0R = vec();
  double* WORKN = (double*)calloc(NTOT * 2, sizeof(double)); /* (nx,ny,2,nblocks) */
  double sumN[2], cgAlpha, cgBeta, cgSigma, cgDelta, cgRhoOld, cgRho = 1.0, rr = 0.0;
  (void)cgRho;
  M.numIterations = maxIt;
  set_a0r(A0R);
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (int b = 0; b < NB; b++) {
    o_btrop_operator(S, X, b);
    for (size_t q = 0; q < M.n2; q++) B2(R, b)[q] = B2(B, b)[q] - B2(S, b)[q];
  }
  halo(R);
  if (use_evp()) { /* :2009-2033: Z = (PC)r, then a halo update of Z */
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) o_preconditioner(Z, R, b);
    halo(Z);
  } else {
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) Z[q] = R[q] * A0R[q];
  }
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (int b = 0; b < NB; b++) {
    double *W1 = WORKN + ((size_t)b * 2) * M.n2, *W2 = W1 + M.n2;
    for (size_t q = 0; q < M.n2; q++) {
      W1[q] = B2(R, b)[q] * B2(Z, b)[q];
      B2(S, b)[q] = B2(Z, b)[q];
    }
    o_btrop_operator(Q, S, b);
    for (size_t q = 0; q < M.n2; q++) W2[q] = B2(S, b)[q] * B2(Q, b)[q];
  }
  halo(Q);
  oracle_global_sum_n(WORKN, 2, POP_LOC_CENTER, M.mMaskTropic, sumN);
  cgRhoOld = sumN[0];
  cgSigma = sumN[1];
  cgAlpha = cgRhoOld / cgSigma;
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (size_t q = 0; q < NTOT; q++) {
    X[q] = X[q] + cgAlpha * S[q];
    R[q] = R[q] - cgAlpha * Q[q];
  }
  for (int m = 1; m <= maxIt; m++) {
    if (use_evp()) {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (int b = 0; b < NB; b++) o_preconditioner(Z, R, b);
    } else {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (size_t q = 0; q < NTOT; q++) Z[q] = R[q] * A0R[q];
    }
    halo(Z);
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) {
      double *W1 = WORKN + ((size_t)b * 2) * M.n2, *W2 = W1 + M.n2;
      o_btrop_operator(AZ, Z, b);
      for (size_t q = 0; q < M.n2; q++) {
        W1[q] = B2(R, b)[q] * B2(Z, b)[q];
        W2[q] = B2(AZ, b)[q] * B2(Z, b)[q];
      }
    }
    oracle_global_sum_n(WORKN, 2, POP_LOC_CENTER, M.mMaskTropic, sumN);
    cgRho = sumN[0];
    cgDelta = sumN[1];
    cgBeta = cgRho / cgRhoOld;
    cgSigma = cgDelta - (cgBeta * cgBeta) * cgSigma;
    cgAlpha = cgRho / cgSigma;
    cgRhoOld = cgRho;
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) {
      for (size_t q = 0; q < M.n2; q++) {
        B2(S, b)[q] = B2(Z, b)[q] + cgBeta * B2(S, b)[q];
        B2(Q, b)[q] = B2(AZ, b)[q] + cgBeta * B2(Q, b)[q];
        B2(X, b)[q] = B2(X, b)[q] + cgAlpha * B2(S, b)[q];
        B2(R, b)[q] = B2(R, b)[q] - cgAlpha * B2(Q, b)[q];
      }
      if (m % freq == 0) {
        o_btrop_operator(Z, X, b);
        for (size_t q = 0; q < M.n2; q++) {
          B2(R, b)[q] = B2(B, b)[q] - B2(Z, b)[q];
          B2(WORK0, b)[q] = B2(R, b)[q] * B2(R, b)[q];
        }
      }
    }
    if (m % freq == 0) {
      halo(R);
      rr = gsum(WORK0);
      if (rr < M.convergenceCriterion) {
        M.numIterations = m;
        break;
      }
    }
  }
  M.rmsResidual = sqrt(rr * M.residualNorm);
  free(R); free(S); free(Q); free(Z); free(AZ); free(WORK0); free(A0R); free(WORKN);
  if (M.numIterations == maxIt && M.convergenceCriterion != 0.0) return -1;
  return 0;
}

/* PCSI :1510-1835 */
static int pcsi(double* X, const double* B) {
  const int maxIt = M.cfg.max_iterations, freq = M.cfg.convergence_check_freq,
            start = M.cfg.convergence_check_start;
  double *R = vec(), *S = vec(), *Q = vec(), *work0 = vec(), *A0R = vec();
  double rr = 0.0;
  set_a0r(A0R);
  double csalpha = 2.0 / (M.PcsiMaxEigs - M.PcsiMinEigs);
  double csbeta = (M.PcsiMaxEigs + M.PcsiMinEigs) / (M.PcsiMaxEigs - M.PcsiMinEigs);
  double csy = csbeta / csalpha;
  double csomga = 2.0 / csy;
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (int b = 0; b < NB; b++) {
    o_btrop_operator(S, X, b);
    for (size_t q = 0; q < M.n2; q++) B2(R, b)[q] = B2(B, b)[q] - B2(S, b)[q];
  }
  if (use_evp()) { /* :1651-1659: R = (PC)R (zero outside the sub-block interiors) */
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) {
      o_preconditioner(work0, R, b);
      memcpy(B2(R, b), B2(work0, b), sizeof(double) * M.n2);
    }
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) Q[q] = (1.0 / csy) * R[q];
  } else {
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) {
      R[q] = R[q] * A0R[q];
      Q[q] = (1.0 / csy) * R[q];
    }
  }
  halo(Q);
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (int b = 0; b < NB; b++) {
    for (size_t q = 0; q < M.n2; q++) B2(X, b)[q] = B2(X, b)[q] + B2(Q, b)[q];
    o_btrop_operator(S, X, b);
    for (size_t q = 0; q < M.n2; q++) B2(R, b)[q] = B2(B, b)[q] - B2(S, b)[q];
  }
  halo(R);
  M.numIterations = maxIt;
  for (int m = 1; m <= maxIt; m++) {
    csomga = 1.0 / (csy - csomga / (4.0 * csalpha * csalpha));
    if (use_evp()) { /* :1734-1742 */
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (int b = 0; b < NB; b++) {
        o_preconditioner(S, R, b); /* S is free here: it is recomputed as A X below */
        memcpy(B2(R, b), B2(S, b), sizeof(double) * M.n2);
      }
    } else {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (size_t q = 0; q < NTOT; q++) R[q] = R[q] * A0R[q];
    }
    halo(R);
    int check = (m % freq == 0) && (m >= start);
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) {
      for (size_t q = 0; q < M.n2; q++) {
        B2(Q, b)[q] = csomga * B2(R, b)[q] + (csy * csomga - 1.0) * B2(Q, b)[q];
        B2(X, b)[q] = B2(X, b)[q] + B2(Q, b)[q];
      }
      o_btrop_operator(S, X, b);
      for (size_t q = 0; q < M.n2; q++) B2(R, b)[q] = B2(B, b)[q] - B2(S, b)[q];
      if (check)
        for (size_t q = 0; q < M.n2; q++) B2(work0, b)[q] = B2(R, b)[q] * B2(R, b)[q];
    }
    if (check) {
      rr = gsum(work0);
      if (rr < M.convergenceCriterion) {
        M.numIterations = m;
        break;
      }
    }
  }
  M.rmsResidual = sqrt(rr * M.residualNorm);
  free(R); free(S); free(Q); free(work0); free(A0R);
  return 0; /* PCSI returns silently when not converged (:1828-1830) */
}

/* ratqr :3122-3222: smallest eigenvalue of a symmetric tridiagonal matrix */
static int ratqr(int n, double eps1, const double* d, const double* e, double* mineig) {
  double *bd = (double*)calloc(n + 1, sizeof(double)), *w = (double*)calloc(n + 1, sizeof(double));
  double f, ep, delta = 0.0, err = 0.0, p, q = 0.0, qp, r, s = 0.0, tot;
  for (int i = 1; i <= n; i++) w[i] = d[i - 1];
  tot = w[1];
  for (int i = 1; i <= n; i++) {
    p = q;
    bd[i] = e[i - 1] * e[i - 1];
    q = 0.0;
    if (i != n) q = fabs(e[i]);
    double c = w[i] - p - q;
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
      int i = n - ii;
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
    if (tot + s <= tot) {
      free(bd); free(w);
      return 1;
    }
  }
  w[1] = tot;
  err = err + fabs(delta);
  bd[1] = err;
  *mineig = w[1];
  free(bd); free(w);
  return 0;
}

/* PcsiLanczos :2699-2990 (diagonal preconditioner) */
static int pcsi_lanczos(void) {
  const int maxstep = M.cfg.max_lanczos_step;
  double *R = vec(), *S = vec(), *Q = vec(), *Q1 = vec(), *P = vec(), *A0R = vec(), *WORK = vec(),
         *WORK1 = vec();
  double* vcsa = (double*)calloc(maxstep + 1, sizeof(double));
  double* vcsb = (double*)calloc(maxstep + 1, sizeof(double));
  double csa, csb, csc, mineig, u, v;
  int rc = 0;
  set_a0r(A0R);
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (size_t q = 0; q < NTOT; q++) { R[q] = 1.0; Q[q] = 0.0; Q1[q] = 0.0; }
  if (use_evp()) { /* :2797-2799 */
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) o_preconditioner(S, R, b);
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) WORK[q] = S[q] * R[q];
  } else {
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) { S[q] = R[q] * A0R[q]; WORK[q] = S[q] * R[q]; }
  }
  csc = -gsum(WORK);
  if (csc > 0.0) {
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) Q[q] = (1 / sqrt(csc)) * R[q];
  } else {
    rc = -1;
    goto done;
  }
  halo(Q);
  csb = 0.0; u = 0.0; v = 0.0; mineig = 1.0;
  M.lanczos_steps = 0;
  for (int m = 1; m <= maxstep; m++) {
    M.lanczos_steps = m;
    if (use_evp()) {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (int b = 0; b < NB; b++) o_preconditioner(P, Q, b);
    } else {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (size_t q = 0; q < NTOT; q++) P[q] = Q[q] * A0R[q];
    }
    halo(P);
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (int b = 0; b < NB; b++) {
      o_btrop_operator(WORK1, P, b);
      for (size_t q = 0; q < M.n2; q++) {
        B2(R, b)[q] = B2(WORK1, b)[q] - csb * B2(Q1, b)[q];
        B2(WORK, b)[q] = B2(P, b)[q] * B2(R, b)[q];
      }
    }
    csa = -gsum(WORK);
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) R[q] = R[q] - csa * Q[q];
    halo(R);
    if (use_evp()) {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (int b = 0; b < NB; b++) o_preconditioner(S, R, b);
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (size_t q = 0; q < NTOT; q++) WORK[q] = S[q] * R[q];
    } else {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (size_t q = 0; q < NTOT; q++) { S[q] = R[q] * A0R[q]; WORK[q] = S[q] * R[q]; }
    }
    csc = -gsum(WORK);
    csb = sqrt(csc);
    vcsa[m] = csa;
    vcsb[m] = csb;
    if (m == 1) u = vcsa[1] + vcsb[1];
    else {
      double c = vcsa[m] + vcsb[m] + vcsb[m - 1];
      u = (u > c) ? u : c;
    }
    if (csb != 0.0) {
      _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
      for (size_t q = 0; q < NTOT; q++) { Q1[q] = Q[q]; Q[q] = (1 / csb) * R[q]; }
    } else {
      rc = -1;
      goto done;
    }
    if (m % 10 == 0 || m == maxstep) {
      double *mcsa = (double*)calloc(m + 1, sizeof(double)), *mcsb = (double*)calloc(m + 1, sizeof(double));
      for (int i = 1; i <= m - 1; i++) { mcsa[i - 1] = vcsa[i]; mcsb[i] = vcsb[i]; }
      mcsa[m - 1] = vcsa[m];
      mcsb[0] = 0.0;
      int info = ratqr(m, 1.0e-8, mcsa, mcsb, &v);
      free(mcsa); free(mcsb);
      if (info != 0) { rc = -1; goto done; }
      if (fabs(1 - v / mineig) < M.cfg.lanczos_convergence_criterion) break;
      mineig = v;
    }
  }
  M.PcsiMaxEigs = u;
  M.PcsiMinEigs = v;
done:
  free(R); free(S); free(Q); free(Q1); free(P); free(A0R); free(WORK); free(WORK1);
  free(vcsa); free(vcsb);
  return rc;
}

/* POP_SolversPrep :165-320: btropWgtCenter <- centerWgtClinic, then Lanczos for PCSI */
int o_solvers_prep(void) {
  memcpy(M.btropWgtCenter, M.centerWgtClinic, sizeof(double) * NTOT);
  if (use_evp() && evp_prep() != 0) return -1; /* :252-292, before the Lanczos estimate that uses it */
  if (M.cfg.solver_choice == POP_SOLVER_PCSI) return pcsi_lanczos();
  return 0;
}

/* POP_SolversRun :327-495 */
int o_solvers_run(double* sfcPressure, const double* rhs) {
  double t0 = o_now();
  int rc;
  memcpy(M.btropWgtCenter, M.centerWgtClinic, sizeof(double) * NTOT);
  switch (M.cfg.solver_choice) {
    case POP_SOLVER_PCG: rc = pcg(sfcPressure, rhs); break;
    case POP_SOLVER_PCSI: rc = pcsi(sfcPressure, rhs); break;
    default: rc = chrongear(sfcPressure, rhs); break;
  }
  M.timer[OT_SOLVER] += o_now() - t0;
  return rc;
}

int oracle_num_iterations(void) { return M.numIterations; }
double oracle_rms_residual(void) { return M.rmsResidual; }
void oracle_get_eigs(double* mn, double* mx) { *mn = M.PcsiMinEigs; *mx = M.PcsiMaxEigs; }
