/*
 * o_solver.c -- ORACLE (test infrastructure): barotropic elliptic solver.
 * Restates source/POP_SolversMod.F90: POP_SolversInit :783-820,895-906; POP_SolversDiagonal
 * :1110-1151; POP_SolversRun :327-495 (clinic == tropic distribution, redistribution = copy);
 * pcg :1200-1503; PCSI :1510-1835; ChronGear :1841-2266; btropOperator :2376-2431;
 * PcsiLanczos :2699-2990; ratqr :3122-3222.  Diagonal preconditioner only.
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
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) {
      work1[q] = (M.btropWgtCenter[q] != 0.0) ? R[q] / M.btropWgtCenter[q] : 0.0;
      work0[q] = R[q] * work1[q];
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
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (size_t q = 0; q < NTOT; q++) Z[q] = R[q] * A0R[q];
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
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) Z[q] = R[q] * A0R[q];
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
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (size_t q = 0; q < NTOT; q++) {
    R[q] = R[q] * A0R[q];
    Q[q] = (1.0 / csy) * R[q];
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
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) R[q] = R[q] * A0R[q];
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
  _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
  for (size_t q = 0; q < NTOT; q++) { S[q] = R[q] * A0R[q]; WORK[q] = S[q] * R[q]; }
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
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) P[q] = Q[q] * A0R[q];
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
    _Pragma("omp parallel for schedule(static) if (NTOT > 100000)")
    for (size_t q = 0; q < NTOT; q++) { S[q] = R[q] * A0R[q]; WORK[q] = S[q] * R[q]; }
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
