/*
 * o_grid.c -- ORACLE (test infrastructure): grid-derived fields and the init-time coefficient
 * tables of the hot path.  Restates source/grid.F90:587-647 (reciprocals/areas, closed-boundary
 * fix-up), :786-803 (vertical grid), :978-1041 (KMU, HU/HUR), :1131-1139 (uarea_equator),
 * :2537-2596 (landmasks), :2882-2932 (cf_area_avg); source/hmix_del2.F90:97-421,428-663;
 * source/hmix_del4.F90:94-387,394-590; source/advection.F90:387-396,420-562;
 * source/state_mod.F90:1040-1063,1765-1766; source/pressure_grad.F90:168-175;
 * source/vertical_mix.F90:383-448; source/time_management.F90:391-439, step_mod.F90:302-320.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pop_oracle.h"

void* o_alloc_d(size_t n);
void* o_alloc_i(size_t n);
#define D2ALLOC() ((double*)o_alloc_d(M.n2 * M.nblocks))
#define I2ALLOC() ((int*)o_alloc_i(M.n2 * M.nblocks))

/* eoshift(A,dim,shift): end-off shift with zero fill */
static inline double eo(const double* a, int i, int j) {
  return (i < 1 || i > M.nxb || j < 1 || j > M.nyb) ? 0.0 : a[IX2(i, j)];
}
static inline int eoi(const int* a, int i, int j) {
  return (i < 1 || i > M.nxb || j < 1 || j > M.nyb) ? 0 : a[IX2(i, j)];
}

static void vertical_grid(const double* dz_in) {
  int km = M.km;
  M.dz = o_alloc_d(km + 2); M.dzw = o_alloc_d(km + 2); M.dzr = o_alloc_d(km + 2);
  M.dz2r = o_alloc_d(km + 2); M.dzwr = o_alloc_d(km + 2); M.c2dz = o_alloc_d(km + 2);
  M.zt = o_alloc_d(km + 2); M.zw = o_alloc_d(km + 2);
  for (int k = 1; k <= km; k++) M.dz[k] = dz_in[k - 1];
  /* grid.F90:786-803 */
  M.dzw[0] = 0.5 * M.dz[1];
  M.dzw[km] = 0.5 * M.dz[km];
  M.dzwr[0] = 1.0 / M.dzw[0];
  M.zw[1] = M.dz[1];
  M.zt[1] = M.dzw[0];
  for (int k = 1; k <= km - 1; k++) {
    M.dzw[k] = 0.5 * (M.dz[k] + M.dz[k + 1]);
    M.zw[k + 1] = M.zw[k] + M.dz[k + 1];
    M.zt[k + 1] = M.zt[k] + M.dzw[k];
  }
  for (int k = 1; k <= km; k++) {
    M.c2dz[k] = 2.0 * M.dz[k];
    M.dzr[k] = 1.0 / M.dz[k];
    M.dz2r[k] = 1.0 / M.c2dz[k];
    M.dzwr[k] = 1.0 / M.dzw[k];
  }
}

/* state_mod.F90:1765-1766 */
static double pressure(double depth) {
  return 0.059808 * (exp(-0.025 * depth) - 1.0) + 0.100766 * depth + 2.28405e-7 * (depth * depth);
}

static void init_state_tables(void) {
  int km = M.km;
  M.pressz = o_alloc_d(km + 2); M.tmin = o_alloc_d(km + 2); M.tmax = o_alloc_d(km + 2);
  M.smin = o_alloc_d(km + 2); M.smax = o_alloc_d(km + 2); M.bouss = o_alloc_d(km + 2);
  for (int k = 1; k <= km; k++) {
    M.pressz[k] = pressure(M.zt[k] * 0.01); /* mpercm */
    if (M.cfg.state_itype == POP_STATE_MWJF) { /* state_mod.F90:1060-1063 */
      M.tmin[k] = -2.0; M.tmax[k] = 999.0; M.smin[k] = 0.0; M.smax[k] = 0.999;
    } else { /* linear: :1144-1147 */
      M.tmin[k] = -2.0; M.tmax[k] = 40.0; M.smin[k] = 0.0; M.smax[k] = 0.042;
    }
    /* pressure_grad.F90:168-175 */
    if (M.cfg.lbouss_correct)
      M.bouss[k] = 1.0 / (1.02819 + 4.4004e-5 * M.pressz[k] - 2.93161e-4 * exp(-0.05 * M.pressz[k]));
    else
      M.bouss[k] = 1.0;
  }
}

/* masked global minval over physical points (POP_GlobalMinval with lMask) */
static double global_minval(const double* a, const double* rmask) {
  double v = 1.0e300;
  for (int b = 0; b < M.nblocks; b++)
    for (int j = M.jb[b]; j <= M.je[b]; j++)
      for (int i = M.ib[b]; i <= M.ie[b]; i++)
        if (B2(rmask, b)[IX2(i, j)] != 0.0 && B2(a, b)[IX2(i, j)] < v) v = B2(a, b)[IX2(i, j)];
  return v;
}

static void init_del2u_or_del4u(void);
static void init_del2t_or_del4t(void);
static void init_advection(void);
static void alloc_state(void);

/* read_bottom_cell (grid.F90:2116-2172): DZBC scattered as a centre scalar */
static double* g_dzbc_global = NULL;
int oracle_set_bottom_cells(const double* DZBC_G) {
  size_t n = (size_t)M.cfg.nx_global * M.cfg.ny_global;
  free(g_dzbc_global);
  g_dzbc_global = (double*)malloc(sizeof(double) * n);
  memcpy(g_dzbc_global, DZBC_G, sizeof(double) * n);
  return 0;
}

int oracle_set_grid(const double* ULAT_G, const double* HTN_G, const double* HTE_G,
                    const double* HUS_G, const double* HUW_G, const double* DXU_G,
                    const double* DYU_G, const double* DXT_G, const double* DYT_G,
                    const int* KMT_G, const double* dz_in) {
  const int nxb = M.nxb, nyb = M.nyb;
  M.TLAT = D2ALLOC();
  M.ULAT = D2ALLOC(); M.HTN = D2ALLOC(); M.HTE = D2ALLOC(); M.HUS = D2ALLOC(); M.HUW = D2ALLOC();
  M.DXU = D2ALLOC(); M.DYU = D2ALLOC(); M.DXT = D2ALLOC(); M.DYT = D2ALLOC();
  M.DXUR = D2ALLOC(); M.DYUR = D2ALLOC(); M.DXTR = D2ALLOC(); M.DYTR = D2ALLOC();
  M.UAREA = D2ALLOC(); M.TAREA = D2ALLOC(); M.UAREA_R = D2ALLOC(); M.TAREA_R = D2ALLOC();
  M.HU = D2ALLOC(); M.HUR = D2ALLOC(); M.HT = D2ALLOC(); M.FCOR = D2ALLOC();
  M.RCALCT = D2ALLOC(); M.RCALCU = D2ALLOC();
  M.AU0 = D2ALLOC(); M.AUN = D2ALLOC(); M.AUE = D2ALLOC(); M.AUNE = D2ALLOC();
  M.AT0 = D2ALLOC(); M.ATS = D2ALLOC(); M.ATW = D2ALLOC(); M.ATSW = D2ALLOC();
  M.KXU = D2ALLOC(); M.KYU = D2ALLOC();
  M.KMT = I2ALLOC(); M.KMU = I2ALLOC(); M.KMTN = I2ALLOC(); M.KMTS = I2ALLOC();
  M.KMTE = I2ALLOC(); M.KMTW = I2ALLOC(); M.KMTEE = I2ALLOC(); M.KMTNN = I2ALLOC();

  /* scatter with the field locations of read_horiz_grid (grid.F90:1420-1520) */
  oracle_scatter_2d(M.ULAT, ULAT_G, POP_LOC_NECORNER, POP_KIND_SCALAR);
  oracle_scatter_2d(M.HTN, HTN_G, POP_LOC_NFACE, POP_KIND_SCALAR);
  oracle_scatter_2d(M.DXU, DXU_G, POP_LOC_NECORNER, POP_KIND_SCALAR);
  oracle_scatter_2d(M.DXT, DXT_G, POP_LOC_CENTER, POP_KIND_SCALAR);
  oracle_scatter_2d(M.HTE, HTE_G, POP_LOC_EFACE, POP_KIND_SCALAR);
  oracle_scatter_2d(M.DYT, DYT_G, POP_LOC_CENTER, POP_KIND_SCALAR);
  oracle_scatter_2d(M.DYU, DYU_G, POP_LOC_NECORNER, POP_KIND_SCALAR);
  oracle_scatter_2d(M.HUS, HUS_G, POP_LOC_EFACE, POP_KIND_SCALAR);
  oracle_scatter_2d(M.HUW, HUW_G, POP_LOC_NFACE, POP_KIND_SCALAR);
  oracle_scatter_2d_i4(M.KMT, KMT_G);
  /* grid.F90:1531-1538 */
  double* fix[8] = {M.HTN, M.HTE, M.HUS, M.HUW, M.DXU, M.DYU, M.DXT, M.DYT};
  for (int f = 0; f < 8; f++)
    for (size_t q = 0; q < M.n2 * M.nblocks; q++)
      if (fix[f][q] <= 0.0) fix[f][q] = 1.0;

  vertical_grid(dz_in);

  /* init_grid2: closed-boundary fix-up and reciprocals, grid.F90:587-647 */
  for (int b = 0; b < M.nblocks; b++) {
    double *DXU = B2(M.DXU, b), *DYU = B2(M.DYU, b), *DXT = B2(M.DXT, b), *DYT = B2(M.DYT, b);
    const int *ig = M.i_glob + (size_t)b * nxb, *jg = M.j_glob + (size_t)b * nyb;
    int ib = M.ib[b], ie = M.ie[b], jb = M.jb[b], je = M.je[b];
    if (ig[0] == 0)
      for (int j = 1; j <= nyb; j++)
        for (int i = 1; i <= ib - 1; i++) {
          DXU[IX2(i, j)] = DXU[IX2(ib, j)]; DYU[IX2(i, j)] = DYU[IX2(ib, j)];
          DXT[IX2(i, j)] = DXT[IX2(ib, j)]; DYT[IX2(i, j)] = DYT[IX2(ib, j)];
        }
    if (ig[ie] == 0) /* i_glob(ie+1) */
      for (int j = 1; j <= nyb; j++)
        for (int i = ie + 1; i <= nxb; i++) {
          DXU[IX2(i, j)] = DXU[IX2(ie, j)]; DYU[IX2(i, j)] = DYU[IX2(ie, j)];
          DXT[IX2(i, j)] = DXT[IX2(ie, j)]; DYT[IX2(i, j)] = DYT[IX2(ie, j)];
        }
    if (jg[0] == 0)
      for (int j = 1; j <= jb - 1; j++)
        for (int i = 1; i <= nxb; i++) {
          DXU[IX2(i, j)] = DXU[IX2(i, jb)]; DYU[IX2(i, j)] = DYU[IX2(i, jb)];
          DXT[IX2(i, j)] = DXT[IX2(i, jb)]; DYT[IX2(i, j)] = DYT[IX2(i, jb)];
        }
    if (jg[je] == 0)
      for (int j = je + 1; j <= nyb; j++)
        for (int i = 1; i <= nxb; i++) {
          DXU[IX2(i, j)] = DXU[IX2(i, je)]; DYU[IX2(i, j)] = DYU[IX2(i, je)];
          DXT[IX2(i, j)] = DXT[IX2(i, je)]; DYT[IX2(i, j)] = DYT[IX2(i, je)];
        }
    for (size_t q = 0; q < M.n2; q++) {
      B2(M.DXUR, b)[q] = 1.0 / DXU[q];
      B2(M.DYUR, b)[q] = 1.0 / DYU[q];
      B2(M.UAREA, b)[q] = DXU[q] * DYU[q];
      B2(M.UAREA_R, b)[q] = 1.0 / B2(M.UAREA, b)[q];
      B2(M.DXTR, b)[q] = 1.0 / DXT[q];
      B2(M.DYTR, b)[q] = 1.0 / DYT[q];
      B2(M.TAREA, b)[q] = DXT[q] * DYT[q];
      B2(M.TAREA_R, b)[q] = 1.0 / B2(M.TAREA, b)[q];
    }
  }
  /* cf_area_avg: grid.F90:2908-2928 */
  for (int b = 0; b < M.nblocks; b++) {
    const double *TA = B2(M.TAREA, b), *UR = B2(M.UAREA_R, b);
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++) {
        size_t q = IX2(i, j);
        B2(M.AT0, b)[q] = 0.25; B2(M.ATS, b)[q] = 0.25; B2(M.ATW, b)[q] = 0.25;
        B2(M.ATSW, b)[q] = 0.25;
        B2(M.AU0, b)[q] = TA[q] * 0.25 * UR[q];
        B2(M.AUN, b)[q] = eo(TA, i, j + 1) * 0.25 * UR[q];
        B2(M.AUE, b)[q] = eo(TA, i + 1, j) * 0.25 * UR[q];
        B2(M.AUNE, b)[q] = eo(TA, i + 1, j + 1) * 0.25 * UR[q];
      }
  }
  /* KMU: grid.F90:978-990 (+ halo update, NEcorner scalar) */
  for (int b = 0; b < M.nblocks; b++) {
    const int* KMT = (const int*)(M.KMT + (size_t)b * M.n2);
    int* KMU = M.KMU + (size_t)b * M.n2;
    for (int j = 1; j <= nyb - 1; j++)
      for (int i = 1; i <= nxb - 1; i++) {
        int m = KMT[IX2(i, j)];
        if (KMT[IX2(i + 1, j)] < m) m = KMT[IX2(i + 1, j)];
        if (KMT[IX2(i, j + 1)] < m) m = KMT[IX2(i, j + 1)];
        if (KMT[IX2(i + 1, j + 1)] < m) m = KMT[IX2(i + 1, j + 1)];
        KMU[IX2(i, j)] = m;
      }
  }
  oracle_halo_2d_i4(M.KMU, POP_LOC_NECORNER, POP_KIND_SCALAR, 0);
  if (PBC) { /* grid.F90:917-966: DZT = dz except the bottom cell, DZU = min of the four surrounding DZT */
    if (!g_dzbc_global) return -1;
    M.DZBC = D2ALLOC();
    oracle_scatter_2d(M.DZBC, g_dzbc_global, POP_LOC_CENTER, POP_KIND_SCALAR);
    size_t n3p = M.n2 * (size_t)(M.km + 2);
    M.DZT = (double*)calloc(n3p * M.nblocks, sizeof(double));
    M.DZU = (double*)calloc(n3p * M.nblocks, sizeof(double));
    for (int b = 0; b < M.nblocks; b++) {
      const int* KMT = M.KMT + (size_t)b * M.n2;
      for (int k = 1; k <= M.km; k++) {
        double *T = DZT3(b, k), *U = DZU3(b, k);
        for (size_t q = 0; q < M.n2; q++) T[q] = (KMT[q] == k) ? B2(M.DZBC, b)[q] : M.dz[k];
        for (int j = 1; j <= nyb - 1; j++)
          for (int i = 1; i <= nxb - 1; i++) {
            double m = T[IX2(i, j)];
            if (T[IX2(i + 1, j)] < m) m = T[IX2(i + 1, j)];
            if (T[IX2(i, j + 1)] < m) m = T[IX2(i, j + 1)];
            if (T[IX2(i + 1, j + 1)] < m) m = T[IX2(i + 1, j + 1)];
            U[IX2(i, j)] = m;
          }
      }
    }
    /* POP_HaloUpdate(DZU, NEcorner, scalar) over the levels 0:km+1; the halo works on [nblocks][nz] slabs */
    oracle_halo_3d(M.DZU, M.km + 2, POP_LOC_NECORNER, POP_KIND_SCALAR, 0.0);
  }
  /* HT, HU, HUR: grid.F90:1001-1041 ; landmasks :2537-2596 */
  for (int b = 0; b < M.nblocks; b++) {
    const int *KMT = M.KMT + (size_t)b * M.n2, *KMU = M.KMU + (size_t)b * M.n2;
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++) {
        size_t q = IX2(i, j);
        int kt = KMT[q], ku = KMU[q];
        if (PBC) { /* :1001-1022 */
          B2(M.HT, b)[q] = (kt >= 1 && kt <= M.km) ? M.zw[kt - 1] + DZT3(b, kt)[q] : 0.0;
          if (ku >= 1 && ku <= M.km) {
            B2(M.HU, b)[q] = M.zw[ku - 1] + DZU3(b, ku)[q];
            B2(M.HUR, b)[q] = 1.0 / B2(M.HU, b)[q];
          } else {
            B2(M.HU, b)[q] = 0.0;
            B2(M.HUR, b)[q] = 0.0;
          }
          for (int k = 1; k <= M.km; k++)
            if (k > ku) DZU3(b, k)[q] = M.dz[k]; /* to prevent divide by zero */
        } else {
          B2(M.HT, b)[q] = (kt >= 1 && kt <= M.km) ? M.zw[kt] : 0.0;
          B2(M.HU, b)[q] = (ku >= 1 && ku <= M.km) ? M.zw[ku] : 0.0;
          B2(M.HUR, b)[q] = (ku >= 1 && ku <= M.km) ? 1.0 / M.zw[ku] : 0.0;
        }
        B2(M.RCALCT, b)[q] = (kt >= 1) ? 1.0 : 0.0;
        B2(M.RCALCU, b)[q] = (ku >= 1) ? 1.0 : 0.0;
        (M.KMTN + (size_t)b * M.n2)[q] = eoi(KMT, i, j + 1);
        (M.KMTS + (size_t)b * M.n2)[q] = eoi(KMT, i, j - 1);
        (M.KMTE + (size_t)b * M.n2)[q] = eoi(KMT, i + 1, j);
        (M.KMTW + (size_t)b * M.n2)[q] = eoi(KMT, i - 1, j);
        (M.KMTEE + (size_t)b * M.n2)[q] = eoi(KMT, i + 2, j);
        (M.KMTNN + (size_t)b * M.n2)[q] = eoi(KMT, i, j + 2);
      }
  }
  /* uarea_equator: grid.F90:1130-1139 ; FCOR :1146 */
  {
    double* W = (double*)malloc(sizeof(double) * M.n2 * M.nblocks);
    for (size_t q = 0; q < M.n2 * M.nblocks; q++) W[q] = fabs(M.ULAT[q]);
    double m = global_minval(W, M.RCALCU);
    for (size_t q = 0; q < M.n2 * M.nblocks; q++) W[q] = (W[q] == m) ? M.UAREA[q] : 1.0e20;
    M.uarea_equator = global_minval(W, M.RCALCU);
    free(W);
    for (size_t q = 0; q < M.n2 * M.nblocks; q++) M.FCOR[q] = 2.0 * O_OMEGA * sin(M.ULAT[q]);
  }
  init_state_tables();
  init_advection();
  init_del2t_or_del4t();
  init_del2u_or_del4u();
  alloc_state();
  if (o_solvers_init() != 0) return -1;
  o_init_barotropic();
  return 0;
}

/* init_advection: advection.F90:387-396 (KXU,KYU), :420-562 (upwind3 tables) */
static void init_advection(void) {
  const int nxb = M.nxb, nyb = M.nyb, km = M.km;
  for (int b = 0; b < M.nblocks; b++) {
    const double *HUW = B2(M.HUW, b), *HUS = B2(M.HUS, b), *UR = B2(M.UAREA_R, b);
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++) {
        B2(M.KXU, b)[IX2(i, j)] = (eo(HUW, i + 1, j) - HUW[IX2(i, j)]) * UR[IX2(i, j)];
        B2(M.KYU, b)[IX2(i, j)] = (eo(HUS, i, j + 1) - HUS[IX2(i, j)]) * UR[IX2(i, j)];
      }
  }
  int use_upwind3 = 0;
  for (int n = 0; n < M.nt; n++)
    if (M.cfg.tadvect_itype[n] == POP_TADVECT_UPWIND3) use_upwind3 = 1;
  M.AUX = o_alloc_d(M.n2 * M.nt * M.nblocks);
  M.use_lw_lim = 0;
  for (int n = 0; n < M.nt; n++)
    if (M.cfg.tadvect_itype[n] == POP_TADVECT_LW_LIM) M.use_lw_lim = 1;
  if (M.use_lw_lim) { /* advection.F90:566-694 */
    const int kk = PBC ? km : 1;
    M.p5_dz_ph_r = o_alloc_d(km + 2);
    M.p5_DXT_ph_R = D2ALLOC(); M.p5_DYT_ph_R = D2ALLOC();
    M.UTE_to_UVEL_E = o_alloc_d(M.n2 * kk * M.nblocks);
    M.VTN_to_VVEL_N = o_alloc_d(M.n2 * kk * M.nblocks);
    M.UTE_jbm2 = o_alloc_d((size_t)nxb * km * M.nblocks); M.WTKB_jbm2 = o_alloc_d((size_t)nxb * km * M.nblocks);
    M.WTKB_jep2 = o_alloc_d((size_t)nxb * km * M.nblocks);
    M.WTKB_ibm2 = o_alloc_d((size_t)nyb * km * M.nblocks); M.WTKB_iep2 = o_alloc_d((size_t)nyb * km * M.nblocks);
    M.FLUX_VEL_prev = o_alloc_d(M.n2 * 5 * M.nblocks);
    for (int k = 1; k <= km - 1; k++) M.p5_dz_ph_r[k] = 1.0 / (M.dz[k] + M.dz[k + 1]);
    M.p5_dz_ph_r[km] = 0.5 / M.dz[km];
    for (int b = 0; b < M.nblocks; b++) {
      const double *DXT = B2(M.DXT, b), *DYT = B2(M.DYT, b), *HTE = B2(M.HTE, b), *HTN = B2(M.HTN, b);
      for (int j = M.jb[b] - 2; j <= M.je[b] + 2; j++)
        for (int i = M.ib[b] - 2; i <= M.ie[b] + 1; i++) {
          B2(M.p5_DXT_ph_R, b)[IX2(i, j)] = 1.0 / (DXT[IX2(i, j)] + DXT[IX2(i + 1, j)]);
          for (int k = 1; k <= kk; k++) {
            double* E = M.UTE_to_UVEL_E + ((size_t)b * kk + (k - 1)) * M.n2;
            if (PBC) E[IX2(i, j)] = 1.0 / HTE[IX2(i, j)] / fmin(DZT3(b, k)[IX2(i, j)], DZT3(b, k)[IX2(i + 1, j)]);
            else E[IX2(i, j)] = 1.0 / HTE[IX2(i, j)];
          }
        }
      for (int j = M.jb[b] - 2; j <= M.je[b] + 1; j++)
        for (int i = M.ib[b]; i <= M.ie[b]; i++) {
          B2(M.p5_DYT_ph_R, b)[IX2(i, j)] = 1.0 / (DYT[IX2(i, j)] + DYT[IX2(i, j + 1)]);
          for (int k = 1; k <= kk; k++) {
            double* N = M.VTN_to_VVEL_N + ((size_t)b * kk + (k - 1)) * M.n2;
            if (PBC) N[IX2(i, j)] = 1.0 / HTN[IX2(i, j)] / fmin(DZT3(b, k)[IX2(i, j)], DZT3(b, k)[IX2(i, j + 1)]);
            else N[IX2(i, j)] = 1.0 / HTN[IX2(i, j)];
          }
        }
    }
  }
  if (!use_upwind3) return;
  M.talfzp = o_alloc_d(km + 2); M.tbetzp = o_alloc_d(km + 2); M.tgamzp = o_alloc_d(km + 2);
  M.talfzm = o_alloc_d(km + 2); M.tbetzm = o_alloc_d(km + 2); M.tdelzm = o_alloc_d(km + 2);
  M.TALFXP = D2ALLOC(); M.TBETXP = D2ALLOC(); M.TGAMXP = D2ALLOC();
  M.TALFYP = D2ALLOC(); M.TBETYP = D2ALLOC(); M.TGAMYP = D2ALLOC();
  M.TALFXM = D2ALLOC(); M.TBETXM = D2ALLOC(); M.TDELXM = D2ALLOC();
  M.TALFYM = D2ALLOC(); M.TBETYM = D2ALLOC(); M.TDELYM = D2ALLOC();
  double* dzc = (double*)calloc(km + 3, sizeof(double)); /* dzc(0:km+1) */
  const double* dz = M.dz;
  dzc[0] = dz[1];
  for (int k = 1; k <= km; k++) dzc[k] = dz[k];
  dzc[km + 1] = dzc[km];
  for (int k = 1; k <= km - 1; k++) {
    M.talfzp[k] = dz[k] * (2.0 * dz[k] + dzc[k - 1]) /
                  ((dz[k] + dz[k + 1]) * (dzc[k - 1] + 2.0 * dz[k] + dz[k + 1]));
    M.tbetzp[k] = dz[k + 1] * (2.0 * dz[k] + dzc[k - 1]) /
                  ((dz[k] + dz[k + 1]) * (dz[k] + dzc[k - 1]));
    M.tgamzp[k] = -(dz[k] * dz[k + 1]) /
                  ((dz[k] + dzc[k - 1]) * (dz[k + 1] + dzc[k - 1] + 2.0 * dz[k]));
  }
  M.tbetzp[1] = M.tbetzp[1] + M.tgamzp[1];
  M.tgamzp[1] = 0.0;
  M.talfzp[km] = 0.0; M.tbetzp[km] = 0.0; M.tgamzp[km] = 0.0;
  for (int k = 1; k <= km - 1; k++) {
    M.talfzm[k] = dz[k] * (2.0 * dz[k + 1] + dzc[k + 2]) /
                  ((dz[k] + dz[k + 1]) * (dz[k + 1] + dzc[k + 2]));
    M.tbetzm[k] = dz[k + 1] * (2.0 * dz[k + 1] + dzc[k + 2]) /
                  ((dz[k] + dz[k + 1]) * (dz[k] + dzc[k + 2] + 2.0 * dz[k + 1]));
    M.tdelzm[k] = -(dz[k] * dz[k + 1]) /
                  ((dz[k + 1] + dzc[k + 2]) * (dz[k] + dzc[k + 2] + 2.0 * dz[k + 1]));
  }
  M.talfzm[km - 1] = M.talfzm[km - 1] + M.tdelzm[km - 1];
  M.tdelzm[km - 1] = 0.0;
  M.talfzm[km] = 0.0; M.tbetzm[km] = 0.0; M.tdelzm[km] = 0.0;
  free(dzc);
  for (int b = 0; b < M.nblocks; b++) {
    const double *DXT = B2(M.DXT, b), *DYT = B2(M.DYT, b);
    for (int j = M.jb[b]; j <= M.je[b]; j++)
      for (int i = M.ib[b] - 1; i <= M.ie[b]; i++) {
        double dxc = DXT[IX2(i, j)], dxcw = DXT[IX2(i - 1, j)], dxce = DXT[IX2(i + 1, j)],
               dxce2 = DXT[IX2(i + 2, j)];
        size_t q = IX2(i, j);
        B2(M.TALFXP, b)[q] = dxc * (2.0 * dxc + dxcw) / ((dxc + dxce) * (dxcw + 2.0 * dxc + dxce));
        B2(M.TBETXP, b)[q] = dxce * (2.0 * dxc + dxcw) / ((dxc + dxcw) * (dxc + dxce));
        B2(M.TGAMXP, b)[q] = -(dxc * dxce) / ((dxc + dxcw) * (dxcw + 2.0 * dxc + dxce));
        B2(M.TALFXM, b)[q] = dxc * (2.0 * dxce + dxce2) / ((dxc + dxce) * (dxce + dxce2));
        B2(M.TBETXM, b)[q] = dxce * (2.0 * dxce + dxce2) / ((dxc + dxce) * (dxc + 2.0 * dxce + dxce2));
        B2(M.TDELXM, b)[q] = -(dxc * dxce) / ((dxce2 + dxce) * (dxc + 2.0 * dxce + dxce2));
      }
    for (int j = M.jb[b] - 1; j <= M.je[b]; j++)
      for (int i = M.ib[b]; i <= M.ie[b]; i++) {
        double dyc = DYT[IX2(i, j)], dycs = DYT[IX2(i, j - 1)], dycn = DYT[IX2(i, j + 1)],
               dycn2 = DYT[IX2(i, j + 2)];
        size_t q = IX2(i, j);
        B2(M.TALFYP, b)[q] = dyc * (2.0 * dyc + dycs) / ((dyc + dycn) * (dycs + 2.0 * dyc + dycn));
        B2(M.TBETYP, b)[q] = dycn * (2.0 * dyc + dycs) / ((dyc + dycn) * (dycs + dyc));
        B2(M.TGAMYP, b)[q] = -(dyc * dycn) / ((dyc + dycs) * (dycs + 2.0 * dyc + dycn));
        B2(M.TALFYM, b)[q] = dyc * (2.0 * dycn + dycn2) / ((dyc + dycn) * (dycn + dycn2));
        B2(M.TBETYM, b)[q] = dycn * (2.0 * dycn + dycn2) / ((dyc + dycn) * (dyc + 2.0 * dycn + dycn2));
        B2(M.TDELYM, b)[q] = -(dyc * dycn) / ((dycn2 + dycn) * (dyc + 2.0 * dycn + dycn2));
      }
  }
}

/* init_del2t (hmix_del2.F90:428-663) / init_del4t (hmix_del4.F90:394-590) */
static void init_del2t_or_del4t(void) {
  const int nxb = M.nxb, nyb = M.nyb;
  const int del4 = (M.cfg.hmix_tracer_itype == POP_HMIX_DEL4);
  M.DTN = D2ALLOC(); M.DTS = D2ALLOC(); M.DTE = D2ALLOC(); M.DTW = D2ALLOC(); M.AHF = D2ALLOC();
  M.ah = M.cfg.ah;
  if (M.cfg.lauto_hmixt)
    M.ah = del4 ? -0.2e20 * (1280.0 / (double)M.cfg.nx_global)
                : 1.0e7 * (720.0 / (double)M.cfg.nx_global);
  for (size_t q = 0; q < M.n2 * M.nblocks; q++) M.AHF[q] = 1.0;
  if (M.cfg.lvariable_hmixt) {
    double den = 2.0 * O_PI * O_RADIUS / M.cfg.nx_global;
    den = den * den;
    for (size_t q = 0; q < M.n2 * M.nblocks; q++)
      M.AHF[q] = del4 ? pow(M.TAREA[q] / M.uarea_equator, 1.5) : sqrt(M.TAREA[q] / den);
    oracle_halo_2d(M.AHF, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0);
  }
  double* W = (double*)malloc(sizeof(double) * M.n2);
  for (int b = 0; b < M.nblocks; b++) {
    const double *HTN = B2(M.HTN, b), *HUW = B2(M.HUW, b), *HTE = B2(M.HTE, b),
                 *HUS = B2(M.HUS, b), *TR = B2(M.TAREA_R, b), *AHF = B2(M.AHF, b);
    /* del4: WORK1 = HTN/HUW ; del2: WORK1 = (HTN/HUW)*p5*(AHF + eoshift(AHF,dim=2,shift=1)) */
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++)
        W[IX2(i, j)] = del4 ? HTN[IX2(i, j)] / HUW[IX2(i, j)]
                            : (HTN[IX2(i, j)] / HUW[IX2(i, j)]) * 0.5 *
                                  (AHF[IX2(i, j)] + eo(AHF, i, j + 1));
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++) {
        B2(M.DTN, b)[IX2(i, j)] = W[IX2(i, j)] * TR[IX2(i, j)];
        B2(M.DTS, b)[IX2(i, j)] = eo(W, i, j - 1) * TR[IX2(i, j)];
      }
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++)
        W[IX2(i, j)] = del4 ? HTE[IX2(i, j)] / HUS[IX2(i, j)]
                            : (HTE[IX2(i, j)] / HUS[IX2(i, j)]) * 0.5 *
                                  (AHF[IX2(i, j)] + eo(AHF, i + 1, j));
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++) {
        B2(M.DTE, b)[IX2(i, j)] = W[IX2(i, j)] * TR[IX2(i, j)];
        B2(M.DTW, b)[IX2(i, j)] = eo(W, i - 1, j) * TR[IX2(i, j)];
      }
  }
  free(W);
}

/* init_del2u (hmix_del2.F90:97-421) / init_del4u (hmix_del4.F90:94-387) */
static void init_del2u_or_del4u(void) {
  const int nxb = M.nxb, nyb = M.nyb;
  const int del4 = (M.cfg.hmix_momentum_itype == POP_HMIX_DEL4);
  M.DUC = D2ALLOC(); M.DUN = D2ALLOC(); M.DUS = D2ALLOC(); M.DUE = D2ALLOC(); M.DUW = D2ALLOC();
  M.DMC = D2ALLOC(); M.DMN = D2ALLOC(); M.DMS = D2ALLOC(); M.DME = D2ALLOC(); M.DMW = D2ALLOC();
  M.DUM = D2ALLOC(); M.AMF = D2ALLOC();
  M.am = M.cfg.am;
  if (M.cfg.lauto_hmixu)
    M.am = del4 ? -0.6e20 * (1280.0 / (double)M.cfg.nx_global)
                : 1.0e7 * (720.0 / (double)M.cfg.nx_global);
  for (size_t q = 0; q < M.n2 * M.nblocks; q++) M.AMF[q] = 1.0;
  if (M.cfg.lvariable_hmixu) {
    double den = 2.0 * O_PI * O_RADIUS / M.cfg.nx_global;
    den = den * den;
    for (size_t q = 0; q < M.n2 * M.nblocks; q++)
      M.AMF[q] = del4 ? pow(M.UAREA[q] / M.uarea_equator, 1.5) : sqrt(M.UAREA[q] / den);
    oracle_halo_2d(M.AMF, POP_LOC_NECORNER, POP_KIND_SCALAR, 0.0);
  }
  size_t n2 = M.n2;
  double *W1 = malloc(8 * n2), *W2 = malloc(8 * n2), *KXU = malloc(8 * n2), *KYU = malloc(8 * n2),
         *DXKX = malloc(8 * n2), *DYKY = malloc(8 * n2), *DXKY = malloc(8 * n2),
         *DYKX = malloc(8 * n2);
#define LOOP2 for (int j = 1; j <= nyb; j++) for (int i = 1; i <= nxb; i++)
  for (int b = 0; b < M.nblocks; b++) {
    const double *HUS = B2(M.HUS, b), *HTE = B2(M.HTE, b), *HUW = B2(M.HUW, b), *HTN = B2(M.HTN, b),
                 *UR = B2(M.UAREA_R, b), *TR = B2(M.TAREA_R, b), *DXUR = B2(M.DXUR, b),
                 *DYUR = B2(M.DYUR, b), *AMF = B2(M.AMF, b);
    double *DUS = B2(M.DUS, b), *DUN = B2(M.DUN, b), *DUW = B2(M.DUW, b), *DUE = B2(M.DUE, b);
    if (del4) {
      LOOP2 W1[IX2(i, j)] = HUS[IX2(i, j)] / HTE[IX2(i, j)];
    } else {
      LOOP2 W1[IX2(i, j)] = (HUS[IX2(i, j)] / HTE[IX2(i, j)]) * 0.5 * (AMF[IX2(i, j)] + eo(AMF, i, j - 1));
    }
    LOOP2 { DUS[IX2(i, j)] = W1[IX2(i, j)] * UR[IX2(i, j)]; DUN[IX2(i, j)] = eo(W1, i, j + 1) * UR[IX2(i, j)]; }
    if (del4) {
      LOOP2 W1[IX2(i, j)] = HUW[IX2(i, j)] / HTN[IX2(i, j)];
    } else {
      LOOP2 W1[IX2(i, j)] = (HUW[IX2(i, j)] / HTN[IX2(i, j)]) * 0.5 * (AMF[IX2(i, j)] + eo(AMF, i - 1, j));
    }
    LOOP2 { DUW[IX2(i, j)] = W1[IX2(i, j)] * UR[IX2(i, j)]; DUE[IX2(i, j)] = eo(W1, i + 1, j) * UR[IX2(i, j)]; }
    LOOP2 {
      KXU[IX2(i, j)] = (eo(HUW, i + 1, j) - HUW[IX2(i, j)]) * UR[IX2(i, j)];
      KYU[IX2(i, j)] = (eo(HUS, i, j + 1) - HUS[IX2(i, j)]) * UR[IX2(i, j)];
    }
    /* KXT */
    LOOP2 W1[IX2(i, j)] = (HTE[IX2(i, j)] - eo(HTE, i - 1, j)) * TR[IX2(i, j)];
    if (del4) {
      LOOP2 W2[IX2(i, j)] = eo(W1, i + 1, j) - W1[IX2(i, j)];
      LOOP2 DXKX[IX2(i, j)] = 0.5 * (W2[IX2(i, j)] + eo(W2, i, j + 1)) * DXUR[IX2(i, j)];
      LOOP2 W2[IX2(i, j)] = eo(W1, i, j + 1) - W1[IX2(i, j)];
      LOOP2 DYKX[IX2(i, j)] = 0.5 * (W2[IX2(i, j)] + eo(W2, i + 1, j)) * DYUR[IX2(i, j)];
    } else {
      LOOP2 W2[IX2(i, j)] = 0.5 * (W1[IX2(i, j)] + eo(W1, i, j + 1)) * 0.5 * (eo(AMF, i - 1, j) + AMF[IX2(i, j)]);
      LOOP2 DXKX[IX2(i, j)] = (eo(W2, i + 1, j) - W2[IX2(i, j)]) * DXUR[IX2(i, j)];
      LOOP2 W2[IX2(i, j)] = 0.5 * (W1[IX2(i, j)] + eo(W1, i + 1, j)) * 0.5 * (eo(AMF, i, j - 1) + AMF[IX2(i, j)]);
      LOOP2 DYKX[IX2(i, j)] = (eo(W2, i, j + 1) - W2[IX2(i, j)]) * DYUR[IX2(i, j)];
    }
    /* KYT */
    LOOP2 W1[IX2(i, j)] = (HTN[IX2(i, j)] - eo(HTN, i, j - 1)) * TR[IX2(i, j)];
    if (del4) {
      LOOP2 W2[IX2(i, j)] = eo(W1, i, j + 1) - W1[IX2(i, j)];
      LOOP2 DYKY[IX2(i, j)] = 0.5 * (W2[IX2(i, j)] + eo(W2, i + 1, j)) * DYUR[IX2(i, j)];
      LOOP2 W2[IX2(i, j)] = eo(W1, i + 1, j) - W1[IX2(i, j)];
      LOOP2 DXKY[IX2(i, j)] = 0.5 * (W2[IX2(i, j)] + eo(W2, i, j + 1)) * DXUR[IX2(i, j)];
    } else {
      LOOP2 W2[IX2(i, j)] = 0.5 * (W1[IX2(i, j)] + eo(W1, i + 1, j)) * 0.5 * (eo(AMF, i, j - 1) + AMF[IX2(i, j)]);
      LOOP2 DYKY[IX2(i, j)] = (eo(W2, i, j + 1) - W2[IX2(i, j)]) * DYUR[IX2(i, j)];
      LOOP2 W2[IX2(i, j)] = 0.5 * (W1[IX2(i, j)] + eo(W1, i, j + 1)) * 0.5 * (eo(AMF, i - 1, j) + AMF[IX2(i, j)]);
      LOOP2 DXKY[IX2(i, j)] = (eo(W2, i + 1, j) - W2[IX2(i, j)]) * DXUR[IX2(i, j)];
    }
    if (del4) {
      LOOP2 {
        size_t q = IX2(i, j);
        B2(M.DUM, b)[q] = -(DXKX[q] + DYKY[q] + 2.0 * (KXU[q] * KXU[q] + KYU[q] * KYU[q]));
        B2(M.DMC, b)[q] = DXKY[q] - DYKX[q];
        B2(M.DME, b)[q] = 2.0 * KYU[q] / (HTN[q] + eo(HTN, i + 1, j));
        B2(M.DMN, b)[q] = -2.0 * KXU[q] / (HTE[q] + eo(HTE, i, j + 1));
      }
    } else {
      LOOP2 {
        size_t q = IX2(i, j);
        B2(M.DUM, b)[q] = -(DXKX[q] + DYKY[q] + 2.0 * AMF[q] * (KXU[q] * KXU[q] + KYU[q] * KYU[q]));
        B2(M.DMC, b)[q] = DXKY[q] - DYKX[q];
      }
      LOOP2 W1[IX2(i, j)] = (eo(AMF, i, j + 1) - eo(AMF, i, j - 1)) / (HTE[IX2(i, j)] + eo(HTE, i, j + 1));
      LOOP2 {
        size_t q = IX2(i, j);
        B2(M.DME, b)[q] = (2.0 * AMF[q] * KYU[q] + W1[q]) / (HTN[q] + eo(HTN, i + 1, j));
      }
      LOOP2 W1[IX2(i, j)] = (eo(AMF, i + 1, j) - eo(AMF, i - 1, j)) / (HTN[IX2(i, j)] + eo(HTN, i + 1, j));
      LOOP2 {
        size_t q = IX2(i, j);
        B2(M.DMN, b)[q] = -(2.0 * AMF[q] * KXU[q] + W1[q]) / (HTE[q] + eo(HTE, i, j + 1));
      }
    }
    LOOP2 {
      size_t q = IX2(i, j);
      B2(M.DUC, b)[q] = -(DUN[q] + DUS[q] + DUE[q] + DUW[q]);
      B2(M.DMW, b)[q] = -B2(M.DME, b)[q];
      B2(M.DMS, b)[q] = -B2(M.DMN, b)[q];
    }
  }
#undef LOOP2
  free(W1); free(W2); free(KXU); free(KYU); free(DXKX); free(DYKY); free(DXKY); free(DYKX);
}

static void alloc_state(void) {
  const pop_config* c = &M.cfg;
  size_t nb = M.nblocks, n2 = M.n2, n3 = M.n3, nt = M.nt;
  for (int t = 0; t < 3; t++) {
    M.TRACER[t] = o_alloc_d(n3 * nt * nb);
    M.UVEL[t] = o_alloc_d(n3 * nb); M.VVEL[t] = o_alloc_d(n3 * nb); M.RHO[t] = o_alloc_d(n3 * nb);
    M.PSURF[t] = o_alloc_d(n2 * nb); M.GRADPX[t] = o_alloc_d(n2 * nb);
    M.GRADPY[t] = o_alloc_d(n2 * nb); M.UBTROP[t] = o_alloc_d(n2 * nb);
    M.VBTROP[t] = o_alloc_d(n2 * nb);
  }
  M.PGUESS = o_alloc_d(n2 * nb);
  M.STF = o_alloc_d(n2 * nt * nb); M.TFW = o_alloc_d(n2 * nt * nb); M.SMF = o_alloc_d(n2 * 2 * nb);
  M.SHF_QSW = o_alloc_d(n2 * nb); M.FW = o_alloc_d(n2 * nb); M.FW_OLD = o_alloc_d(n2 * nb);
  M.VTF = o_alloc_d(n2 * nt * nb); M.VUF = o_alloc_d(n2 * nb); M.VVF = o_alloc_d(n2 * nb);
  M.SUMX = o_alloc_d(n2 * nb); M.SUMY = o_alloc_d(n2 * nb);
  M.RHOKMX = o_alloc_d(n2 * nb); M.RHOKMY = o_alloc_d(n2 * nb);
  M.UTK = o_alloc_d(n2 * nb); M.VTK = o_alloc_d(n2 * nb);
  M.DH = o_alloc_d(n2 * nb); M.DHU = o_alloc_d(n2 * nb); M.ZX = o_alloc_d(n2 * nb);
  M.ZY = o_alloc_d(n2 * nb);
  M.c2dtt = o_alloc_d(M.km + 2); M.dt = o_alloc_d(M.km + 2);
  /* vertical_mix.F90:383-419: VDC/VVC shapes */
  if (c->vmix_itype == POP_VMIX_GIVEN) {
    M.vdc_nk = c->vdc_kdim_halo ? M.km + 2 : M.km;
    M.vdc_k0 = c->vdc_kdim_halo ? 0 : 1;
    M.vdc_nd = c->vdc_ndim > 0 ? c->vdc_ndim : 1;
    M.vvc_nk = M.km;
  } else if (c->vmix_itype == POP_VMIX_CONST) {
    M.vdc_nk = c->implicit_vertical_mix ? M.km : 1; M.vdc_k0 = 1; M.vdc_nd = 1;
    M.vvc_nk = c->implicit_vertical_mix ? M.km : 1;
  } else {
    M.vdc_nk = c->implicit_vertical_mix ? M.km : 1; M.vdc_k0 = 1; M.vdc_nd = 1;
    M.vvc_nk = M.km;
  }
  M.VDC = o_alloc_d(n2 * M.vdc_nk * M.vdc_nd * nb);
  M.VVC = o_alloc_d(n2 * M.vvc_nk * nb);
  if (c->vmix_itype == POP_VMIX_CONST) { /* vmix_const.F90 init: VVC=const_vvc, VDC=const_vdc */
    for (size_t q = 0; q < n2 * M.vdc_nk * M.vdc_nd * nb; q++) M.VDC[q] = c->const_vdc;
    for (size_t q = 0; q < n2 * M.vvc_nk * nb; q++) M.VVC[q] = c->const_vvc;
  }
  /* vertical_mix.F90:444-448 */
  M.afac_t = o_alloc_d(M.km + 2); M.afac_u = o_alloc_d(M.km + 2);
  for (int k = 1; k <= M.km; k++) {
    M.afac_u[k] = c->aidif * M.dzwr[k];
    M.afac_t[k] = c->aidif * M.dzwr[k];
  }
  /* time_management.F90:962-964,1004,434-439 */
  M.dtt = c->dtt; M.dtu = c->dtt; M.dtp = c->dtt;
  for (int k = 1; k <= M.km; k++) M.dt[k] = M.dtt;
  M.alpha = 1.0 / 3.0; M.theta = 0.5; M.gamma_ = 1.0 - 2.0 * M.alpha;
  oracle_set_timestep(POP_TS_LEAPFROG);
}

/* step_mod.F90:302-320 */
void oracle_set_timestep(int ts_type) {
  M.leapfrogts = (ts_type == POP_TS_LEAPFROG || ts_type == POP_TS_AVG || ts_type == POP_TS_ROBERT);
  M.f_euler_ts = (ts_type == POP_TS_EULER);
  M.avg_ts = (ts_type == POP_TS_AVG);
  M.mix_pass = 0;
  if (M.leapfrogts) {
    M.mixtime = M.oldtime;
    M.beta = M.alpha;
    for (int k = 1; k <= M.km; k++) M.c2dtt[k] = 2.0 * M.dt[k];
    M.c2dtu = 2.0 * M.dtu;
    M.c2dtp = 2.0 * M.dtp;
  } else {
    M.mixtime = M.curtime;
    M.beta = M.theta;
    for (int k = 1; k <= M.km; k++) M.c2dtt[k] = M.dt[k];
    M.c2dtu = M.dtu;
    M.c2dtp = M.dtp;
  }
}

/* ---- field registry for the python harness ---- */
void* oracle_field(const char* name, int tlev) {
  int t = (tlev == POP_TIME_OLD) ? M.oldtime : (tlev == POP_TIME_NEW) ? M.newtime : M.curtime;
#define F(n, p) if (!strcmp(name, n)) return (void*)(p)
  F("TRACER", M.TRACER[t]); F("UVEL", M.UVEL[t]); F("VVEL", M.VVEL[t]); F("RHO", M.RHO[t]);
  F("PSURF", M.PSURF[t]); F("GRADPX", M.GRADPX[t]); F("GRADPY", M.GRADPY[t]);
  F("UBTROP", M.UBTROP[t]); F("VBTROP", M.VBTROP[t]); F("PGUESS", M.PGUESS);
  F("VDC", M.VDC); F("VVC", M.VVC); F("STF", M.STF); F("SMF", M.SMF); F("SHF_QSW", M.SHF_QSW);
  F("FW", M.FW); F("FW_OLD", M.FW_OLD); F("TFW", M.TFW);
  F("DH", M.DH); F("DHU", M.DHU); F("ZX", M.ZX); F("ZY", M.ZY);
  F("KMT", M.KMT); F("KMU", M.KMU); F("KMTN", M.KMTN); F("KMTS", M.KMTS); F("KMTE", M.KMTE);
  F("KMTW", M.KMTW); F("KMTEE", M.KMTEE); F("KMTNN", M.KMTNN);
  F("TLAT", M.TLAT); F("ULAT", M.ULAT); F("HTN", M.HTN); F("HTE", M.HTE); F("HUS", M.HUS); F("HUW", M.HUW);
  F("DXU", M.DXU); F("DYU", M.DYU); F("DXT", M.DXT); F("DYT", M.DYT);
  F("DXUR", M.DXUR); F("DYUR", M.DYUR); F("DXTR", M.DXTR); F("DYTR", M.DYTR);
  F("UAREA", M.UAREA); F("TAREA", M.TAREA); F("UAREA_R", M.UAREA_R); F("TAREA_R", M.TAREA_R);
  F("HU", M.HU); F("HUR", M.HUR); F("HT", M.HT); F("FCOR", M.FCOR);
  F("DZT", M.DZT); F("DZU", M.DZU); F("DZBC", M.DZBC);
  F("RCALCT", M.RCALCT); F("RCALCU", M.RCALCU);
  F("AU0", M.AU0); F("AUN", M.AUN); F("AUE", M.AUE); F("AUNE", M.AUNE);
  F("KXU", M.KXU); F("KYU", M.KYU);
  F("DTN", M.DTN); F("DTS", M.DTS); F("DTE", M.DTE); F("DTW", M.DTW); F("AHF", M.AHF);
  F("DUC", M.DUC); F("DUN", M.DUN); F("DUS", M.DUS); F("DUE", M.DUE); F("DUW", M.DUW);
  F("DMC", M.DMC); F("DMN", M.DMN); F("DMS", M.DMS); F("DME", M.DME); F("DMW", M.DMW);
  F("DUM", M.DUM); F("AMF", M.AMF);
  F("TALFXP", M.TALFXP); F("TBETXP", M.TBETXP); F("TGAMXP", M.TGAMXP);
  F("TALFYP", M.TALFYP); F("TBETYP", M.TBETYP); F("TGAMYP", M.TGAMYP);
  F("TALFXM", M.TALFXM); F("TBETXM", M.TBETXM); F("TDELXM", M.TDELXM);
  F("TALFYM", M.TALFYM); F("TBETYM", M.TBETYM); F("TDELYM", M.TDELYM);
  F("btropWgtCenter", M.btropWgtCenter); F("btropWgtNorth", M.btropWgtNorth);
  F("btropWgtEast", M.btropWgtEast); F("btropWgtNE", M.btropWgtNE);
  F("centerWgtClinicIndep", M.centerWgtClinicIndep); F("mMaskTropic", M.mMaskTropic);
  F("CHECKER", M.CHECKER); F("CONSTNT", M.CONSTNT);
  F("dz", M.dz); F("zt", M.zt); F("zw", M.zw); F("dzw", M.dzw); F("pressz", M.pressz);
  F("bouss", M.bouss); F("c2dtt", M.c2dtt);
  F("talfzp", M.talfzp); F("tbetzp", M.tbetzp); F("tgamzp", M.tgamzp);
  F("talfzm", M.talfzm); F("tbetzm", M.tbetzm); F("tdelzm", M.tdelzm);
#undef F
  return o_gm_field(name);
}
