/*
 * o_domain.c -- ORACLE (test infrastructure): block decomposition, halo update, global sums,
 * scatter/gather.  Restates source/blocks.F90:90-275 (create_blocks), the address-list semantics
 * of mpi/POP_HaloMod.F90:1732-2071 + 5632-6100 (POP_HaloUpdate / POP_HaloMsgCreate, including the
 * tripole fold :1947-2048), mpi/POP_ReductionsMod.F90:144-389 (POP_GlobalSum2DR8) and
 * source/POP_DistributionMod.F90:637-820 (cartesian distribution), as a serial multi-block code.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "pop_oracle.h"

omodel M;

double o_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
double oracle_timer(int id) { return (id >= 0 && id < OT_N) ? M.timer[id] : 0.0; }
void oracle_timers_reset(void) { memset(M.timer, 0, sizeof(M.timer)); }
/* OpenMP team size of the block loops (bench.py sets it explicitly: torchrun exports OMP_NUM_THREADS=1) */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_get_max_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_get_max_threads(void) { return 1; }
#endif

static void* zalloc(size_t n, size_t sz) {
  void* p = calloc(n ? n : 1, sz);
  if (!p) {
    fprintf(stderr, "oracle: out of memory (%zu x %zu)\n", n, sz);
    abort();
  }
  return p;
}
#define DALLOC(n) ((double*)zalloc((n), sizeof(double)))
#define IALLOC(n) ((int*)zalloc((n), sizeof(int)))

/* create_blocks: source/blocks.F90:90-275; the last-physical-point test carries the extra
   `< ie` guard of source/POP_BlocksMod.F90:325-328,373-376 (the legacy blocks.F90 variant mis-sets
   ie when the padded last block is no wider than the halo) */
static void create_blocks(void) {
  const pop_config* c = &M.cfg;
  int nxg = c->nx_global, nyg = c->ny_global, bsx = c->block_size_x, bsy = c->block_size_y;
  M.nbx = (nxg - 1) / bsx + 1;
  M.nby = (nyg - 1) / bsy + 1;
  M.nblocks = M.nbx * M.nby;
  M.nxb = bsx + 2 * NGHOST;
  M.nyb = bsy + 2 * NGHOST;
  int nb = M.nblocks;
  M.ib = IALLOC(nb); M.ie = IALLOC(nb); M.jb = IALLOC(nb); M.je = IALLOC(nb);
  M.iblk = IALLOC(nb); M.jblk = IALLOC(nb); M.active = IALLOC(nb);
  M.i_glob = IALLOC((size_t)nb * M.nxb);
  M.j_glob = IALLOC((size_t)nb * M.nyb);
  int n = 0;
  for (int jblock = 1; jblock <= M.nby; jblock++) {
    int js = (jblock - 1) * bsy + 1;
    for (int iblock = 1; iblock <= M.nbx; iblock++) {
      int is = (iblock - 1) * bsx + 1;
      M.iblk[n] = iblock; M.jblk[n] = jblock; M.active[n] = 1;
      M.ib[n] = NGHOST + 1; M.jb[n] = NGHOST + 1;
      M.ie[n] = M.nxb - NGHOST; M.je[n] = M.nyb - NGHOST;
      int* jg = M.j_glob + (size_t)n * M.nyb;
      for (int j = 1; j <= M.nyb; j++) {
        int v = js - NGHOST + j - 1;
        if (v < 1) {
          if (c->ns_boundary_type == POP_BNDY_CYCLIC) v += nyg; else v = 0;
        }
        if (v > nyg + NGHOST) v = 0;
        else if (v > nyg) {
          if (c->ns_boundary_type == POP_BNDY_CYCLIC) v -= nyg;
          else if (c->ns_boundary_type == POP_BNDY_TRIPOLE) v = -v;
          else v = 0;
        } else if (v == nyg && j > M.jb[n] && j < M.je[n]) M.je[n] = j; /* POP_BlocksMod.F90:325-328 */
        jg[j - 1] = v;
      }
      int* ig = M.i_glob + (size_t)n * M.nxb;
      for (int i = 1; i <= M.nxb; i++) {
        int v = is - NGHOST + i - 1;
        if (v < 1) {
          if (c->ew_boundary_type == POP_BNDY_CYCLIC) v += nxg; else v = 0;
        }
        if (v > nxg + NGHOST) v = 0;
        else if (v > nxg) {
          if (c->ew_boundary_type == POP_BNDY_CYCLIC) v -= nxg; else v = 0;
        } else if (v == nxg && i > M.ib[n] && i < M.ie[n]) M.ie[n] = i; /* POP_BlocksMod.F90:373-376 */
        ig[i - 1] = v;
      }
      n++;
    }
  }
}

int oracle_block_info(int b, int* out8, int* iglob, int* jglob) {
  if (b < 0 || b >= M.nblocks) return -1;
  out8[0] = b + 1; out8[1] = b + 1; out8[2] = M.ib[b]; out8[3] = M.ie[b];
  out8[4] = M.jb[b]; out8[5] = M.je[b]; out8[6] = M.iblk[b]; out8[7] = M.jblk[b];
  if (iglob) memcpy(iglob, M.i_glob + (size_t)b * M.nxb, sizeof(int) * M.nxb);
  if (jglob) memcpy(jglob, M.j_glob + (size_t)b * M.nyb, sizeof(int) * M.nyb);
  return 0;
}

int oracle_set_active(const int* active) {
  for (int b = 0; b < M.nblocks; b++) M.active[b] = active[b] ? 1 : 0;
  return 0;
}

/* which block owns global point (ig,jg) (1-based), and its local address */
static inline int owner(int ig, int jg, int* il, int* jl) {
  int bsx = M.cfg.block_size_x, bsy = M.cfg.block_size_y;
  int ibk = (ig - 1) / bsx, jbk = (jg - 1) / bsy;
  *il = ig - ibk * bsx + NGHOST;
  *jl = jg - jbk * bsy + NGHOST;
  return jbk * M.nbx + ibk;
}

/*
 * One 2-d halo update over all blocks (generic element type via macro).  Ghost cells whose global
 * index is 0 (closed boundary / padding) are NOT touched, as in the reference, where no message or
 * local copy targets them (POP_HaloMsgCreate returns for dstProc==0).  Ghost cells facing an
 * eliminated (land) block receive `fill` (mpi/POP_HaloMod.F90:1911-1913).  Tripole:
 * bufTripole(nxGlobal, haloWidth+1) gathers the top haloWidth+1 physical rows, is symmetrised for
 * NEcorner / Nface fields, then copied out with sign and offsets (:1947-2048, :5846-5882).
 */
#define HALO_IMPL(NAME, T, ABS, SIGNFIX)                                                          \
  static void NAME(T* a, size_t bstride, int loc, int kind, T fill) {                             \
    const int nxb = M.nxb, nyb = M.nyb, nxg = M.cfg.nx_global, nyg = M.cfg.ny_global;             \
    const int tripole = (M.cfg.ns_boundary_type == POP_BNDY_TRIPOLE);                             \
    /* regular copies: physical source cells only, so update order is irrelevant */              \
    _Pragma("omp parallel for schedule(static) if (nxg * nyg > 100000)")                        \
    for (int b = 0; b < M.nblocks; b++) {                                                         \
      if (!M.active[b]) continue;                                                                 \
      T* ab = a + (size_t)b * bstride;                                                            \
      const int *ig = M.i_glob + (size_t)b * nxb, *jg = M.j_glob + (size_t)b * nyb;               \
      for (int j = 1; j <= nyb; j++)                                                              \
        for (int i = 1; i <= nxb; i++) {                                                          \
          int phys = (i >= M.ib[b] && i <= M.ie[b] && j >= M.jb[b] && j <= M.je[b]);              \
          if (phys) continue;                                                                     \
          /* ghost ring around the physical domain only (padding beyond ie+2/je+2 has glob 0) */ \
          if (i > M.ie[b] + NGHOST || j > M.je[b] + NGHOST) continue;                             \
          int gi = ig[i - 1], gj = jg[j - 1];                                                     \
          if (gi <= 0 || gj <= 0) continue; /* closed / padding / tripole (below) */              \
          int il, jl, sb = owner(gi, gj, &il, &jl);                                               \
          ab[IX2(i, j)] = M.active[sb] ? a[(size_t)sb * bstride + IX2(il, jl)] : fill;            \
        }                                                                                         \
    }                                                                                             \
    if (!tripole) return;                                                                         \
    T* buf = (T*)malloc(sizeof(T) * (size_t)nxg * (NGHOST + 1));                                  \
    for (size_t q = 0; q < (size_t)nxg * (NGHOST + 1); q++) buf[q] = fill;                        \
    /* copy-in: rows je-2..je of every top-row block -> buf(iGlobal, 1..3) */                     \
    for (int b = 0; b < M.nblocks; b++) {                                                         \
      if (!M.active[b] || M.jblk[b] != M.nby) continue;                                           \
      const int* ig = M.i_glob + (size_t)b * nxb;                                                 \
      for (int j = 1; j <= NGHOST + 1; j++)                                                       \
        for (int i = M.ib[b]; i <= M.ie[b]; i++)                                                  \
          buf[(size_t)(j - 1) * nxg + (ig[i - 1] - 1)] =                                          \
              a[(size_t)b * bstride + IX2(i, M.je[b] - 1 - NGHOST + j)];                          \
    }                                                                                             \
    (void)nyg;                                                                                    \
    int isign = (kind == POP_KIND_SCALAR) ? 1 : -1;                                               \
    int ioff = 0, joff = 0;                                                                       \
    T* top = buf + (size_t)NGHOST * nxg; /* buf(:,haloWidth+1) */                                 \
    if (loc == POP_LOC_NECORNER) {                                                                \
      ioff = 1; joff = 1;                                                                         \
      for (int i = 1; i <= nxg / 2; i++) {                                                        \
        int id = nxg - i;                                                                         \
        T x1 = top[i - 1], x2 = top[id - 1];                                                      \
        SIGNFIX(top[i - 1], top[id - 1], x1, x2)                                                  \
      }                                                                                           \
      top[nxg - 1] = (T)(isign * top[nxg - 1]);                                                   \
    } else if (loc == POP_LOC_EFACE) {                                                            \
      ioff = 1; joff = 0;                                                                         \
    } else if (loc == POP_LOC_NFACE) {                                                            \
      ioff = 0; joff = 1;                                                                         \
      for (int i = 1; i <= nxg / 2; i++) {                                                        \
        int id = nxg + 1 - i;                                                                     \
        T x1 = top[i - 1], x2 = top[id - 1];                                                      \
        SIGNFIX(top[i - 1], top[id - 1], x1, x2)                                                  \
      }                                                                                           \
    }                                                                                             \
    /* copy-out into rows je..je+2 of every top-row block, all columns 1..ie+2 */                \
    for (int b = 0; b < M.nblocks; b++) {                                                         \
      if (!M.active[b] || M.jblk[b] != M.nby) continue;                                           \
      const int* ig = M.i_glob + (size_t)b * nxb;                                                 \
      for (int j = 1; j <= NGHOST + 1; j++)                                                       \
        for (int i = 1; i <= M.ie[b] + NGHOST; i++) {                                             \
          int is = nxg - ig[i - 1] + 1 - ioff;                                                    \
          int js = NGHOST + 3 - j - joff;                                                         \
          if (is == 0) is = nxg;                                                                  \
          if (js <= NGHOST + 1)                                                                   \
            a[(size_t)b * bstride + IX2(i, M.je[b] + j - 1)] =                                    \
                (T)(isign * buf[(size_t)(js - 1) * nxg + (is - 1)]);                              \
        }                                                                                         \
    }                                                                                             \
    free(buf);                                                                                    \
  }

#define SIGNFIX_R8(D1, D2, x1, x2)                          \
  {                                                         \
    double xavg = 0.5 * (fabs(x1) + fabs(x2));              \
    D1 = isign * copysign(xavg, x2);                        \
    D2 = isign * copysign(xavg, x1);                        \
  }
/* integer variant: mpi/POP_HaloMod.F90 2DI4 uses nint(0.5*(abs(x1)+abs(x2))) */
#define SIGNFIX_I4(D1, D2, x1, x2)                                   \
  {                                                                  \
    int xavg = (int)lround(0.5 * (abs(x1) + abs(x2)));               \
    D1 = isign * ((x2) < 0 ? -xavg : xavg);                          \
    D2 = isign * ((x1) < 0 ? -xavg : xavg);                          \
  }
HALO_IMPL(halo2_r8, double, fabs, SIGNFIX_R8)
HALO_IMPL(halo2_i4, int, abs, SIGNFIX_I4)

void oracle_halo_2d(double* a, int loc, int kind, double fill) {
  double t0 = o_now();
  halo2_r8(a, M.n2, loc, kind, fill);
  M.timer[OT_HALO] += o_now() - t0;
}
void oracle_halo_2d_i4(int* a, int loc, int kind, int fill) { halo2_i4(a, M.n2, loc, kind, fill); }
/* 3-d / 4-d: the same update for every level (and tracer); array(nxb,nyb,nz[,nt],nblocks) */
void oracle_halo_3d(double* a, int nz, int loc, int kind, double fill) {
  double t0 = o_now();
  for (int k = 0; k < nz; k++) halo2_r8(a + (size_t)k * M.n2, M.n2 * nz, loc, kind, fill);
  M.timer[OT_HALO] += o_now() - t0;
}
void oracle_halo_4d(double* a, int nz, int nt, int loc, int kind, double fill) {
  double t0 = o_now();
  for (int q = 0; q < nz * nt; q++)
    halo2_r8(a + (size_t)q * M.n2, M.n2 * nz * nt, loc, kind, fill);
  M.timer[OT_HALO] += o_now() - t0;
}

/* The reference's REPRODUCIBLE build (mpi/POP_ReductionsMod.F90:279-283, :357-363) accumulates blockSum and
   localSum in real(r16), reduces in r16 and rounds once to r8: the result is the correctly rounded exact sum and
   does not depend on the block decomposition or the summation order.  oracle_set_reproducible(1) selects that
   branch; the 113-bit accumulator is restated as a double-double (106 bits, error-free TwoSum), which rounds to the
   same r8 value unless the exact sum lies within ~n * 2^-106 of a rounding boundary.  Products array*mask are formed
   in r8 first, as in the Fortran expression. */
static int g_repro = 0;
void oracle_set_reproducible(int on) { g_repro = on ? 1 : 0; }
int oracle_get_reproducible(void) { return g_repro; }
typedef struct { double hi, lo; } odd;
static inline odd odd_add_d(odd a, double b) {
  double s = a.hi + b;
  double bb = s - a.hi;
  double e = (a.hi - (s - bb)) + (b - bb);
  e += a.lo;
  double hi = s + e;
  double lo = e - (hi - s);
  odd r = {hi, lo};
  return r;
}
static inline odd odd_add(odd a, odd b) {
  double s = a.hi + b.hi;
  double bb = s - a.hi;
  double e = (a.hi - (s - bb)) + (b.hi - bb);
  e += a.lo + b.lo;
  double hi = s + e;
  double lo = e - (hi - s);
  odd r = {hi, lo};
  return r;
}
/* one block's contribution in r16-equivalent precision (tripole de-duplication included) */
static odd block_sum_repro(const double* ab, const double* mb, int b, int loc) {
  odd acc = {0.0, 0.0};
  for (int j = M.jb[b]; j <= M.je[b]; j++)
    for (int i = M.ib[b]; i <= M.ie[b]; i++)
      acc = odd_add_d(acc, mb ? ab[IX2(i, j)] * mb[IX2(i, j)] : ab[IX2(i, j)]);
  if (M.cfg.ns_boundary_type == POP_BNDY_TRIPOLE && M.jblk[b] == M.nby &&
      (loc == POP_LOC_NFACE || loc == POP_LOC_NECORNER)) {
    int j = M.je[b];
    const int* ig = M.i_glob + (size_t)b * M.nxb;
    for (int i = M.ib[b]; i <= M.ie[b]; i++)
      if (ig[i - 1] > M.cfg.nx_global / 2)
        acc = odd_add_d(acc, -(mb ? ab[IX2(i, j)] * mb[IX2(i, j)] : ab[IX2(i, j)]));
  }
  return acc;
}

/* POP_GlobalSum2DR8: mpi/POP_ReductionsMod.F90:255-353 (per-block sums, j outer / i inner,
   tripole de-duplication :312-341, then the sum over blocks in block order) */
double oracle_global_sum(const double* a, int loc, const double* mask) {
  double localSum = 0.0;
  int tripole = (M.cfg.ns_boundary_type == POP_BNDY_TRIPOLE);
  if (g_repro) {
    odd* pr = (odd*)calloc(M.nblocks, sizeof(odd));
#pragma omp parallel for schedule(static) if (M.nblocks * M.n2 > 100000)
    for (int b = 0; b < M.nblocks; b++)
      if (M.active[b]) pr[b] = block_sum_repro(B2(a, b), mask ? B2(mask, b) : NULL, b, loc);
    odd tot = {0.0, 0.0};
    for (int b = 0; b < M.nblocks; b++)
      if (M.active[b]) tot = odd_add(tot, pr[b]);
    free(pr);
    return tot.hi + tot.lo;
  }
  double* part = (double*)calloc(M.nblocks, sizeof(double));
#pragma omp parallel for schedule(static) if (M.nblocks * M.n2 > 100000)
  for (int b = 0; b < M.nblocks; b++) {
    if (!M.active[b]) continue;
    const double* ab = B2(a, b);
    const double* mb = mask ? B2(mask, b) : NULL;
    double blockSum = 0.0;
    for (int j = M.jb[b]; j <= M.je[b]; j++)
      for (int i = M.ib[b]; i <= M.ie[b]; i++)
        blockSum = blockSum + (mb ? ab[IX2(i, j)] * mb[IX2(i, j)] : ab[IX2(i, j)]);
    if (tripole && M.jblk[b] == M.nby && (loc == POP_LOC_NFACE || loc == POP_LOC_NECORNER)) {
      int j = M.je[b];
      const int* ig = M.i_glob + (size_t)b * M.nxb;
      for (int i = M.ib[b]; i <= M.ie[b]; i++)
        if (ig[i - 1] > M.cfg.nx_global / 2)
          blockSum = blockSum - (mb ? ab[IX2(i, j)] * mb[IX2(i, j)] : ab[IX2(i, j)]);
    }
    part[b] = blockSum;
  }
  for (int b = 0; b < M.nblocks; b++) /* block order, as the serial loop */
    if (M.active[b]) localSum = localSum + part[b];
  free(part);
  return localSum;
}
/* POP_GlobalSumNfields2DR8 (:823): array(nxb,nyb,nfields,nblocks) */
void oracle_global_sum_n(const double* a, int nf, int loc, const double* mask, double* out) {
  for (int f = 0; f < nf; f++) out[f] = 0.0;
  if (g_repro) {
    for (int f = 0; f < nf; f++) {
      odd tot = {0.0, 0.0};
      for (int b = 0; b < M.nblocks; b++)
        if (M.active[b])
          tot = odd_add(tot, block_sum_repro(a + ((size_t)b * nf + f) * M.n2, mask ? B2(mask, b) : NULL, b, POP_LOC_CENTER));
      out[f] = tot.hi + tot.lo;
    }
    (void)loc;
    return;
  }
  double* part = (double*)calloc((size_t)M.nblocks * nf, sizeof(double));
#pragma omp parallel for schedule(static) if (M.nblocks * M.n2 > 100000)
  for (int b = 0; b < M.nblocks; b++) {
    if (!M.active[b]) continue;
    for (int f = 0; f < nf; f++) {
      const double* ab = a + ((size_t)b * nf + f) * M.n2;
      const double* mb = mask ? B2(mask, b) : NULL;
      double blockSum = 0.0;
      for (int j = M.jb[b]; j <= M.je[b]; j++)
        for (int i = M.ib[b]; i <= M.ie[b]; i++)
          blockSum = blockSum + (mb ? ab[IX2(i, j)] * mb[IX2(i, j)] : ab[IX2(i, j)]);
      part[(size_t)b * nf + f] = blockSum;
    }
  }
  for (int b = 0; b < M.nblocks; b++)
    if (M.active[b])
      for (int f = 0; f < nf; f++) out[f] = out[f] + part[(size_t)b * nf + f];
  free(part);
}

/* scatter physical cells of a global (nx_global, ny_global [,nz]) array; ghost cells untouched */
void oracle_scatter(double* dst, const double* glob, int nz) {
  int nxg = M.cfg.nx_global, nyg = M.cfg.ny_global;
  for (int b = 0; b < M.nblocks; b++) {
    const int *ig = M.i_glob + (size_t)b * M.nxb, *jg = M.j_glob + (size_t)b * M.nyb;
    for (int k = 0; k < nz; k++) {
      double* d = dst + ((size_t)b * nz + k) * M.n2;
      const double* g = glob + (size_t)k * nxg * nyg;
      for (int j = M.jb[b]; j <= M.je[b]; j++)
        for (int i = M.ib[b]; i <= M.ie[b]; i++)
          d[IX2(i, j)] = g[(size_t)(jg[j - 1] - 1) * nxg + (ig[i - 1] - 1)];
    }
  }
}
void oracle_gather(double* glob, const double* src, int nz) {
  int nxg = M.cfg.nx_global, nyg = M.cfg.ny_global;
  for (int b = 0; b < M.nblocks; b++) {
    const int *ig = M.i_glob + (size_t)b * M.nxb, *jg = M.j_glob + (size_t)b * M.nyb;
    for (int k = 0; k < nz; k++) {
      const double* s = src + ((size_t)b * nz + k) * M.n2;
      double* g = glob + (size_t)k * nxg * nyg;
      for (int j = M.jb[b]; j <= M.je[b]; j++)
        for (int i = M.ib[b]; i <= M.ie[b]; i++)
          g[(size_t)(jg[j - 1] - 1) * nxg + (ig[i - 1] - 1)] = s[IX2(i, j)];
    }
  }
}
void oracle_gather_i4(int* glob, const int* src) {
  int nxg = M.cfg.nx_global;
  for (int b = 0; b < M.nblocks; b++) {
    const int *ig = M.i_glob + (size_t)b * M.nxb, *jg = M.j_glob + (size_t)b * M.nyb;
    const int* s = src + (size_t)b * M.n2;
    for (int j = M.jb[b]; j <= M.je[b]; j++)
      for (int i = M.ib[b]; i <= M.ie[b]; i++)
        glob[(size_t)(jg[j - 1] - 1) * nxg + (ig[i - 1] - 1)] = s[IX2(i, j)];
  }
}
/* scatter_global semantics (gather_scatter.F90): physical cells + ghost cells via halo update */
void oracle_scatter_2d(double* dst, const double* glob, int loc, int kind) {
  memset(dst, 0, sizeof(double) * M.n2 * M.nblocks);
  oracle_scatter(dst, glob, 1);
  halo2_r8(dst, M.n2, loc, kind, 0.0);
}
void oracle_scatter_2d_i4(int* dst, const int* glob) {
  int nxg = M.cfg.nx_global;
  memset(dst, 0, sizeof(int) * M.n2 * M.nblocks);
  for (int b = 0; b < M.nblocks; b++) {
    const int *ig = M.i_glob + (size_t)b * M.nxb, *jg = M.j_glob + (size_t)b * M.nyb;
    int* d = dst + (size_t)b * M.n2;
    for (int j = M.jb[b]; j <= M.je[b]; j++)
      for (int i = M.ib[b]; i <= M.ie[b]; i++)
        d[IX2(i, j)] = glob[(size_t)(jg[j - 1] - 1) * nxg + (ig[i - 1] - 1)];
  }
  halo2_i4(dst, M.n2, POP_LOC_CENTER, POP_KIND_SCALAR, 0);
}

/*
 * POP_DistributionCreateCartesian: source/POP_DistributionMod.F90:637-820 with
 * POP_DistributionProcDecomp :1655.  blockLocation(n) = processor (1-based) or 0 for blocks with
 * no work.  Processor grid nprocsX x nprocsY chosen to make blocks-per-proc as square as possible.
 */
int oracle_distribution_cartesian(int nprocs, int nbx, int nby, const int* work, int* loc) {
  /* POP_DistributionProcDecomp (:1655): start at nint(sqrt(n)) and walk down; a factor pair that
     divides the block grid evenly (either orientation) wins; else the first valid pair */
  int nprocsX = 0, nprocsY = 0;
  int iguess = (int)lround(sqrt((double)nprocs));
  while (iguess >= 1) {
    int jguess = nprocs / iguess;
    if (iguess * jguess == nprocs) {
      if (nbx % iguess == 0 && nby % jguess == 0) { nprocsX = iguess; nprocsY = jguess; break; }
      else if (nbx % jguess == 0 && nby % iguess == 0) { nprocsX = jguess; nprocsY = iguess; break; }
      else if (nprocsX == 0) { nprocsX = iguess; nprocsY = jguess; }
    }
    iguess--;
  }
  if (nprocsX == 0) return -1;
  int numBlocksXPerProc = (nbx - 1) / nprocsX + 1;
  int numBlocksYPerProc = (nby - 1) / nprocsY + 1;
  for (int n = 0; n < nbx * nby; n++) loc[n] = 0;
  for (int j = 1; j <= nprocsY; j++)
    for (int i = 1; i <= nprocsX; i++) {
      int processor = (j - 1) * nprocsX + i;
      int is = (i - 1) * numBlocksXPerProc + 1, ie = i * numBlocksXPerProc;
      int js = (j - 1) * numBlocksYPerProc + 1, je = j * numBlocksYPerProc;
      if (ie > nbx) ie = nbx;
      if (je > nby) je = nby;
      for (int jb = js; jb <= je; jb++)
        for (int ib = is; ib <= ie; ib++) {
          int n = (jb - 1) * nbx + ib - 1;
          loc[n] = (work[n] != 0) ? processor : 0;
        }
    }
  return nprocsX * 1000 + nprocsY;
}

int oracle_init(const pop_config* cfg) {
  oracle_finalize();
  memset(&M, 0, sizeof(M));
  M.cfg = *cfg;
  M.km = cfg->km;
  M.nt = cfg->nt;
  create_blocks();
  M.n2 = (size_t)M.nxb * M.nyb;
  M.n3 = M.n2 * M.km;
  M.oldtime = 0; M.curtime = 1; M.newtime = 2; M.mixtime = 0;
  return 0;
}

/* every pointer stored in M is registered here so finalize can free them */
#define MAXPTR 512
static void* g_ptrs[MAXPTR];
static int g_nptr = 0;
void* o_alloc_d(size_t n) {
  void* p = DALLOC(n);
  if (g_nptr < MAXPTR) g_ptrs[g_nptr++] = p;
  return p;
}
void* o_alloc_i(size_t n) {
  void* p = IALLOC(n);
  if (g_nptr < MAXPTR) g_ptrs[g_nptr++] = p;
  return p;
}
void oracle_finalize(void) {
  o_gm_reset();
  for (int i = 0; i < g_nptr; i++) free(g_ptrs[i]);
  g_nptr = 0;
  free(M.ib); free(M.ie); free(M.jb); free(M.je); free(M.iblk); free(M.jblk); free(M.active);
  free(M.i_glob); free(M.j_glob);
  memset(&M, 0, sizeof(M));
}
