// pop_tracer.cu -- the tracer half of the baroclinic step.
//
//   tracer_column : ONE fused column kernel for everything tracer_update does per level
//                   (baroclinic.F90:1902-2306): hdifft del2/del4 (hmix_del2.F90:970-1144,
//                   hmix_del4.F90:889-1106), comp_flux_vel + advt centered/upwind3
//                   (advection.F90:1970-2132, :2139-2306, :2313-2677), vdifft
//                   (vertical_mix.F90:691-847) and the TNEW assembly (baroclinic.F90:2212-2300).
//                   The reference calls these once per (block, level) slab and carries WTK, VTF and
//                   AUX between calls; here a thread owns an (i,j) column, marches k and keeps the
//                   carried quantities in registers.  The same kernel, restricted to one level and
//                   one term, implements the slab entry points pop_advt / pop_hdifft / pop_vdifft.
//   impvmixt_dev  : implicit vertical mixing of tracers (vertical_mix.F90:1164-1382) and its
//                   corrector (:1460-1672): one column per thread, Thomas coefficients E(k) of the CTA's 128
//                   columns in shared memory (km x 128 doubles), operands prefetched through a register ring,
//                   the solution streamed in place through the output array; an opt-in variant stages the
//                   operands with TMA (POP_B200_THOMAS_TMA=1).
//   lw_lim tracers: their advective tendency is computed by pop_lwlim.cu into the output array, where the
//                   column kernels pick it up.
//
// Data movement: per level a CTA stages the (32+4) x (8+4) halo tile of each tracer (and the
// U*DYU, V*DXU flux operands) in shared memory; the k-invariant masks/coefficients (KMT, DTN..DTW,
// AHF) are staged once per CTA for the whole column march, so 2-d arrays cost ~1/km of a 3-d field.
#include "pop_state.cuh"

#define NTC 2  // tracers processed per pass of the column kernel

struct TracerArgs {
  GridView g;
  const double *TCUR, *TMIX, *TOLD, *UCUR, *VCUR, *STF, *TFW, *DH, *POLD, *PCUR;
  double* OUT;   // FULL: TRACER(new) (nxb,nyb,km,nt); slab modes: (nxb,nyb,nt)
  double* WTK;   // slab modes: carried vertical velocity at the top of level k (in/out)
  double *VTF, *AUX;  // slab modes: carried fluxes (nxb,nyb,nt)
  const double* HDT;  // GM: precomputed horizontal-mixing tendency (nxb,nyb,km,nt), pop_gm.cu; else null
  int lwl;            // lw_lim: the advective tendency of such tracers is waiting in OUT (ocean cells), pop_lwlim.cu
  // STF(i,j,n) = STF[n * stf_ns + (j - stf_j0) * stf_pitch + (i - stf_i0)]: the padded field (n2, nxb, 0, 0) or a strip of
  // physical cells in pinned host memory (nx_global * ny_local, nx_global, ib-1, jb-1)
  size_t stf_ns;
  int stf_pitch, stf_i0, stf_j0;
  int k0, k1;    // level range (1-based, inclusive)
  int n0, nn;    // tracers n0 .. n0+nn-1 (0-based)
  int adv[NTC];  // advection scheme of each tracer of this pass
  int hmix, lvariable_hmixt, varthick, implicit_vmix, predictor;
  double ah;
  PopTmap tm_tcur, tm_tmix, tm_u, tm_v;  // TMA descriptors (FULL mode, even nx_block)
};
#define TR_NS 3  // TMA pipeline depth (levels in flight per CTA)

// masked 5-point coefficients at tile point (ii,jj): hmix_del2.F90:1064-1078 / hmix_del4.F90:1014-1023
struct Coef5 {
  double cc, cn, cs, ce, cw;
};
// dzt: level k of DZT (partial bottom cells: hmix_del2.F90:1034-1062, hmix_del4.F90:964-988 -- a face is as thick as
// the thinner of its two cells) or null; gq/nxb: array index of the tile point and the row pitch; inb: the point and
// its four neighbours lie inside the padded block
__device__ __forceinline__ Coef5 tracer_coef(const int* s_kmt, const double* s_dtn, const double* s_dts,
                                             const double* s_dte, const double* s_dtw, int ii, int jj,
                                             int k, const double* dzt = nullptr, size_t gq = 0, int nxb = 0,
                                             bool inb = false) {
  const int q = TIX(ii, jj);
  const int kmt = s_kmt[q];
  Coef5 c;
  double dn = s_dtn[q], ds = s_dts[q], de = s_dte[q], dw = s_dtw[q];
  if (dzt) {
    if (inb) {
      const double z = dzt[gq];
      const RcpD rz = rcp_prepare(z);  // four IEEE quotients by the same thickness: one reciprocal refinement (div_by)
      dn = div_by(dn * fmin(z, dzt[gq + nxb]), rz);
      ds = div_by(ds * fmin(z, dzt[gq - nxb]), rz);
      de = div_by(de * fmin(z, dzt[gq + 1]), rz);
      dw = div_by(dw * fmin(z, dzt[gq - 1]), rz);
    } else {
      dn = ds = de = dw = 0.0;
    }
  }
  c.cn = (k <= s_kmt[TIX(ii, jj + 1)] && k <= kmt) ? dn : 0.0;
  c.cs = (k <= s_kmt[TIX(ii, jj - 1)] && k <= kmt) ? ds : 0.0;
  c.ce = (k <= s_kmt[TIX(ii + 1, jj)] && k <= kmt) ? de : 0.0;
  c.cw = (k <= s_kmt[TIX(ii - 1, jj)] && k <= kmt) ? dw : 0.0;
  c.cc = -(c.cn + c.cs + c.ce + c.cw);
  return c;
}
__device__ __forceinline__ double lap5(const Coef5& c, const double* t, int ii, int jj) {
  return c.cc * t[TIX(ii, jj)] + c.cn * t[TIX(ii, jj + 1)] + c.cs * t[TIX(ii, jj - 1)] +
         c.ce * t[TIX(ii + 1, jj)] + c.cw * t[TIX(ii - 1, jj)];
}

// TMA = true : levels are staged by cp.async.bulk.tensor into a TR_NS-deep ring of stages, one mbarrier
//              per stage; a single elected thread issues the copies, so no thread spends instructions
//              or registers on halo loads and DRAM latency is hidden by the ring, not by occupancy.
// TMA = false: cooperative plain loads into a single stage (slab entry points, odd nx_block).
template <int MODE, bool DEL4, bool UPW, bool TMA>
__global__ void __launch_bounds__(POP_NTHREADS, 2)
tracer_column_kernel(const POP_GRID_CONSTANT TracerArgs a) {
  POP_DYN_SMEM(smem_raw);
  constexpr int NS = TMA ? TR_NS : 1;
  constexpr int STAGE_TILES = 2 * NTC + 2;  // tc[NTC], tm[NTC], u, v
  double* sm = (double*)smem_raw;
  double* s_stage = sm;                                   // [NS][STAGE_TILES][TN]
  double* s_dtn = s_stage + NS * STAGE_TILES * POP_TN;
  double* s_dts = s_dtn + POP_TN;
  double* s_dte = s_dts + POP_TN;
  double* s_dtw = s_dte + POP_TN;
  double* s_ahf = s_dtw + POP_TN;
  double* s_dyu = s_ahf + POP_TN;
  double* s_dxu = s_dyu + POP_TN;
  double* s_d2base = s_dxu + POP_TN;       // [TMA ? 2 : 1][NTC][TN]
  int* s_kmt = (int*)(s_d2base + (TMA ? 2 : 1) * NTC * POP_TN);
  uint64_t* s_bar = (uint64_t*)(s_kmt + POP_TN);  // [NS]

  const GridView& g = a.g;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * POP_BX + tx;
  const int i0 = (g.ib - 1) + blockIdx.x * POP_BX;  // 0-based index of tile column 0
  const int j0 = (g.jb - 1) + blockIdx.y * POP_BY;
  const int i = i0 + tx, j = j0 + ty;                // 0-based
  const bool active = (i <= g.ie - 1) && (j <= g.je - 1);
  const size_t q = (size_t)j * g.nxb + i;
  const size_t n2 = g.n2;
  const int km = g.km, nxb = g.nxb, nyb = g.nyb;
  const bool same_mix = (a.TMIX == a.TCUR);
  constexpr bool DO_ADV = (MODE == TR_FULL || MODE == TR_ADVT);
  constexpr bool DO_HMIX = (MODE == TR_FULL || MODE == TR_HDIFFT);
  constexpr bool DO_VDIF = (MODE == TR_FULL || MODE == TR_VDIFFT);
  constexpr int HA = UPW ? 2 : 1;   // halo of the advected tracer tile
  constexpr int HM = DEL4 ? 2 : 1;  // halo of the mixed tracer tile

  // ---- k-invariant staging
  if (DO_ADV) {
    tile_load(s_dyu, g.DYU, i0, j0, nxb, nyb, -1, POP_BX - 1, -1, POP_BY - 1, tid);
    tile_load(s_dxu, g.DXU, i0, j0, nxb, nyb, -1, POP_BX - 1, -1, POP_BY - 1, tid);
  }
  if (DO_HMIX) {
    tile_load_i(s_kmt, g.KMT, i0, j0, nxb, nyb, -2, POP_BX + 1, -2, POP_BY + 1, tid);
    tile_load(s_dtn, g.DTN, i0, j0, nxb, nyb, -1, POP_BX, -1, POP_BY, tid);
    tile_load(s_dts, g.DTS, i0, j0, nxb, nyb, -1, POP_BX, -1, POP_BY, tid);
    tile_load(s_dte, g.DTE, i0, j0, nxb, nyb, -1, POP_BX, -1, POP_BY, tid);
    tile_load(s_dtw, g.DTW, i0, j0, nxb, nyb, -1, POP_BX, -1, POP_BY, tid);
    if (DEL4 && a.lvariable_hmixt) tile_load(s_ahf, g.AHF, i0, j0, nxb, nyb, -1, POP_BX, -1, POP_BY, tid);
  }
  int kmt = 0;
  double tarea_r = 0.0, tarea_rw = 0.0, tarea_rs = 0.0;
  if (active) {
    kmt = g.KMT[q];
    tarea_r = g.TAREA_R[q];
    if (UPW) {
      tarea_rw = g.TAREA_R[q - 1];
      tarea_rs = g.TAREA_R[q - nxb];
    }
  }
  // ---- carried state
  double wtk = 0.0;
  double tc_m[NTC], told_c[NTC], vtf[NTC], aux[NTC];
#pragma unroll
  for (int m = 0; m < NTC; m++) tc_m[m] = told_c[m] = vtf[m] = aux[m] = 0.0;
  if (active) {
    if (DO_ADV) wtk = (a.k0 == 1 && MODE == TR_FULL) ? a.DH[q] : (MODE == TR_FULL ? 0.0 : a.WTK[q]);
#pragma unroll
    for (int m = 0; m < NTC; m++) {
      if (m >= a.nn) continue;
      const int n = a.n0 + m;
      if (DO_ADV && a.k0 > 1) tc_m[m] = a.TCUR[((size_t)n * km + (a.k0 - 2)) * n2 + q];
      if (DO_VDIF) {
        told_c[m] = a.TOLD[((size_t)n * km + (a.k0 - 1)) * n2 + q];
        if (a.k0 > 1) vtf[m] = a.VTF[(size_t)n * n2 + q];
      }
      if (DO_ADV && UPW && a.k0 > 1) aux[m] = a.AUX[(size_t)n * n2 + q];
    }
  }

  const bool mix_alias = same_mix && DO_ADV && (TMA || HA >= HM);  // TMIX tile == TCUR tile
  const uint32_t stage_bytes = (uint32_t)((a.nn * (mix_alias ? 1 : 2) + 2) * POP_TILE_BYTES);
  // issue the TMA copies of level kk into ring slot (kk - k0) % NS (one thread)
  auto issue = [&](int kk) {
    const int sl = (kk - a.k0) % NS;
    double* st = s_stage + (size_t)sl * STAGE_TILES * POP_TN;
    mbar_expect_tx(&s_bar[sl], stage_bytes);
    for (int m = 0; m < a.nn; m++) {
      const int z = (a.n0 + m) * km + (kk - 1);
      tma_load_tile(st + m * POP_TN, &a.tm_tcur, i0 - POP_H, j0 - POP_H, z, &s_bar[sl]);
      if (!mix_alias) tma_load_tile(st + (NTC + m) * POP_TN, &a.tm_tmix, i0 - POP_H, j0 - POP_H, z, &s_bar[sl]);
    }
    tma_load_tile(st + 2 * NTC * POP_TN, &a.tm_u, i0 - POP_H, j0 - POP_H, kk - 1, &s_bar[sl]);
    tma_load_tile(st + (2 * NTC + 1) * POP_TN, &a.tm_v, i0 - POP_H, j0 - POP_H, kk - 1, &s_bar[sl]);
  };
  if (TMA) {
    if (tid == 0) {
      for (int sl = 0; sl < NS; sl++) mbar_init(&s_bar[sl], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0)
      for (int kk = a.k0; kk <= a.k1 && kk < a.k0 + NS; kk++) issue(kk);
  }

  // D2TK = AHF * L(T) on the first halo ring (hmix_del4.F90:1025-1046)
  auto compute_d2 = [&](int kk, const double* tmix, double* d2) {
    constexpr int w = POP_BX + 2, npts = w * (POP_BY + 2);
    for (int p = tid; p < npts; p += POP_NTHREADS) {
      const int jj = p / w - 1, ii = p % w - 1;
      const int gi = i0 + ii, gj = j0 + jj;
      const Coef5 c = tracer_coef(s_kmt, s_dtn, s_dts, s_dte, s_dtw, ii, jj, kk,
                                  g.DZT ? g.DZT + (size_t)kk * n2 : nullptr, (size_t)gj * nxb + gi, nxb,
                                  gi >= 1 && gi <= nxb - 2 && gj >= 1 && gj <= nyb - 2);
#pragma unroll
      for (int m = 0; m < NTC; m++) {
        if (m >= a.nn) continue;
        const double v = lap5(c, tmix + m * POP_TN, ii, jj);
        d2[m * POP_TN + TIX(ii, jj)] = a.lvariable_hmixt ? s_ahf[TIX(ii, jj)] * v : v;
      }
    }
  };
  if (TMA && DO_HMIX && DEL4) {  // prologue of the software pipeline: D2 of the first level
    mbar_wait(&s_bar[0], 0u);
    compute_d2(a.k0, mix_alias ? s_stage : s_stage + NTC * POP_TN, s_d2base);
    __syncthreads();
  }
  const bool told_is_mix = (a.TOLD == a.TMIX);
  // VDC of the next level is fetched one level ahead (its latency hides behind a whole level)
  double vdc_nx[NTC];
#pragma unroll
  for (int m = 0; m < NTC; m++) vdc_nx[m] = 0.0;
  auto vdc_at = [&](int m, int kk) {
    const int n = a.n0 + m;
    const int mt2 = (n + 1 < g.vdc_nd) ? n + 1 : g.vdc_nd;
    const int kq = (g.vdc_nk == 1) ? 1 : kk;
    return g.VDC[((size_t)(mt2 - 1) * g.vdc_nk + (kq - g.vdc_k0)) * n2 + q];
  };
  if (DO_VDIF && active) {
#pragma unroll
    for (int m = 0; m < NTC; m++)
      if (m < a.nn) vdc_nx[m] = vdc_at(m, a.k0);
  }

  for (int k = a.k0; k <= a.k1; k++) {
    const int slot = (k - a.k0) % NS;
    double* s_tc = s_stage + (size_t)slot * STAGE_TILES * POP_TN;  // [NTC][TN]
    double* s_tm = mix_alias ? s_tc : s_tc + NTC * POP_TN;          // [NTC][TN]
    double* s_u = s_tc + 2 * NTC * POP_TN;
    double* s_v = s_u + POP_TN;
    if (TMA) {
      mbar_wait(&s_bar[slot], (uint32_t)(((k - a.k0) / NS) & 1));
    } else {
      __syncthreads();
      // ---- stage level k
      if (DO_ADV) {
        tile_load(s_u, a.UCUR + (size_t)(k - 1) * n2, i0, j0, nxb, nyb, -1, POP_BX - 1, -1, POP_BY - 1, tid);
        tile_load(s_v, a.VCUR + (size_t)(k - 1) * n2, i0, j0, nxb, nyb, -1, POP_BX - 1, -1, POP_BY - 1, tid);
      }
#pragma unroll
      for (int m = 0; m < NTC; m++) {
        if (m >= a.nn) continue;
        const size_t lev = ((size_t)(a.n0 + m) * km + (k - 1)) * n2;
        if (DO_ADV) tile_load(s_tc + m * POP_TN, a.TCUR + lev, i0, j0, nxb, nyb, -HA, POP_BX - 1 + HA, -HA, POP_BY - 1 + HA, tid);
        if (DO_HMIX && !mix_alias)
          tile_load(s_tm + m * POP_TN, a.TMIX + lev, i0, j0, nxb, nyb, -HM, POP_BX - 1 + HM, -HM, POP_BY - 1 + HM, tid);
      }
      __syncthreads();
    }
    double* s_d2 = s_d2base + (TMA ? ((k - a.k0) & 1) * NTC * POP_TN : 0);
    if (DO_HMIX && DEL4 && !TMA) {
      compute_d2(k, s_tm, s_d2);
      __syncthreads();
    }
    // TMA path: the next level has (normally) landed already; it provides T(k+1) at the column and is
    // the input of the D2 pass that is overlapped with this level's output pass
    const bool have_next = TMA && (k < a.k1);
    const int nslot = (k + 1 - a.k0) % NS;
    double* n_tc = s_stage + (size_t)nslot * STAGE_TILES * POP_TN;
    double* n_tm = mix_alias ? n_tc : n_tc + NTC * POP_TN;
    if (have_next) mbar_wait(&s_bar[nslot], (uint32_t)(((k + 1 - a.k0) / NS) & 1));
    if (active) {
    // ---- flux velocities and the vertical velocity at the bottom of the level
    double ute = 0.0, utw = 0.0, vtn = 0.0, vts = 0.0, wtkb = 0.0;
    const bool pbc = (g.DZT != nullptr);
    const double dzt_c = pbc ? g.DZT[(size_t)k * n2 + q] : 0.0;  // thickness of this cell
    const RcpD rzt = pbc ? rcp_prepare(dzt_c) : RcpD{1.0, 1.0};
    const double h_dzt = pbc ? div_by(0.5, rzt) : 0.0;  // 0.5 / DZT
    if (DO_ADV) {
#define UD(di, dj) (s_u[TIX(tx + (di), ty + (dj))] * s_dyu[TIX(tx + (di), ty + (dj))])
#define VD(di, dj) (s_v[TIX(tx + (di), ty + (dj))] * s_dxu[TIX(tx + (di), ty + (dj))])
      if (pbc) {  // advection.F90:2040-2066: the flux velocities carry the thickness of the U cells
        const double* zu = g.DZU + (size_t)k * n2 + q;
        const double z00 = zu[0], z0m = zu[-(ptrdiff_t)nxb], zm0 = zu[-1], zmm = zu[-(ptrdiff_t)nxb - 1];
        ute = 0.5 * (UD(0, 0) * z00 + UD(0, -1) * z0m);
        utw = 0.5 * (UD(-1, 0) * zm0 + UD(-1, -1) * zmm);
        vtn = 0.5 * (VD(0, 0) * z00 + VD(-1, 0) * zm0);
        vts = 0.5 * (VD(0, -1) * z0m + VD(-1, -1) * zmm);
      } else {
      ute = 0.5 * (UD(0, 0) + UD(0, -1));
      utw = 0.5 * (UD(-1, 0) + UD(-1, -1));
      vtn = 0.5 * (VD(0, 0) + VD(-1, 0));
      vts = 0.5 * (VD(0, -1) + VD(-1, -1));
      }
#undef UD
#undef VD
      if (k < km) {
        const double FC = (vtn - vts + ute - utw) * tarea_r;
        if (pbc) wtkb = (k < kmt) ? wtk + FC : 0.0;  // :2110-2111
        else wtkb = (k < kmt) ? wtk + c_vc.dz[k] * FC : 0.0;
      }
    }
    Coef5 cc5;
    if (DO_HMIX) cc5 = tracer_coef(s_kmt, s_dtn, s_dts, s_dte, s_dtw, tx, ty, k, pbc ? g.DZT + (size_t)k * n2 : nullptr, q,
                                   nxb, true);

#pragma unroll
    for (int m = 0; m < NTC; m++) {
      if (m >= a.nn) continue;
      const int n = a.n0 + m;  // 0-based tracer index
      const size_t lev = ((size_t)n * km + (k - 1)) * n2 + q;
      const double* tc = s_tc + m * POP_TN;
      // ---- horizontal mixing
      double hd = 0.0;
      if (DO_HMIX) {
        if (a.HDT) hd = a.HDT[lev];
        else if (DEL4) hd = a.ah * lap5(cc5, s_d2 + m * POP_TN, tx, ty);
        else {
          hd = a.ah * lap5(cc5, s_tm + m * POP_TN, tx, ty);
        }
      }
      // ---- advection
      double L = 0.0;
      if (DO_ADV) {
        const double T = tc[TIX(tx, ty)];
        const double Tp = (k < km) ? (have_next ? n_tc[m * POP_TN + TIX(tx, ty)] : a.TCUR[lev + n2]) : 0.0;
        if (a.adv[m] == POP_TADVECT_LW_LIM) {  // advection.F90:2684-3280, left in the output cell by lw_lim_kernel
          L = (k <= kmt) ? a.OUT[(MODE == TR_FULL) ? lev : (size_t)n * n2 + q] : 0.0;
        } else if (a.adv[m] == POP_TADVECT_CENTERED) {  // advection.F90:2243-2301
          L = 0.5 *
              ((vtn - vts + ute - utw) * T + vtn * tc[TIX(tx, ty + 1)] - vts * tc[TIX(tx, ty - 1)] +
               ute * tc[TIX(tx + 1, ty)] - utw * tc[TIX(tx - 1, ty)]) *
              tarea_r;
          if (pbc) L = div_by(L, rzt);  // :2223-2238
          if (k == 1) {
            if (!a.varthick) L = L + c_vc.dzr[k] * wtk * T;
          } else {
            if (pbc) L = L + h_dzt * wtk * (tc_m[m] + T);  // :2278-2280
            else L = L + c_vc.dz2r[k] * wtk * (tc_m[m] + T);
          }
          if (k < km) {
            if (pbc) L = L - h_dzt * wtkb * (T + Tp);
            else L = L - c_vc.dz2r[k] * wtkb * (T + Tp);
          }
        } else if (UPW) {  // upwind3: advection.F90:2387-2476 + hupw3 :2543-2672
          const double CE = ute * tarea_r, CW = -utw * tarea_r, CN = vtn * tarea_r, CS = -vts * tarea_r;
          const double CEw = utw * tarea_rw;  // CE(i-1,j) = UTE(i-1,j)*TAREA_R(i-1,j)
          const double CNs = vts * tarea_rs;  // CN(i,j-1)
          double TE[2], TN[2];
#pragma unroll
          for (int s = 0; s < 2; s++) {  // s=0: face (i,j); s=1: face (i-1,j)
            const int ii = tx - s;
            const size_t qq = q - s;
            const int kE = g.KMT[qq + 1], kW = g.KMT[qq - 1], kEE = g.KMT[qq + 2];
            double work, ap, bp, gp, am, bm, dm;
            if (k <= kE) { work = g.TBETXP[qq]; ap = g.TALFXP[qq]; }
            else { work = g.TBETXP[qq] + g.TALFXP[qq]; ap = 0.0; }
            if (k <= kW) { bp = work; gp = g.TGAMXP[qq]; }
            else { bp = work + g.TGAMXP[qq]; gp = 0.0; }
            if (k <= kEE) { am = g.TALFXM[qq]; dm = g.TDELXM[qq]; }
            else { am = g.TALFXM[qq] + g.TDELXM[qq]; dm = 0.0; }
            bm = g.TBETXM[qq];
            const double cface = s ? CEw : CE;
            if (cface > 0.0) TE[s] = ap * tc[TIX(ii + 1, ty)] + bp * tc[TIX(ii, ty)] + gp * tc[TIX(ii - 1, ty)];
            else TE[s] = am * tc[TIX(ii + 1, ty)] + bm * tc[TIX(ii, ty)] + dm * tc[TIX(ii + 2, ty)];
          }
#pragma unroll
          for (int s = 0; s < 2; s++) {  // s=0: face (i,j); s=1: face (i,j-1)
            const int jj = ty - s;
            const size_t qq = q - (size_t)s * nxb;
            const int kN = g.KMT[qq + nxb], kS = g.KMT[qq - nxb], kNN = g.KMT[qq + 2 * (size_t)nxb];
            double work, ap, bp, gp, am, bm, dm;
            if (k <= kN) { work = g.TBETYP[qq]; ap = g.TALFYP[qq]; }
            else { work = g.TBETYP[qq] + g.TALFYP[qq]; ap = 0.0; }
            if (k <= kS) { bp = work; gp = g.TGAMYP[qq]; }
            else { bp = work + g.TGAMYP[qq]; gp = 0.0; }
            if (k <= kNN) { am = g.TALFYM[qq]; dm = g.TDELYM[qq]; }
            else { am = g.TALFYM[qq] + g.TDELYM[qq]; dm = 0.0; }
            bm = g.TBETYM[qq];
            const double cface = s ? CNs : CN;
            if (cface > 0.0) TN[s] = ap * tc[TIX(tx, jj + 1)] + bp * tc[TIX(tx, jj)] + gp * tc[TIX(tx, jj - 1)];
            else TN[s] = am * tc[TIX(tx, jj + 1)] + bm * tc[TIX(tx, jj)] + dm * tc[TIX(tx, jj + 2)];
          }
          L = CE * TE[0] + CW * TE[1];
          L = L + CN * TN[0] + CS * TN[1];
          // vertical: advection.F90:2402-2476
          double AZM, DZM;
          if (k < kmt - 1) { AZM = c_vc.talfzm[k]; DZM = c_vc.tdelzm[k]; }
          else { AZM = c_vc.talfzm[k] + c_vc.tdelzm[k]; DZM = 0.0; }
          double auxb;
          if (k < km - 1 && k > 1) {
            const double Tpp = a.TCUR[lev + 2 * n2];
            const double TPLUS = c_vc.talfzp[k] * Tp + c_vc.tbetzp[k] * T + c_vc.tgamzp[k] * tc_m[m];
            const double TMINUS = AZM * Tp + c_vc.tbetzm[k] * T + DZM * Tpp;
            auxb = (wtkb - fabs(wtkb)) * TPLUS + (wtkb + fabs(wtkb)) * TMINUS;
          } else if (k == 1) {
            const double Tpp = (k + 2 <= km) ? a.TCUR[lev + 2 * n2] : 0.0;
            const double TPLUS = c_vc.talfzp[k] * Tp + c_vc.tbetzp[k] * T;
            const double TMINUS = AZM * Tp + c_vc.tbetzm[k] * T + DZM * Tpp;
            auxb = (wtkb - fabs(wtkb)) * TPLUS + (wtkb + fabs(wtkb)) * TMINUS;
          } else if (k == km - 1) {
            const double TPLUS = c_vc.talfzp[k] * Tp + c_vc.tbetzp[k] * T + c_vc.tgamzp[k] * tc_m[m];
            const double TMINUS = AZM * Tp + c_vc.tbetzm[k] * T;
            auxb = (wtkb - fabs(wtkb)) * TPLUS + (wtkb + fabs(wtkb)) * TMINUS;
          } else {
            auxb = 0.0;
          }
          if (k == 1) {
            if (!a.varthick) {
              const double FLUX_T = c_vc.dzr[k] * wtk * T;
              L = L + FLUX_T - c_vc.dz2r[k] * auxb;
            } else {
              L = L - c_vc.dz2r[k] * auxb;
            }
          } else {
            L = L + c_vc.dz2r[k] * (aux[m] - auxb);
          }
          aux[m] = auxb;
        }
        tc_m[m] = T;
      }
      // ---- explicit vertical diffusion (top/bottom fluxes): vertical_mix.F90:779-838
      double vd = 0.0;
      if (DO_VDIF) {
        const double vdc = vdc_nx[m];
        if (k < a.k1) vdc_nx[m] = vdc_at(m, k + 1);
        const double told_p = (k < km) ? ((have_next && told_is_mix) ? n_tm[m * POP_TN + TIX(tx, ty)] : a.TOLD[lev + n2])
                                       : told_c[m];
        if (k == 1) vtf[m] = (kmt >= 1) ? a.STF[(size_t)n * a.stf_ns + (size_t)(j - a.stf_j0) * a.stf_pitch + (i - a.stf_i0)] : 0.0;
        double VTFB;
        if (pbc) {  // vertical_mix.F90:790-803
          const double dzt_p = g.DZT[(size_t)((k < km) ? k + 1 : km) * n2 + q];
          VTFB = (kmt > k) ? vdc * (told_c[m] - told_p) / (0.5 * (dzt_c + dzt_p)) : 0.0;
          vd = (k <= kmt) ? div_by(vtf[m] - VTFB, rzt) : 0.0;
        } else {
          VTFB = (kmt > k) ? vdc * (told_c[m] - told_p) * c_vc.dzwr[k] : 0.0;
          vd = (k <= kmt) ? (vtf[m] - VTFB) * c_vc.dzr[k] : 0.0;
        }
        vtf[m] = VTFB;
        told_c[m] = told_p;
      }
      // ---- output
      if (MODE == TR_HDIFFT) a.OUT[(size_t)n * n2 + q] = hd;
      else if (MODE == TR_ADVT) a.OUT[(size_t)n * n2 + q] = L;
      else if (MODE == TR_VDIFFT) a.OUT[(size_t)n * n2 + q] = vd;
      else {  // tracer_update: baroclinic.F90:1993-2300
        double FT = 0.0 + hd;
        FT = FT - L;
        FT = FT + vd;
        if (k == 1 && a.varthick) FT = FT + c_vc.dzr[1] * a.TFW[(size_t)n * n2 + q];
        FT = FT + 0.0;  // source terms are zero on this path
        const bool pred = a.predictor && k == 1 && n < 2;
        double* out = a.OUT + lev;
        if (a.implicit_vmix) {
          if (pred) {
            if (kmt > 0)
              *out = c_vc.c2dtt[1] * FT -
                     2.0 * tc[TIX(tx, ty)] * (a.PCUR[q] - a.POLD[q]) / (POP_GRAV * c_vc.dz[1]);
          } else {
            *out = (k <= kmt) ? c_vc.c2dtt[k] * FT : 0.0;
          }
        } else {
          const double To = a.TOLD[lev];
          if (a.varthick && k == 1) {
            if (pred)
              *out = (kmt > 0) ? To + (1.0 / (1.0 + a.PCUR[q] / (POP_GRAV * c_vc.dz[1]))) *
                                          (c_vc.c2dtt[1] * FT - 2.0 * tc[TIX(tx, ty)] * (a.PCUR[q] - a.POLD[q]) /
                                                                    (POP_GRAV * c_vc.dz[1]))
                               : 0.0;
            else
              *out = (k <= kmt) ? c_vc.c2dtt[k] * FT : 0.0;
          } else {
            *out = (k <= kmt) ? To + c_vc.c2dtt[k] * FT : 0.0;
          }
        }
      }
    }
    if (DO_ADV) wtk = wtkb;  // advection.F90:1960
    }  // active
    if (TMA) {
      if (DO_HMIX && DEL4 && have_next)
        compute_d2(k + 1, n_tm, s_d2base + ((k + 1 - a.k0) & 1) * NTC * POP_TN);
      __syncthreads();  // every thread is done with this ring slot; D2(k+1) is complete
      if (tid == 0 && k + NS <= a.k1) issue(k + NS);
    }
  }
  // ---- hand the carried state back to the slab caller
  if (MODE != TR_FULL && active) {
    if (MODE == TR_ADVT) a.WTK[q] = wtk;
#pragma unroll
    for (int m = 0; m < NTC; m++) {
      if (m >= a.nn) continue;
      const int n = a.n0 + m;
      if (MODE == TR_VDIFFT) a.VTF[(size_t)n * n2 + q] = vtf[m];
      if (MODE == TR_ADVT && UPW) a.AUX[(size_t)n * n2 + q] = aux[m];
    }
  }
}

// =====================================================================================
// Fast path of tracer_update for the production configuration of the leapfrog step: a pair of
// centred-advection tracers, implicit vertical mixing, TMIX == TOLD != TCUR, even nx_block.
// Same arithmetic, operation for operation, as tracer_column_kernel<TR_FULL> (which stays the
// general path and the reference for every other option); what changes is where the operands
// come from:
//   * every per-level input, VDC included, arrives by TMA in a 3-deep mbarrier ring, so the level
//     loop issues no global loads at all (only the two output stores);
//   * the k-invariant stencil coefficients of a thread's own column, its neighbours' KMT and its
//     area factor live in registers;
//   * the flux velocities UTE/VTN and the first del4 Laplacian are built once per tile point per
//     level (ring-1 tiles) instead of once per use.
// =====================================================================================
struct TracerFastArgs {
  GridView g;
  const double *STF, *TFW, *DH, *POLD, *PCUR;
  double* OUT;
  int n0;                      // first tracer (0-based) of the pair
  int lw[NTC];                 // tracer m is advected by lw_lim: its L(T) waits in OUT (ocean cells), pop_lwlim.cu
  size_t stf_ns;               // STF addressing, as in TracerArgs
  int stf_pitch, stf_i0, stf_j0;
  int vdc_lev0[NTC], vdc_kstr;  // VDC level of tracer m at level k: vdc_lev0[m] + k*vdc_kstr
  int lvariable_hmixt, varthick, predictor;
  double ah;
  PopTmap tm_tcur, tm_tmix, tm_u, tm_v, tm_vdc;
  PopTmap tm_dzt, tm_dzu;  // partial bottom cells: DZT, DZU (levels 0..km+1) as two more halo tiles of a stage
};
#define TF_NS 3
#define TF_NS_PBC 2  // pipeline depth with the two extra tiles (two CTAs per SM must still fit)
#define TF_STAGE (6 * POP_TN + NTC * POP_NTHREADS)  // tc[2], tm[2], u, v halo tiles + vdc[2] column tiles
#define TF_FIXED 11                                  // ring-1 tiles: dtn dts dte dtw ahf dyu dxu ute vtn d2[2]
template <bool DEL4, bool PBC>
__global__ void __launch_bounds__(POP_NTHREADS, 2)
tracer_fast_kernel(const POP_GRID_CONSTANT TracerFastArgs a) {
  POP_DYN_SMEM(smem_raw);
  constexpr int NS = PBC ? TF_NS_PBC : TF_NS;
  constexpr int NTL = PBC ? 8 : 6;                               // halo tiles per stage (PBC: + DZT, DZU)
  constexpr int STAGE = NTL * POP_TN + NTC * POP_NTHREADS;      // + the vdc column tiles
  double* s_stage = (double*)smem_raw;  // [NS][STAGE]
  double* s_dtn = s_stage + NS * STAGE;
  double* s_dts = s_dtn + POP_T1N;
  double* s_dte = s_dts + POP_T1N;
  double* s_dtw = s_dte + POP_T1N;
  double* s_ahf = s_dtw + POP_T1N;
  double* s_dyu = s_ahf + POP_T1N;
  double* s_dxu = s_dyu + POP_T1N;
  double* s_ute = s_dxu + POP_T1N;
  double* s_vtn = s_ute + POP_T1N;
  double* s_d2 = s_vtn + POP_T1N;  // [NTC][T1N]
  int* s_kmt = (int*)(s_d2 + NTC * POP_T1N);  // halo tile (TIX)
  uint64_t* s_bar = (uint64_t*)(s_kmt + POP_TN);

  const GridView& g = a.g;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * POP_BX + tx;
  const int i0 = (g.ib - 1) + blockIdx.x * POP_BX, j0 = (g.jb - 1) + blockIdx.y * POP_BY;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = (i <= g.ie - 1) && (j <= g.je - 1);
  const size_t q = (size_t)j * g.nxb + i, n2 = g.n2;
  const int km = g.km, nxb = g.nxb, nyb = g.nyb;

  auto issue = [&](int kk) {
    const int sl = (kk - 1) % NS;
    double* st = s_stage + (size_t)sl * STAGE;
    mbar_expect_tx(&s_bar[sl], (uint32_t)(NTL * POP_TILE_BYTES + NTC * POP_NTHREADS * 8));
    if (PBC) {
      tma_load_tile(st + 6 * POP_TN, &a.tm_dzt, i0 - POP_H, j0 - POP_H, kk, &s_bar[sl]);
      tma_load_tile(st + 7 * POP_TN, &a.tm_dzu, i0 - POP_H, j0 - POP_H, kk, &s_bar[sl]);
    }
#pragma unroll
    for (int m = 0; m < NTC; m++) {
      const int z = (a.n0 + m) * km + (kk - 1);
      tma_load_tile(st + m * POP_TN, &a.tm_tcur, i0 - POP_H, j0 - POP_H, z, &s_bar[sl]);
      tma_load_tile(st + (NTC + m) * POP_TN, &a.tm_tmix, i0 - POP_H, j0 - POP_H, z, &s_bar[sl]);
      tma_load_tile(st + NTL * POP_TN + m * POP_NTHREADS, &a.tm_vdc, i0, j0, a.vdc_lev0[m] + kk * a.vdc_kstr,
                    &s_bar[sl]);
    }
    tma_load_tile(st + 4 * POP_TN, &a.tm_u, i0 - POP_H, j0 - POP_H, kk - 1, &s_bar[sl]);
    tma_load_tile(st + 5 * POP_TN, &a.tm_v, i0 - POP_H, j0 - POP_H, kk - 1, &s_bar[sl]);
  };
  if (tid == 0) {
    for (int sl = 0; sl < NS; sl++) mbar_init(&s_bar[sl], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int kk = 1; kk <= km && kk <= NS; kk++) issue(kk);

  // ---- k-invariant staging (ring-1 coefficient tiles, KMT halo tile)
  tile_load_i(s_kmt, g.KMT, i0, j0, nxb, nyb, -2, POP_BX + 1, -2, POP_BY + 1, tid);
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int p = tid + s * POP_NTHREADS;
    if (p < POP_T1N) {
      const int jj = p / POP_T1W - 1, ii = p % POP_T1W - 1;
      const int gi = i0 + ii, gj = j0 + jj;
      const bool in = (gi >= 0 && gi < nxb && gj >= 0 && gj < nyb);
      const size_t qq = (size_t)gj * nxb + gi;
      s_dtn[p] = in ? g.DTN[qq] : 0.0;
      s_dts[p] = in ? g.DTS[qq] : 0.0;
      s_dte[p] = in ? g.DTE[qq] : 0.0;
      s_dtw[p] = in ? g.DTW[qq] : 0.0;
      s_ahf[p] = (in && DEL4 && a.lvariable_hmixt) ? g.AHF[qq] : 0.0;
      s_dyu[p] = in ? g.DYU[qq] : 0.0;
      s_dxu[p] = in ? g.DXU[qq] : 0.0;
    }
  }
  const int o1 = TIX1(tx, ty), oT = TIX(tx, ty);
  __syncthreads();
  // own column: registers (from the staged tiles, which are zero outside the block, so that threads whose
  // column lies beyond the physical domain still rebuild D2 of their ghost cell with the right masks)
  const int kmt = s_kmt[oT];
  const int k_n = s_kmt[oT + POP_TW], k_s = s_kmt[oT - POP_TW], k_e = s_kmt[oT + 1], k_w = s_kmt[oT - 1];
  const int kmn = k_n < kmt ? k_n : kmt, kms = k_s < kmt ? k_s : kmt;  // kmX = min(KMT, KMT of that neighbour)
  const int kme = k_e < kmt ? k_e : kmt, kmw = k_w < kmt ? k_w : kmt;
  const double o_dtn = s_dtn[o1], o_dts = s_dts[o1], o_dte = s_dte[o1], o_dtw = s_dtw[o1], o_ahf = s_ahf[o1];
  const double tarea_r = active ? g.TAREA_R[q] : 0.0;
  // per-level intermediates of level kk from its stage: UTE, VTN (advection.F90:2031-2060) and, for
  // del4, D2 = AHF*L(TMIX) (hmix_del4.F90:1025-1046) on the ring-1 tile
  // partial bottom cells: thickness of the U cell (gi,gj) at level kk (0 outside the block), and the mixing
  // coefficients of a T cell scaled by the face thicknesses (hmix_del2.F90:1034-1062, hmix_del4.F90:964-988)
  // (st: the stage of that level; its DZT / DZU halo tiles are zero outside the block like every staged tile)
  auto zu_at = [&](const double* st, int ii, int jj) { return st[7 * POP_TN + TIX(ii, jj)]; };
  auto pbc_scale = [&](const double* st, int ii, int jj, double& dn, double& ds, double& de, double& dw) {
    const int gi = i0 + ii, gj = j0 + jj;
    if (gi >= 1 && gi <= nxb - 2 && gj >= 1 && gj <= nyb - 2) {
      const double* dzt = st + 6 * POP_TN + TIX(ii, jj);
      const double z = dzt[0];
      const RcpD rz = rcp_prepare(z);
      dn = div_by(dn * fmin(z, dzt[POP_TW]), rz);
      ds = div_by(ds * fmin(z, dzt[-POP_TW]), rz);
      de = div_by(de * fmin(z, dzt[1]), rz);
      dw = div_by(dw * fmin(z, dzt[-1]), rz);
    } else {
      dn = ds = de = dw = 0.0;
    }
  };
  double p_dn = o_dtn, p_ds = o_dts, p_de = o_dte, p_dw = o_dtw;  // PBC: own-point coefficients of the level pre() saw last
  auto pre = [&](int kk, const double* st) {
    const double* su = st + 4 * POP_TN;
    const double* sv = st + 5 * POP_TN;
    if (PBC) {  // advection.F90:2040-2066: (U*DYU)*DZU
      s_ute[o1] = 0.5 * (su[oT] * s_dyu[o1] * zu_at(st, tx, ty) + su[oT - POP_TW] * s_dyu[o1 - POP_T1W] * zu_at(st, tx, ty - 1));
      s_vtn[o1] = 0.5 * (sv[oT] * s_dxu[o1] * zu_at(st, tx, ty) + sv[oT - 1] * s_dxu[o1 - 1] * zu_at(st, tx - 1, ty));
      if (tid < POP_BY) {
        const int t1 = TIX1(-1, tid), tt = TIX(-1, tid);
        s_ute[t1] = 0.5 * (su[tt] * s_dyu[t1] * zu_at(st, -1, tid) + su[tt - POP_TW] * s_dyu[t1 - POP_T1W] * zu_at(st, -1, tid - 1));
      } else if (tid >= 32 && tid < 32 + POP_BX) {
        const int t1 = TIX1(tid - 32, -1), tt = TIX(tid - 32, -1);
        s_vtn[t1] = 0.5 * (sv[tt] * s_dxu[t1] * zu_at(st, tid - 32, -1) + sv[tt - 1] * s_dxu[t1 - 1] * zu_at(st, tid - 33, -1));
      }
      p_dn = o_dtn; p_ds = o_dts; p_de = o_dte; p_dw = o_dtw;
      pbc_scale(st, tx, ty, p_dn, p_ds, p_de, p_dw);
    } else {
    // own point
    s_ute[o1] = 0.5 * (su[oT] * s_dyu[o1] + su[oT - POP_TW] * s_dyu[o1 - POP_T1W]);
    s_vtn[o1] = 0.5 * (sv[oT] * s_dxu[o1] + sv[oT - 1] * s_dxu[o1 - 1]);
    // the west column of UTE and the south row of VTN
    if (tid < POP_BY) {
      const int t1 = TIX1(-1, tid), tt = TIX(-1, tid);
      s_ute[t1] = 0.5 * (su[tt] * s_dyu[t1] + su[tt - POP_TW] * s_dyu[t1 - POP_T1W]);
    } else if (tid >= 32 && tid < 32 + POP_BX) {
      const int t1 = TIX1(tid - 32, -1), tt = TIX(tid - 32, -1);
      s_vtn[t1] = 0.5 * (sv[tt] * s_dxu[t1] + sv[tt - 1] * s_dxu[t1 - 1]);
    }
    }
    if (DEL4) {
      {  // own point: coefficients from registers
        const double cn = (kk <= kmn) ? (PBC ? p_dn : o_dtn) : 0.0, cs = (kk <= kms) ? (PBC ? p_ds : o_dts) : 0.0;
        const double ce = (kk <= kme) ? (PBC ? p_de : o_dte) : 0.0, cw = (kk <= kmw) ? (PBC ? p_dw : o_dtw) : 0.0;
        const double cc = -(cn + cs + ce + cw);
#pragma unroll
        for (int m = 0; m < NTC; m++) {
          const double* t = st + (NTC + m) * POP_TN + oT;
          const double v = cc * t[0] + cn * t[POP_TW] + cs * t[-POP_TW] + ce * t[1] + cw * t[-1];
          s_d2[m * POP_T1N + o1] = a.lvariable_hmixt ? o_ahf * v : v;
        }
      }
      constexpr int NRING = 2 * POP_T1W + 2 * POP_BY;  // the ring around the CTA's columns
      if (tid < NRING) {
        int ii, jj;
        if (tid < 2 * POP_T1W) { ii = tid % POP_T1W - 1; jj = (tid < POP_T1W) ? -1 : POP_BY; }
        else { const int t = tid - 2 * POP_T1W; ii = (t < POP_BY) ? -1 : POP_BX; jj = t % POP_BY; }
        const int t1 = TIX1(ii, jj), tt = TIX(ii, jj);
        const int kc = s_kmt[tt];
        double rn = s_dtn[t1], rs = s_dts[t1], re = s_dte[t1], rw = s_dtw[t1];
        if (PBC) pbc_scale(st, ii, jj, rn, rs, re, rw);
        const double cn = (kk <= s_kmt[tt + POP_TW] && kk <= kc) ? rn : 0.0;
        const double cs = (kk <= s_kmt[tt - POP_TW] && kk <= kc) ? rs : 0.0;
        const double ce = (kk <= s_kmt[tt + 1] && kk <= kc) ? re : 0.0;
        const double cw = (kk <= s_kmt[tt - 1] && kk <= kc) ? rw : 0.0;
        const double cc = -(cn + cs + ce + cw);
#pragma unroll
        for (int m = 0; m < NTC; m++) {
          const double* t = st + (NTC + m) * POP_TN + tt;
          const double v = cc * t[0] + cn * t[POP_TW] + cs * t[-POP_TW] + ce * t[1] + cw * t[-1];
          s_d2[m * POP_T1N + t1] = a.lvariable_hmixt ? s_ahf[t1] * v : v;
        }
      }
    }
  };
  // ---- carried state
  double wtk = 0.0, tc_m[NTC], told_c[NTC], vtf[NTC], lw_next[NTC];
#pragma unroll
  for (int m = 0; m < NTC; m++)  // lw_lim tendency of level 1 (requested before the first wait)
    lw_next[m] = (a.lw[m] && active) ? a.OUT[((size_t)(a.n0 + m) * km) * n2 + q] : 0.0;
  mbar_wait(&s_bar[0], 0u);
  __syncthreads();  // invariants staged
#pragma unroll
  for (int m = 0; m < NTC; m++) {
    tc_m[m] = 0.0;
    vtf[m] = 0.0;
    told_c[m] = s_stage[(NTC + m) * POP_TN + oT];  // TOLD (== TMIX) at level 1
  }
  if (active) wtk = a.DH[q];
  pre(1, s_stage);
  __syncthreads();

  for (int k = 1; k <= km; k++) {
    const int slot = (k - 1) % NS;
    const double* st = s_stage + (size_t)slot * STAGE;
    const bool have_next = (k < km);
    const int nslot = k % NS;
    const double* nst = s_stage + (size_t)nslot * STAGE;
    if (have_next) mbar_wait(&s_bar[nslot], (uint32_t)((k / NS) & 1));
    if (active) {
      const double ute = s_ute[o1], utw = s_ute[o1 - 1], vtn = s_vtn[o1], vts = s_vtn[o1 - POP_T1W];
      double wtkb = 0.0;
      if (k < km) {
        const double FC = (vtn - vts + ute - utw) * tarea_r;
        if (PBC) wtkb = (k < kmt) ? wtk + FC : 0.0;  // advection.F90:2110-2111
        else wtkb = (k < kmt) ? wtk + c_vc.dz[k] * FC : 0.0;
      }
      const double cn = (k <= kmn) ? (PBC ? p_dn : o_dtn) : 0.0, cs = (k <= kms) ? (PBC ? p_ds : o_dts) : 0.0;
      const double ce = (k <= kme) ? (PBC ? p_de : o_dte) : 0.0, cw = (k <= kmw) ? (PBC ? p_dw : o_dtw) : 0.0;
      const double cc = -(cn + cs + ce + cw);
      const double dzt_c = PBC ? st[6 * POP_TN + oT] : 0.0;
      const RcpD rzt = PBC ? rcp_prepare(dzt_c) : RcpD{1.0, 1.0};
      const double h_dzt = PBC ? div_by(0.5, rzt) : 0.0;
      const double dzt_p = PBC ? ((k < km) ? nst[6 * POP_TN + oT] : dzt_c) : 0.0;  // DZT(min(k+1, km))
#pragma unroll
      for (int m = 0; m < NTC; m++) {
        const int n = a.n0 + m;
        const size_t lev = ((size_t)n * km + (k - 1)) * n2 + q;
        // ---- horizontal mixing
        double hd;
        if (DEL4) {
          const double* d = s_d2 + m * POP_T1N + o1;
          hd = a.ah * (cc * d[0] + cn * d[POP_T1W] + cs * d[-POP_T1W] + ce * d[1] + cw * d[-1]);
        } else {
          const double* t = st + (NTC + m) * POP_TN + oT;
          hd = a.ah * (cc * t[0] + cn * t[POP_TW] + cs * t[-POP_TW] + ce * t[1] + cw * t[-1]);
        }
        // ---- centred advection: advection.F90:2243-2301
        const double* tc = st + m * POP_TN + oT;
        const double T = tc[0];
        const double Tp = (k < km) ? nst[m * POP_TN + oT] : 0.0;
        double L;
        if (a.lw[m]) {  // lw_lim (advection.F90:2684-3280): computed by lw_lim_kernel, one level requested ahead
          L = (k <= kmt) ? lw_next[m] : 0.0;
          if (k < km) lw_next[m] = a.OUT[lev + n2];
        } else {
          L = 0.5 *
              ((vtn - vts + ute - utw) * T + vtn * tc[POP_TW] - vts * tc[-POP_TW] + ute * tc[1] - utw * tc[-1]) *
              tarea_r;
          if (PBC) L = div_by(L, rzt);  // advection.F90:2223-2238
          if (k == 1) {
            if (!a.varthick) L = L + c_vc.dzr[k] * wtk * T;
          } else {
            if (PBC) L = L + h_dzt * wtk * (tc_m[m] + T);  // :2278-2280
            else L = L + c_vc.dz2r[k] * wtk * (tc_m[m] + T);
          }
          if (k < km) {
            if (PBC) L = L - h_dzt * wtkb * (T + Tp);
            else L = L - c_vc.dz2r[k] * wtkb * (T + Tp);
          }
        }
        tc_m[m] = T;
        // ---- vertical diffusion (top/bottom fluxes): vertical_mix.F90:779-838
        const double vdc = st[NTL * POP_TN + m * POP_NTHREADS + tid];
        const double told_p = (k < km) ? nst[(NTC + m) * POP_TN + oT] : told_c[m];
        if (k == 1) vtf[m] = (kmt >= 1) ? a.STF[(size_t)n * a.stf_ns + (size_t)(j - a.stf_j0) * a.stf_pitch + (i - a.stf_i0)] : 0.0;
        double VTFB, vd;
        if constexpr (PBC) {  // vertical_mix.F90:790-803
          const double wz = 0.5 * (dzt_c + dzt_p);
          VTFB = (kmt > k) ? vdc * (told_c[m] - told_p) / wz : 0.0;
          vd = (k <= kmt) ? div_by(vtf[m] - VTFB, rzt) : 0.0;
        } else {
          VTFB = (kmt > k) ? vdc * (told_c[m] - told_p) * c_vc.dzwr[k] : 0.0;
          vd = (k <= kmt) ? (vtf[m] - VTFB) * c_vc.dzr[k] : 0.0;
        }
        vtf[m] = VTFB;
        told_c[m] = told_p;
        // ---- tracer_update: baroclinic.F90:1993-2300 (implicit vertical mixing)
        double FT = 0.0 + hd;
        FT = FT - L;
        FT = FT + vd;
        if (k == 1 && a.varthick) FT = FT + c_vc.dzr[1] * a.TFW[(size_t)n * n2 + q];
        FT = FT + 0.0;
        double* out = a.OUT + lev;
        if (a.predictor && k == 1 && n < 2) {
          if (kmt > 0) *out = c_vc.c2dtt[1] * FT - 2.0 * T * (a.PCUR[q] - a.POLD[q]) / (POP_GRAV * c_vc.dz[1]);
        } else {
          *out = (k <= kmt) ? c_vc.c2dtt[k] * FT : 0.0;
        }
      }
      wtk = wtkb;
    }
    __syncthreads();  // every thread is done with ring slot k and with the intermediates of level k
    if (tid == 0 && k + NS <= km) issue(k + NS);
    if (have_next) {
      pre(k + 1, nst);
      __syncthreads();
    }
  }
}

static int launch_tracer_fast(const TracerFastArgs& a, bool del4) {
  void (*kfn)(const TracerFastArgs);
  if (a.g.DZT) kfn = del4 ? tracer_fast_kernel<true, true> : tracer_fast_kernel<false, true>;
  else kfn = del4 ? tracer_fast_kernel<true, false> : tracer_fast_kernel<false, false>;
  const bool pbc = (a.g.DZT != nullptr);
  const size_t stage = (size_t)(pbc ? 8 : 6) * POP_TN + (size_t)NTC * POP_NTHREADS;
  const size_t smem = sizeof(double) * ((size_t)(pbc ? TF_NS_PBC : TF_NS) * stage + (size_t)TF_FIXED * POP_T1N) +
                      sizeof(int) * POP_TN + 8 * TF_NS;
#ifndef POP_EMUL
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
  POP_LAUNCH(kfn, col_grid(G.nxg, G.ny_local), col_block(), smem, a);
  return POP_SUCCESS;
}

static size_t tracer_smem_bytes(bool tma) {
  const int ns = tma ? TR_NS : 1;
  return sizeof(double) * POP_TN * (ns * (2 * NTC + 2) + 7 + (tma ? 2 : 1) * NTC) + sizeof(int) * POP_TN + 8 * TR_NS;
}

template <int MODE>
static int launch_tracer(const TracerArgs& a, bool del4, bool upw, bool tma) {
  void (*kfn)(const TracerArgs) = nullptr;
  if (tma) {
    if (del4 && upw) kfn = tracer_column_kernel<MODE, true, true, true>;
    else if (del4) kfn = tracer_column_kernel<MODE, true, false, true>;
    else if (upw) kfn = tracer_column_kernel<MODE, false, true, true>;
    else kfn = tracer_column_kernel<MODE, false, false, true>;
  } else {
    if (del4 && upw) kfn = tracer_column_kernel<MODE, true, true, false>;
    else if (del4) kfn = tracer_column_kernel<MODE, true, false, false>;
    else if (upw) kfn = tracer_column_kernel<MODE, false, true, false>;
    else kfn = tracer_column_kernel<MODE, false, false, false>;
  }
  const size_t smem = tracer_smem_bytes(tma);
#ifndef POP_EMUL
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
  POP_LAUNCH(kfn, col_grid(G.nxg, G.ny_local), col_block(), smem, a);
  return POP_SUCCESS;
}

int tracer_column(int mode, int k, const TracerIO& io) {
  TracerArgs a;
  memset(&a, 0, sizeof(a));
  a.g = grid_view();
  a.TCUR = io.TCUR; a.TMIX = io.TMIX; a.TOLD = io.TOLD; a.UCUR = io.UCUR; a.VCUR = io.VCUR;
  a.STF = io.STF; a.TFW = io.TFW; a.DH = io.DH; a.POLD = io.POLD; a.PCUR = io.PCUR;
  if (io.stf_strip) { a.stf_ns = (size_t)G.nxg * G.ny_local; a.stf_pitch = G.nxg; a.stf_i0 = G.ib - 1; a.stf_j0 = G.jb - 1; }
  else { a.stf_ns = G.n2; a.stf_pitch = G.nxb; a.stf_i0 = 0; a.stf_j0 = 0; }
  a.OUT = io.TNEW;
  a.WTK = io.WTK;
  a.VTF = fld("VTF");
  a.AUX = fld("AUX");
  a.hmix = G.cfg.hmix_tracer_itype;
  const bool gm = (a.hmix == POP_HMIX_GM);
  if (gm) {
    POP_REQUIRE(mode == TR_FULL, "tracer_column: GM tendencies come from gm_tendency_dev");
    POP_TRY(gm_tendency_dev(io.TMIX));  // hdifft_gm for every level; also VDC += VDC_GM before the column kernel reads VDC
    a.HDT = fld("GM_HDT");
  }
  a.lvariable_hmixt = G.cfg.lvariable_hmixt;
  a.varthick = (G.cfg.sfc_layer_type == POP_SFC_VARTHICK);
  a.implicit_vmix = G.cfg.implicit_vertical_mix;
  a.predictor = (a.varthick && G.cfg.lpressure_avg && G.leapfrogts);
  a.ah = G.ah;
  // lw_lim needs the flux velocities two cells out (comp_flux_vel_ghost, baroclinic.F90:667): the fused driver makes
  // them here, the slab entry point pop_advt expects pop_comp_flux_vel_ghost to have been called
  if (mode == TR_FULL && G.use_lw_lim) POP_TRY(lw_flux_prepare_dev(io.UCUR, io.VCUR, io.DH));
  if (mode == TR_FULL) { a.k0 = 1; a.k1 = G.km; }
  else {
    POP_REQUIRE(k >= 1 && k <= G.km, "tracer slab operator: k=%d out of range", k);
    a.k0 = a.k1 = k;
  }
  const bool del4 = (G.cfg.hmix_tracer_itype == POP_HMIX_DEL4);
  for (int n0 = 0; n0 < G.nt; n0 += NTC) {
    a.n0 = n0;
    a.nn = (G.nt - n0 < NTC) ? G.nt - n0 : NTC;
    bool upw = false, lwl = false;
    int lw_slots[NTC];
    for (int m = 0; m < NTC; m++) lw_slots[m] = -1;
    for (int m = 0; m < a.nn; m++) {
      a.adv[m] = G.cfg.tadvect_itype[n0 + m];
      if (a.adv[m] == POP_TADVECT_UPWIND3) upw = true;
      if (a.adv[m] == POP_TADVECT_LW_LIM) { lwl = true; lw_slots[m] = n0 + m; }
    }
    a.lwl = 0;
    if (lwl && (mode == TR_FULL || mode == TR_ADVT)) {  // X of lw_lim is the mix-time field (advection.F90:2806)
      if (mode == TR_FULL) POP_TRY(lw_lim_dev(lw_slots, io.TMIX, a.k0, a.k1, a.OUT, G.n3, G.n2));
      else POP_TRY(lw_lim_dev(lw_slots, io.TMIX, a.k0, a.k1, a.OUT, G.n2, 0));
      a.lwl = 1;
    }
    switch (mode) {
      case TR_FULL: {
        // fast path: a full pair of centred tracers, implicit vertical mixing, leapfrog-type levels
        const bool fast_ok = !gm && !(G.cfg.partial_bottom_cells && G.no_pbc_fast) && !G.no_tma && !G.no_fast_tracer && a.nn == NTC && !upw && a.implicit_vmix &&
                             a.TMIX == a.TOLD && a.TMIX != a.TCUR && (G.nxb % 2) == 0 && G.km >= TF_NS;
        if (fast_ok) {
          TracerFastArgs f;
          memset(&f, 0, sizeof(f));
          f.stf_ns = a.stf_ns; f.stf_pitch = a.stf_pitch; f.stf_i0 = a.stf_i0; f.stf_j0 = a.stf_j0;
          f.g = a.g; f.STF = a.STF; f.TFW = a.TFW; f.DH = a.DH; f.POLD = a.POLD; f.PCUR = a.PCUR; f.OUT = a.OUT;
          for (int m = 0; m < NTC; m++) f.lw[m] = (a.adv[m] == POP_TADVECT_LW_LIM) ? 1 : 0;
          f.n0 = n0; f.lvariable_hmixt = a.lvariable_hmixt; f.varthick = a.varthick; f.predictor = a.predictor;
          f.ah = a.ah;
          f.vdc_kstr = (G.vdc_nk == 1) ? 0 : 1;
          for (int m = 0; m < NTC; m++) {
            const int n = n0 + m;
            const int mt2 = (n + 1 < G.vdc_nd) ? n + 1 : G.vdc_nd;
            f.vdc_lev0[m] = (mt2 - 1) * G.vdc_nk + ((G.vdc_nk == 1) ? 0 : -G.vdc_k0);
          }
          if (make_tmap(&f.tm_tcur, a.TCUR, G.km * G.nt) && make_tmap(&f.tm_tmix, a.TMIX, G.km * G.nt) &&
              make_tmap(&f.tm_u, a.UCUR, G.km) && make_tmap(&f.tm_v, a.VCUR, G.km) &&
              make_tmap_box(&f.tm_vdc, a.g.VDC, G.vdc_nk * G.vdc_nd, POP_BX, POP_BY) &&
              (!a.g.DZT || (make_tmap(&f.tm_dzt, a.g.DZT, G.km + 2) && make_tmap(&f.tm_dzu, a.g.DZU, G.km + 2)))) {
            POP_TRY(launch_tracer_fast(f, del4));
            break;
          }
        }
        const bool tma = !G.no_tma && make_tmap(&a.tm_tcur, a.TCUR, G.km * G.nt) &&
                         make_tmap(&a.tm_tmix, a.TMIX, G.km * G.nt) && make_tmap(&a.tm_u, a.UCUR, G.km) &&
                         make_tmap(&a.tm_v, a.VCUR, G.km);
        POP_TRY(launch_tracer<TR_FULL>(a, del4, upw, tma));
        break;
      }
      case TR_ADVT: {
        // every pass restarts from the caller's WTK; only the last pass stores WTKB back
        TracerArgs b = a;
        if (n0 + NTC < G.nt) b.WTK = fld("WTK_C");  // scratch copy so the input survives
        if (n0 + NTC < G.nt)
          POP_CHECK_CUDA(cudaMemcpyAsync(b.WTK, a.WTK, sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
        POP_TRY(launch_tracer<TR_ADVT>(b, false, upw, false));
        break;
      }
      case TR_HDIFFT: POP_TRY(launch_tracer<TR_HDIFFT>(a, del4, false, false)); break;
      case TR_VDIFFT: POP_TRY(launch_tracer<TR_VDIFFT>(a, false, false, false)); break;
      default: POP_REQUIRE(false, "tracer_column: bad mode %d", mode);
    }
  }
  return pop_post_launch("tracer_column");
}

// =====================================================================================
// impvmixt / impvmixt_correct: vertical_mix.F90:1164-1382, :1460-1672
// =====================================================================================
// One thread per physical column, 128 columns (one row segment) per CTA.  The Thomas coefficients
// E(k) of the CTA's columns live in shared memory (km x 128 doubles, conflict-free), the eliminated
// right-hand side F(k) is streamed in place through the output array (it is re-read by the back
// substitution while still L2-resident), and every global load of a chunk of IV_CH levels is issued
// one chunk ahead of the (division-bound) recurrence that consumes it, so the dependent chain never
// waits on DRAM.  Level addresses are running pointers (one IMAD.WIDE per level and array; the first
// version recomputed (k-1)*n2 in 64 bits per access, which was half of the kernel's instructions:
// profiles/r1_ncu_summary.md), full chunks carry no bounds predicates, and the two quotients of a level
// share one reciprocal refinement (div_by: the compiler's own division sequence, bit-identical).
// ncu (profiles/): 12 warps/SM (E(k) footprint), HBM-bound target 64 B/cell for two tracers.
#ifndef IV_CH
#define IV_CH 4   // levels per chunk
#endif
#ifndef IV_NB
#define IV_NB 4   // register ring depth: IV_NB - 1 chunks of loads are in flight ahead of the recurrence
#endif
#ifndef IV_MINB
#define IV_MINB 3
#endif
#ifndef IV_PD
#define IV_PD 8   // L2 prefetch distance in levels (0: off)
#endif
#define IV_THREADS 128
template <bool CORRECT, bool PBC>
__global__ void __launch_bounds__(IV_THREADS, IV_MINB)
impvmixt_kernel(GridView g, double* __restrict__ TNEW, const double* __restrict__ TOLD,
                const double* __restrict__ PSFC, const double* __restrict__ RHS, double* FB,
                int nfirst, int nlast, int varthick) {
  POP_DYN_SMEM(smem_raw);
  double* sE = (double*)smem_raw + threadIdx.x;  // E(k) at sE[(k-1)*IV_THREADS]
  const int i = (g.ib - 1) + blockIdx.x * IV_THREADS + threadIdx.x;
  const int j = (g.jb - 1) + blockIdx.y;
  if (i > g.ie - 1 || j > g.je - 1) return;
  const size_t q = (size_t)j * g.nxb + i, n2 = g.n2;
  const int n2i = (int)n2;  // elements per level (< 2^31): the level stride of every running pointer
  const int km = g.km, kmt = g.KMT[q];
  const double hfac1 = c_vc.hfac_t[1];
  const double H1 = varthick ? hfac1 + PSFC[q] / (POP_GRAV * c_vc.c2dtt[1]) : hfac1;
  const int vstr = (g.vdc_nk == 1) ? 0 : n2i;
  for (int n = nfirst; n <= nlast; n++) {  // 1-based tracer index
    const int mt2 = (n < g.vdc_nd) ? n : g.vdc_nd;
    // VDC(:,:,k,mt2) = VDC1[(k-1) * vstr] (a single level when vdc_nk == 1)
    const double* VDC1 = g.VDC + (size_t)(mt2 - 1) * g.vdc_nk * n2 + q + (size_t)(1 - g.vdc_k0) * n2;
    double* Tn = TNEW + (size_t)(n - 1) * km * n2 + q;
    const double* Bs = CORRECT ? Tn : TOLD + (size_t)(n - 1) * km * n2 + q;  // what the solution is added to
    double* Fb = CORRECT ? FB + q : Tn;
    double A, B, C, Fm;
    {
      A = c_vc.afac_t[1] * VDC1[0];
      const RcpD rd = rcp_prepare(H1 + A);
      const double e = div_by(A, rd);
      sE[0] = e;
      B = H1 * e;
      const double R1 = CORRECT ? RHS[(size_t)(n - 1) * n2 + q] : Tn[0];
      Fm = div_by(hfac1 * R1, rd);
      Fb[0] = Fm;
    }
    // ---- forward elimination, levels 2..km: full chunks of IV_CH levels through the register ring (loads IV_NB-1
    // chunks ahead, L2 prefetch IV_PD levels ahead), then the < IV_CH remaining levels one by one
    const double* pv = VDC1 + vstr;  // VDC of the next level to load (level 2)
    const double* pr = Tn + n2i;     // right-hand side of the next level to load
    double* pf = Fb + n2i;           // F of the next level to store
    const int nfull = (km - 1) / IV_CH;  // full chunks of either sweep
    int nld = nfull;                     // chunks not loaded yet
    double rb[IV_NB][IV_CH], vb[IV_NB][IV_CH];  // ring of chunk buffers (static indices after unrolling)
#pragma unroll
    for (int s = 0; s < IV_NB; s++)
#pragma unroll
      for (int c = 0; c < IV_CH; c++) { rb[s][c] = 0.0; vb[s][c] = 0.0; }
    const ptrdiff_t pdv = (ptrdiff_t)IV_PD * vstr, pdn = (ptrdiff_t)IV_PD * n2i;
    int kpf = 2 + IV_PD;  // level the next forward L2 prefetch targets
    auto load_fwd = [&](double* v, double* r) {
      if (nld > 0) {
        if (IV_PD && kpf + IV_CH - 1 <= km) {
#pragma unroll
          for (int c = 0; c < IV_CH; c++) {
            if (vstr) prefetch_l2(pv + pdv + (ptrdiff_t)c * vstr);
            if (!CORRECT) prefetch_l2(pr + pdn + (ptrdiff_t)c * n2i);
          }
        }
        kpf += IV_CH;
#pragma unroll
        for (int c = 0; c < IV_CH; c++) {
          v[c] = *pv;
          pv += vstr;
          if (!CORRECT) { r[c] = *pr; pr += n2i; }
        }
        nld--;
      }
    };
    // partial bottom cells (vertical_mix.F90:1279-1286, :1577-1582): the thicknesses ride two levels ahead of the
    // recurrence in registers (zc = DZT(k), zn = DZT(k+1), zf = DZT(k+2)) behind an L2 prefetch
    const double* pz = PBC ? g.DZT + q + (size_t)5 * n2 : nullptr;  // next level to load (level 5)
    double zc = 0.0, zn = 0.0, zf = 0.0;
    if (PBC) {
      const double* z2 = g.DZT + q + (size_t)2 * n2;
      zc = z2[0];
      zn = (3 <= km + 1) ? z2[n2] : 0.0;
      zf = (4 <= km + 1) ? z2[2 * n2] : 0.0;
    }
    auto fwd_level = [&](int k, double vdc, double rhs) {
      C = A;
      double hfac;
      if (PBC) {  // level 1 keeps the full-cell A and hfac (:1263-1269); DZT(k+1) is read at k = km too
        const double zk = zc;
        A = g.aidif * vdc / (0.5 * (zk + zn));
        hfac = zk / c_vc.c2dtt[k];
        zc = zn; zn = zf;
        if (k + 3 <= km + 1) {
          zf = *pz;
          if (IV_PD && k + 3 + IV_PD <= km + 1) prefetch_l2(pz + (ptrdiff_t)IV_PD * n2i);
        }
        pz += n2i;
      } else {
        A = c_vc.afac_t[k] * vdc;
        hfac = c_vc.hfac_t[k];
      }
      double F;
      if (k > kmt) {
        F = 0.0;
      } else {
        const RcpD rd = rcp_prepare((k == kmt) ? hfac + B : hfac + A + B);
        const double e = div_by(A, rd);
        sE[(k - 1) * IV_THREADS] = e;
        B = (hfac + B) * e;
        F = CORRECT ? div_by(C * Fm, rd) : div_by(hfac * rhs + C * Fm, rd);
      }
      *pf = F;
      pf += n2i;
      Fm = F;
    };
#pragma unroll
    for (int s = 0; s < IV_NB - 1; s++) load_fwd(vb[s], rb[s]);
    int k = 2;
    for (int ch = 0; ch < nfull; ch += IV_NB) {
#pragma unroll
      for (int s = 0; s < IV_NB; s++)
        if (ch + s < nfull) {
          load_fwd(vb[(s + IV_NB - 1) % IV_NB], rb[(s + IV_NB - 1) % IV_NB]);
#pragma unroll
          for (int c = 0; c < IV_CH; c++) fwd_level(k + c, vb[s][c], rb[s][c]);
          k += IV_CH;
        }
    }
    for (; k <= km; k++) {  // the last (km-1) % IV_CH levels
      const double vdc = *pv, rhs = CORRECT ? 0.0 : *pr;
      pv += vstr;
      pr += n2i;
      fwd_level(k, vdc, rhs);
    }
    // ---- back substitution + final update, levels km..1 (Fm = F(km)): the same ring going up
    double Fp = Fm;
    const size_t top = (size_t)(km - 1) * n2;
    if (IV_PD && !CORRECT) {  // TOLD is touched for the first time on the way up
#pragma unroll
      for (int c = 0; c < IV_PD; c++)
        if (c < km) prefetch_l2(Bs + top - (size_t)c * n2);
    }
    Tn[top] = Bs[top] + Fp;
    const double* qf = Fb + top - n2i;  // F of the next level to load (level km-1), going up
    const double* qb = Bs + top - n2i;
    double* qt = Tn + top - n2i;        // next level to store
    nld = nfull;
    kpf = km - 1 - IV_PD;               // level the next backward L2 prefetch targets
    auto load_bwd = [&](double* f, double* b) {
      if (nld > 0) {
        if (IV_PD && !CORRECT && kpf - IV_CH + 1 >= 1) {
#pragma unroll
          for (int c = 0; c < IV_CH; c++) prefetch_l2(qb - pdn - (ptrdiff_t)c * n2i);
        }
        kpf -= IV_CH;
#pragma unroll
        for (int c = 0; c < IV_CH; c++) {
          f[c] = *qf;
          b[c] = *qb;
          qf -= n2i;
          qb -= n2i;
        }
        nld--;
      }
    };
    auto bwd_level = [&](int kk, double f, double b) {
      double F = f;
      if (kk < kmt) F = F + sE[(kk - 1) * IV_THREADS] * Fp;
      *qt = b + F;
      qt -= n2i;
      Fp = F;
    };
#pragma unroll
    for (int s = 0; s < IV_NB - 1; s++) load_bwd(rb[s], vb[s]);
    k = km - 1;
    for (int ch = 0; ch < nfull; ch += IV_NB) {
#pragma unroll
      for (int s = 0; s < IV_NB; s++)
        if (ch + s < nfull) {
          load_bwd(rb[(s + IV_NB - 1) % IV_NB], vb[(s + IV_NB - 1) % IV_NB]);
#pragma unroll
          for (int c = 0; c < IV_CH; c++) bwd_level(k - c, rb[s][c], vb[s][c]);
          k -= IV_CH;
        }
    }
    for (; k >= 1; k--) {
      const double f = *qf, b2 = *qb;
      qf -= n2i;
      qb -= n2i;
      bwd_level(k, f, b2);
    }
  }
}

// ---- the same solve with the operands staged by TMA (opt-in, POP_B200_THOMAS_TMA=1: the north_star's "warp-cooperative
// path"; bit-identical, but measured SLOWER than the register-ring kernel above on B200: with E(k) and the ring in shared
// memory only 8 warps fit per SM, and the per-chunk barrier + mbarrier wait cost more than the DRAM latency they hide) --
// A warp owns 32 neighbouring columns, a CTA 128 (one 1 KB run per level and array).  One elected thread streams the
// level rows of the CTA -- VDC and the right-hand side on the way down, the field the solution is added to on the way
// up -- as 128 x 1 x 1 TMA boxes into a TVS-deep shared-memory ring, chunk by chunk, across the sweep and tracer
// boundaries (the whole sequence of chunks is known up front), with one mbarrier per ring slot; the recurrence reads
// its operands from shared memory with immediate offsets: no address arithmetic, no scoreboard wait on DRAM in the
// dependent chain.  E(k) stays in shared memory, F streams through the output array as before (re-read from L2 by
// plain loads one chunk ahead: a TMA read of lines this CTA has just written would need a cross-proxy fence).
// Shared memory per CTA: km + 2*TVS*TVC KB (62 + 40 KB): two CTAs per SM.
#ifndef TVC
#define TVC 4   // levels per chunk
#endif
#ifndef TVS
#define TVS 5   // ring slots
#endif
struct IvTmaArgs {
  GridView g;
  double* TNEW;
  const double *TOLD, *PSFC, *RHS;
  double* FB;
  int nfirst, nlast, varthick;
  PopTmap tmT, tmO, tmV;  // TNEW, TOLD (nxb, nyb, km*nt); VDC (nxb, nyb, vdc_nd*vdc_nk): boxes of IV_THREADS x 1 x 1
};
template <bool CORRECT>
__global__ void __launch_bounds__(IV_THREADS, 2)
impvmixt_tma_kernel(const POP_GRID_CONSTANT IvTmaArgs a) {
  POP_DYN_SMEM(smem_raw);
  const GridView& g = a.g;
  const int km = g.km, tid = threadIdx.x;
  double* sEb = (double*)smem_raw;                       // E(k): [km][IV_THREADS]
  double* ring = sEb + (size_t)km * IV_THREADS;          // [TVS][TVC][2][IV_THREADS]
  uint64_t* bar = (uint64_t*)(ring + (size_t)TVS * TVC * 2 * IV_THREADS);
  double* sE = sEb + tid;
  const int i0 = (g.ib - 1) + blockIdx.x * IV_THREADS;
  const int j = (g.jb - 1) + blockIdx.y;
  const int i = i0 + tid;
  const bool active = (i <= g.ie - 1);
  const size_t n2 = g.n2;
  const int n2i = (int)n2;
  const size_t q = (size_t)j * g.nxb + (active ? i : i0);
  const int kmt = active ? g.KMT[q] : 0;
  const double hfac1 = c_vc.hfac_t[1];
  const double H1 = a.varthick ? hfac1 + a.PSFC[q] / (POP_GRAV * c_vc.c2dtt[1]) : hfac1;
  const int nch = (km - 1 + TVC - 1) / TVC;   // chunks per sweep (levels 2..km down, km-1..1 up)
  const int ntr = a.nlast - a.nfirst + 1;
  const int total = ntr * 2 * nch;            // chunks of the whole kernel, in the order they are consumed
  // ---- producer (thread 0): chunk gi of the sequence -> ring slot gi % TVS
  auto issue = [&](int gi) {
    const int tr = gi / (2 * nch), r = gi % (2 * nch);
    const int n = a.nfirst + tr;  // 1-based tracer
    const int slot = gi % TVS;
    double* st = ring + (size_t)slot * TVC * 2 * IV_THREADS;
    if (r < nch) {  // way down: levels k0 .. k0+TVC-1
      const int k0 = 2 + r * TVC;
      const int nl = (km - k0 + 1 < TVC) ? km - k0 + 1 : TVC;
      mbar_expect_tx(&bar[slot], (uint32_t)(nl * (CORRECT ? 1 : 2) * IV_THREADS * 8));
      const int mt2 = (n < g.vdc_nd) ? n : g.vdc_nd;
      for (int c = 0; c < nl; c++) {
        const int k = k0 + c;
        const int zv = (mt2 - 1) * g.vdc_nk + ((g.vdc_nk == 1) ? 0 : k - g.vdc_k0);
        tma_load_tile(st + (size_t)(c * 2) * IV_THREADS, &a.tmV, i0, j, zv, &bar[slot]);
        if (!CORRECT) tma_load_tile(st + (size_t)(c * 2 + 1) * IV_THREADS, &a.tmT, i0, j, (n - 1) * km + (k - 1), &bar[slot]);
      }
    } else {        // way up: levels k0, k0-1, ...
      const int k0 = km - 1 - (r - nch) * TVC;
      const int nl = (k0 < TVC) ? k0 : TVC;
      mbar_expect_tx(&bar[slot], (uint32_t)(nl * IV_THREADS * 8));
      for (int c = 0; c < nl; c++)
        tma_load_tile(st + (size_t)(c * 2) * IV_THREADS, CORRECT ? &a.tmT : &a.tmO, i0, j, (n - 1) * km + (k0 - c - 1),
                      &bar[slot]);
    }
  };
  if (tid == 0) {
    for (int sl = 0; sl < TVS; sl++) mbar_init(&bar[sl], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int gi = 0; gi < TVS && gi < total; gi++) issue(gi);
  int gc = 0;  // chunk being consumed
  auto next_chunk = [&]() {  // every thread is done with the slot of chunk gc: refill it, move on
    __syncthreads();
    if (tid == 0 && gc + TVS < total) issue(gc + TVS);
    gc++;
  };
  for (int n = a.nfirst; n <= a.nlast; n++) {
    const int mt2 = (n < g.vdc_nd) ? n : g.vdc_nd;
    const double* VDC1 = g.VDC + (size_t)(mt2 - 1) * g.vdc_nk * n2 + q + (size_t)(1 - g.vdc_k0) * n2;
    double* Tn = a.TNEW + (size_t)(n - 1) * km * n2 + q;
    double* Fb = CORRECT ? a.FB + q : Tn;
    double A = 0.0, B = 0.0, C = 0.0, Fm = 0.0;
    if (active) {
      A = c_vc.afac_t[1] * VDC1[0];
      const RcpD rd = rcp_prepare(H1 + A);
      const double e = div_by(A, rd);
      sE[0] = e;
      B = H1 * e;
      const double R1 = CORRECT ? a.RHS[(size_t)(n - 1) * n2 + q] : Tn[0];
      Fm = div_by(hfac1 * R1, rd);
      Fb[0] = Fm;
    }
    // ---- forward elimination
    double* pf = Fb + n2i;  // F of the next level to store
    for (int r = 0; r < nch; r++) {
      const double* st = ring + (size_t)(gc % TVS) * TVC * 2 * IV_THREADS + tid;
      mbar_wait(&bar[gc % TVS], (uint32_t)((gc / TVS) & 1));
      const int k0 = 2 + r * TVC;
      if (active) {
#pragma unroll
        for (int c = 0; c < TVC; c++) {
          const int k = k0 + c;
          if (k <= km) {
            const double vdc = st[(size_t)(c * 2) * IV_THREADS];
            const double rhs = CORRECT ? 0.0 : st[(size_t)(c * 2 + 1) * IV_THREADS];
            C = A;
            A = c_vc.afac_t[k] * vdc;
            const double hfac = c_vc.hfac_t[k];
            double F;
            if (k > kmt) {
              F = 0.0;
            } else {
              const RcpD rd = rcp_prepare((k == kmt) ? hfac + B : hfac + A + B);
              const double e = div_by(A, rd);
              sE[(size_t)(k - 1) * IV_THREADS] = e;
              B = (hfac + B) * e;
              F = CORRECT ? div_by(C * Fm, rd) : div_by(hfac * rhs + C * Fm, rd);
            }
            *pf = F;
            pf += n2i;
            Fm = F;
          }
        }
      }
      next_chunk();
    }
    // ---- back substitution + final update (Fm = F(km))
    double Fp = Fm;
    const size_t top = (size_t)(km - 1) * n2;
    const double* Bs = CORRECT ? Tn : a.TOLD + (size_t)(n - 1) * km * n2 + q;
    if (active) Tn[top] = Bs[top] + Fp;
    const double* qf = Fb + top - n2i;  // F of the next level to load (km-1), going up: plain loads, L2-resident
    double* qt = Tn + top - n2i;        // next level to store
    double fa[TVC], fn[TVC];
#pragma unroll
    for (int c = 0; c < TVC; c++) { fa[c] = 0.0; fn[c] = 0.0; }
    int kl = km - 1;  // next level to load F of
    auto load_f = [&](double* f) {
#pragma unroll
      for (int c = 0; c < TVC; c++)
        if (active && kl - c >= 1) f[c] = qf[-(ptrdiff_t)c * n2i];
      qf -= (ptrdiff_t)TVC * n2i;
      kl -= TVC;
    };
    load_f(fa);
    for (int r = 0; r < nch; r++) {
      const double* st = ring + (size_t)(gc % TVS) * TVC * 2 * IV_THREADS + tid;
      load_f(fn);
      mbar_wait(&bar[gc % TVS], (uint32_t)((gc / TVS) & 1));
      const int k0 = km - 1 - r * TVC;
      if (active) {
#pragma unroll
        for (int c = 0; c < TVC; c++) {
          const int k = k0 - c;
          if (k >= 1) {
            double F = fa[c];
            if (k < kmt) F = F + sE[(size_t)(k - 1) * IV_THREADS] * Fp;
            *qt = st[(size_t)(c * 2) * IV_THREADS] + F;
            qt -= n2i;
            Fp = F;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < TVC; c++) fa[c] = fn[c];
      next_chunk();
    }
  }
}

int impvmixt_dev(double* TNEW, const double* TOLD, const double* PSFC, const double* RHS, int nfirst,
                 int nlast, int correct) {
  if (nfirst > nlast || nfirst > G.nt) return POP_SUCCESS;  // vertical_mix.F90:1232
  POP_REQUIRE(nfirst >= 1 && nlast <= G.nt, "impvmixt: tracer range %d..%d", nfirst, nlast);
  ScopedTimer tm("VMIX_TRACER_IMPLICIT");
  GridView g = grid_view();
  const int varthick = (G.cfg.sfc_layer_type == POP_SFC_VARTHICK);
  dim3 block(IV_THREADS, 1, 1), grid((unsigned)((G.nxg + IV_THREADS - 1) / IV_THREADS), (unsigned)G.ny_local, 1);
  double* FB = fld("WORK3D_E");
  // TMA-staged path (needs an even row pitch; TOLD is only dereferenced by the predictor form)
  IvTmaArgs ta;
  const bool tma = !G.no_tma && G.thomas_tma && !G.cfg.partial_bottom_cells && make_tmap_box(&ta.tmT, TNEW, G.km * G.nt, IV_THREADS, 1) &&
                   (correct || make_tmap_box(&ta.tmO, TOLD, G.km * G.nt, IV_THREADS, 1)) &&
                   make_tmap_box(&ta.tmV, g.VDC, g.vdc_nd * g.vdc_nk, IV_THREADS, 1);
  if (tma) {
    ta.g = g; ta.TNEW = TNEW; ta.TOLD = TOLD; ta.PSFC = PSFC; ta.RHS = RHS; ta.FB = FB;
    ta.nfirst = nfirst; ta.nlast = nlast; ta.varthick = varthick;
    if (correct) ta.tmO = ta.tmT;
    const size_t smem = sizeof(double) * IV_THREADS * ((size_t)G.km + (size_t)TVS * TVC * 2) + 8 * TVS + 16;
    auto kc = impvmixt_tma_kernel<true>;
    auto kp = impvmixt_tma_kernel<false>;
#ifndef POP_EMUL
    POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)(correct ? kc : kp), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)(correct ? kc : kp), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
    if (correct) POP_LAUNCH(kc, grid, block, smem, ta);
    else POP_LAUNCH(kp, grid, block, smem, ta);
    return pop_post_launch("impvmixt");
  }
  const size_t smem = sizeof(double) * IV_THREADS * (size_t)G.km;
  const bool pbc = (g.DZT != nullptr);
  auto kc = pbc ? impvmixt_kernel<true, true> : impvmixt_kernel<true, false>;
  auto kp = pbc ? impvmixt_kernel<false, true> : impvmixt_kernel<false, false>;
#ifndef POP_EMUL
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)(correct ? kc : kp), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)(correct ? kc : kp), cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
  if (correct) POP_LAUNCH(kc, grid, block, smem, g, TNEW, TOLD, PSFC, RHS, FB, nfirst, nlast, varthick);
  else POP_LAUNCH(kp, grid, block, smem, g, TNEW, TOLD, PSFC, RHS, FB, nfirst, nlast, varthick);
  return pop_post_launch("impvmixt");
}

// vmix_coeffs (vertical_mix.F90:518-670).  'given' (KPP-shaped VDC/VVC supplied by the caller) needs
// no work.  'const' (vmix_const.F90:144-233): constant coefficients, replaced by the convective
// values where the column is statically unstable (convection_type = 'diffusion').
__global__ void vmix_const_kernel(StateOpt o, const double* __restrict__ TMIX, const int* __restrict__ KMT,
                                  double* __restrict__ VDC, double* __restrict__ VVC, size_t n2, int km,
                                  int k0, double const_vdc, double const_vvc, double convect_diff,
                                  double vvconv) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = k0 + blockIdx.y;
  if (q >= n2) return;
  const int kp1 = (k + 1 < km) ? k + 1 : km;
  const size_t c = (size_t)(k - 1) * n2 + q, cp = (size_t)(kp1 - 1) * n2 + q;
  double rhok, rhokp;
  state_cell(o, kp1, TMIX[c], TMIX[(size_t)km * n2 + c], &rhok, nullptr, nullptr, nullptr);
  state_cell(o, kp1, TMIX[cp], TMIX[(size_t)km * n2 + cp], &rhokp, nullptr, nullptr, nullptr);
  double vdc = const_vdc, vvc = const_vvc;
  if (rhok > rhokp && k < KMT[q]) {
    vdc = convect_diff;
    vvc = vvconv;
  }
  VDC[c] = vdc;
  VVC[c] = vvc;
}

// Richardson-number dependent coefficients (vmix_rich.F90:179-414, full cells).  The reference walks k
// and carries UTK/VTK (velocities averaged to T points) from the previous level; here every (point,
// level) is independent: both levels are averaged on the fly (same four-point weights AT0=ATS=ATW=ATSW
// = 1/4, grid.F90:2908-2928, zero on the first row and column like ugrid_to_tgrid, grid.F90:3297-3355).
struct RichArgs {
  StateOpt o;
  const double *TMIX, *UMIX, *VMIX, *RHOMIX;
  const int *KMT, *KMU;
  const double *AU0, *AUN, *AUE, *AUNE;
  double *VDC, *VVC, *RICH;
  int nxb, nyb, km, k0;
  size_t n2;
  double bckgrnd_vdc, bckgrnd_vvc, rich_mix, convect_diff, convect_visc;
  int convection_diff;
};
__device__ __forceinline__ double u2t(const double* __restrict__ AU, int i, int j, int nxb) {
  if (i == 0 || j == 0) return 0.0;
  const size_t q = (size_t)j * nxb + i;
  return 0.25 * AU[q] + 0.25 * AU[q - nxb] + 0.25 * AU[q - 1] + 0.25 * AU[q - nxb - 1];
}
__global__ void vmix_rich_t_kernel(RichArgs a) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = a.k0 + blockIdx.y;
  if (q >= a.n2) return;
  const int i = (int)(q % a.nxb), j = (int)(q / a.nxb);
  const int kp1 = (k + 1 < a.km) ? k + 1 : a.km;
  const size_t c = (size_t)(k - 1) * a.n2, cp = (size_t)(kp1 - 1) * a.n2;
  const double critnu = a.convection_diff ? a.convect_diff : (0.25 * c_vc.dz[k] * c_vc.dzw[k]) / c_vc.c2dtt[k];
  double rich = 0.0, vdc = 0.0;
  if (k < a.KMT[q]) {
    const double utk = u2t(a.UMIX + c, i, j, a.nxb), utkp = u2t(a.UMIX + cp, i, j, a.nxb);
    const double vtk = u2t(a.VMIX + c, i, j, a.nxb), vtkp = u2t(a.VMIX + cp, i, j, a.nxb);
    double rhok;
    state_cell(a.o, kp1, a.TMIX[c + q], a.TMIX[(size_t)a.km * a.n2 + c + q], &rhok, nullptr, nullptr, nullptr);
    rich = -POP_GRAV * c_vc.dzw[k] * (rhok - a.RHOMIX[cp + q]) /
           ((utk - utkp) * (utk - utkp) + (vtk - vtkp) * (vtk - vtkp) + 1.0e-10);
    const double d = 1.0 + 5.0 * rich;
    const double v = a.bckgrnd_vdc + (a.bckgrnd_vvc + a.rich_mix / (d * d)) / d;
    vdc = (critnu < v) ? critnu : v;
  }
  if (rich < 0.0) vdc = critnu;
  a.RICH[c + q] = rich;
  a.VDC[c + q] = vdc;
}
__global__ void vmix_rich_u_kernel(RichArgs a) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = a.k0 + blockIdx.y;
  if (q >= a.n2) return;
  const int i = (int)(q % a.nxb), j = (int)(q / a.nxb);
  const size_t c = (size_t)(k - 1) * a.n2;
  double critnu = a.convection_diff ? a.convect_diff : (0.25 * c_vc.dz[k] * c_vc.dzw[k]) / c_vc.c2dtt[k];
  if (a.convection_diff) critnu = a.convect_visc;
  double richu = 0.0;  // tgrid_to_ugrid: zero on the last row and column (grid.F90:3403-3415)
  if (i < a.nxb - 1 && j < a.nyb - 1) {
    const double* R = a.RICH + c;
    richu = a.AU0[q] * R[q] + a.AUN[q] * R[q + a.nxb] + a.AUE[q] * R[q + 1] + a.AUNE[q] * R[q + a.nxb + 1];
  }
  double vvc = 0.0;
  if (k < a.KMU[q]) {
    const double d = 1.0 + 5.0 * richu;
    const double v = a.bckgrnd_vvc + a.rich_mix / (d * d);
    vvc = (critnu < v) ? critnu : v;
  } else {
    richu = 0.0;
  }
  if (richu < 0.0) vvc = critnu;
  a.VVC[c + q] = vvc;
}

int vmix_coeffs_dev(int k0, int k1, const double* TMIX, const double* UMIX, const double* VMIX,
                    const double* RHOMIX) {
  (void)UMIX; (void)VMIX; (void)RHOMIX;
  const pop_config& c = G.cfg;
  if (c.vmix_itype == POP_VMIX_GIVEN) return POP_SUCCESS;
  if (c.vmix_itype == POP_VMIX_CONST) {
    if (!c.convection_diff) return POP_SUCCESS;  // vmix_const.F90: coefficients stay at their constants
    POP_REQUIRE(G.vdc_nk == G.km, "vmix_coeffs(const, convective diffusion) needs implicit vertical mixing");
    POP_REQUIRE(k0 >= 1 && k1 <= G.km && k0 <= k1, "vmix_coeffs: bad level range %d..%d", k0, k1);
    ScopedTimer tm("VMIX_COEFFICIENTS_CONSTANT");
    StateOpt o{c.state_itype, c.state_range_iopt};
    const double vvconv = (c.convect_visc != 0.0) ? c.convect_visc : c.const_vvc;
    dim3 grid(ew_grid(G.n2), (unsigned)(k1 - k0 + 1), 1);
    POP_LAUNCH(vmix_const_kernel, grid, POP_EW_THREADS, 0, o, TMIX, fldi("KMT"), fld("VDC"), fld("VVC"),
               G.n2, G.km, k0, c.const_vdc, c.const_vvc, c.convect_diff, vvconv);
    return pop_post_launch("vmix_coeffs");
  }
  if (c.vmix_itype == POP_VMIX_RICH) {  // vmix_rich.F90:179-414 (full cells)
    POP_REQUIRE(G.vdc_nk == G.km && G.vvc_nk == G.km, "vmix_coeffs(rich) needs implicit vertical mixing");
    POP_REQUIRE(k0 >= 1 && k1 <= G.km && k0 <= k1, "vmix_coeffs: bad level range %d..%d", k0, k1);
    ScopedTimer tm("VMIX_COEFFICIENTS_RICH");
    StateOpt o{c.state_itype, c.state_range_iopt};
    RichArgs a;
    a.o = o; a.TMIX = TMIX; a.UMIX = UMIX; a.VMIX = VMIX; a.RHOMIX = RHOMIX; a.KMT = fldi("KMT"); a.KMU = fldi("KMU");
    a.AU0 = fld("AU0"); a.AUN = fld("AUN"); a.AUE = fld("AUE"); a.AUNE = fld("AUNE");
    a.VDC = fld("VDC"); a.VVC = fld("VVC"); a.RICH = fld("WORK3D_E");
    a.nxb = G.nxb; a.nyb = G.nyb; a.km = G.km; a.k0 = k0; a.n2 = G.n2;
    a.bckgrnd_vdc = c.bckgrnd_vdc; a.bckgrnd_vvc = c.bckgrnd_vvc; a.rich_mix = c.rich_mix;
    a.convection_diff = c.convection_diff; a.convect_diff = c.convect_diff; a.convect_visc = c.convect_visc;
    dim3 grid(ew_grid(G.n2), (unsigned)(k1 - k0 + 1), 1);
    POP_LAUNCH(vmix_rich_t_kernel, grid, POP_EW_THREADS, 0, a);
    POP_LAUNCH(vmix_rich_u_kernel, grid, POP_EW_THREADS, 0, a);
    return pop_post_launch("vmix_coeffs");
  }
  POP_REQUIRE(false, "vmix_coeffs: vmix_itype=%d is not implemented", c.vmix_itype);
  return POP_FAIL;
}
