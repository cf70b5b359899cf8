// pop_momentum.cu -- the momentum half of the baroclinic step.
//
//   momentum_column : ONE fused column kernel for everything `clinic` does per level
//                     (baroclinic.F90:1635-1895): advu (advection.F90:1127-1570), Coriolis, gradp + grad
//                     (pressure_grad.F90:187-306, operators.F90:126-192), hdiffu del2/del4
//                     (hmix_del2.F90:670-963, hmix_del4.F90:597-882), vdiffu
//                     (vertical_mix.F90:853-1026), plus the implicit-Coriolis solve and the ZX/ZY
//                     vertical integrals of baroclinic_driver (baroclinic.F90:1013-1057).  A thread
//                     owns an (i,j) column and carries WUK, SUMX/SUMY, RHOKMX/RHOKMY, VUF/VVF in
//                     registers; restricted to one level and one term it implements the slab entry
//                     points pop_advu / pop_hdiffu / pop_gradp / pop_vdiffu.
//   momentum_finish : impvmixu (vertical_mix.F90:1679-1881) + U = Uold + dU + removal of the
//                     vertical mean + KMU mask (baroclinic.F90:1067-1129).
//   grad / div      : operators.F90:126-192, :49-119 (2-d, used by the barotropic driver).
#include "pop_dev.cuh"

struct MomentumArgs {
  GridView g;
  const double *UCUR, *VCUR, *UOLD, *VOLD, *UMIX, *VMIX, *RHOOLD, *RHOCUR, *RHONEW, *SMF, *DHU;
  double *OUT1, *OUT2;  // FULL: UVEL/VVEL(new) (nxb,nyb,km); slab modes: two (nxb,nyb) slabs
  double *ZX, *ZY;
  double* WUK;          // slab ADVU: carried (in/out)
  double *SUMX, *SUMY, *RHOKMX, *RHOKMY, *VUF, *VVF;  // slab carried state
  int k0, k1;
  int lvariable_hmixu, impcor, leapfrog, pavg;
  double am, c2dtu, beta, gamma, bottom_drag;
  PopTmap tm_uc, tm_vc, tm_um, tm_vm, tm_ro, tm_rc, tm_rn;  // TMA descriptors (FULL mode)
  PopTmap tm_dz;                                             // DZU (partial bottom cells), levels 0..km+1
};
#define MO_NS 3  // TMA pipeline depth (levels in flight per CTA)

struct MomCoefTiles {  // ring-1 tiles; DMS = -DMN and DMW = -DME exactly (hmix_del2.F90:395-396): not staged
  const double *cc, *dun, *dus, *due, *duw, *dmc, *dmn, *dme;
};
// 5+5-point momentum stencil (hmix_del2.F90:892-921): s1 on A with the DU* set, s2 on B with DM*.
// R1 = true: A and B are ring-1 tiles (TIX1); false: halo tiles of a pipeline stage (TIX).
// dzu (partial bottom cells, hmix_del2.F90:852-863 / hmix_del4.F90:683-694): the staged halo tile (TIX) of level k of
// DZU, or null; inb: the centre and its four neighbours lie inside the padded block
template <bool R1>
__device__ __forceinline__ double mom_stencil(const MomCoefTiles& c, const double* A, const double* B,
                                              int ii, int jj, bool plus, const double* dzu = nullptr, bool inb = false) {
  const int q = TIX1(ii, jj);
  const int o = R1 ? q : TIX(ii, jj);
  constexpr int W = R1 ? POP_T1W : POP_TW;
  double dn = c.dun[q], ds = c.dus[q], de = c.due[q], dw = c.duw[q];
  if (dzu) {
    if (inb) {
      const double* zc = dzu + TIX(ii, jj);
      const double z = zc[0];
      const RcpD rz = rcp_prepare(z);  // four IEEE quotients by the same thickness: one reciprocal refinement (div_by)
      dn = div_by(dn * fmin(zc[POP_TW], z), rz);
      ds = div_by(ds * fmin(zc[-POP_TW], z), rz);
      de = div_by(de * fmin(zc[1], z), rz);
      dw = div_by(dw * fmin(zc[-1], z), rz);
    } else {
      dn = ds = de = dw = 0.0;
    }
  }
  const double s1 = c.cc[q] * A[o] + dn * A[o + W] + ds * A[o - W] + de * A[o + 1] + dw * A[o - 1];
  const double s2 = c.dmc[q] * B[o] + c.dmn[q] * B[o + W] + (-c.dmn[q]) * B[o - W] + c.dme[q] * B[o + 1] +
                    (-c.dme[q]) * B[o - 1];
  return plus ? (s1 + s2) : (s1 - s2);
}

#define MOM_STAGE_TILES 7   // uc, vc, um, vm, rho old/cur/new (halo tiles, TMA boxes); partial bottom cells: + DZU
#define MO_NS_PBC 2         // pipeline depth with the eighth tile (two CTAs per SM must still fit)
#define MOM_FIXED_TILES 13  // ring-1: cc, dun, dus, due, duw, dmc, dmn, dme, amf, ud, vd, d2u, d2v
// Per level: every thread first finishes the "main" pass of level k (which reads the per-level
// intermediates UD = U*DYU, VD = V*DXU and, for del4, D2U/D2V = AMF*L(U,V) of level k), then the CTA
// rebuilds those intermediates for level k+1 from the stage that has already landed.  The products
// and the first Laplacian are thus computed once per tile point instead of once per stencil use.
template <int MODE, bool DEL4, bool TMA, bool PBC>
__global__ void __launch_bounds__(POP_NTHREADS, 2)
momentum_column_kernel(const POP_GRID_CONSTANT MomentumArgs a) {
  POP_DYN_SMEM(smem_raw);
  constexpr int NS = TMA ? (PBC ? MO_NS_PBC : MO_NS) : 1;
  constexpr int NTL = PBC ? MOM_STAGE_TILES + 1 : MOM_STAGE_TILES;  // tiles per stage
  double* sm = (double*)smem_raw;
  double* s_stage = sm;  // [NS][NTL][TN]
  double* s_cc = s_stage + NS * NTL * POP_TN;
  double* s_dun = s_cc + POP_T1N;
  double* s_dus = s_dun + POP_T1N;
  double* s_due = s_dus + POP_T1N;
  double* s_duw = s_due + POP_T1N;
  double* s_dmc = s_duw + POP_T1N;
  double* s_dmn = s_dmc + POP_T1N;
  double* s_dme = s_dmn + POP_T1N;
  double* s_amf = s_dme + POP_T1N;
  double* s_ud = s_amf + POP_T1N;
  double* s_vd = s_ud + POP_T1N;
  double* s_d2u = s_vd + POP_T1N;
  double* s_d2v = s_d2u + POP_T1N;
  int* s_kmu = (int*)(s_d2v + POP_T1N);
  uint64_t* s_bar = (uint64_t*)(s_kmu + POP_T1N);

  const GridView& g = a.g;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * POP_BX + tx;
  const int i0 = (g.ib - 1) + blockIdx.x * POP_BX;
  const int j0 = (g.jb - 1) + blockIdx.y * POP_BY;
  const int i = i0 + tx, j = j0 + ty;
  const bool active = (i <= g.ie - 1) && (j <= g.je - 1);
  const size_t q = (size_t)j * g.nxb + i;
  const size_t n2 = g.n2;
  const int km = g.km, nxb = g.nxb, nyb = g.nyb;
  constexpr bool DO_ADV = (MODE == MO_FULL || MODE == MO_ADVU);
  constexpr bool DO_HMIX = (MODE == MO_FULL || MODE == MO_HDIFFU);
  constexpr bool DO_GRADP = (MODE == MO_FULL || MODE == MO_GRADP);
  constexpr bool DO_VDIF = (MODE == MO_FULL || MODE == MO_VDIFFU);
  constexpr int HM = DEL4 ? 2 : 1;
  const bool same_mix = (a.UMIX == a.UCUR) && DO_ADV && (TMA || !DEL4);

  // ---- pipeline start (one thread): the first NS levels are in flight while the k-invariants load
  const uint32_t stage_bytes = (uint32_t)((2 + (same_mix ? 0 : 2) + (a.pavg ? 3 : 1) + (PBC ? 1 : 0)) * POP_TILE_BYTES);
  auto issue = [&](int kk) {
    const int sl = (kk - a.k0) % NS;
    double* st = s_stage + (size_t)sl * NTL * POP_TN;
    const int x = i0 - POP_H, y = j0 - POP_H, z = kk - 1;
    mbar_expect_tx(&s_bar[sl], stage_bytes);
    if (PBC) tma_load_tile(st + 7 * POP_TN, &a.tm_dz, x, y, kk, &s_bar[sl]);  // DZU has levels 0..km+1
    tma_load_tile(st, &a.tm_uc, x, y, z, &s_bar[sl]);
    tma_load_tile(st + POP_TN, &a.tm_vc, x, y, z, &s_bar[sl]);
    if (!same_mix) {
      tma_load_tile(st + 2 * POP_TN, &a.tm_um, x, y, z, &s_bar[sl]);
      tma_load_tile(st + 3 * POP_TN, &a.tm_vm, x, y, z, &s_bar[sl]);
    }
    tma_load_tile(st + 5 * POP_TN, &a.tm_rc, x, y, z, &s_bar[sl]);
    if (a.pavg) {
      tma_load_tile(st + 4 * POP_TN, &a.tm_ro, x, y, z, &s_bar[sl]);
      tma_load_tile(st + 6 * POP_TN, &a.tm_rn, x, y, z, &s_bar[sl]);
    }
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

  // ---- k-invariant staging: coefficients on the ring-1 tile; DYU/DXU of the (at most two) ring-1
  // points this thread rebuilds every level stay in registers
  double r_dyu[2] = {0.0, 0.0}, r_dxu[2] = {0.0, 0.0};
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int p = tid + s * POP_NTHREADS;
    if (p < POP_T1N) {
      const int jj = p / POP_T1W - 1, ii = p % POP_T1W - 1;
      const int gi = i0 + ii, gj = j0 + jj;
      const bool in = (gi >= 0 && gi < nxb && gj >= 0 && gj < nyb);
      const size_t qq = (size_t)gj * nxb + gi;
      if (DO_ADV) {
        r_dyu[s] = in ? g.DYU[qq] : 0.0;
        r_dxu[s] = in ? g.DXU[qq] : 0.0;
      }
      if (DO_HMIX) {
        s_cc[p] = in ? g.DUC[qq] + g.DUM[qq] : 0.0;
        s_dun[p] = in ? g.DUN[qq] : 0.0;
        s_dus[p] = in ? g.DUS[qq] : 0.0;
        s_due[p] = in ? g.DUE[qq] : 0.0;
        s_duw[p] = in ? g.DUW[qq] : 0.0;
        s_dmc[p] = in ? g.DMC[qq] : 0.0;
        s_dmn[p] = in ? g.DMN[qq] : 0.0;
        s_dme[p] = in ? g.DME[qq] : 0.0;
        s_amf[p] = in ? g.AMF[qq] : 0.0;
        s_kmu[p] = in ? g.KMU[qq] : 0;
      }
    }
  }
  const MomCoefTiles ct{s_cc, s_dun, s_dus, s_due, s_duw, s_dmc, s_dmn, s_dme};
  int kmu = 0;
  double uarea_r = 0, kxu = 0, kyu = 0, fcor = 0, dxur = 0, dyur = 0;
  if (active) {
    kmu = g.KMU[q];
    uarea_r = g.UAREA_R[q];
    kxu = g.KXU[q];
    kyu = g.KYU[q];
    fcor = g.FCOR[q];
    dxur = g.DXUR[q];
    dyur = g.DYUR[q];
  }
  // implicit Coriolis factors of the column (baroclinic.F90:1013-1030)
  const double cor_w1 = a.c2dtu * a.beta * fcor;
  const double cor_w2 = a.c2dtu / (1.0 + cor_w1 * cor_w1);
  // ---- carried state
  double wuk = 0, u_m = 0, v_m = 0, uo_c = 0, vo_c = 0;
  double sumx = 0, sumy = 0, rmx = 0, rmy = 0, vuf = 0, vvf = 0, zx = 0, zy = 0;
  if (active) {
    if (DO_ADV) {
      wuk = (MODE == MO_FULL) ? a.DHU[q] : a.WUK[q];
      if (a.k0 > 1) {
        u_m = a.UCUR[(size_t)(a.k0 - 2) * n2 + q];
        v_m = a.VCUR[(size_t)(a.k0 - 2) * n2 + q];
      }
    }
    if (DO_VDIF || MODE == MO_FULL) {
      uo_c = a.UOLD[(size_t)(a.k0 - 1) * n2 + q];
      vo_c = a.VOLD[(size_t)(a.k0 - 1) * n2 + q];
    }
    if (DO_VDIF && a.k0 > 1) {
      vuf = a.VUF[q];
      vvf = a.VVF[q];
    }
    if (DO_GRADP && a.k0 > 1) {
      sumx = a.SUMX[q];
      sumy = a.SUMY[q];
      rmx = a.RHOKMX[q];
      rmy = a.RHOKMY[q];
    }
  }

  // per-level intermediates on the ring-1 tile: flux operands (advection.F90:1307-1340) and the first
  // application of the del4 operator (hmix_del4.F90:730-790)
  auto pre = [&](int kk, const double* uc, const double* vc, const double* um, const double* vm, const double* dzu) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int p = tid + s * POP_NTHREADS;
      if (p < POP_T1N) {
        const int jj = p / POP_T1W - 1, ii = p % POP_T1W - 1;
        const int gi = i0 + ii, gj = j0 + jj;
        const bool inb = (gi >= 1 && gi <= nxb - 2 && gj >= 1 && gj <= nyb - 2);
        if (DO_ADV) {
          s_ud[p] = uc[TIX(ii, jj)] * r_dyu[s];
          s_vd[p] = vc[TIX(ii, jj)] * r_dxu[s];
          if (PBC) {  // advection.F90:1245-1303: (U*DYU)*DZU
            const double z = dzu[TIX(ii, jj)];  // zero outside the block, like the velocity tiles
            s_ud[p] = s_ud[p] * z;
            s_vd[p] = s_vd[p] * z;
          }
        }
        if (DO_HMIX && DEL4) {
          double d2u = mom_stencil<false>(ct, um, vm, ii, jj, true, dzu, inb);
          double d2v = mom_stencil<false>(ct, vm, um, ii, jj, false, dzu, inb);
          if (a.lvariable_hmixu) {
            if (kk <= s_kmu[p]) { d2u = s_amf[p] * d2u; d2v = s_amf[p] * d2v; }
            else { d2u = 0.0; d2v = 0.0; }
          } else if (kk > s_kmu[p]) { d2u = 0.0; d2v = 0.0; }
          s_d2u[p] = d2u;
          s_d2v[p] = d2v;
        }
      }
    }
  };
  if (TMA) {
    mbar_wait(&s_bar[0], 0u);
    __syncthreads();  // coefficient tiles are staged
    pre(a.k0, s_stage, s_stage + POP_TN, same_mix ? s_stage : s_stage + 2 * POP_TN,
        same_mix ? s_stage + POP_TN : s_stage + 3 * POP_TN, PBC ? s_stage + 7 * POP_TN : nullptr);
    __syncthreads();
  }
  const bool uold_is_mix = (a.UOLD == a.UMIX) && !same_mix;
  double vvc_nx = 0.0;
  auto vvc_at = [&](int kk) { return g.VVC[(size_t)(((g.vvc_nk == 1) ? 1 : kk) - 1) * n2 + q]; };
  if (DO_VDIF && active) vvc_nx = vvc_at(a.k0);

  for (int k = a.k0; k <= a.k1; k++) {
    const size_t lev = (size_t)(k - 1) * n2;
    const int slot = (k - a.k0) % NS;
    double* s_uc = s_stage + (size_t)slot * NTL * POP_TN;
    double* s_vc = s_uc + POP_TN;
    double* s_um = s_uc + 2 * POP_TN;
    double* s_vm = s_uc + 3 * POP_TN;
    double* s_ro = s_uc + 4 * POP_TN;
    double* s_rc = s_uc + 5 * POP_TN;
    double* s_rn = s_uc + 6 * POP_TN;
    double* s_dz = PBC ? s_uc + 7 * POP_TN : nullptr;  // DZU of level k (halo tile)
    const double* um = same_mix ? s_uc : s_um;
    const double* vm = same_mix ? s_vc : s_vm;
    if (TMA) {
      mbar_wait(&s_bar[slot], (uint32_t)(((k - a.k0) / NS) & 1));
    } else {
      __syncthreads();
      if (DO_ADV) {
        tile_load(s_uc, a.UCUR + lev, i0, j0, nxb, nyb, -1, POP_BX, -1, POP_BY, tid);
        tile_load(s_vc, a.VCUR + lev, i0, j0, nxb, nyb, -1, POP_BX, -1, POP_BY, tid);
      }
      if (DO_HMIX && !same_mix) {
        tile_load(s_um, a.UMIX + lev, i0, j0, nxb, nyb, -HM, POP_BX - 1 + HM, -HM, POP_BY - 1 + HM, tid);
        tile_load(s_vm, a.VMIX + lev, i0, j0, nxb, nyb, -HM, POP_BX - 1 + HM, -HM, POP_BY - 1 + HM, tid);
      }
      if (DO_GRADP) {
        tile_load(s_rc, a.RHOCUR + lev, i0, j0, nxb, nyb, 0, POP_BX, 0, POP_BY, tid);
        if (a.pavg) {
          tile_load(s_ro, a.RHOOLD + lev, i0, j0, nxb, nyb, 0, POP_BX, 0, POP_BY, tid);
          tile_load(s_rn, a.RHONEW + lev, i0, j0, nxb, nyb, 0, POP_BX, 0, POP_BY, tid);
        }
      }
      if (PBC) tile_load(s_dz, g.DZU + (size_t)k * n2, i0, j0, nxb, nyb, -POP_H, POP_BX - 1 + POP_H, -POP_H, POP_BY - 1 + POP_H, tid);
      __syncthreads();
      if (DO_ADV || (DO_HMIX && DEL4)) {
        pre(k, s_uc, s_vc, um, vm, s_dz);
        __syncthreads();
      }
    }
    const bool have_next = TMA && (k < a.k1);
    const int nslot = (k + 1 - a.k0) % NS;
    const double* n_uc = s_stage + (size_t)nslot * NTL * POP_TN;
    const double* n_vc = n_uc + POP_TN;
    const double* n_um = same_mix ? n_uc : n_uc + 2 * POP_TN;
    const double* n_vm = same_mix ? n_vc : n_uc + 3 * POP_TN;
    if (have_next) mbar_wait(&s_bar[nslot], (uint32_t)(((k + 1 - a.k0) / NS) & 1));
    if (active) {
    double fx = 0.0, fy = 0.0;
    constexpr bool pbc = PBC;  // partial bottom cells: g.DZU is set
    const double* dzu_k = pbc ? s_dz : nullptr;
    const double dzu_c = pbc ? s_dz[TIX(tx, ty)] : 0.0;  // thickness of this U cell
    const RcpD rzu = pbc ? rcp_prepare(dzu_c) : RcpD{1.0, 1.0};
    const double h_dzu = pbc ? div_by(0.5, rzu) : 0.0;  // 0.5 / DZU
    const double U = DO_ADV ? s_uc[TIX(tx, ty)] : 0.0, V = DO_ADV ? s_vc[TIX(tx, ty)] : 0.0;
    // ---- advu: advection.F90:1307-1491
    double luk = 0.0, lvk = 0.0;
    if (DO_ADV) {
#define UD(di, dj) s_ud[TIX1(tx + (di), ty + (dj))]
#define VD(di, dj) s_vd[TIX1(tx + (di), ty + (dj))]
      const double uuw = 0.25 * (UD(0, 0) + UD(-1, 0)) + 0.125 * (UD(0, -1) + UD(-1, -1) + UD(0, 1) + UD(-1, 1));
      const double uue = 0.25 * (UD(1, 0) + UD(0, 0)) + 0.125 * (UD(1, -1) + UD(0, -1) + UD(1, 1) + UD(0, 1));
      const double vus = 0.25 * (VD(0, 0) + VD(0, -1)) + 0.125 * (VD(-1, 0) + VD(-1, -1) + VD(1, 0) + VD(1, -1));
      const double vun = 0.25 * (VD(0, 1) + VD(0, 0)) + 0.125 * (VD(-1, 1) + VD(-1, 0) + VD(1, 1) + VD(1, 0));
#undef UD
#undef VD
      const double wukb = pbc ? wuk + (vun - vus + uue - uuw) * uarea_r  // advection.F90:1352-1353
                              : wuk + c_vc.c2dz[k] * 0.5 * (vun - vus + uue - uuw) * uarea_r;
      const double cc = vun - vus + uue - uuw;  // VUS(i,j+1) - VUS(i,j) + UUW(i+1,j) - UUW(i,j)
      luk = 0.5 * (cc * U + vun * s_uc[TIX(tx, ty + 1)] - vus * s_uc[TIX(tx, ty - 1)] +
                   uue * s_uc[TIX(tx + 1, ty)] - uuw * s_uc[TIX(tx - 1, ty)]) * uarea_r;
      lvk = 0.5 * (cc * V + vun * s_vc[TIX(tx, ty + 1)] - vus * s_vc[TIX(tx, ty - 1)] +
                   uue * s_vc[TIX(tx + 1, ty)] - uuw * s_vc[TIX(tx - 1, ty)]) * uarea_r;
      if (pbc) {  // :1381-1405
        luk = div_by(luk, rzu);
        lvk = div_by(lvk, rzu);
      }
      if (k == 1) {
        luk = luk + c_vc.dzr[k] * wuk * U;
        lvk = lvk + c_vc.dzr[k] * wuk * V;
      } else if (pbc) {  // :1443-1447
        luk = luk + h_dzu * wuk * (u_m + U);
        lvk = lvk + h_dzu * wuk * (v_m + V);
      } else {
        luk = luk + c_vc.dz2r[k] * wuk * (u_m + U);
        lvk = lvk + c_vc.dz2r[k] * wuk * (v_m + V);
      }
      if (k < km) {
        const double Up = have_next ? n_uc[TIX(tx, ty)] : a.UCUR[lev + n2 + q];
        const double Vp = have_next ? n_vc[TIX(tx, ty)] : a.VCUR[lev + n2 + q];
        if (pbc) {  // :1462-1466
          luk = luk - h_dzu * wukb * (U + Up);
          lvk = lvk - h_dzu * wukb * (V + Vp);
        } else {
          luk = luk - c_vc.dz2r[k] * wukb * (U + Up);
          lvk = lvk - c_vc.dz2r[k] * wukb * (V + Vp);
        }
      }
      if (k <= kmu) {
        luk = luk + U * V * kyu - (V * V) * kxu;
        lvk = lvk + U * V * kxu - (U * U) * kyu;
      } else {
        luk = 0.0;
        lvk = 0.0;
      }
      wuk = wukb;
      u_m = U;
      v_m = V;
    }
    // ---- gradp: pressure_grad.F90:268-301 with grad of operators.F90:178-187
    double pkx = 0.0, pky = 0.0;
    if (DO_GRADP) {
      double rkx = 0.0, rky = 0.0;
      if (k <= kmu) {
        // RHOAVG at the four T corners: pressure_grad.F90:258-266
        const double bouss = c_vc.bouss[k];
#define RAVG(di, dj)                                                                                   \
  (a.pavg ? 0.25 * (s_rn[TIX(tx + (di), ty + (dj))] + 2.0 * s_rc[TIX(tx + (di), ty + (dj))] +          \
                    s_ro[TIX(tx + (di), ty + (dj))]) * bouss                                           \
          : s_rc[TIX(tx + (di), ty + (dj))] * bouss)
        const double f11 = RAVG(1, 1), f00 = RAVG(0, 0), f01 = RAVG(0, 1), f10 = RAVG(1, 0);
#undef RAVG
        rkx = dxur * 0.5 * (f11 - f00 - f01 + f10);
        rky = dyur * 0.5 * (f11 - f00 + f01 - f10);
      }
      if (k == 1) {
        rmx = rkx;
        rmy = rky;
        sumx = 0.0;
        sumy = 0.0;
      }
      const double factor = c_vc.dzw[k - 1] * POP_GRAV * 0.5;
      sumx = sumx + factor * (rkx + rmx);
      sumy = sumy + factor * (rky + rmy);
      pkx = sumx;
      pky = sumy;
      rmx = rkx;
      rmy = rky;
    }
    // ---- hdiffu
    double hdu = 0.0, hdv = 0.0;
    if (DO_HMIX) {
      if (DEL4) {
        hdu = a.am * mom_stencil<true>(ct, s_d2u, s_d2v, tx, ty, true, dzu_k, true);
        hdv = a.am * mom_stencil<true>(ct, s_d2v, s_d2u, tx, ty, false, dzu_k, true);
      } else {
        hdu = a.am * mom_stencil<false>(ct, um, vm, tx, ty, true, dzu_k, true);
        hdv = a.am * mom_stencil<false>(ct, vm, um, tx, ty, false, dzu_k, true);
      }
      if (k > kmu) { hdu = 0.0; hdv = 0.0; }
    }
    // ---- vdiffu: vertical_mix.F90:935-1010
    double vdu = 0.0, vdv = 0.0;
    double uo_p = 0.0, vo_p = 0.0;
    if (DO_VDIF || MODE == MO_FULL) {
      const bool from_tile = have_next && uold_is_mix;
      uo_p = (k < km) ? (from_tile ? n_um[TIX(tx, ty)] : a.UOLD[lev + n2 + q]) : uo_c;
      vo_p = (k < km) ? (from_tile ? n_vm[TIX(tx, ty)] : a.VOLD[lev + n2 + q]) : vo_c;
    }
    if (DO_VDIF) {
      const double vvc = vvc_nx;
      if (k < a.k1) vvc_nx = vvc_at(k + 1);
      if (k == 1) {
        vuf = (kmu >= 1) ? a.SMF[q] : 0.0;
        vvf = (kmu >= 1) ? a.SMF[n2 + q] : 0.0;
      }
      double vufb, vvfb;
      if (pbc) {  // vertical_mix.F90:946-958: kp1 = min(k+1, km)
        const double zp = g.DZU[(size_t)((k < km) ? k + 1 : km) * n2 + q];
        const double W = (k < km) ? 0.5 * (dzu_c + zp) : 0.5 * zp;
        const RcpD rw = rcp_prepare(W);
        vufb = div_by(vvc * (uo_c - uo_p), rw);
        vvfb = div_by(vvc * (vo_c - vo_p), rw);
      } else {
        vufb = vvc * (uo_c - uo_p) * c_vc.dzwr[k];
        vvfb = vvc * (vo_c - vo_p) * c_vc.dzwr[k];
      }
      if (k == kmu) {
        const double vmag = a.bottom_drag * sqrt(uo_c * uo_c + vo_c * vo_c);
        vufb = vmag * uo_c;
        vvfb = vmag * vo_c;
      }
      if (pbc) {  // :991-995
        vdu = (k <= kmu) ? div_by(vuf - vufb, rzu) : 0.0;
        vdv = (k <= kmu) ? div_by(vvf - vvfb, rzu) : 0.0;
      } else {
        vdu = (k <= kmu) ? (vuf - vufb) * c_vc.dzr[k] : 0.0;
        vdv = (k <= kmu) ? (vvf - vvfb) * c_vc.dzr[k] : 0.0;
      }
      vuf = vufb;
      vvf = vvfb;
    }
    if (MODE == MO_ADVU) { a.OUT1[q] = luk; a.OUT2[q] = lvk; }
    else if (MODE == MO_GRADP) { a.OUT1[q] = pkx; a.OUT2[q] = pky; }
    else if (MODE == MO_HDIFFU) { a.OUT1[q] = hdu; a.OUT2[q] = hdv; }
    else if (MODE == MO_VDIFFU) { a.OUT1[q] = vdu; a.OUT2[q] = vdv; }
    else {
      // ---- clinic assembly: baroclinic.F90:1728-1895
      fx = -luk;
      fy = -lvk;
      if (a.impcor && a.leapfrog) {
        fx = fx + fcor * (a.gamma * V + (1.0 - a.gamma) * vo_c);
        fy = fy - fcor * (a.gamma * U + (1.0 - a.gamma) * uo_c);
      } else if (!a.impcor && a.leapfrog) {
        fx = fx + fcor * V;
        fy = fy - fcor * U;
      } else {
        fx = fx + fcor * vo_c;
        fy = fy - fcor * uo_c;
      }
      fx = fx - pkx;
      fy = fy - pky;
      fx = fx + hdu;
      fy = fy + hdv;
      fx = fx + vdu;
      fy = fy + vdv;
      if (k > kmu) { fx = 0.0; fy = 0.0; }
      // ---- implicit Coriolis and vertical integrals: baroclinic.F90:1013-1045
      double un, vn;
      if (a.impcor) {
        un = (fx + cor_w1 * fy) * cor_w2;
        vn = (fy - cor_w1 * fx) * cor_w2;
      } else {
        un = a.c2dtu * fx;
        vn = a.c2dtu * fy;
      }
      a.OUT1[lev + q] = un;
      a.OUT2[lev + q] = vn;
      zx = zx + fx * (pbc ? dzu_c : c_vc.dz[k]);  // baroclinic.F90:1037-1043
      zy = zy + fy * (pbc ? dzu_c : c_vc.dz[k]);
    }
    uo_c = uo_p;
    vo_c = vo_p;
    }  // active
    if (TMA) {
      __syncthreads();  // every thread is done with ring slot k and with the intermediates of level k
      if (tid == 0 && k + NS <= a.k1) issue(k + NS);
      if (have_next) {
        pre(k + 1, n_uc, n_vc, n_um, n_vm, PBC ? n_uc + 7 * POP_TN : nullptr);
        __syncthreads();
      }
    }
  }
  if (!active) return;
  if (MODE == MO_FULL) {
    const double hur = g.HUR[q];
    a.ZX[q] = zx * hur;
    a.ZY[q] = zy * hur;
  } else if (MODE == MO_ADVU) {
    a.WUK[q] = wuk;
  } else if (MODE == MO_GRADP) {
    a.SUMX[q] = sumx; a.SUMY[q] = sumy; a.RHOKMX[q] = rmx; a.RHOKMY[q] = rmy;
  } else if (MODE == MO_VDIFFU) {
    a.VUF[q] = vuf; a.VVF[q] = vvf;
  }
}

template <int MODE>
static int launch_momentum(const MomentumArgs& a, bool del4, bool tma) {
  void (*kfn)(const MomentumArgs) = nullptr;
  const bool pbc = (a.g.DZU != nullptr);
  if (pbc) {
    if (tma) kfn = del4 ? momentum_column_kernel<MODE, true, true, true> : momentum_column_kernel<MODE, false, true, true>;
    else kfn = del4 ? momentum_column_kernel<MODE, true, false, true> : momentum_column_kernel<MODE, false, false, true>;
  } else {
    if (tma) kfn = del4 ? momentum_column_kernel<MODE, true, true, false> : momentum_column_kernel<MODE, false, true, false>;
    else kfn = del4 ? momentum_column_kernel<MODE, true, false, false> : momentum_column_kernel<MODE, false, false, false>;
  }
  const int ns = tma ? (pbc ? MO_NS_PBC : MO_NS) : 1;
  const size_t smem = sizeof(double) * ((size_t)POP_TN * ns * (MOM_STAGE_TILES + (pbc ? 1 : 0)) + (size_t)POP_T1N * MOM_FIXED_TILES) +
                      sizeof(int) * POP_T1N + 8 * MO_NS;
#ifndef POP_EMUL
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
  POP_LAUNCH(kfn, col_grid(G.nxg, G.ny_local), col_block(), smem, a);
  return POP_SUCCESS;
}

int momentum_column(int mode, int k, const MomentumIO& io) {
  MomentumArgs a;
  memset(&a, 0, sizeof(a));
  a.g = grid_view();
  a.UCUR = io.UCUR; a.VCUR = io.VCUR; a.UOLD = io.UOLD; a.VOLD = io.VOLD; a.UMIX = io.UMIX;
  a.VMIX = io.VMIX; a.RHOOLD = io.RHOOLD; a.RHOCUR = io.RHOCUR; a.RHONEW = io.RHONEW;
  a.SMF = io.SMF; a.DHU = io.DHU;
  a.OUT1 = io.UNEW; a.OUT2 = io.VNEW; a.ZX = io.ZX; a.ZY = io.ZY; a.WUK = io.WUK;
  a.SUMX = fld("SUMX"); a.SUMY = fld("SUMY"); a.RHOKMX = fld("RHOKMX"); a.RHOKMY = fld("RHOKMY");
  a.VUF = fld("VUF"); a.VVF = fld("VVF");
  a.lvariable_hmixu = G.cfg.lvariable_hmixu;
  a.impcor = G.cfg.impcor;
  a.leapfrog = G.leapfrogts;
  a.pavg = (G.cfg.lpressure_avg && G.leapfrogts);
  a.am = G.am; a.c2dtu = G.c2dtu; a.beta = G.beta; a.gamma = G.gamma;
  a.bottom_drag = G.cfg.bottom_drag;
  const bool del4 = (G.cfg.hmix_momentum_itype == POP_HMIX_DEL4);
  if (mode == MO_FULL) { a.k0 = 1; a.k1 = G.km; }
  else {
    POP_REQUIRE(k >= 1 && k <= G.km, "momentum slab operator: k=%d out of range", k);
    a.k0 = a.k1 = k;
    // 2-d slab arguments of the reference signatures are addressed as "level k" of a virtual 3-d array
    const size_t off = (size_t)(k - 1) * G.n2;
    if (mode == MO_HDIFFU) { a.UMIX -= off; a.VMIX -= off; }
    if (mode == MO_GRADP) {
      POP_REQUIRE(a.RHOCUR && (!a.pavg || (a.RHOOLD && a.RHONEW)), "gradp: null density slab");
      a.RHOCUR -= off;
      if (a.RHOOLD) a.RHOOLD -= off;
      if (a.RHONEW) a.RHONEW -= off;
    }
  }
  switch (mode) {
    case MO_FULL: {
      const bool tma = !G.no_tma && make_tmap(&a.tm_uc, a.UCUR, G.km) && make_tmap(&a.tm_vc, a.VCUR, G.km) &&
                       make_tmap(&a.tm_um, a.UMIX, G.km) && make_tmap(&a.tm_vm, a.VMIX, G.km) &&
                       make_tmap(&a.tm_rc, a.RHOCUR, G.km) &&
                       (!a.pavg || (make_tmap(&a.tm_ro, a.RHOOLD, G.km) && make_tmap(&a.tm_rn, a.RHONEW, G.km))) &&
                       (!a.g.DZU || make_tmap(&a.tm_dz, a.g.DZU, G.km + 2));
      POP_TRY(launch_momentum<MO_FULL>(a, del4, tma));
      break;
    }
    case MO_ADVU: POP_TRY(launch_momentum<MO_ADVU>(a, false, false)); break;
    case MO_HDIFFU: POP_TRY(launch_momentum<MO_HDIFFU>(a, del4, false)); break;
    case MO_GRADP: POP_TRY(launch_momentum<MO_GRADP>(a, false, false)); break;
    case MO_VDIFFU: POP_TRY(launch_momentum<MO_VDIFFU>(a, false, false)); break;
    default: POP_REQUIRE(false, "momentum_column: bad mode %d", mode);
  }
  return pop_post_launch("momentum_column");
}

// =====================================================================================
// impvmixu + velocity finish
// =====================================================================================
// Same layout as impvmixt (pop_tracer.cu): one thread per column, 128 columns per CTA, E(k) in shared memory,
// running level pointers, loads issued one chunk ahead of the recurrence, one reciprocal refinement shared by the
// three quotients of a level (div_by).  Passes over the column:
//   1 (down) forward elimination: reads RHS u,v and VVC, leaves F1,F2 in place
//   2 (up)   back substitution fused with U = Uold + dU: the E(k) slot of a finished level is recycled for U(k),
//            V(k) goes back to its (L2-resident) place
//   3 (down) vertical means in the reference's summation order (U from shared memory)
//   4 (down) mean removal + KMU mask [+ UBTROP/VBTROP(new): the `add barotropic` pass of step_mod.F90:581-592
//            when the caller runs this kernel after the barotropic solve]
// so DRAM sees RHS u,v, VVC, Uold, Vold in and U,V out (56 B/cell, 72 with the fused barotropic add that used to be
// a 32 B/cell kernel of its own).
#ifndef MF_CH
#define MF_CH 4    // levels per chunk, forward sweep (three arrays)
#endif
#ifndef MF_NB
#define MF_NB 4    // register ring depth, forward sweep
#endif
#ifndef MF_BCH
#define MF_BCH 2   // levels per chunk, backward sweep (four arrays)
#endif
#ifndef MF_BNB
#define MF_BNB 4
#endif
#ifndef MF_MINB
#define MF_MINB 3
#endif
#ifndef MF_PD
#define MF_PD 8    // L2 prefetch distance in levels (0: off)
#endif
#define MF_THREADS 128
template <bool PBC>
__global__ void __launch_bounds__(MF_THREADS, MF_MINB)
momentum_finish_kernel(GridView g, double* __restrict__ UNEW, double* __restrict__ VNEW,
                       const double* __restrict__ UOLD, const double* __restrict__ VOLD,
                       const double* __restrict__ UB, const double* __restrict__ VB, int bt_skip_row,
                       int implicit_vmix, int finish, double c2dtu, double* sfc_uv, int nxg) {
  POP_DYN_SMEM(smem_raw);
  double* sE = (double*)smem_raw + threadIdx.x;
  const int i = (g.ib - 1) + blockIdx.x * MF_THREADS + threadIdx.x;
  const int j = (g.jb - 1) + blockIdx.y;
  if (i > g.ie - 1 || j > g.je - 1) return;
  const size_t q = (size_t)j * g.nxb + i, n2 = g.n2;
  const int n2i = (int)n2;
  const int km = g.km, kmu = g.KMU[q];
  double* Un = UNEW + q;
  double* Vn = VNEW + q;
  const int vstr = (g.vvc_nk == 1) ? 0 : n2i;
  const size_t top = (size_t)(km - 1) * n2;
  if (implicit_vmix) {
    double A, B, C, F1, F2;
    {
      const double hfac = c_vc.hfac_u[1];
      A = c_vc.afac_u[1] * g.VVC[q];
      const RcpD rd = rcp_prepare(hfac + A);
      const double e = div_by(A, rd);
      sE[0] = e;
      B = hfac * e;
      F1 = div_by(hfac * Un[0], rd);
      F2 = div_by(hfac * Vn[0], rd);
      Un[0] = F1;
      Vn[0] = F2;
    }
    const double* pc = g.VVC + q + vstr;  // next level to load (level 2)
    const double* pu = Un + n2i;
    const double* pv = Vn + n2i;
    double* su = Un + n2i;                // next level to store
    double* sv = Vn + n2i;
    const int nfull = (km - 1) / MF_CH;   // full chunks of the forward sweep
    int nld = nfull;                      // chunks not loaded yet
    double ub[MF_NB][MF_CH], wb[MF_NB][MF_CH], cb[MF_NB][MF_CH];  // ring of chunk buffers
#pragma unroll
    for (int s = 0; s < MF_NB; s++)
#pragma unroll
      for (int c = 0; c < MF_CH; c++) { ub[s][c] = 0.0; wb[s][c] = 0.0; cb[s][c] = 0.0; }
    const ptrdiff_t pdc = (ptrdiff_t)MF_PD * vstr, pdn = (ptrdiff_t)MF_PD * n2i;
    int kpf = 2 + MF_PD;  // level the next forward L2 prefetch targets
    auto load_fwd = [&](double* cc, double* uu, double* ww) {
      if (nld > 0) {
        if (MF_PD && kpf + MF_CH - 1 <= km) {
#pragma unroll
          for (int c = 0; c < MF_CH; c++) {
            if (vstr) prefetch_l2(pc + pdc + (ptrdiff_t)c * vstr);
            prefetch_l2(pu + pdn + (ptrdiff_t)c * n2i);
            prefetch_l2(pv + pdn + (ptrdiff_t)c * n2i);
          }
        }
        kpf += MF_CH;
#pragma unroll
        for (int c = 0; c < MF_CH; c++) {
          cc[c] = *pc; uu[c] = *pu; ww[c] = *pv;
          pc += vstr; pu += n2i; pv += n2i;
        }
        nld--;
      }
    };
    // partial bottom cells (vertical_mix.F90:1777-1784): zc = DZU(k), zn = DZU(k+1), zf = DZU(k+2) ride ahead in registers
    const double* pz = PBC ? g.DZU + q + (size_t)5 * n2 : nullptr;
    double zc = 0.0, zn = 0.0, zf = 0.0;
    if (PBC) {
      const double* z2 = g.DZU + q + (size_t)2 * n2;
      zc = z2[0];
      zn = (3 <= km + 1) ? z2[n2] : 0.0;
      zf = (4 <= km + 1) ? z2[2 * n2] : 0.0;
    }
    auto fwd_level = [&](int k, double vvc, double ru, double rw) {
      double hfac;
      C = A;
      if (PBC) {
        const double zk = zc;
        hfac = zk / c2dtu;
        A = g.aidif * vvc / (0.5 * (zk + zn));
        zc = zn; zn = zf;
        if (k + 3 <= km + 1) {
          zf = *pz;
          if (MF_PD && k + 3 + MF_PD <= km + 1) prefetch_l2(pz + (ptrdiff_t)MF_PD * n2i);
        }
        pz += n2i;
      } else {
        hfac = c_vc.hfac_u[k];
        A = c_vc.afac_u[k] * vvc;
      }
      if (k <= kmu) {
        const RcpD rd = rcp_prepare((k < kmu) ? hfac + A + B : hfac + B);
        const double e = div_by(A, rd);
        sE[(k - 1) * MF_THREADS] = e;
        B = (hfac + B) * e;
        F1 = div_by(hfac * ru + C * F1, rd);
        F2 = div_by(hfac * rw + C * F2, rd);
      } else {
        F1 = 0.0;
        F2 = 0.0;
      }
      *su = F1;
      *sv = F2;
      su += n2i;
      sv += n2i;
    };
#pragma unroll
    for (int s = 0; s < MF_NB - 1; s++) load_fwd(cb[s], ub[s], wb[s]);
    int k = 2;
    for (int ch = 0; ch < nfull; ch += MF_NB) {
#pragma unroll
      for (int s = 0; s < MF_NB; s++)
        if (ch + s < nfull) {
          load_fwd(cb[(s + MF_NB - 1) % MF_NB], ub[(s + MF_NB - 1) % MF_NB], wb[(s + MF_NB - 1) % MF_NB]);
#pragma unroll
          for (int c = 0; c < MF_CH; c++) fwd_level(k + c, cb[s][c], ub[s][c], wb[s][c]);
          k += MF_CH;
        }
    }
    for (; k <= km; k++) {  // the last (km-1) % MF_CH levels
      const double vvc = *pc, ru = *pu, rw = *pv;
      pc += vstr; pu += n2i; pv += n2i;
      fwd_level(k, vvc, ru, rw);
    }
    // ---- back substitution (only levels k < kmu change; F1,F2 = F(km)), fused with U = Uold + dU when finishing
    if (MF_PD && finish) {  // Uold, Vold are touched for the first time on the way up
#pragma unroll
      for (int c = 1; c <= MF_PD; c++)
        if (c < km) { prefetch_l2(UOLD + q + top - (size_t)c * n2); prefetch_l2(VOLD + q + top - (size_t)c * n2); }
    }
    if (finish) {
      const double u = UOLD[top + q] + F1, v = VOLD[top + q] + F2;
      sE[(km - 1) * MF_THREADS] = u;
      Vn[top] = v;
    }
    const double* qu = Un + top - n2i;  // next level to load (km-1), going up
    const double* qv = Vn + top - n2i;
    const double* qo = finish ? UOLD + q + top - n2i : Un;
    const double* qp = finish ? VOLD + q + top - n2i : Vn;
    double* tu = Un + top - n2i;        // next level to store
    double* tv = Vn + top - n2i;
    const int nfullb = (km - 1) / MF_BCH;
    nld = nfullb;
    kpf = km - 1 - MF_PD;
    double f1[MF_BNB][MF_BCH], f2[MF_BNB][MF_BCH], o1[MF_BNB][MF_BCH], o2[MF_BNB][MF_BCH];
#pragma unroll
    for (int s = 0; s < MF_BNB; s++)
#pragma unroll
      for (int c = 0; c < MF_BCH; c++) { f1[s][c] = 0.0; f2[s][c] = 0.0; o1[s][c] = 0.0; o2[s][c] = 0.0; }
    auto load_bwd = [&](double* a1, double* a2, double* b1, double* b2) {
      if (nld > 0) {
        if (MF_PD && finish && kpf - MF_BCH + 1 >= 1) {
#pragma unroll
          for (int c = 0; c < MF_BCH; c++) {
            prefetch_l2(qo - pdn - (ptrdiff_t)c * n2i);
            prefetch_l2(qp - pdn - (ptrdiff_t)c * n2i);
          }
        }
        kpf -= MF_BCH;
#pragma unroll
        for (int c = 0; c < MF_BCH; c++) {
          a1[c] = *qu; a2[c] = *qv;
          qu -= n2i; qv -= n2i;
          if (finish) { b1[c] = *qo; b2[c] = *qp; qo -= n2i; qp -= n2i; }
        }
        nld--;
      }
    };
    auto bwd_level = [&](int kk, double a1, double a2, double b1, double b2) {
      if (kk < kmu) {
        const double e = sE[(kk - 1) * MF_THREADS];
        a1 = a1 + e * F1;
        a2 = a2 + e * F2;
        if (!finish) { *tu = a1; *tv = a2; }
      }
      F1 = a1;
      F2 = a2;
      if (finish) {
        sE[(kk - 1) * MF_THREADS] = b1 + a1;  // U(k); E(k) is not needed any more
        *tv = b2 + a2;
      }
      tu -= n2i;
      tv -= n2i;
    };
#pragma unroll
    for (int s = 0; s < MF_BNB - 1; s++) load_bwd(f1[s], f2[s], o1[s], o2[s]);
    k = km - 1;
    for (int ch = 0; ch < nfullb; ch += MF_BNB) {
#pragma unroll
      for (int s = 0; s < MF_BNB; s++)
        if (ch + s < nfullb) {
          load_bwd(f1[(s + MF_BNB - 1) % MF_BNB], f2[(s + MF_BNB - 1) % MF_BNB], o1[(s + MF_BNB - 1) % MF_BNB],
                   o2[(s + MF_BNB - 1) % MF_BNB]);
#pragma unroll
          for (int c = 0; c < MF_BCH; c++) bwd_level(k - c, f1[s][c], f2[s][c], o1[s][c], o2[s][c]);
          k -= MF_BCH;
        }
    }
    for (; k >= 1; k--) {
      const double a1 = *qu, a2 = *qv, b1 = finish ? *qo : 0.0, b2 = finish ? *qp : 0.0;
      qu -= n2i; qv -= n2i;
      if (finish) { qo -= n2i; qp -= n2i; }
      bwd_level(k, a1, a2, b1, b2);
    }
    if (!finish) return;
  } else {
    if (!finish) return;
    // explicit vertical mixing: U = Uold + dU (baroclinic.F90:1077-1080)
    const double* po = UOLD + q;
    const double* pp = VOLD + q;
    double* tv = Vn;
    const double* pu = Un;
#pragma unroll 4
    for (int k = 1; k <= km; k++) {
      sE[(k - 1) * MF_THREADS] = *po + *pu;
      *tv = *pp + *tv;
      po += n2i; pp += n2i; pu += n2i; tv += n2i;
    }
  }
  // ---- remove the vertical mean, KMU mask (baroclinic.F90:1085-1129) [+ barotropic velocity, step_mod.F90:581-592]
  const double hur = g.HUR[q];
  double w1 = 0.0, w2 = 0.0;
  if (PBC) {  // baroclinic.F90:1097-1107
    const double* pv = Vn;
    const double* pz = g.DZU + q + n2;
#pragma unroll 4
    for (int k = 1; k <= km; k++) {
      const double dzk = *pz;
      w1 = w1 + sE[(k - 1) * MF_THREADS] * dzk;
      w2 = w2 + *pv * dzk;
      pv += n2i;
      pz += n2i;
    }
  } else {
    const double* pv = Vn;
#pragma unroll 8
    for (int k = 1; k <= km; k++) {
      w1 = w1 + sE[(k - 1) * MF_THREADS] * c_vc.dz[k];
      w2 = w2 + *pv * c_vc.dz[k];
      pv += n2i;
    }
  }
  w1 = w1 * hur;
  w2 = w2 * hur;
  const bool addbt = (UB != nullptr) && (j != bt_skip_row);
  const double ub = addbt ? UB[q] : 0.0, vb = addbt ? VB[q] : 0.0;
  {
    double* tu = Un;
    double* tv = Vn;
#pragma unroll 8
    for (int k = 1; k <= km; k++) {
      double u = 0.0, v = 0.0;
      if (k <= kmu) {
        u = sE[(k - 1) * MF_THREADS] - w1;
        v = *tv - w2;
        if (addbt) { u = u + ub; v = v + vb; }
      }
      *tu = u;
      *tv = v;
      tu += n2i;
      tv += n2i;
      if (k == 1 && sfc_uv != nullptr && j != bt_skip_row) {
        // pop_step_coupled: the surface velocities go straight to the (pinned, device-mapped) host buffer of the coupler
        // while the rest of the column is still being written; the tripole seam row is final only after the halo updates
        const size_t sq = (size_t)(j - (g.jb - 1)) * nxg + (i - (g.ib - 1));
        sfc_uv[sq] = u;
        sfc_uv[(size_t)(g.je - g.jb + 1) * nxg + sq] = v;
      }
    }
  }
}

// ---- the velocity finish with the operands staged by TMA (opt-in, POP_B200_THOMAS_TMA=1; bit-identical, measured slower:
// 14.3 vs 11.5 ms at tx0.1v3) ---
// Same passes as above; the level rows of the CTA's 128 columns (VVC and the two right-hand sides on the way down,
// Uold and Vold on the way up) arrive as 128 x 1 x 1 TMA boxes in an MTS-deep shared-memory ring filled by one elected
// thread across the sweep boundary; F1,F2 are re-read from L2 by plain loads one chunk ahead.
// Shared memory per CTA: km + 3*MTS*MTC KB (62 + 48 KB): two CTAs per SM.
#ifndef MTC
#define MTC 4
#endif
#ifndef MTS
#define MTS 4
#endif
struct MfTmaArgs {
  GridView g;
  double *UNEW, *VNEW;
  const double *UOLD, *VOLD, *UB, *VB;
  int bt_skip_row;
  PopTmap tmU, tmV, tmUo, tmVo, tmC;  // UNEW, VNEW, UOLD, VOLD (nxb,nyb,km); VVC (nxb,nyb,vvc_nk)
};
__global__ void __launch_bounds__(MF_THREADS, 2)
momentum_finish_tma_kernel(const POP_GRID_CONSTANT MfTmaArgs a) {
  POP_DYN_SMEM(smem_raw);
  const GridView& g = a.g;
  const int km = g.km, tid = threadIdx.x;
  double* sEb = (double*)smem_raw;
  double* ring = sEb + (size_t)km * MF_THREADS;  // [MTS][MTC][3][MF_THREADS]
  uint64_t* bar = (uint64_t*)(ring + (size_t)MTS * MTC * 3 * MF_THREADS);
  double* sE = sEb + tid;
  const int i0 = (g.ib - 1) + blockIdx.x * MF_THREADS;
  const int j = (g.jb - 1) + blockIdx.y;
  const int i = i0 + tid;
  const bool active = (i <= g.ie - 1);
  const size_t n2 = g.n2;
  const int n2i = (int)n2;
  const size_t q = (size_t)j * g.nxb + (active ? i : i0);
  const int kmu = active ? g.KMU[q] : 0;
  double* Un = a.UNEW + q;
  double* Vn = a.VNEW + q;
  const size_t top = (size_t)(km - 1) * n2;
  const int nch = (km - 1 + MTC - 1) / MTC;
  const int total = 2 * nch;
  auto issue = [&](int gi) {
    const int slot = gi % MTS;
    double* st = ring + (size_t)slot * MTC * 3 * MF_THREADS;
    if (gi < nch) {
      const int k0 = 2 + gi * MTC;
      const int nl = (km - k0 + 1 < MTC) ? km - k0 + 1 : MTC;
      mbar_expect_tx(&bar[slot], (uint32_t)(nl * 3 * MF_THREADS * 8));
      for (int c = 0; c < nl; c++) {
        const int k = k0 + c;
        tma_load_tile(st + (size_t)(c * 3) * MF_THREADS, &a.tmC, i0, j, (g.vvc_nk == 1) ? 0 : k - 1, &bar[slot]);
        tma_load_tile(st + (size_t)(c * 3 + 1) * MF_THREADS, &a.tmU, i0, j, k - 1, &bar[slot]);
        tma_load_tile(st + (size_t)(c * 3 + 2) * MF_THREADS, &a.tmV, i0, j, k - 1, &bar[slot]);
      }
    } else {
      const int k0 = km - 1 - (gi - nch) * MTC;
      const int nl = (k0 < MTC) ? k0 : MTC;
      mbar_expect_tx(&bar[slot], (uint32_t)(nl * 2 * MF_THREADS * 8));
      for (int c = 0; c < nl; c++) {
        tma_load_tile(st + (size_t)(c * 3) * MF_THREADS, &a.tmUo, i0, j, k0 - c - 1, &bar[slot]);
        tma_load_tile(st + (size_t)(c * 3 + 1) * MF_THREADS, &a.tmVo, i0, j, k0 - c - 1, &bar[slot]);
      }
    }
  };
  if (tid == 0) {
    for (int sl = 0; sl < MTS; sl++) mbar_init(&bar[sl], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int gi = 0; gi < MTS && gi < total; gi++) issue(gi);
  int gc = 0;
  auto next_chunk = [&]() {
    __syncthreads();
    if (tid == 0 && gc + MTS < total) issue(gc + MTS);
    gc++;
  };
  double A = 0.0, B = 0.0, C = 0.0, F1 = 0.0, F2 = 0.0;
  if (active) {
    const double hfac = c_vc.hfac_u[1];
    A = c_vc.afac_u[1] * g.VVC[q];
    const RcpD rd = rcp_prepare(hfac + A);
    const double e = div_by(A, rd);
    sE[0] = e;
    B = hfac * e;
    F1 = div_by(hfac * Un[0], rd);
    F2 = div_by(hfac * Vn[0], rd);
    Un[0] = F1;
    Vn[0] = F2;
  }
  // ---- forward elimination
  {
    double* su = Un + n2i;
    double* sv = Vn + n2i;
    for (int r = 0; r < nch; r++) {
      const double* st = ring + (size_t)(gc % MTS) * MTC * 3 * MF_THREADS + tid;
      mbar_wait(&bar[gc % MTS], (uint32_t)((gc / MTS) & 1));
      const int k0 = 2 + r * MTC;
      if (active) {
#pragma unroll
        for (int c = 0; c < MTC; c++) {
          const int k = k0 + c;
          if (k <= km) {
            const double vvc = st[(size_t)(c * 3) * MF_THREADS];
            const double ru = st[(size_t)(c * 3 + 1) * MF_THREADS], rw = st[(size_t)(c * 3 + 2) * MF_THREADS];
            const double hfac = c_vc.hfac_u[k];
            C = A;
            A = c_vc.afac_u[k] * vvc;
            if (k <= kmu) {
              const RcpD rd = rcp_prepare((k < kmu) ? hfac + A + B : hfac + B);
              const double e = div_by(A, rd);
              sE[(size_t)(k - 1) * MF_THREADS] = e;
              B = (hfac + B) * e;
              F1 = div_by(hfac * ru + C * F1, rd);
              F2 = div_by(hfac * rw + C * F2, rd);
            } else {
              F1 = 0.0;
              F2 = 0.0;
            }
            *su = F1;
            *sv = F2;
            su += n2i;
            sv += n2i;
          }
        }
      }
      next_chunk();
    }
  }
  // ---- back substitution fused with U = Uold + dU (U(k) recycles the E(k) slot, V(k) goes back in place)
  if (active) {
    const double u = a.UOLD[top + q] + F1, v = a.VOLD[top + q] + F2;
    sE[(size_t)(km - 1) * MF_THREADS] = u;
    Vn[top] = v;
  }
  {
    const double* qu = Un + top - n2i;  // F1, F2 of the next level to load (km-1), going up
    const double* qv = Vn + top - n2i;
    double* tv = Vn + top - n2i;
    double fa1[MTC], fa2[MTC], fn1[MTC], fn2[MTC];
#pragma unroll
    for (int c = 0; c < MTC; c++) { fa1[c] = fa2[c] = fn1[c] = fn2[c] = 0.0; }
    int kl = km - 1;
    auto load_f = [&](double* f1, double* f2) {
#pragma unroll
      for (int c = 0; c < MTC; c++)
        if (active && kl - c >= 1) { f1[c] = qu[-(ptrdiff_t)c * n2i]; f2[c] = qv[-(ptrdiff_t)c * n2i]; }
      qu -= (ptrdiff_t)MTC * n2i;
      qv -= (ptrdiff_t)MTC * n2i;
      kl -= MTC;
    };
    load_f(fa1, fa2);
    for (int r = 0; r < nch; r++) {
      const double* st = ring + (size_t)(gc % MTS) * MTC * 3 * MF_THREADS + tid;
      load_f(fn1, fn2);
      mbar_wait(&bar[gc % MTS], (uint32_t)((gc / MTS) & 1));
      const int k0 = km - 1 - r * MTC;
      if (active) {
#pragma unroll
        for (int c = 0; c < MTC; c++) {
          const int k = k0 - c;
          if (k >= 1) {
            double a1 = fa1[c], a2 = fa2[c];
            if (k < kmu) {
              const double e = sE[(size_t)(k - 1) * MF_THREADS];
              a1 = a1 + e * F1;
              a2 = a2 + e * F2;
            }
            F1 = a1;
            F2 = a2;
            sE[(size_t)(k - 1) * MF_THREADS] = st[(size_t)(c * 3) * MF_THREADS] + a1;
            *tv = st[(size_t)(c * 3 + 1) * MF_THREADS] + a2;
            tv -= n2i;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < MTC; c++) { fa1[c] = fn1[c]; fa2[c] = fn2[c]; }
      next_chunk();
    }
  }
  if (!active) return;
  // ---- remove the vertical mean, KMU mask [+ barotropic velocity]
  const double hur = g.HUR[q];
  double w1 = 0.0, w2 = 0.0;
  {
    const double* pv = Vn;
#pragma unroll 8
    for (int k = 1; k <= km; k++) {
      w1 = w1 + sE[(size_t)(k - 1) * MF_THREADS] * c_vc.dz[k];
      w2 = w2 + *pv * c_vc.dz[k];
      pv += n2i;
    }
  }
  w1 = w1 * hur;
  w2 = w2 * hur;
  const bool addbt = (a.UB != nullptr) && (j != a.bt_skip_row);
  const double ub = addbt ? a.UB[q] : 0.0, vb = addbt ? a.VB[q] : 0.0;
  {
    double* tu = Un;
    double* tv = Vn;
#pragma unroll 8
    for (int k = 1; k <= km; k++) {
      double u = 0.0, v = 0.0;
      if (k <= kmu) {
        u = sE[(size_t)(k - 1) * MF_THREADS] - w1;
        v = *tv - w2;
        if (addbt) { u = u + ub; v = v + vb; }
      }
      *tu = u;
      *tv = v;
      tu += n2i;
      tv += n2i;
    }
  }
}

static int launch_finish(double* UNEW, double* VNEW, const double* UOLD, const double* VOLD, const double* UB,
                         const double* VB, int bt_skip_row, int implicit_vmix, int finish) {
  GridView g = grid_view();
  dim3 block(MF_THREADS, 1, 1), grid((unsigned)((G.nxg + MF_THREADS - 1) / MF_THREADS), (unsigned)G.ny_local, 1);
  if (implicit_vmix && finish && !G.no_tma && G.thomas_tma && !G.cfg.partial_bottom_cells) {
    MfTmaArgs ta;
    if (make_tmap_box(&ta.tmU, UNEW, G.km, MF_THREADS, 1) && make_tmap_box(&ta.tmV, VNEW, G.km, MF_THREADS, 1) &&
        make_tmap_box(&ta.tmUo, UOLD, G.km, MF_THREADS, 1) && make_tmap_box(&ta.tmVo, VOLD, G.km, MF_THREADS, 1) &&
        make_tmap_box(&ta.tmC, g.VVC, g.vvc_nk, MF_THREADS, 1)) {
      ta.g = g; ta.UNEW = UNEW; ta.VNEW = VNEW; ta.UOLD = UOLD; ta.VOLD = VOLD; ta.UB = UB; ta.VB = VB;
      ta.bt_skip_row = bt_skip_row;
      const size_t smem_t = sizeof(double) * MF_THREADS * ((size_t)G.km + (size_t)MTS * MTC * 3) + 8 * MTS + 16;
#ifndef POP_EMUL
      POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)momentum_finish_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)momentum_finish_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
      POP_LAUNCH(momentum_finish_tma_kernel, grid, block, smem_t, ta);
      return pop_post_launch("momentum_finish");
    }
  }
  const size_t smem = sizeof(double) * MF_THREADS * (size_t)G.km;
  auto kfn = g.DZU ? momentum_finish_kernel<true> : momentum_finish_kernel<false>;
#ifndef POP_EMUL
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  POP_CHECK_CUDA(cudaFuncSetAttribute((const void*)kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
#endif
  // (only the call that also adds the barotropic velocities produces the final surface level)
  double* sfc_uv = (UB != nullptr && finish) ? G.cio.uv_dev : nullptr;
  POP_LAUNCH(kfn, grid, block, smem, g, UNEW, VNEW, UOLD, VOLD, UB, VB, bt_skip_row, implicit_vmix,
             finish, G.c2dtu, sfc_uv, G.nxg);
  if (sfc_uv) G.cio.uv_sent = true;
  return pop_post_launch("momentum_finish");
}

int impvmixu_dev(double* UNEW, double* VNEW) {
  ScopedTimer tm("VMIX_MOMENTUM_IMPLICIT");
  return launch_finish(UNEW, VNEW, nullptr, nullptr, nullptr, nullptr, -1, 1, 0);
}
int momentum_finish(double* UNEW, double* VNEW, const double* UOLD, const double* VOLD, const double* UB,
                    const double* VB, int bt_skip_row) {
  ScopedTimer tm("MOMENTUM_FINISH");
  return launch_finish(UNEW, VNEW, UOLD, VOLD, UB, VB, bt_skip_row, G.cfg.implicit_vertical_mix, 1);
}

// =====================================================================================
// grad / div (operators.F90:126-192, :49-119)
// =====================================================================================
__global__ void grad_kernel(int k, double* __restrict__ GX, double* __restrict__ GY,
                            const double* __restrict__ F, const int* __restrict__ KMU,
                            const double* __restrict__ DXUR, const double* __restrict__ DYUR, int nxb,
                            int nyb) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)nxb * nyb) return;
  const int i = (int)(q % nxb), j = (int)(q / nxb);
  double gx = 0.0, gy = 0.0;
  if (i < nxb - 1 && j < nyb - 1 && k <= KMU[q]) {
    const double f11 = F[q + nxb + 1], f00 = F[q], f01 = F[q + nxb], f10 = F[q + 1];
    gx = DXUR[q] * 0.5 * (f11 - f00 - f01 + f10);
    gy = DYUR[q] * 0.5 * (f11 - f00 + f01 - f10);
  }
  GX[q] = gx;
  GY[q] = gy;
}
__global__ void div_kernel(int k, double* __restrict__ D, const double* __restrict__ UX,
                           const double* __restrict__ UY, const int* __restrict__ KMT,
                           const double* __restrict__ DXU, const double* __restrict__ DYU, int nxb,
                           int nyb) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)nxb * nyb) return;
  const int i = (int)(q % nxb), j = (int)(q / nxb);
  double d = 0.0;
  if (i >= 1 && j >= 1 && k <= KMT[q]) {
    const size_t s = q - nxb, w = q - 1, sw = q - nxb - 1;
    d = 0.5 * (UX[q] * DYU[q] + UX[s] * DYU[s] - UX[w] * DYU[w] - UX[sw] * DYU[sw] + UY[q] * DXU[q] +
               UY[w] * DXU[w] - UY[s] * DXU[s] - UY[sw] * DXU[sw]);
  }
  D[q] = d;
}

int grad_dev(int k, double* GX, double* GY, const double* F) {
  POP_LAUNCH(grad_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, k, GX, GY, F, fldi("KMU"), fld("DXUR"),
             fld("DYUR"), G.nxb, G.nyb);
  return pop_post_launch("grad");
}
int div_dev(int k, double* D, const double* UX, const double* UY) {
  POP_LAUNCH(div_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, k, D, UX, UY, fldi("KMT"), fld("DXU"),
             fld("DYU"), G.nxb, G.nyb);
  return pop_post_launch("div");
}
