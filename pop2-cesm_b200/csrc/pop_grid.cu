// pop_grid.cu -- init-time derivation of every grid-dependent array of the hot path from the nine
// primary metric fields + KMT + dz (host code; runs once, results live in HBM).
//   grid.F90:587-647 (reciprocals/areas, closed-boundary fix-up), :786-803 (vertical grid),
//   :978-1041 (KMU, HU/HUR), :1131-1146 (uarea_equator, FCOR), :2882-2932 (cf_area_avg);
//   hmix_del2.F90:97-421,428-663 and hmix_del4.F90:94-387,394-590 (mixing coefficients);
//   advection.F90:387-396,420-562 (KXU/KYU, upwind3 tables); state_mod.F90:1040-1063,1765-1766;
//   pressure_grad.F90:168-175; vertical_mix.F90:383-448; POP_SolversMod.F90:783-820,895-906;
//   barotropic.F90:170-252 (checkerboard masks, initial diagonal).
// Ghost cells of the primaries are filled by the device halo update with the field location the
// reference's scatter uses (grid.F90:1420-1520), so the multi-rank strips see their neighbours.
#include <cmath>
#include <cstring>
#include <vector>
#include "pop_state.cuh"

namespace {
typedef std::vector<double> HV;
typedef std::vector<int> HI;

struct Host2 {  // 1-based accessors on a padded (nxb, nyb) host array with end-off (zero) shifts
  int nxb, nyb;
  inline size_t ix(int i, int j) const { return (size_t)(j - 1) * nxb + (i - 1); }
  inline double eo(const HV& a, int i, int j) const {
    return (i < 1 || i > nxb || j < 1 || j > nyb) ? 0.0 : a[ix(i, j)];
  }
  inline int eoi(const HI& a, int i, int j) const {
    return (i < 1 || i > nxb || j < 1 || j > nyb) ? 0 : a[ix(i, j)];
  }
};

int put(const char* name, const HV& v) {
  POP_TRY(alloc_field(name, (int)(v.size() / G.n2), false));
  POP_CHECK_CUDA(cudaMemcpyAsync(fld(name), v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
int puti(const char* name, const HI& v) {
  POP_TRY(alloc_field(name, (int)(v.size() / G.n2), true));
  POP_CHECK_CUDA(cudaMemcpyAsync(fldi(name), v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
int get(const char* name, HV& v) {
  POP_CHECK_CUDA(cudaMemcpyAsync(v.data(), fld(name), v.size() * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
// scatter_global semantics: physical strip -> padded block, ghost cells by halo update
int scatter_halo(const char* name, const double* strip, int loc, HV& out) {
  out.assign(G.n2, 0.0);
  for (int j = 0; j < G.ny_local; j++)
    for (int i = 0; i < G.nxg; i++)
      out[(size_t)(j + POP_NGHOST) * G.nxb + (i + POP_NGHOST)] = strip[(size_t)j * G.nxg + i];
  POP_TRY(put(name, out));
  POP_TRY(halo_update(fld(name), 1, loc, POP_KIND_SCALAR, 0.0));
  return get(name, out);
}
int halo_host(const char* name, HV& v, int loc) {
  POP_TRY(put(name, v));
  POP_TRY(halo_update(fld(name), 1, loc, POP_KIND_SCALAR, 0.0));
  return get(name, v);
}
double pressure(double depth) {  // state_mod.F90:1765-1766
  return 0.059808 * (exp(-0.025 * depth) - 1.0) + 0.100766 * depth + 2.28405e-7 * (depth * depth);
}
}  // namespace

int set_grid_host(const double* ULAT_S, const double* HTN_S, const double* HTE_S, const double* HUS_S,
                  const double* HUW_S, const double* DXU_S, const double* DYU_S, const double* DXT_S,
                  const double* DYT_S, const int* KMT_S, const double* dz_in) {
  const int nxb = G.nxb, nyb = G.nyb, km = G.km;
  const size_t n2 = G.n2;
  const pop_config& c = G.cfg;
  Host2 H{nxb, nyb};
#define IX(i, j) H.ix(i, j)
#define ALL2 for (int j = 1; j <= nyb; j++) for (int i = 1; i <= nxb; i++)

  HV ULAT, HTN, HTE, HUS, HUW, DXU, DYU, DXT, DYT;
  POP_TRY(scatter_halo("ULAT", ULAT_S, POP_LOC_NECORNER, ULAT));
  POP_TRY(scatter_halo("HTN", HTN_S, POP_LOC_NFACE, HTN));
  POP_TRY(scatter_halo("DXU", DXU_S, POP_LOC_NECORNER, DXU));
  POP_TRY(scatter_halo("DXT", DXT_S, POP_LOC_CENTER, DXT));
  POP_TRY(scatter_halo("HTE", HTE_S, POP_LOC_EFACE, HTE));
  POP_TRY(scatter_halo("DYT", DYT_S, POP_LOC_CENTER, DYT));
  POP_TRY(scatter_halo("DYU", DYU_S, POP_LOC_NECORNER, DYU));
  POP_TRY(scatter_halo("HUS", HUS_S, POP_LOC_EFACE, HUS));
  POP_TRY(scatter_halo("HUW", HUW_S, POP_LOC_NFACE, HUW));
  HI KMT(n2, 0);
  for (int j = 0; j < G.ny_local; j++)
    for (int i = 0; i < G.nxg; i++)
      KMT[(size_t)(j + POP_NGHOST) * nxb + (i + POP_NGHOST)] = KMT_S[(size_t)j * G.nxg + i];
  POP_TRY(puti("KMT", KMT));
  POP_TRY(halo_update_i4(fldi("KMT"), 1, POP_LOC_CENTER, POP_KIND_SCALAR, 0));
  // (ordered on the library stream: G.stream is non-blocking, a legacy-stream copy would race the halo)
  POP_CHECK_CUDA(cudaMemcpyAsync(KMT.data(), fldi("KMT"), n2 * sizeof(int), cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  for (HV* f : {&HTN, &HTE, &HUS, &HUW, &DXU, &DYU, &DXT, &DYT})  // grid.F90:1531-1538
    for (size_t q = 0; q < n2; q++)
      if ((*f)[q] <= 0.0) (*f)[q] = 1.0;

  // ---- vertical grid: grid.F90:786-803
  VertConst& vc = G.vc;
  memset(&vc, 0, sizeof(vc));
  for (int k = 1; k <= km; k++) vc.dz[k] = dz_in[k - 1];
  vc.dzw[0] = 0.5 * vc.dz[1];
  vc.dzw[km] = 0.5 * vc.dz[km];
  vc.dzwr[0] = 1.0 / vc.dzw[0];
  vc.zw[1] = vc.dz[1];
  vc.zt[1] = vc.dzw[0];
  for (int k = 1; k <= km - 1; k++) {
    vc.dzw[k] = 0.5 * (vc.dz[k] + vc.dz[k + 1]);
    vc.zw[k + 1] = vc.zw[k] + vc.dz[k + 1];
    vc.zt[k + 1] = vc.zt[k] + vc.dzw[k];
  }
  for (int k = 1; k <= km; k++) {
    vc.c2dz[k] = 2.0 * vc.dz[k];
    vc.dzr[k] = 1.0 / vc.dz[k];
    vc.dz2r[k] = 1.0 / vc.c2dz[k];
    vc.dzwr[k] = 1.0 / vc.dzw[k];
  }

  // ---- closed-boundary fix-up and reciprocals: grid.F90:587-647
  const int ib = G.ib, ie = G.ie, jb = G.jb, je = G.je;
  const std::vector<int>&ig = G.i_glob, &jg = G.j_glob;
  HV* four[4] = {&DXU, &DYU, &DXT, &DYT};
  if (ig[0] == 0)
    for (int j = 1; j <= nyb; j++)
      for (int i = 1; i <= ib - 1; i++)
        for (HV* f : four) (*f)[IX(i, j)] = (*f)[IX(ib, j)];
  if (ig[ie] == 0)
    for (int j = 1; j <= nyb; j++)
      for (int i = ie + 1; i <= nxb; i++)
        for (HV* f : four) (*f)[IX(i, j)] = (*f)[IX(ie, j)];
  if (jg[0] == 0)
    for (int j = 1; j <= jb - 1; j++)
      for (int i = 1; i <= nxb; i++)
        for (HV* f : four) (*f)[IX(i, j)] = (*f)[IX(i, jb)];
  if (jg[je] == 0)
    for (int j = je + 1; j <= nyb; j++)
      for (int i = 1; i <= nxb; i++)
        for (HV* f : four) (*f)[IX(i, j)] = (*f)[IX(i, je)];
  HV DXUR(n2), DYUR(n2), UAREA(n2), UAREA_R(n2), DXTR(n2), DYTR(n2), TAREA(n2), TAREA_R(n2);
  for (size_t q = 0; q < n2; q++) {
    DXUR[q] = 1.0 / DXU[q];
    DYUR[q] = 1.0 / DYU[q];
    UAREA[q] = DXU[q] * DYU[q];
    UAREA_R[q] = 1.0 / UAREA[q];
    DXTR[q] = 1.0 / DXT[q];
    DYTR[q] = 1.0 / DYT[q];
    TAREA[q] = DXT[q] * DYT[q];
    TAREA_R[q] = 1.0 / TAREA[q];
  }
  // ---- cf_area_avg: grid.F90:2908-2928
  HV AU0(n2), AUN(n2), AUE(n2), AUNE(n2);
  ALL2 {
    const size_t q = IX(i, j);
    AU0[q] = TAREA[q] * 0.25 * UAREA_R[q];
    AUN[q] = H.eo(TAREA, i, j + 1) * 0.25 * UAREA_R[q];
    AUE[q] = H.eo(TAREA, i + 1, j) * 0.25 * UAREA_R[q];
    AUNE[q] = H.eo(TAREA, i + 1, j + 1) * 0.25 * UAREA_R[q];
  }
  // ---- KMU: grid.F90:978-990 (+ halo update, NEcorner scalar)
  HI KMU(n2, 0);
  for (int j = 1; j <= nyb - 1; j++)
    for (int i = 1; i <= nxb - 1; i++) {
      int m = KMT[IX(i, j)];
      if (KMT[IX(i + 1, j)] < m) m = KMT[IX(i + 1, j)];
      if (KMT[IX(i, j + 1)] < m) m = KMT[IX(i, j + 1)];
      if (KMT[IX(i + 1, j + 1)] < m) m = KMT[IX(i + 1, j + 1)];
      KMU[IX(i, j)] = m;
    }
  POP_TRY(puti("KMU", KMU));
  POP_TRY(halo_update_i4(fldi("KMU"), 1, POP_LOC_NECORNER, POP_KIND_SCALAR, 0));
  POP_CHECK_CUDA(cudaMemcpyAsync(KMU.data(), fldi("KMU"), n2 * sizeof(int), cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  // ---- partial bottom cells: DZT, DZU (grid.F90:917-966)
  const bool pbc = G.cfg.partial_bottom_cells != 0;
  HV DZT, DZU;
  if (pbc) {
    POP_REQUIRE(G.dzbc_strip.size() == (size_t)G.nxg * G.ny_local,
                "pop_set_grid: partial_bottom_cells needs pop_set_bottom_cells (DZBC) first");
    HV DZBC;
    POP_TRY(scatter_halo("DZBC", G.dzbc_strip.data(), POP_LOC_CENTER, DZBC));
    DZT.assign(n2 * (size_t)(km + 2), 0.0);
    DZU.assign(n2 * (size_t)(km + 2), 0.0);
    for (int k = 1; k <= km; k++) {
      double *T = DZT.data() + (size_t)k * n2, *U = DZU.data() + (size_t)k * n2;
      for (size_t q = 0; q < n2; q++) T[q] = (KMT[q] == k) ? DZBC[q] : vc.dz[k];
      for (int j = 1; j <= nyb - 1; j++)
        for (int i = 1; i <= nxb - 1; i++) {
          double m = T[IX(i, j)];
          if (T[IX(i + 1, j)] < m) m = T[IX(i + 1, j)];
          if (T[IX(i, j + 1)] < m) m = T[IX(i, j + 1)];
          if (T[IX(i + 1, j + 1)] < m) m = T[IX(i + 1, j + 1)];
          U[IX(i, j)] = m;
        }
    }
    POP_TRY(put("DZU", DZU));
    POP_TRY(halo_update(fld("DZU"), km + 2, POP_LOC_NECORNER, POP_KIND_SCALAR, 0.0));
    POP_TRY(get("DZU", DZU));
  }
  // ---- HT, HU, HUR, masks: grid.F90:1001-1041
  HV HT(n2), HU(n2), HUR(n2), RCALCT(n2), RCALCU(n2), FCOR(n2);
  for (size_t q = 0; q < n2; q++) {
    const int kt = KMT[q], ku = KMU[q];
    if (pbc) {
      HT[q] = (kt >= 1 && kt <= km) ? vc.zw[kt - 1] + DZT[(size_t)kt * n2 + q] : 0.0;
      if (ku >= 1 && ku <= km) {
        HU[q] = vc.zw[ku - 1] + DZU[(size_t)ku * n2 + q];
        HUR[q] = 1.0 / HU[q];
      } else {
        HU[q] = 0.0;
        HUR[q] = 0.0;
      }
      for (int k = 1; k <= km; k++)
        if (k > ku) DZU[(size_t)k * n2 + q] = vc.dz[k];  // "to prevent divide by zero", grid.F90:1013
    } else {
    HT[q] = (kt >= 1 && kt <= km) ? vc.zw[kt] : 0.0;
    HU[q] = (ku >= 1 && ku <= km) ? vc.zw[ku] : 0.0;
    HUR[q] = (ku >= 1 && ku <= km) ? 1.0 / vc.zw[ku] : 0.0;
    }
    RCALCT[q] = (kt >= 1) ? 1.0 : 0.0;
    RCALCU[q] = (ku >= 1) ? 1.0 : 0.0;
    FCOR[q] = 2.0 * POP_OMEGA * sin(ULAT[q]);  // grid.F90:1146
  }
  // ---- uarea_equator: grid.F90:1130-1139
  {
    double m = 1.0e300;
    for (int j = jb; j <= je; j++)
      for (int i = ib; i <= ie; i++)
        if (RCALCU[IX(i, j)] != 0.0 && fabs(ULAT[IX(i, j)]) < m) m = fabs(ULAT[IX(i, j)]);
    POP_TRY(comm_allreduce_min(&m));
    double a = 1.0e300;
    for (int j = jb; j <= je; j++)
      for (int i = ib; i <= ie; i++)
        if (RCALCU[IX(i, j)] != 0.0) {
          const double w = (fabs(ULAT[IX(i, j)]) == m) ? UAREA[IX(i, j)] : 1.0e20;
          if (w < a) a = w;
        }
    POP_TRY(comm_allreduce_min(&a));
    G.uarea_equator = a;
  }
  // ---- state tables (state_mod.F90:1040-1063) and Boussinesq correction (pressure_grad.F90:168-175)
  for (int k = 1; k <= km; k++) {
    vc.pressz[k] = pressure(vc.zt[k] * 0.01);
    mwjf_level_coefficients(vc.pressz[k], &vc.eos_n0t0[k], &vc.eos_n0t2[k], &vc.eos_n1t0[k], &vc.eos_d0t0[k],
                            &vc.eos_d0t1[k], &vc.eos_d0t3[k]);
    if (c.state_itype == POP_STATE_MWJF) { vc.tmin[k] = -2.0; vc.tmax[k] = 999.0; vc.smin[k] = 0.0; vc.smax[k] = 0.999; }
    else { vc.tmin[k] = -2.0; vc.tmax[k] = 40.0; vc.smin[k] = 0.0; vc.smax[k] = 0.042; }
    vc.bouss[k] = c.lbouss_correct
                      ? 1.0 / (1.02819 + 4.4004e-5 * vc.pressz[k] - 2.93161e-4 * exp(-0.05 * vc.pressz[k]))
                      : 1.0;
    vc.afac_u[k] = c.aidif * vc.dzwr[k];  // vertical_mix.F90:444-448
    vc.afac_t[k] = c.aidif * vc.dzwr[k];
  }
  // ---- init_advection: advection.F90:387-396 (KXU, KYU), :420-562 (upwind3 tables)
  HV KXU(n2), KYU(n2);
  ALL2 {
    KXU[IX(i, j)] = (H.eo(HUW, i + 1, j) - HUW[IX(i, j)]) * UAREA_R[IX(i, j)];
    KYU[IX(i, j)] = (H.eo(HUS, i, j + 1) - HUS[IX(i, j)]) * UAREA_R[IX(i, j)];
  }
  if (G.use_upwind3) {
    const double* dz = vc.dz;
    std::vector<double> dzc(km + 3, 0.0);
    dzc[0] = dz[1];
    for (int k = 1; k <= km; k++) dzc[k] = dz[k];
    dzc[km + 1] = dzc[km];
    for (int k = 1; k <= km - 1; k++) {
      vc.talfzp[k] = dz[k] * (2.0 * dz[k] + dzc[k - 1]) / ((dz[k] + dz[k + 1]) * (dzc[k - 1] + 2.0 * dz[k] + dz[k + 1]));
      vc.tbetzp[k] = dz[k + 1] * (2.0 * dz[k] + dzc[k - 1]) / ((dz[k] + dz[k + 1]) * (dz[k] + dzc[k - 1]));
      vc.tgamzp[k] = -(dz[k] * dz[k + 1]) / ((dz[k] + dzc[k - 1]) * (dz[k + 1] + dzc[k - 1] + 2.0 * dz[k]));
    }
    vc.tbetzp[1] = vc.tbetzp[1] + vc.tgamzp[1];
    vc.tgamzp[1] = 0.0;
    vc.talfzp[km] = 0.0; vc.tbetzp[km] = 0.0; vc.tgamzp[km] = 0.0;
    for (int k = 1; k <= km - 1; k++) {
      vc.talfzm[k] = dz[k] * (2.0 * dz[k + 1] + dzc[k + 2]) / ((dz[k] + dz[k + 1]) * (dz[k + 1] + dzc[k + 2]));
      vc.tbetzm[k] = dz[k + 1] * (2.0 * dz[k + 1] + dzc[k + 2]) / ((dz[k] + dz[k + 1]) * (dz[k] + dzc[k + 2] + 2.0 * dz[k + 1]));
      vc.tdelzm[k] = -(dz[k] * dz[k + 1]) / ((dz[k + 1] + dzc[k + 2]) * (dz[k] + dzc[k + 2] + 2.0 * dz[k + 1]));
    }
    vc.talfzm[km - 1] = vc.talfzm[km - 1] + vc.tdelzm[km - 1];
    vc.tdelzm[km - 1] = 0.0;
    vc.talfzm[km] = 0.0; vc.tbetzm[km] = 0.0; vc.tdelzm[km] = 0.0;
    HV AXP(n2, 0.0), BXP(n2, 0.0), GXP(n2, 0.0), AYP(n2, 0.0), BYP(n2, 0.0), GYP(n2, 0.0);
    HV AXM(n2, 0.0), BXM(n2, 0.0), DXM(n2, 0.0), AYM(n2, 0.0), BYM(n2, 0.0), DYM(n2, 0.0);
    for (int j = jb; j <= je; j++)
      for (int i = ib - 1; i <= ie; i++) {
        const double dxc = DXT[IX(i, j)], dxcw = DXT[IX(i - 1, j)], dxce = DXT[IX(i + 1, j)], dxce2 = DXT[IX(i + 2, j)];
        const size_t q = IX(i, j);
        AXP[q] = dxc * (2.0 * dxc + dxcw) / ((dxc + dxce) * (dxcw + 2.0 * dxc + dxce));
        BXP[q] = dxce * (2.0 * dxc + dxcw) / ((dxc + dxcw) * (dxc + dxce));
        GXP[q] = -(dxc * dxce) / ((dxc + dxcw) * (dxcw + 2.0 * dxc + dxce));
        AXM[q] = dxc * (2.0 * dxce + dxce2) / ((dxc + dxce) * (dxce + dxce2));
        BXM[q] = dxce * (2.0 * dxce + dxce2) / ((dxc + dxce) * (dxc + 2.0 * dxce + dxce2));
        DXM[q] = -(dxc * dxce) / ((dxce2 + dxce) * (dxc + 2.0 * dxce + dxce2));
      }
    for (int j = jb - 1; j <= je; j++)
      for (int i = ib; i <= ie; i++) {
        const double dyc = DYT[IX(i, j)], dycs = DYT[IX(i, j - 1)], dycn = DYT[IX(i, j + 1)], dycn2 = DYT[IX(i, j + 2)];
        const size_t q = IX(i, j);
        AYP[q] = dyc * (2.0 * dyc + dycs) / ((dyc + dycn) * (dycs + 2.0 * dyc + dycn));
        BYP[q] = dycn * (2.0 * dyc + dycs) / ((dyc + dycn) * (dycs + dyc));
        GYP[q] = -(dyc * dycn) / ((dyc + dycs) * (dycs + 2.0 * dyc + dycn));
        AYM[q] = dyc * (2.0 * dycn + dycn2) / ((dyc + dycn) * (dycn + dycn2));
        BYM[q] = dycn * (2.0 * dycn + dycn2) / ((dyc + dycn) * (dyc + 2.0 * dycn + dycn2));
        DYM[q] = -(dyc * dycn) / ((dycn2 + dycn) * (dyc + 2.0 * dycn + dycn2));
      }
    POP_TRY(put("TALFXP", AXP)); POP_TRY(put("TBETXP", BXP)); POP_TRY(put("TGAMXP", GXP));
    POP_TRY(put("TALFYP", AYP)); POP_TRY(put("TBETYP", BYP)); POP_TRY(put("TGAMYP", GYP));
    POP_TRY(put("TALFXM", AXM)); POP_TRY(put("TBETXM", BXM)); POP_TRY(put("TDELXM", DXM));
    POP_TRY(put("TALFYM", AYM)); POP_TRY(put("TBETYM", BYM)); POP_TRY(put("TDELYM", DYM));
  }

  // ---- init_del2t (hmix_del2.F90:428-663) / init_del4t (hmix_del4.F90:394-590)
  {
    const bool del4 = (c.hmix_tracer_itype == POP_HMIX_DEL4);
    G.ah = c.ah;
    if (c.lauto_hmixt) G.ah = del4 ? -0.2e20 * (1280.0 / (double)c.nx_global) : 1.0e7 * (720.0 / (double)c.nx_global);
    HV AHF(n2, 1.0), DTN(n2), DTS(n2), DTE(n2), DTW(n2), W(n2);
    if (c.lvariable_hmixt) {
      double den = 2.0 * POP_PI * POP_RADIUS / c.nx_global;
      den = den * den;
      for (size_t q = 0; q < n2; q++) AHF[q] = del4 ? pow(TAREA[q] / G.uarea_equator, 1.5) : sqrt(TAREA[q] / den);
      POP_TRY(halo_host("AHF", AHF, POP_LOC_CENTER));
    }
    ALL2 W[IX(i, j)] = del4 ? HTN[IX(i, j)] / HUW[IX(i, j)]
                            : (HTN[IX(i, j)] / HUW[IX(i, j)]) * 0.5 * (AHF[IX(i, j)] + H.eo(AHF, i, j + 1));
    ALL2 {
      DTN[IX(i, j)] = W[IX(i, j)] * TAREA_R[IX(i, j)];
      DTS[IX(i, j)] = H.eo(W, i, j - 1) * TAREA_R[IX(i, j)];
    }
    ALL2 W[IX(i, j)] = del4 ? HTE[IX(i, j)] / HUS[IX(i, j)]
                            : (HTE[IX(i, j)] / HUS[IX(i, j)]) * 0.5 * (AHF[IX(i, j)] + H.eo(AHF, i + 1, j));
    ALL2 {
      DTE[IX(i, j)] = W[IX(i, j)] * TAREA_R[IX(i, j)];
      DTW[IX(i, j)] = H.eo(W, i - 1, j) * TAREA_R[IX(i, j)];
    }
    POP_TRY(put("AHF", AHF)); POP_TRY(put("DTN", DTN)); POP_TRY(put("DTS", DTS));
    POP_TRY(put("DTE", DTE)); POP_TRY(put("DTW", DTW));
  }
  // ---- init_del2u (hmix_del2.F90:97-421) / init_del4u (hmix_del4.F90:94-387)
  {
    const bool del4 = (c.hmix_momentum_itype == POP_HMIX_DEL4);
    G.am = c.am;
    if (c.lauto_hmixu) G.am = del4 ? -0.6e20 * (1280.0 / (double)c.nx_global) : 1.0e7 * (720.0 / (double)c.nx_global);
    HV AMF(n2, 1.0);
    if (c.lvariable_hmixu) {
      double den = 2.0 * POP_PI * POP_RADIUS / c.nx_global;
      den = den * den;
      for (size_t q = 0; q < n2; q++) AMF[q] = del4 ? pow(UAREA[q] / G.uarea_equator, 1.5) : sqrt(UAREA[q] / den);
      POP_TRY(halo_host("AMF", AMF, POP_LOC_NECORNER));
    }
    HV DUC(n2), DUN(n2), DUS(n2), DUE(n2), DUW(n2), DMC(n2), DMN(n2), DMS(n2), DME(n2), DMW(n2), DUM(n2);
    HV W1(n2), W2(n2), DXKX(n2), DYKY(n2), DXKY(n2), DYKX(n2);
    if (del4) { ALL2 W1[IX(i, j)] = HUS[IX(i, j)] / HTE[IX(i, j)]; }
    else { ALL2 W1[IX(i, j)] = (HUS[IX(i, j)] / HTE[IX(i, j)]) * 0.5 * (AMF[IX(i, j)] + H.eo(AMF, i, j - 1)); }
    ALL2 { DUS[IX(i, j)] = W1[IX(i, j)] * UAREA_R[IX(i, j)]; DUN[IX(i, j)] = H.eo(W1, i, j + 1) * UAREA_R[IX(i, j)]; }
    if (del4) { ALL2 W1[IX(i, j)] = HUW[IX(i, j)] / HTN[IX(i, j)]; }
    else { ALL2 W1[IX(i, j)] = (HUW[IX(i, j)] / HTN[IX(i, j)]) * 0.5 * (AMF[IX(i, j)] + H.eo(AMF, i - 1, j)); }
    ALL2 { DUW[IX(i, j)] = W1[IX(i, j)] * UAREA_R[IX(i, j)]; DUE[IX(i, j)] = H.eo(W1, i + 1, j) * UAREA_R[IX(i, j)]; }
    // KXT
    ALL2 W1[IX(i, j)] = (HTE[IX(i, j)] - H.eo(HTE, i - 1, j)) * TAREA_R[IX(i, j)];
    if (del4) {
      ALL2 W2[IX(i, j)] = H.eo(W1, i + 1, j) - W1[IX(i, j)];
      ALL2 DXKX[IX(i, j)] = 0.5 * (W2[IX(i, j)] + H.eo(W2, i, j + 1)) * DXUR[IX(i, j)];
      ALL2 W2[IX(i, j)] = H.eo(W1, i, j + 1) - W1[IX(i, j)];
      ALL2 DYKX[IX(i, j)] = 0.5 * (W2[IX(i, j)] + H.eo(W2, i + 1, j)) * DYUR[IX(i, j)];
    } else {
      ALL2 W2[IX(i, j)] = 0.5 * (W1[IX(i, j)] + H.eo(W1, i, j + 1)) * 0.5 * (H.eo(AMF, i - 1, j) + AMF[IX(i, j)]);
      ALL2 DXKX[IX(i, j)] = (H.eo(W2, i + 1, j) - W2[IX(i, j)]) * DXUR[IX(i, j)];
      ALL2 W2[IX(i, j)] = 0.5 * (W1[IX(i, j)] + H.eo(W1, i + 1, j)) * 0.5 * (H.eo(AMF, i, j - 1) + AMF[IX(i, j)]);
      ALL2 DYKX[IX(i, j)] = (H.eo(W2, i, j + 1) - W2[IX(i, j)]) * DYUR[IX(i, j)];
    }
    // KYT
    ALL2 W1[IX(i, j)] = (HTN[IX(i, j)] - H.eo(HTN, i, j - 1)) * TAREA_R[IX(i, j)];
    if (del4) {
      ALL2 W2[IX(i, j)] = H.eo(W1, i, j + 1) - W1[IX(i, j)];
      ALL2 DYKY[IX(i, j)] = 0.5 * (W2[IX(i, j)] + H.eo(W2, i + 1, j)) * DYUR[IX(i, j)];
      ALL2 W2[IX(i, j)] = H.eo(W1, i + 1, j) - W1[IX(i, j)];
      ALL2 DXKY[IX(i, j)] = 0.5 * (W2[IX(i, j)] + H.eo(W2, i, j + 1)) * DXUR[IX(i, j)];
    } else {
      ALL2 W2[IX(i, j)] = 0.5 * (W1[IX(i, j)] + H.eo(W1, i + 1, j)) * 0.5 * (H.eo(AMF, i, j - 1) + AMF[IX(i, j)]);
      ALL2 DYKY[IX(i, j)] = (H.eo(W2, i, j + 1) - W2[IX(i, j)]) * DYUR[IX(i, j)];
      ALL2 W2[IX(i, j)] = 0.5 * (W1[IX(i, j)] + H.eo(W1, i, j + 1)) * 0.5 * (H.eo(AMF, i - 1, j) + AMF[IX(i, j)]);
      ALL2 DXKY[IX(i, j)] = (H.eo(W2, i + 1, j) - W2[IX(i, j)]) * DXUR[IX(i, j)];
    }
    if (del4) {
      ALL2 {
        const size_t q = IX(i, j);
        DUM[q] = -(DXKX[q] + DYKY[q] + 2.0 * (KXU[q] * KXU[q] + KYU[q] * KYU[q]));
        DMC[q] = DXKY[q] - DYKX[q];
        DME[q] = 2.0 * KYU[q] / (HTN[q] + H.eo(HTN, i + 1, j));
        DMN[q] = -2.0 * KXU[q] / (HTE[q] + H.eo(HTE, i, j + 1));
      }
    } else {
      ALL2 {
        const size_t q = IX(i, j);
        DUM[q] = -(DXKX[q] + DYKY[q] + 2.0 * AMF[q] * (KXU[q] * KXU[q] + KYU[q] * KYU[q]));
        DMC[q] = DXKY[q] - DYKX[q];
      }
      ALL2 W1[IX(i, j)] = (H.eo(AMF, i, j + 1) - H.eo(AMF, i, j - 1)) / (HTE[IX(i, j)] + H.eo(HTE, i, j + 1));
      ALL2 { const size_t q = IX(i, j); DME[q] = (2.0 * AMF[q] * KYU[q] + W1[q]) / (HTN[q] + H.eo(HTN, i + 1, j)); }
      ALL2 W1[IX(i, j)] = (H.eo(AMF, i + 1, j) - H.eo(AMF, i - 1, j)) / (HTN[IX(i, j)] + H.eo(HTN, i + 1, j));
      ALL2 { const size_t q = IX(i, j); DMN[q] = -(2.0 * AMF[q] * KXU[q] + W1[q]) / (HTE[q] + H.eo(HTE, i, j + 1)); }
    }
    ALL2 {
      const size_t q = IX(i, j);
      DUC[q] = -(DUN[q] + DUS[q] + DUE[q] + DUW[q]);
      DMW[q] = -DME[q];
      DMS[q] = -DMN[q];
    }
    POP_TRY(put("AMF", AMF)); POP_TRY(put("DUC", DUC)); POP_TRY(put("DUN", DUN)); POP_TRY(put("DUS", DUS));
    POP_TRY(put("DUE", DUE)); POP_TRY(put("DUW", DUW)); POP_TRY(put("DMC", DMC)); POP_TRY(put("DMN", DMN));
    POP_TRY(put("DMS", DMS)); POP_TRY(put("DME", DME)); POP_TRY(put("DMW", DMW)); POP_TRY(put("DUM", DUM));
  }

  // ---- POP_SolversInit: 9-point weights, POP_SolversMod.F90:783-820
  HV WNE(n2, 0.0), WE(n2, 0.0), WN(n2, 0.0), CIND(n2, 0.0), MASK(n2, 0.0);
  for (int j = 2; j <= nyb; j++)
    for (int i = 2; i <= nxb; i++) {
      const double xne = 0.25 * HU[IX(i, j)] * DXUR[IX(i, j)] * DYU[IX(i, j)];
      const double xse = 0.25 * HU[IX(i, j - 1)] * DXUR[IX(i, j - 1)] * DYU[IX(i, j - 1)];
      const double xnw = 0.25 * HU[IX(i - 1, j)] * DXUR[IX(i - 1, j)] * DYU[IX(i - 1, j)];
      const double xsw = 0.25 * HU[IX(i - 1, j - 1)] * DXUR[IX(i - 1, j - 1)] * DYU[IX(i - 1, j - 1)];
      const double yne = 0.25 * HU[IX(i, j)] * DYUR[IX(i, j)] * DXU[IX(i, j)];
      const double yse = 0.25 * HU[IX(i, j - 1)] * DYUR[IX(i, j - 1)] * DXU[IX(i, j - 1)];
      const double ynw = 0.25 * HU[IX(i - 1, j)] * DYUR[IX(i - 1, j)] * DXU[IX(i - 1, j)];
      const double ysw = 0.25 * HU[IX(i - 1, j - 1)] * DYUR[IX(i - 1, j - 1)] * DXU[IX(i - 1, j - 1)];
      const size_t q = IX(i, j);
      const double ne = xne + yne, ase = xse + yse, anw = xnw + ynw, asw = xsw + ysw;
      WNE[q] = ne;
      WE[q] = xne + xse - yne - yse;
      WN[q] = yne + ynw - xne - xnw;
      CIND[q] = -(ne + ase + anw + asw);
      MASK[q] = RCALCT[q];
    }
  // ---- init_barotropic: barotropic.F90:170-252
  HI CHECKER(n2, 0), CONSTNT(n2, 0);
  HV CHK(n2, 0.0), CST(n2, 0.0), CHKA(n2, 0.0), CSTA(n2, 0.0), DC(n2, 0.0), CWC(n2);
  if (c.sfc_layer_type == POP_SFC_VARTHICK) {
    ALL2 {
      const size_t q = IX(i, j);
      const int n = ig[i - 1] + abs(jg[j - 1]);
      const int chk = 2 * (n % 2) - 1;
      if (KMT[q] > 0) {
        CHECKER[q] = chk; CONSTNT[q] = 1;
        CHKA[q] = chk * TAREA[q]; CSTA[q] = TAREA[q];
      }
      CHK[q] = CHECKER[q];
      CST[q] = CONSTNT[q];
    }
  }
  for (size_t q = 0; q < n2; q++) {
    if (c.sfc_layer_type == POP_SFC_RIGID) DC[q] = 0.0;
    else DC[q] = (RCALCT[q] != 0.0) ? TAREA[q] / (G.alpha * 2.0 * G.dtp * G.dtp * POP_GRAV) : 0.0;
    CWC[q] = CIND[q] - DC[q];
  }

  // ---- upload
  POP_TRY(put("DXU", DXU)); POP_TRY(put("DYU", DYU)); POP_TRY(put("DXT", DXT)); POP_TRY(put("DYT", DYT));
  POP_TRY(put("HTN", HTN)); POP_TRY(put("HTE", HTE)); POP_TRY(put("HUS", HUS)); POP_TRY(put("HUW", HUW));
  POP_TRY(put("DXUR", DXUR)); POP_TRY(put("DYUR", DYUR)); POP_TRY(put("DXTR", DXTR)); POP_TRY(put("DYTR", DYTR));
  POP_TRY(put("UAREA", UAREA)); POP_TRY(put("UAREA_R", UAREA_R)); POP_TRY(put("TAREA", TAREA));
  POP_TRY(put("TAREA_R", TAREA_R)); POP_TRY(put("AU0", AU0)); POP_TRY(put("AUN", AUN)); POP_TRY(put("AUE", AUE));
  POP_TRY(put("AUNE", AUNE)); POP_TRY(put("HT", HT)); POP_TRY(put("HU", HU)); POP_TRY(put("HUR", HUR));
  if (pbc) { POP_TRY(put("DZT", DZT)); POP_TRY(put("DZU", DZU)); }
  POP_TRY(put("RCALCT", RCALCT)); POP_TRY(put("RCALCU", RCALCU)); POP_TRY(put("FCOR", FCOR));
  POP_TRY(put("KXU", KXU)); POP_TRY(put("KYU", KYU));
  POP_TRY(put("btropWgtNE", WNE)); POP_TRY(put("btropWgtEast", WE)); POP_TRY(put("btropWgtNorth", WN));
  POP_TRY(put("centerWgtClinicIndep", CIND)); POP_TRY(put("mMaskTropic", MASK));
  POP_TRY(put("centerWgtClinic", CWC)); POP_TRY(put("btropWgtCenter", CWC));
  POP_TRY(puti("CHECKER", CHECKER)); POP_TRY(puti("CONSTNT", CONSTNT));
  // the operator weights are derived from HU one row/column beyond each cell: on the outermost ghost row
  // of a strip that neighbour does not exist locally, so take the owner's values (the two-iteration
  // P-CSI pass evaluates the operator on the first ghost ring and must reproduce the owner's bits)
  // (a cyclic north-south boundary on one rank is the same situation: the strip is its own neighbour)
  if (G.nranks > 1 || c.ns_boundary_type == POP_BNDY_CYCLIC)
    for (const char* w : {"btropWgtNE", "btropWgtEast", "btropWgtNorth", "centerWgtClinicIndep"})
      POP_TRY(halo_rows_only(fld(w), 1));
  for (const char* w : {"BT_R", "BT_S", "BT_Q", "BT_Z", "BT_AZ", "BT_A0R"}) POP_TRY(alloc_field(w, 1, false));
  POP_TRY(alloc_field("BT_PCSI", 6, false));  // three (X, Q) pairs of the fused PCSI passes (two in rotation, one held by a pending check)
  // vmix_const init: VVC = const_vvc, VDC = const_vdc
  if (c.vmix_itype == POP_VMIX_CONST) {
    HV v((size_t)G.vdc_nk * G.vdc_nd * n2, c.const_vdc);
    POP_TRY(put("VDC", v));
    HV w((size_t)G.vvc_nk * n2, c.const_vvc);
    POP_TRY(put("VVC", w));
  }
  G.grid_set = true;
  G.rf_ready = false;  // budget areas depend on KMT/TAREA
  G.gm_dirty = true;
  G.lw_coef_dirty = true;
  G.lw_flux_ready = false;
  POP_TRY(set_timestep(POP_TS_LEAPFROG));
  POP_TRY(upload_vert_const());
  POP_TRY(solvers_init_dev());
  if (c.sfc_layer_type == POP_SFC_VARTHICK) {
    // sum_check / sum_const are integer sums in the reference (barotropic.F90:200-215)
    double s[4];
    POP_TRY(put("W2A", CHK)); POP_TRY(global_sum_dev(fld("W2A"), 1, n2, POP_LOC_CENTER, nullptr, &s[0]));
    POP_TRY(put("W2A", CST)); POP_TRY(global_sum_dev(fld("W2A"), 1, n2, POP_LOC_CENTER, nullptr, &s[1]));
    POP_TRY(put("W2A", CHKA)); POP_TRY(global_sum_dev(fld("W2A"), 1, n2, POP_LOC_CENTER, nullptr, &s[2]));
    POP_TRY(put("W2A", CSTA)); POP_TRY(global_sum_dev(fld("W2A"), 1, n2, POP_LOC_CENTER, nullptr, &s[3]));
    const long sum_check = lround(s[0]), sum_const = lround(s[1]);
    const double acheck = s[2] / s[3];
    G.rcheck = acheck / ((double)sum_const - acheck * (double)sum_check);
    G.rconst = 1.0 / ((double)sum_const - acheck * (double)sum_check);
  }
  return POP_SUCCESS;
#undef IX
#undef ALL2
}
