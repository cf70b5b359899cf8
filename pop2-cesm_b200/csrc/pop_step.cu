// pop_step.cu -- time-step drivers on the library-resident state.
//   dhdt                        surface_hgt.F90:131-286 (+ tgrid_to_ugrid, grid.F90:3362-3415)
//   baroclinic_driver           baroclinic.F90:578-1210
//   baroclinic_correct_adjust   baroclinic.F90:1217-1497
//   step                        step_mod.F90:296-626, :634-640, :663-832
// The reference loops over blocks and levels on the host and calls one operator per slab; here each
// driver is a short sequence of fused column kernels plus the halo updates the reference issues at
// the same points of the step.
#include <utility>
#include "pop_dev.cuh"
#include "pop_state.cuh"

// ------------------------------------------------------------------ dhdt
__global__ void dhdt_kernel(double* __restrict__ DH, const double* __restrict__ Pc,
                            const double* __restrict__ Po, const double* __restrict__ FW_OLD, double dtp,
                            int sfc, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  double v;
  if (sfc == POP_SFC_VARTHICK) v = (Pc[q] - Po[q]) / (POP_GRAV * dtp) - FW_OLD[q];
  else if (sfc == POP_SFC_RIGID) v = 0.0;
  else v = (Pc[q] - Po[q]) / (POP_GRAV * dtp);
  DH[q] = v;
}
__global__ void t2u_kernel(double* __restrict__ AU, const double* __restrict__ AT,
                           const double* __restrict__ AU0, const double* __restrict__ AUN,
                           const double* __restrict__ AUE, const double* __restrict__ AUNE,
                           const double* __restrict__ RCALCU, int nxb, int nyb) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (size_t)nxb * nyb) return;
  const int i = (int)(q % nxb), j = (int)(q / nxb);
  double v = 0.0;
  if (i < nxb - 1 && j < nyb - 1)
    v = AU0[q] * AT[q] + AUN[q] * AT[q + nxb] + AUE[q] * AT[q + 1] + AUNE[q] * AT[q + nxb + 1];
  if (RCALCU[q] == 0.0) v = 0.0;
  AU[q] = v;
}
int dhdt_dev() {
  POP_LAUNCH(dhdt_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("DH"), fld_t("PSURF", G.curtime),
             fld_t("PSURF", G.oldtime), fld("FW_OLD"), G.dtp, G.cfg.sfc_layer_type, G.n2);
  POP_LAUNCH(t2u_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("DHU"), fld("DH"), fld("AU0"), fld("AUN"),
             fld("AUE"), fld("AUNE"), fld("RCALCU"), G.nxb, G.nyb);
  return pop_post_launch("dhdt");
}

// ------------------------------------------------------------------ baroclinic_driver
// array row (0-based) of the tripole seam on the rank that owns it, else -1
static int tripole_seam_row() {
  return (G.cfg.ns_boundary_type == POP_BNDY_TRIPOLE && G.rank == G.nranks - 1) ? G.je - 1 : -1;
}
int momentum_finish_new(bool add_barotropic) {
  return momentum_finish(fld_t("UVEL", G.newtime), fld_t("VVEL", G.newtime), fld_t("UVEL", G.oldtime),
                         fld_t("VVEL", G.oldtime), add_barotropic ? fld_t("UBTROP", G.newtime) : nullptr,
                         add_barotropic ? fld_t("VBTROP", G.newtime) : nullptr, tripole_seam_row());
}

int baroclinic_driver_dev(bool defer_finish) {
  ScopedTimer tm("BAROCLINIC");
  const int o = G.oldtime, c = G.curtime, n_ = G.newtime, mx = G.mixtime;
  const bool pavg = G.cfg.lpressure_avg && G.leapfrogts;
  const bool varthick = (G.cfg.sfc_layer_type == POP_SFC_VARTHICK);
  POP_TRY(gm_begin_step());
  POP_TRY(vmix_coeffs_dev(1, G.km, fld_t("TRACER", mx), fld_t("UVEL", mx), fld_t("VVEL", mx), fld_t("RHO", mx)));
  {
    ScopedTimer t2("TRACER_UPDATE");
    TracerIO io;
    io.TCUR = fld_t("TRACER", c); io.TMIX = fld_t("TRACER", mx); io.TOLD = fld_t("TRACER", o);
    io.UCUR = fld_t("UVEL", c); io.VCUR = fld_t("VVEL", c);
    io.STF = fld("STF"); io.TFW = fld("TFW"); io.DH = fld("DH");
    if (G.cio.stf_dev) { io.STF = G.cio.stf_dev; io.stf_strip = true; }  // pop_step_coupled, pinned host buffer
    io.POLD = fld_t("PSURF", o); io.PCUR = fld_t("PSURF", c);
    io.TNEW = fld_t("TRACER", n_);
    io.WTK = nullptr;
    POP_TRY(tracer_column(TR_FULL, 0, io));
  }
  if (G.cfg.implicit_vertical_mix) {  // baroclinic.F90:878-896
    if (!varthick)
      POP_TRY(impvmixt_dev(fld_t("TRACER", n_), fld_t("TRACER", o), fld_t("PSURF", c), nullptr, 1, G.nt, 0));
    else if (pavg)
      POP_TRY(impvmixt_dev(fld_t("TRACER", n_), fld_t("TRACER", o), fld_t("PSURF", c), nullptr, 1, 2, 0));
  }
  if (pavg) {
    // T,S at the new time are needed in the ghost cells for the pressure average: baroclinic.F90:919-935
    POP_TRY(halo_update(fld_t("TRACER", n_), 2 * G.km, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
    POP_TRY(state_3d(fld_t("TRACER", n_), fld_t("RHO", n_)));  // baroclinic.F90:981-985
  }
  POP_TRY(coupled_wait_forcing());  // SMF is first read by the momentum kernel
  {
    ScopedTimer t2("CLINIC");
    MomentumIO io;
    io.UCUR = fld_t("UVEL", c); io.VCUR = fld_t("VVEL", c); io.UOLD = fld_t("UVEL", o); io.VOLD = fld_t("VVEL", o);
    io.UMIX = fld_t("UVEL", mx); io.VMIX = fld_t("VVEL", mx);
    io.RHOOLD = fld_t("RHO", o); io.RHOCUR = fld_t("RHO", c); io.RHONEW = fld_t("RHO", n_);
    io.SMF = fld("SMF"); io.DHU = fld("DHU");
    io.UNEW = fld_t("UVEL", n_); io.VNEW = fld_t("VVEL", n_); io.ZX = fld("ZX"); io.ZY = fld("ZY");
    io.WUK = nullptr;
    {
      ScopedTimer t3("MOMENTUM_COLUMN");
      POP_TRY(momentum_column(MO_FULL, 0, io));
    }
    if (!defer_finish) POP_TRY(momentum_finish(io.UNEW, io.VNEW, io.UOLD, io.VOLD));
  }
  return POP_SUCCESS;
}

// convad (vertical_mix.F90:1888-2027): nconvad passes over the odd, then the even level pairs of a column; a pair
// that is statically unstable after adiabatic displacement is mixed to its thickness-weighted mean. One thread
// per column: the passes are sequential in k by construction. dttxcel = 1 (time_management.F90:1005-1010).
__global__ void convad_kernel(StateOpt so, double* T, const int* __restrict__ KMT, size_t n2, int km,
                              int nt, int nconvad, const double* __restrict__ DZT) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2) return;
  const int kmt = KMT[q];
  double* S = T + (size_t)km * n2;
  for (int nc = 1; nc <= nconvad; nc++)
    for (int ks = 1; ks <= 2; ks++)
      for (int k = ks; k <= km - 1; k += 2) {
        if (!(k < kmt)) break;  // nothing below the bottom is touched (the where mask), k only grows
        const size_t c = (size_t)(k - 1) * n2 + q, c1 = c + n2;
        double rhok, rhokp;
        state_cell(so, k + 1, T[c], S[c], &rhok, nullptr, nullptr, nullptr);
        state_cell(so, k + 1, T[c1], S[c1], &rhokp, nullptr, nullptr, nullptr);
        if (rhok > rhokp) {
          double dztk = c_vc.dz[k] / 1.0, dztk1 = c_vc.dz[k + 1] / 1.0;
          if (DZT) {  // partial bottom cells: vertical_mix.F90:1960-1968
            dztk = DZT[(size_t)k * n2 + q] / 1.0;
            dztk1 = DZT[(size_t)(k + 1) * n2 + q] / 1.0;
          }
          const double dzwx = 1.0 / (dztk + dztk1);
          for (int n = 0; n < nt; n++) {
            double* t = T + (size_t)n * km * n2;
            const double v = dzwx * (dztk * t[c] + dztk1 * t[c1]);
            t[c] = v;
            t[c1] = v;
          }
        }
      }
}

// ------------------------------------------------------------------ baroclinic_correct_adjust
struct SfcArgs {
  double *Tn;
  const double *To, *Tc, *Pn, *Pc, *Po, *Pm;
  double* RHS;
  const int* KMT;
  size_t n2, n3;
  double dz1;
  int nt, mode;
};
// surface-layer corrections of baroclinic.F90:1283-1420; one thread per (i,j), loop over tracers
__global__ void sfc_correct_kernel(SfcArgs a) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n2) return;
  const bool ocean = a.KMT[q] > 0;
  const double Pn = a.Pn[q], Pc = a.Pc[q], Po = a.Po[q], Pm = a.Pm[q];
  for (int n = 0; n < a.nt; n++) {
    double* tn = a.Tn + (size_t)n * a.n3 + q;
    const double to = a.To[(size_t)n * a.n3 + q], tc = a.Tc[(size_t)n * a.n3 + q];
    switch (a.mode) {
      case 0:  // implicit vmix, pressure averaging: RHS for T,S (:1283-1296), passive tracers (:1303-1312)
        if (n < 2)
          a.RHS[(size_t)n * a.n2 + q] =
              ocean ? ((2.0 * tc - to) * (Pc - Po) - *tn * (Pn - Pc)) / (POP_GRAV * a.dz1) : 0.0;
        else if (ocean)
          *tn = *tn - to * (Pn - Po) / (POP_GRAV * a.dz1);
        break;
      case 1:  // implicit vmix, no pressure averaging (:1327-1337)
        if (ocean) *tn = *tn - to * (Pn - Pm) / (POP_GRAV * a.dz1);
        break;
      case 2:  // explicit vmix, pressure averaging (:1355-1392)
        if (n < 2)
          *tn = ocean ? (*tn * (a.dz1 + Pc / POP_GRAV) + (2.0 * tc - to) * (Pc - Po) / POP_GRAV) /
                            (a.dz1 + Pn / POP_GRAV)
                      : 0.0;
        else
          *tn = ocean ? (to * (a.dz1 + Po / POP_GRAV) + a.dz1 * *tn) / (a.dz1 + Pn / POP_GRAV) : 0.0;
        break;
      default:  // explicit vmix, no pressure averaging (:1398-1412)
        *tn = ocean ? (to * (a.dz1 + Pm / POP_GRAV) + a.dz1 * *tn) / (a.dz1 + Pn / POP_GRAV) : 0.0;
        break;
    }
  }
}

int baroclinic_correct_adjust_dev() {
  ScopedTimer tm("BAROCLINIC");
  const int o = G.oldtime, c = G.curtime, n_ = G.newtime, mx = G.mixtime;
  const bool pavg = G.cfg.lpressure_avg && G.leapfrogts;
  double* Tn = fld_t("TRACER", n_);
  if (G.cfg.sfc_layer_type == POP_SFC_VARTHICK) {
    SfcArgs a;
    a.Tn = Tn; a.To = fld_t("TRACER", o); a.Tc = fld_t("TRACER", c);
    a.Pn = fld_t("PSURF", n_); a.Pc = fld_t("PSURF", c); a.Po = fld_t("PSURF", o); a.Pm = fld_t("PSURF", mx);
    a.RHS = fld("RHS1"); a.KMT = fldi("KMT"); a.n2 = G.n2; a.n3 = G.n3; a.dz1 = G.vc.dz[1]; a.nt = G.nt;
    a.mode = G.cfg.implicit_vertical_mix ? (pavg ? 0 : 1) : (pavg ? 2 : 3);
    POP_LAUNCH(sfc_correct_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, a);
    if (G.cfg.implicit_vertical_mix) {
      if (pavg) {
        POP_TRY(impvmixt_dev(Tn, nullptr, a.Pn, a.RHS, 1, 2, 1));
        POP_TRY(impvmixt_dev(Tn, a.To, a.Pn, nullptr, 3, G.nt, 0));
      } else {
        POP_TRY(impvmixt_dev(Tn, a.To, a.Pn, nullptr, 1, G.nt, 0));
      }
    }
  }
  // convad: convection_type = 'diffusion' returns immediately (vertical_mix.F90:1925)
  if (!G.cfg.convection_diff && G.cfg.nconvad > 0)
    POP_LAUNCH(convad_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, StateOpt{G.cfg.state_itype, G.cfg.state_range_iopt}, Tn,
               fldi("KMT"), G.n2, G.km, G.nt, G.cfg.nconvad,
               G.cfg.partial_bottom_cells ? (const double*)fld("DZT") : (const double*)nullptr);
  POP_TRY(state_3d(Tn, fld_t("RHO", n_)));  // baroclinic.F90:1476-1482
  return pop_post_launch("baroclinic_correct_adjust");
}

// ------------------------------------------------------------------ step
__global__ void add_barotropic_kernel(double* __restrict__ U, double* __restrict__ V,
                                      const double* __restrict__ UB, const double* __restrict__ VB,
                                      const int* __restrict__ KMU, size_t n2, size_t q0, size_t nq) {
  const size_t q = q0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y + 1;
  if (q >= q0 + nq) return;
  if (k <= KMU[q]) {
    const size_t c = (size_t)(k - 1) * n2 + q;
    U[c] = U[c] + UB[q];
    V[c] = V[c] + VB[q];
  }
}
__global__ void pguess_kernel(double* __restrict__ PG, const double* __restrict__ Pn,
                              const double* __restrict__ Pc, const double* __restrict__ Po, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) PG[q] = 3.0 * (Pn[q] - Pc[q]) + Po[q];
}
// averaging step (step_mod.F90:663-796): old <- (old+cur)/2 then cur <- (cur+new)/2
__global__ void avg_kernel(double* __restrict__ Fo, double* __restrict__ Fc, const double* __restrict__ Fn,
                           size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double c = Fc[q];
  Fo[q] = 0.5 * (Fo[q] + c);
  Fc[q] = 0.5 * (c + Fn[q]);
}
__global__ void avg_fw_kernel(double* __restrict__ FW_OLD, const double* __restrict__ FW, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) FW_OLD[q] = 0.5 * (FW[q] + FW_OLD[q]);
}
struct AvgSfc {
  double *To, *Tc, *Po, *Pc, *PG;
  const double *Tn, *Pn;
  size_t n2, n3;
  double dz1;
  int nt, varthick;
};
__global__ void avg_surface_kernel(AvgSfc a) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n2) return;
  const double Po = a.Po[q], Pc = a.Pc[q], Pn = a.Pn[q];
  if (a.varthick) {
    const double PFO = 0.5 * (Po + Pc), PFC = 0.5 * (Pc + Pn);
    for (int n = 0; n < a.nt; n++) {
      const size_t c = (size_t)n * a.n3 + q;
      const double to = a.To[c], tc = a.Tc[c], tn = a.Tn[c];
      const double wmin = fmin(to, tc), wmax = fmax(to, tc);
      double v = 0.5 * ((a.dz1 + Po / POP_GRAV) * to + (a.dz1 + Pc / POP_GRAV) * tc);
      v = v / (a.dz1 + PFO / POP_GRAV);
      if (v < wmin) v = wmin;
      if (v > wmax) v = wmax;
      const double wmin2 = fmin(tc, tn), wmax2 = fmax(tc, tn);
      double w = 0.5 * ((a.dz1 + Pc / POP_GRAV) * tc + (a.dz1 + Pn / POP_GRAV) * tn);
      w = w / (a.dz1 + PFC / POP_GRAV);
      if (w < wmin2) w = wmin2;
      if (w > wmax2) w = wmax2;
      a.To[c] = v;
      a.Tc[c] = w;
    }
    a.Po[q] = PFO;
    a.Pc[q] = PFC;
  } else {
    for (int n = 0; n < a.nt; n++) {
      const size_t c = (size_t)n * a.n3 + q;
      const double tc = a.Tc[c];
      a.To[c] = 0.5 * (a.To[c] + tc);
      a.Tc[c] = 0.5 * (tc + a.Tn[c]);
    }
    a.Po[q] = 0.5 * (Po + Pc);
    a.Pc[q] = 0.5 * (Pc + Pn);
  }
  a.PG[q] = 0.5 * (a.PG[q] + Pn);
}
// tracer levels 2..km of the averaging step, all tracers
__global__ void avg_tracer_kernel(double* __restrict__ To, double* __restrict__ Tc,
                                  const double* __restrict__ Tn, size_t n2, int km) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2) return;
  const int z = blockIdx.y;  // 0 .. nt*km-1
  if (z % km == 0) return;   // level 1 is handled by avg_surface_kernel
  const size_t c = (size_t)z * n2 + q;
  const double tc = Tc[c];
  To[c] = 0.5 * (To[c] + tc);
  Tc[c] = 0.5 * (tc + Tn[c]);
}

// ------------------------------------------------------------------ Robert-Asselin-Williams filter
// step_RF, step_mod.F90:919-1354, for the configuration this library covers (variable-thickness surface
// layer, no ice formation, no passive-tracer resets, no marginal seas, so that MASK_TRBUDGET(k) and
// RCALCT_OPEN_OCEAN_3D(k) are both (KMT >= k), init_step :1570-1600). STORE_RF is a diagnostic-only array
// in the reference; here the S terms live in registers and only their column integrals reach memory.
__global__ void rf_filter_kernel(double* __restrict__ Fo, double* __restrict__ Fc, double* __restrict__ Fn,
                                 size_t n, double rc, double rn, int nonzero_new) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double fc = Fc[q], fn = Fn[q];
  const double w = Fo[q] + fn - 2.0 * fc;
  if (nonzero_new) Fn[q] = fn + rn * w;
  Fc[q] = fc + rc * w;
}
struct RfTracer {
  const double *To, *Po, *Pc, *Pn, *TAREA, *RCALCT;
  double *Tc, *Tn, *WI, *WS;  // WI/WS: (n2, nt) interior / surface volume-weighted S
  const int* KMT;
  size_t n2, n3;
  double rc, rn, dz1;
  int km, nonzero_new;
};
// one thread per (column, tracer): levels 2..km (:1031-1066) then the thickness-weighted surface (:1070-1094)
__global__ void rf_tracer_kernel(RfTracer a) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n2) return;
  const int n = blockIdx.y;
  const size_t base = (size_t)n * a.n3 + q;
  const int kmt = a.KMT[q];
  double w = 0.0;
  for (int k = 2; k <= a.km; k++) {
    const size_t c = base + (size_t)(k - 1) * a.n2;
    const double tc = a.Tc[c], tn = a.Tn[c];
    const double s = a.To[c] + tn - 2.0 * tc;
    if (a.nonzero_new) a.Tn[c] = tn + a.rn * s;
    a.Tc[c] = tc + a.rc * s;
    w = w + c_vc.dz[k] * ((kmt >= k) ? 1.0 : 0.0) * s;
  }
  const double tarea = a.TAREA[q];
  a.WI[(size_t)n * a.n2 + q] = tarea * w;
  const double ho = a.dz1 + a.Po[q] / POP_GRAV, hc = a.dz1 + a.Pc[q] / POP_GRAV, hn = a.dz1 + a.Pn[q] / POP_GRAV;
  const double tc = a.Tc[base], tn = a.Tn[base];
  const double s = ho * a.To[base] + hn * tn - 2.0 * hc * tc;
  if (a.nonzero_new) a.Tn[base] = hn * tn + a.rn * s;
  a.Tc[base] = hc * tc + a.rc * s;
  a.WS[(size_t)n * a.n2 + q] = tarea * a.RCALCT[q] * s;
}
// PSURF filter (:1099-1112); WB = S_p * TAREA for the conservation term
__global__ void rf_psurf_kernel(const double* __restrict__ Po, double* __restrict__ Pc, double* __restrict__ Pn,
                                const double* __restrict__ TAREA, double* __restrict__ WB, size_t n,
                                double rc, double rn, int nonzero_new) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double pc = Pc[q], pn = Pn[q];
  const double w = Po[q] + pn - 2.0 * pc;
  if (nonzero_new) Pn[q] = pn + rn * w;
  Pc[q] = pc + rc * w;
  WB[q] = w * TAREA[q];
}
struct RfSurface {
  double *Pc, *Pn, *Tc, *Tn, *WB;
  const double *TAREA, *RCALCT;
  size_t n2, n3;
  double rc, rn, dz1, rf_sump;
  int nt, nonzero_new;
};
// PSURF conservation term, surface tracers back from tracer*thickness (:1120-1141), surface volume (:1160)
__global__ void rf_surface_kernel(RfSurface a) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n2) return;
  const double w2 = (a.RCALCT[q] != 0.0) ? a.rf_sump : 0.0;
  double pn = a.Pn[q];
  if (a.nonzero_new) { pn = pn - a.rn * w2; a.Pn[q] = pn; }
  const double pc = a.Pc[q] - a.rc * w2;
  a.Pc[q] = pc;
  const double hn = a.dz1 + pn / POP_GRAV, hc = a.dz1 + pc / POP_GRAV;
  for (int n = 0; n < a.nt; n++) {
    const size_t c = (size_t)n * a.n3 + q;
    if (a.nonzero_new) a.Tn[c] = a.Tn[c] / hn;
    a.Tc[c] = a.Tc[c] / hc;
  }
  a.WB[q] = a.TAREA[q] * hc;
}
struct RfConserve {
  double f_new[POP_MAX_NT], f_cur[POP_MAX_NT];
};
// conservation adjustment at all levels (:1193-1205)
__global__ void rf_conserve_kernel(double* __restrict__ Tc, double* __restrict__ Tn, const int* __restrict__ KMT,
                                   size_t n2, int km, int nonzero_new, const POP_GRID_CONSTANT RfConserve f) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2) return;
  const int z = blockIdx.y;  // 0 .. nt*km-1
  const int n = z / km, k = z % km + 1;
  const double m3 = (KMT[q] >= k) ? 1.0 : 0.0;
  const size_t c = (size_t)z * n2 + q;
  if (nonzero_new) Tn[c] = Tn[c] - f.f_new[n] * m3;
  Tc[c] = Tc[c] - f.f_cur[n] * m3;
}
__global__ void rf_levelmask_kernel(double* __restrict__ mk, const int* __restrict__ KMT, size_t n, int k) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) mk[q] = (KMT[q] >= k) ? 1.0 : 0.0;
}

static int step_rf_dev() {
  ScopedTimer tm("ROBERT_FILTER");
  POP_REQUIRE(G.cfg.sfc_layer_type == POP_SFC_VARTHICK,
              "pop_step: the Robert filter needs sfc_layer_type varthick (step_mod.F90:1150-1154)");
  const int km = G.km, nt = G.nt;
  const int o = G.oldtime, c = G.curtime, n_ = G.newtime;
  const double rc = 0.5 * G.cfg.robert_nu * G.cfg.robert_alpha;  // time_management.F90:897-898
  const double rn = 0.5 * G.cfg.robert_nu * (G.cfg.robert_alpha - 1.0);
  const int nonzero_new = !(rn == 0.0);
  const double dz1 = G.vc.dz[1];
  if (!G.rf_ready) {  // init_step, step_mod.F90:1558-1600
    G.rf_volume_2_km = 0.0;
    for (int k = 1; k <= km; k++) {
      double area = 0.0;
      POP_LAUNCH(rf_levelmask_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("W2A"), fldi("KMT"), G.n2, k);
      POP_TRY(global_sum_dev(fld("TAREA"), 1, 0, POP_LOC_CENTER, fld("W2A"), &area));
      if (k == 1) G.rf_bgtarea1 = area;
      else G.rf_volume_2_km = G.rf_volume_2_km + area * G.vc.dz[k];
    }
    for (int n = 0; n < POP_MAX_NT; n++) { G.rf_S_prev[n] = 0.0; G.rf_S_prev_valid[n] = false; }
    G.rf_ready = true;
  }
  for (const char* f : {"UBTROP", "VBTROP", "GRADPX", "GRADPY"})
    POP_LAUNCH(rf_filter_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld_t(f, o), fld_t(f, c), fld_t(f, n_), G.n2,
               rc, rn, nonzero_new);
  for (const char* f : {"UVEL", "VVEL"})
    POP_LAUNCH(rf_filter_kernel, ew_grid(G.n3), POP_EW_THREADS, 0, fld_t(f, o), fld_t(f, c), fld_t(f, n_), G.n3,
               rc, rn, nonzero_new);
  RfTracer a;
  a.To = fld_t("TRACER", o); a.Tc = fld_t("TRACER", c); a.Tn = fld_t("TRACER", n_);
  a.Po = fld_t("PSURF", o); a.Pc = fld_t("PSURF", c); a.Pn = fld_t("PSURF", n_);
  a.TAREA = fld("TAREA"); a.RCALCT = fld("RCALCT"); a.KMT = fldi("KMT");
  a.WI = fld("AUX"); a.WS = fld("RHS1");
  a.n2 = G.n2; a.n3 = G.n3; a.rc = rc; a.rn = rn; a.dz1 = dz1; a.km = km; a.nonzero_new = nonzero_new;
  POP_LAUNCH(rf_tracer_kernel, dim3(ew_grid(G.n2), (unsigned)nt, 1), POP_EW_THREADS, 0, a);
  double rf_Svol[POP_MAX_NT], tmp[POP_RED_NF];
  for (int n0 = 0; n0 < nt; n0 += POP_RED_NF) {  // :1066 and :1096
    const int nf = (nt - n0 < POP_RED_NF) ? nt - n0 : POP_RED_NF;
    POP_TRY(global_sum_dev(a.WI + (size_t)n0 * G.n2, nf, G.n2, POP_LOC_CENTER, nullptr, tmp));
    for (int f = 0; f < nf; f++) rf_Svol[n0 + f] = tmp[f];
    POP_TRY(global_sum_dev(a.WS + (size_t)n0 * G.n2, nf, G.n2, POP_LOC_CENTER, nullptr, tmp));
    for (int f = 0; f < nf; f++) rf_Svol[n0 + f] = rf_Svol[n0 + f] + tmp[f];
  }
  double *Pc = fld_t("PSURF", c), *Pn = fld_t("PSURF", n_);
  POP_LAUNCH(rf_psurf_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, a.Po, Pc, Pn, a.TAREA, fld("W2A"), G.n2, rc, rn,
             nonzero_new);
  double rf_sump = 0.0;
  POP_TRY(global_sum_dev(fld("W2A"), 1, 0, POP_LOC_CENTER, a.RCALCT, &rf_sump));  // :1116-1118
  rf_sump = rf_sump / G.rf_bgtarea1;
  RfSurface s;
  s.Pc = Pc; s.Pn = Pn; s.Tc = a.Tc; s.Tn = a.Tn; s.WB = fld("W2A"); s.TAREA = a.TAREA; s.RCALCT = a.RCALCT;
  s.n2 = G.n2; s.n3 = G.n3; s.rc = rc; s.rn = rn; s.dz1 = dz1; s.rf_sump = rf_sump; s.nt = nt;
  s.nonzero_new = nonzero_new;
  POP_LAUNCH(rf_surface_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, s);
  double sfc_vol = 0.0;
  POP_TRY(global_sum_dev(fld("W2A"), 1, 0, POP_LOC_CENTER, a.RCALCT, &sfc_vol));  // :1162-1171
  const double rf_ocean_norm = G.rf_volume_2_km + sfc_vol;
  RfConserve f;
  double rf_S[POP_MAX_NT];
  for (int n = 0; n < nt; n++) {  // :1173-1187
    rf_S[n] = rf_Svol[n] / rf_ocean_norm;
    const double factor =
        (!G.rf_S_prev_valid[n] || nonzero_new) ? rf_S[n] : 0.5 * (rf_S[n] + G.rf_S_prev[n]);
    f.f_new[n] = factor * rn;
    f.f_cur[n] = factor * rc;
  }
  POP_LAUNCH(rf_conserve_kernel, dim3(ew_grid(G.n2), (unsigned)(km * nt), 1), POP_EW_THREADS, 0, a.Tc, a.Tn, a.KMT,
             G.n2, km, nonzero_new, f);
  // :1230-1290
  POP_CHECK_CUDA(cudaMemcpyAsync(fld("FW_OLD"), fld("FW"), sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
  POP_TRY(state_3d(a.Tc, fld_t("RHO", c)));
  POP_TRY(state_3d(a.Tn, fld_t("RHO", n_)));
  POP_LAUNCH(pguess_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("PGUESS"), Pn, Pc, a.Po, G.n2);  // :1305-1311
  const int t = G.oldtime;
  G.oldtime = G.curtime;
  G.curtime = G.newtime;
  G.newtime = t;
  if (!nonzero_new)
    for (int n = 0; n < nt; n++) { G.rf_S_prev[n] = rf_S[n]; G.rf_S_prev_valid[n] = true; }
  return pop_post_launch("step_RF");
}

int step_dev(int ts_type) {
  POP_REQUIRE(G.grid_set, "pop_step: grid not set");
  POP_REQUIRE(ts_type == POP_TS_LEAPFROG || ts_type == POP_TS_EULER || ts_type == POP_TS_AVG ||
                  ts_type == POP_TS_ROBERT,
              "pop_step: unknown time-step type %d (matsuno steps are not supported)", ts_type);
  ScopedTimer tm("STEP");
  POP_TRY(set_timestep(ts_type));
  const int km = G.km;
  POP_TRY(dhdt_dev());
  // Velocity finish (impvmixu + Uold + vertical-mean removal + KMU mask, baroclinic.F90:1067-1129): nothing reads
  // UVEL/VVEL(new) before the halo updates below, so by default it runs as ONE kernel after the barotropic solve and
  // adds UBTROP/VBTROP(new) in its last pass (the reference's separate add at step_mod.F90:581-592; ghost cells get
  // the same bits through the halo update, which copies / sign-flips sums exactly as it copies the addends)
  const bool overlap = (G.finish_mode == 1), fused = (G.finish_mode == 0);
  POP_TRY(baroclinic_driver_dev(overlap || fused));
  if (overlap) {
    // fork: the barotropic solve and the tracer corrector proceed on the main stream meanwhile
    POP_CHECK_CUDA(cudaEventRecord(G.ev_fork, G.stream));
    POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream2, G.ev_fork, 0));
    std::swap(G.stream, G.stream2);
    const int rc = momentum_finish_new(false);
    std::swap(G.stream, G.stream2);
    POP_TRY(rc);
    POP_CHECK_CUDA(cudaEventRecord(G.ev_join, G.stream2));
  }
  POP_TRY(halo_update(fld("ZX"), 1, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));  // step_mod.F90:405-417
  POP_TRY(halo_update(fld("ZY"), 1, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));
  POP_TRY(barotropic_driver_dev());
  POP_TRY(coupled_after_barotropic());
  if (fused) POP_TRY(momentum_finish_new(true));
  POP_TRY(baroclinic_correct_adjust_dev());
  POP_TRY(coupled_after_corrector());
  if (overlap) POP_CHECK_CUDA(cudaStreamWaitEvent(G.stream, G.ev_join, 0));  // join
  const int n_ = G.newtime, c = G.curtime, o = G.oldtime;
  // step_mod.F90:467-560
  POP_TRY(halo_update(fld_t("UBTROP", n_), 1, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));
  POP_TRY(halo_update(fld_t("VBTROP", n_), 1, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));
  POP_TRY(halo_update(fld_t("UVEL", n_), km, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));
  POP_TRY(halo_update(fld_t("VVEL", n_), km, POP_LOC_NECORNER, POP_KIND_VECTOR, 0.0));
  POP_TRY(halo_update(fld_t("RHO", n_), km, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
  POP_TRY(halo_update(fld_t("TRACER", n_), km * G.nt, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0));
  if (!fused) {  // step_mod.F90:581-592
    dim3 grid(ew_grid(G.n2), (unsigned)km, 1);
    POP_LAUNCH(add_barotropic_kernel, grid, POP_EW_THREADS, 0, fld_t("UVEL", n_), fld_t("VVEL", n_),
               fld_t("UBTROP", n_), fld_t("VBTROP", n_), fldi("KMU"), G.n2, (size_t)0, G.n2);
  } else if (tripole_seam_row() >= 0) {  // the seam row, symmetrised by the halo updates above, gets its add now
    dim3 grid(ew_grid((size_t)G.nxb), (unsigned)km, 1);
    POP_LAUNCH(add_barotropic_kernel, grid, POP_EW_THREADS, 0, fld_t("UVEL", n_), fld_t("VVEL", n_),
               fld_t("UBTROP", n_), fld_t("VBTROP", n_), fldi("KMU"), G.n2, (size_t)tripole_seam_row() * G.nxb,
               (size_t)G.nxb);
  }
  POP_LAUNCH(pguess_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("PGUESS"), fld_t("PSURF", n_),
             fld_t("PSURF", c), fld_t("PSURF", o), G.n2);  // step_mod.F90:634-640
  if (ts_type == POP_TS_ROBERT) {
    POP_TRY(step_rf_dev());  // step_mod.F90:798-802
  } else if (G.avg_ts) {
    for (const char* f : {"UBTROP", "VBTROP", "GRADPX", "GRADPY"})
      POP_LAUNCH(avg_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld_t(f, o), fld_t(f, c), fld_t(f, n_), G.n2);
    POP_LAUNCH(avg_fw_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, fld("FW_OLD"), fld("FW"), G.n2);
    for (const char* f : {"UVEL", "VVEL"})
      POP_LAUNCH(avg_kernel, ew_grid(G.n3), POP_EW_THREADS, 0, fld_t(f, o), fld_t(f, c), fld_t(f, n_), G.n3);
    dim3 gt(ew_grid(G.n2), (unsigned)(km * G.nt), 1);
    POP_LAUNCH(avg_tracer_kernel, gt, POP_EW_THREADS, 0, fld_t("TRACER", o), fld_t("TRACER", c),
               fld_t("TRACER", n_), G.n2, km);
    AvgSfc a;
    a.To = fld_t("TRACER", o); a.Tc = fld_t("TRACER", c); a.Tn = fld_t("TRACER", n_);
    a.Po = fld_t("PSURF", o); a.Pc = fld_t("PSURF", c); a.Pn = fld_t("PSURF", n_); a.PG = fld("PGUESS");
    a.n2 = G.n2; a.n3 = G.n3; a.dz1 = G.vc.dz[1]; a.nt = G.nt;
    a.varthick = (G.cfg.sfc_layer_type == POP_SFC_VARTHICK);
    POP_LAUNCH(avg_surface_kernel, ew_grid(G.n2), POP_EW_THREADS, 0, a);
    POP_TRY(state_3d(fld_t("TRACER", o), fld_t("RHO", o)));
    POP_TRY(state_3d(fld_t("TRACER", c), fld_t("RHO", c)));
  } else {
    // FW_OLD = FW; rotate the time levels (step_mod.F90:804-830): an index swap, no data moves
    POP_CHECK_CUDA(cudaMemcpyAsync(fld("FW_OLD"), fld("FW"), sizeof(double) * G.n2, cudaMemcpyDeviceToDevice, G.stream));
    const int tmp = G.oldtime;
    G.oldtime = G.curtime;
    G.curtime = G.newtime;
    G.newtime = tmp;
  }
  POP_TRY(p2p_check());
  return pop_post_launch("step");
}
