// pop_state.cuh -- device function of the equation of state (source/state_mod.F90:258-683), shared by
// pop_state.cu (state slabs / whole blocks) and pop_tracer.cu (convective vertical-mixing coefficients).
#pragma once
#include "pop_dev.cuh"

namespace {
// McDougall, Wright, Jackett & Feistel (2003) coefficients as tabulated in state_mod.F90:136-165
constexpr double p001 = 0.001;
constexpr double mwjfnp0s0t0 = 9.99843699e+2 * p001, mwjfnp0s0t1 = 7.35212840e+0 * p001,
                 mwjfnp0s0t2 = -5.45928211e-2 * p001, mwjfnp0s0t3 = 3.98476704e-4 * p001,
                 mwjfnp0s1t0 = 2.96938239e+0 * p001, mwjfnp0s1t1 = -7.23268813e-3 * p001,
                 mwjfnp0s2t0 = 2.12382341e-3 * p001, mwjfnp1s0t0 = 1.04004591e-2 * p001,
                 mwjfnp1s0t2 = 1.03970529e-7 * p001, mwjfnp1s1t0 = 5.18761880e-6 * p001,
                 mwjfnp2s0t0 = -3.24041825e-8 * p001, mwjfnp2s0t2 = -1.23869360e-11 * p001;
constexpr double mwjfdp0s0t0 = 1.0e+0, mwjfdp0s0t1 = 7.28606739e-3, mwjfdp0s0t2 = -4.60835542e-5,
                 mwjfdp0s0t3 = 3.68390573e-7, mwjfdp0s0t4 = 1.80809186e-10,
                 mwjfdp0s1t0 = 2.14691708e-3, mwjfdp0s1t1 = -9.27062484e-6,
                 mwjfdp0s1t3 = -1.78343643e-10, mwjfdp0sqt0 = 4.76534122e-6,
                 mwjfdp0sqt2 = 1.63410736e-9, mwjfdp1s0t0 = 5.30848875e-6,
                 mwjfdp2s0t3 = -3.03175128e-16, mwjfdp3s0t1 = -1.27934137e-17;
// linear EOS, state_mod.F90:177-182
constexpr double T_leos_ref = 19.0, S_leos_ref = 0.035, rho_leos_ref = 1.025022, alf = 2.55e-4,
                 bet = 7.64e-1;
}  // namespace

struct StateOpt {
  int itype, range;
};

// host: the pressure-dependent coefficient sets of level kk, p = 10*pressz(kk) (state_mod.F90:420-452)
inline void mwjf_level_coefficients(double pressz, double* n0t0, double* n0t2, double* n1t0, double* d0t0,
                                    double* d0t1, double* d0t3) {
  const double p = 10.0 * pressz;
  *n0t0 = mwjfnp0s0t0 + p * (mwjfnp1s0t0 + p * mwjfnp2s0t0);
  *n0t2 = mwjfnp0s0t2 + p * (mwjfnp1s0t2 + p * mwjfnp2s0t2);
  *n1t0 = mwjfnp0s1t0 + p * mwjfnp1s1t0;
  *d0t0 = mwjfdp0s0t0 + p * mwjfdp1s0t0;
  *d0t1 = mwjfdp0s0t1 + (p * p * p) * mwjfdp3s0t1;
  *d0t3 = mwjfdp0s0t3 + (p * p) * mwjfdp2s0t3;
}

// one cell of `state` at pressure level kk; any output pointer may be null
__device__ __forceinline__ void state_cell(StateOpt o, int kk, double T, double S, double* rho,
                                           double* rhofull, double* drhodt, double* drhods) {
  if (o.itype == POP_STATE_LINEAR) {
    if (rho) *rho = bet * (S - S_leos_ref) - alf * (T - T_leos_ref);
    if (rhofull) *rhofull = rho_leos_ref + bet * (S - S_leos_ref) - alf * (T - T_leos_ref);
    if (drhodt) *drhodt = -alf;
    if (drhods) *drhods = bet;
    return;
  }
  // the six pressure-dependent sets come from per-level tables (mwjf_level_coefficients below, same
  // expressions evaluated once per level on the host)
  const double n0t0 = c_vc.eos_n0t0[kk];
  const double n0t1 = mwjfnp0s0t1;
  const double n0t2 = c_vc.eos_n0t2[kk];
  const double n0t3 = mwjfnp0s0t3;
  const double n1t0 = c_vc.eos_n1t0[kk];
  const double n1t1 = mwjfnp0s1t1;
  const double n2t0 = mwjfnp0s2t0;
  const double d0t0 = c_vc.eos_d0t0[kk];
  const double d0t1 = c_vc.eos_d0t1[kk];
  const double d0t2 = mwjfdp0s0t2;
  const double d0t3 = c_vc.eos_d0t3[kk];
  const double d0t4 = mwjfdp0s0t4;
  const double d1t0 = mwjfdp0s1t0, d1t1 = mwjfdp0s1t1, d1t3 = mwjfdp0s1t3;
  const double dqt0 = mwjfdp0sqt0, dqt2 = mwjfdp0sqt2;
  double TQ, SQ;
  if (o.range == POP_STATE_RANGE_ENFORCE) {  // :394-398
    TQ = fmin(T, c_vc.tmax[kk]);
    TQ = fmax(TQ, c_vc.tmin[kk]);
    SQ = fmin(S, c_vc.smax[kk]);
    SQ = fmax(SQ, c_vc.smin[kk]);
  } else {  // 'ignore' :355-358 -- SQ = max(SALTK,0) discards the preceding min
    TQ = fmin(T, 1000.0);
    TQ = fmax(TQ, -1000.0);
    SQ = fmax(S, 0.0);
  }
  SQ = 1000.0 * SQ;
  const double SQR = sqrt(SQ);
  const double WORK1 = n0t0 + TQ * (n0t1 + TQ * (n0t2 + n0t3 * TQ)) + SQ * (n1t0 + n1t1 * TQ + n2t0 * SQ);
  const double WORK2 = d0t0 + TQ * (d0t1 + TQ * (d0t2 + TQ * (d0t3 + d0t4 * TQ))) +
                       SQ * (d1t0 + TQ * (d1t1 + TQ * TQ * d1t3) + SQR * (dqt0 + TQ * TQ * dqt2));
  const double DENOMK = 1.0 / WORK2;
  if (rho) *rho = WORK1 * DENOMK;
  if (rhofull) *rhofull = WORK1 * DENOMK;
  if (drhodt) {
    const double WORK3 = n0t1 + TQ * (2.0 * n0t2 + 3.0 * n0t3 * TQ) + n1t1 * SQ;
    const double WORK4 = d0t1 + SQ * d1t1 +
                         TQ * (2.0 * (d0t2 + SQ * SQR * dqt2) +
                               TQ * (3.0 * (d0t3 + SQ * d1t3) + TQ * 4.0 * d0t4));
    *drhodt = (WORK3 - WORK1 * DENOMK * WORK4) * DENOMK;
  }
  if (drhods) {
    const double WORK3 = n1t0 + n1t1 * TQ + 2.0 * n2t0 * SQ;
    const double WORK4 = d1t0 + TQ * (d1t1 + TQ * TQ * d1t3) + 1.5 * SQR * (dqt0 + TQ * TQ * dqt2);
    *drhods = (WORK3 - WORK1 * DENOMK * WORK4) * DENOMK * 1000.0;
  }
}

