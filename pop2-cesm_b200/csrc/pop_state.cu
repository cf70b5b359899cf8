// pop_state.cu -- equation of state (source/state_mod.F90:258-683): MWJF (:418-500) and linear
// (:664-672), with the 'enforce' / 'ignore' range options (:350-398).  One thread per cell, the
// pressure-dependent coefficient sets are evaluated per level from the __constant__ pressz table
// (uniform across a warp).  24 B/cell (T, S in; rho out), ~45 flop + 1 sqrt + 1 div.
#include "pop_state.cuh"

// slab form of the reference signature: state(k, kk, TEMPK, SALTK, ..., RHOOUT, RHOFULL, DRHODT, DRHODS)
__global__ void state_slab_kernel(StateOpt o, int kk, const double* __restrict__ T,
                                  const double* __restrict__ S, double* RHOOUT, double* RHOFULL,
                                  double* DRHODT, double* DRHODS, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  state_cell(o, kk, T[q], S[q], RHOOUT ? RHOOUT + q : nullptr, RHOFULL ? RHOFULL + q : nullptr,
             DRHODT ? DRHODT + q : nullptr, DRHODS ? DRHODS + q : nullptr);
}

// all levels of a block: RHO(:,:,k) = state(k, k, TRACER(:,:,k,1), TRACER(:,:,k,2))
__global__ void state_3d_kernel(StateOpt o, const double* __restrict__ TR, double* __restrict__ RHO,
                                size_t n2, int km) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y + 1;
  if (q >= n2) return;
  const size_t c = (size_t)(k - 1) * n2 + q;
  double r;
  state_cell(o, k, TR[c], TR[(size_t)km * n2 + c], &r, nullptr, nullptr, nullptr);
  RHO[c] = r;
}

int state_slab(int k, int kk, const double* T, const double* S, double* RHOOUT, double* RHOFULL,
               double* DRHODT, double* DRHODS, size_t n) {
  (void)k;
  POP_REQUIRE(kk >= 1 && kk <= G.km, "state: kk=%d out of range", kk);
  StateOpt o{G.cfg.state_itype, G.cfg.state_range_iopt};
  POP_LAUNCH(state_slab_kernel, ew_grid(n), POP_EW_THREADS, 0, o, kk, T, S, RHOOUT, RHOFULL, DRHODT,
             DRHODS, n);
  return pop_post_launch("state");
}

int state_3d(const double* TRACER, double* RHO) {
  ScopedTimer tm("STATE");
  StateOpt o{G.cfg.state_itype, G.cfg.state_range_iopt};
  dim3 grid(ew_grid(G.n2), (unsigned)G.km, 1);
  POP_LAUNCH(state_3d_kernel, grid, POP_EW_THREADS, 0, o, TRACER, RHO, G.n2, G.km);
  return pop_post_launch("state_3d");
}
