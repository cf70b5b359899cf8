// pop_reduce.cu -- POP_GlobalSum2DR8 / POP_GlobalSumNfields2DR8 (mpi/POP_ReductionsMod.F90:144-389,
// :823) on the device: masked sums over the physical domain of this rank's strip, then over ranks.
//
// Determinism (SURVEY section 7 "bit-exact iteration counts"): every reduction uses the same fixed
// grid (G.red_blocks CTAs of POP_EW_THREADS threads), warp-shuffle + block trees in a fixed order,
// a single-CTA combine of the block partials in block order, and a combine over ranks in rank
// order (all-gather of the per-rank sums, no NCCL reduction whose order could change).  Partials
// are carried in double-double, so the result equals the exactly rounded sum to ~1e-32 relative and
// is independent of the decomposition; the reference's b4b mode (:373-383) has the same purpose.
#include "pop_dev.cuh"

#ifndef POP_EMUL
#include <nccl.h>
#endif

int reduce_alloc() {
  G.red_blocks = 8 * G.sm_count;  // fixed grid of every reduction: 8 CTAs of 256 threads per SM (full occupancy)
  if (G.red_blocks > 2048) G.red_blocks = 2048;
  cudaFree(G.d_partials);
  cudaFree(G.d_sums);
  cudaFree(G.d_local);
  cudaFree(G.d_gather);
  cudaFree(G.d_scal);
  if (G.h_sums) cudaFreeHost(G.h_sums);
  POP_CHECK_CUDA(cudaMalloc(&G.d_partials, sizeof(double) * 2 * POP_RED_NF * G.red_blocks));
  POP_CHECK_CUDA(cudaMalloc(&G.d_sums, sizeof(double) * POP_RED_NF));
  POP_CHECK_CUDA(cudaMalloc(&G.d_local, sizeof(double) * 2 * POP_RED_NF));
  POP_CHECK_CUDA(cudaMalloc(&G.d_gather, sizeof(double) * 2 * POP_RED_NF * G.nranks));
  POP_CHECK_CUDA(cudaMalloc(&G.d_scal, sizeof(SolverScalars)));
  POP_CHECK_CUDA(cudaMemset(G.d_scal, 0, sizeof(SolverScalars)));
  POP_CHECK_CUDA(cudaMallocHost(&G.h_sums, sizeof(double) * POP_RED_NF));
  return POP_SUCCESS;
}

// stage 1: per-CTA dd partial sums of a(:,:,f) * mask over the physical domain
__global__ void sum_fields_kernel(const double* __restrict__ a, int nf, size_t stride,
                                  const double* __restrict__ mask, int nxb, int ib, int jb, int nxp,
                                  int nyp, int dedup_top, int nxg_half, const int* __restrict__ iglob,
                                  double* __restrict__ partials, int red_blocks) {
  dd acc[POP_RED_NF];
  for (int f = 0; f < POP_RED_NF; f++) acc[f] = dd{0.0, 0.0};
  const size_t np = (size_t)nxp * nyp;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < np;
       p += (size_t)gridDim.x * blockDim.x) {
    const int jj = (int)(p / nxp), ii = (int)(p % nxp);
    const int i = ib - 1 + ii, j = jb - 1 + jj;  // 0-based array indices
    // tripole: the redundant half of the top row of NFace/NECorner fields does not count
    // (mpi/POP_ReductionsMod.F90:312-341)
    if (dedup_top && jj == nyp - 1 && iglob[i] > nxg_half) continue;
    const size_t q = (size_t)j * nxb + i;
    const double m = mask ? mask[q] : 1.0;
    for (int f = 0; f < nf; f++) acc[f] = dd_add_d(acc[f], a[f * stride + q] * m);
  }
  for (int f = 0; f < nf; f++) {
    dd r = block_reduce_dd(acc[f]);
    if (threadIdx.x == 0) {
      partials[((size_t)f * red_blocks + blockIdx.x) * 2 + 0] = r.hi;
      partials[((size_t)f * red_blocks + blockIdx.x) * 2 + 1] = r.lo;
    }
  }
}

// stage 2: one CTA combines the block partials of each field in block order
__global__ void sum_blocks_kernel(const double* __restrict__ partials, int nf, int red_blocks,
                                  double* __restrict__ local) {
  for (int f = 0; f < nf; f++) {
    dd acc{0.0, 0.0};
    for (int b = threadIdx.x; b < red_blocks; b += blockDim.x)
      acc = dd_add(acc, dd{partials[((size_t)f * red_blocks + b) * 2], partials[((size_t)f * red_blocks + b) * 2 + 1]});
    dd r = block_reduce_dd(acc);
    if (threadIdx.x == 0) {
      local[f * 2] = r.hi;
      local[f * 2 + 1] = r.lo;
    }
  }
}

// stage 3: combine ranks in rank order, round to double, update the device-resident solver scalars
__global__ void sum_ranks_kernel(const double* __restrict__ gathered, int nranks, int nf, int postop,
                                 double* __restrict__ sums, SolverScalars* __restrict__ sc) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[POP_RED_NF];
  for (int f = 0; f < nf; f++) {
    dd acc{0.0, 0.0};
    for (int r = 0; r < nranks; r++)
      acc = dd_add(acc, dd{gathered[((size_t)r * POP_RED_NF + f) * 2], gathered[((size_t)r * POP_RED_NF + f) * 2 + 1]});
    s[f] = acc.hi + acc.lo;
    sums[f] = s[f];
  }
  switch (postop) {
    case RED_POST_CG_INIT:  // POP_SolversMod.F90:2090-2100
      sc->cgRhoOld = s[0];
      sc->cgSigma = s[1];
      sc->cgAlpha = sc->cgRhoOld / sc->cgSigma;
      break;
    case RED_POST_CG_ITER: {  // POP_SolversMod.F90:2166-2178
      const double cgRho = s[0], cgDelta = s[1];
      sc->cgBeta = cgRho / sc->cgRhoOld;
      sc->cgSigma = cgDelta - (sc->cgBeta * sc->cgBeta) * sc->cgSigma;
      sc->cgAlpha = cgRho / sc->cgSigma;
      sc->cgRhoOld = cgRho;
      break;
    }
    case RED_POST_PCG_ETA1:  // POP_SolversMod.F90:1357
      sc->eta1 = s[0];
      break;
    case RED_POST_PCG_ETA2:  // POP_SolversMod.F90:1409-1410
      sc->eta0 = sc->eta1;
      sc->eta1 = sc->eta0 / s[0];
      break;
    case RED_POST_RR:
      sc->rr = s[0];
      break;
    default:
      break;
  }
}

int reduce_reserve_partials(int nblocks) {
  if (nblocks <= G.partials_big_n) return POP_SUCCESS;
  cudaFree(G.d_partials_big);
  G.d_partials_big = nullptr;
  G.partials_big_n = 0;
  POP_CHECK_CUDA(cudaMalloc(&G.d_partials_big, sizeof(double) * 2 * POP_RED_NF * (size_t)nblocks));
  G.partials_big_n = nblocks;
  return POP_SUCCESS;
}

int reduce_finish(int nfields, int postop, double* out_host) {
  return reduce_finish_n(nfields, postop, out_host, G.d_partials, G.red_blocks);
}

int reduce_finish_n(int nfields, int postop, double* out_host, const double* partials, int nblocks) {
  POP_REQUIRE(nfields >= 1 && nfields <= POP_RED_NF, "reduce_finish: nfields=%d", nfields);
  POP_LAUNCH(sum_blocks_kernel, 1, POP_EW_THREADS, 0, partials, nfields, nblocks, G.d_local);
  const double* gathered = G.d_local;
#ifndef POP_EMUL
  if (G.nranks > 1) {
    ncclResult_t r = ncclAllGather(G.d_local, G.d_gather, 2 * POP_RED_NF, ncclDouble,
                                   (ncclComm_t)G.nccl_comm, G.stream);
    POP_REQUIRE(r == ncclSuccess, "ncclAllGather failed: %s", ncclGetErrorString(r));
    gathered = G.d_gather;
  }
#endif
  POP_LAUNCH(sum_ranks_kernel, 1, 32, 0, gathered, G.nranks, nfields, postop, G.d_sums,
             (SolverScalars*)G.d_scal);
  POP_TRY(pop_post_launch("reduce_finish"));
  if (out_host) {
    POP_CHECK_CUDA(cudaMemcpyAsync(G.h_sums, G.d_sums, sizeof(double) * nfields,
                                   cudaMemcpyDeviceToHost, G.stream));
    POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
    for (int f = 0; f < nfields; f++) out_host[f] = G.h_sums[f];
  }
  return POP_SUCCESS;
}

int global_sum_dev(const double* a, int nfields, size_t field_stride, int loc, const double* mask,
                   double* out_host) {
  POP_REQUIRE(nfields >= 1 && nfields <= POP_RED_NF, "global_sum: nfields=%d", nfields);
  const int dedup = (G.cfg.ns_boundary_type == POP_BNDY_TRIPOLE && G.rank == G.nranks - 1 &&
                     (loc == POP_LOC_NFACE || loc == POP_LOC_NECORNER))
                        ? 1
                        : 0;
  POP_LAUNCH(sum_fields_kernel, G.red_blocks, POP_EW_THREADS, 0, a, nfields, field_stride, mask,
             G.nxb, G.ib, G.jb, G.ie - G.ib + 1, G.je - G.jb + 1, dedup, G.nxg / 2, G.d_iglob,
             G.d_partials, G.red_blocks);
  return reduce_finish(nfields, RED_POST_NONE, out_host);
}
