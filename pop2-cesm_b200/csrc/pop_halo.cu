// pop_halo.cu -- POP_HaloUpdate (mpi/POP_HaloMod.F90:1732-2071; 3-d :2766, 4-d :4122; 2-d I4) for the
// 1 x P j-strip decomposition: one block per rank spanning all of i.
//
//   * north/south ghost rows come from the neighbouring rank: the two boundary rows of every level
//     are packed into one message per neighbour (k fastest in the reference's messages; here a
//     message is [nz][2][nx_global], contiguous runs of a full row) and exchanged with grouped
//     ncclSend/ncclRecv over NVLink; with one rank and a cyclic north-south boundary it is a local
//     copy.  Ghost rows that face a closed boundary are left untouched, exactly like the reference
//     (no message targets them, mpi/POP_HaloMod.F90:5632-5700).
//   * east/west ghost columns are a GPU-local cyclic wrap (done after the row exchange so that the
//     corner cells receive the diagonal neighbour's values).
//   * the tripole fold (mpi/POP_HaloMod.F90:1947-2048, :5846-5882) is local to the last rank: the top
//     haloWidth+1 physical rows are gathered into bufTripole, symmetrised for NEcorner / Nface
//     fields, and copied out with the sign of the field kind and the (ioffset, joffset) of the
//     field location.
#include "pop_dev.cuh"

#ifndef POP_EMUL
#include <nccl.h>
#endif

// ------------------------------------------------------------------ kernels
// pack rows (jrow0, jrow0+1) of physical columns of every level: buf[z][r][ig]
template <typename T>
__global__ void halo_pack_rows(const T* __restrict__ a, T* __restrict__ buf, int nz, int nxb, size_t n2,
                               int nxg, int jrow0 /*0-based*/) {
  const size_t n = (size_t)nz * 2 * nxg;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (size_t)gridDim.x * blockDim.x) {
    const int ig = (int)(p % nxg), r = (int)((p / nxg) % 2);
    const size_t z = p / ((size_t)2 * nxg);
    buf[p] = a[z * n2 + (size_t)(jrow0 + r) * nxb + (POP_NGHOST + ig)];
  }
}
template <typename T>
__global__ void halo_unpack_rows(T* __restrict__ a, const T* __restrict__ buf, int nz, int nxb,
                                 size_t n2, int nxg, int jrow0) {
  const size_t n = (size_t)nz * 2 * nxg;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (size_t)gridDim.x * blockDim.x) {
    const int ig = (int)(p % nxg), r = (int)((p / nxg) % 2);
    const size_t z = p / ((size_t)2 * nxg);
    a[z * n2 + (size_t)(jrow0 + r) * nxb + (POP_NGHOST + ig)] = buf[p];
  }
}
// single rank, cyclic north-south: ghost rows from the opposite physical rows (physical columns)
template <typename T>
__global__ void halo_ns_cyclic_local(T* __restrict__ a, int nz, int nxb, int nyb, size_t n2, int nxg) {
  const size_t n = (size_t)nz * 4 * nxg;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (size_t)gridDim.x * blockDim.x) {
    const int ig = (int)(p % nxg), r = (int)((p / nxg) % 4);
    const size_t z = p / ((size_t)4 * nxg);
    // r = 0,1: south ghost rows 0,1 <- rows nyb-4, nyb-3 ; r = 2,3: north ghost rows <- rows 2,3
    const int jd = (r < 2) ? r : nyb - 4 + r;
    const int js = (r < 2) ? nyb - 4 + r : r;
    T* az = a + z * n2;
    az[(size_t)jd * nxb + POP_NGHOST + ig] = az[(size_t)js * nxb + POP_NGHOST + ig];
  }
}
// east-west cyclic wrap of every row whose global j index is positive (closed-boundary ghost rows and
// tripole ghost rows are not touched here)
template <typename T>
__global__ void halo_ew_wrap(T* __restrict__ a, int nz, int nxb, int nyb, size_t n2,
                             const int* __restrict__ jglob) {
  const size_t n = (size_t)nz * nyb * 4;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(p % 4), j = (int)((p / 4) % nyb);
    const size_t z = p / ((size_t)4 * nyb);
    if (jglob[j] <= 0) continue;
    T* row = a + z * n2 + (size_t)j * nxb;
    // c = 0,1: west ghost columns 0,1 <- columns nxb-4, nxb-3 ; c = 2,3: east ghost <- columns 2,3
    if (c < 2) row[c] = row[nxb - 4 + c];
    else row[nxb - 4 + c] = row[c];
  }
}

__device__ __forceinline__ double sym_val(double xp, double xq, int isign) {
  const double xavg = 0.5 * (fabs(xp) + fabs(xq));
  return isign * copysign(xavg, xq);
}
__device__ __forceinline__ int sym_val(int xp, int xq, int isign) {
  const int xavg = (int)lround(0.5 * (abs(xp) + abs(xq)));
  return isign * (xq < 0 ? -xavg : xavg);
}
// tripole copy-in: bufTripole(ig, r, z) <- physical rows je-2..je (r = 0..2), with the top row
// symmetrised for NEcorner / Nface fields (mpi/POP_HaloMod.F90:1961-2030)
template <typename T>
__global__ void halo_tripole_in(const T* __restrict__ a, T* __restrict__ buf, int nz, int nxb,
                                size_t n2, int nxg, int je0 /*0-based row of je*/, int loc, int isign) {
  const size_t n = (size_t)nz * 3 * nxg;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (size_t)gridDim.x * blockDim.x) {
    const int ig = (int)(p % nxg) + 1, r = (int)((p / nxg) % 3);  // ig 1-based
    const size_t z = p / ((size_t)3 * nxg);
    const T* row = a + z * n2 + (size_t)(je0 - 2 + r) * nxb + (POP_NGHOST - 1);  // row[ig]
    T v = row[ig];
    if (r == 2) {
      if (loc == POP_LOC_NECORNER) {
        if (ig == nxg) v = (T)(isign * v);
        else v = sym_val(v, row[nxg - ig], isign);
      } else if (loc == POP_LOC_NFACE) {
        const int partner = nxg + 1 - ig;
        if (partner != ig) v = sym_val(v, row[partner], isign);
      }
    }
    buf[p] = v;
  }
}
// tripole copy-out into rows je..je+2, all columns (mpi/POP_HaloMod.F90:2032-2048)
template <typename T>
__global__ void halo_tripole_out(T* __restrict__ a, const T* __restrict__ buf, int nz, int nxb,
                                 size_t n2, int nxg, int je0, int ioff, int joff, int isign,
                                 const int* __restrict__ iglob) {
  const size_t n = (size_t)nz * 3 * nxb;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(p % nxb), j = (int)((p / nxb) % 3) + 1;  // j = 1..3
    const size_t z = p / ((size_t)3 * nxb);
    int is = nxg - iglob[i] + 1 - ioff;
    const int js = POP_NGHOST + 3 - j - joff;
    if (is == 0) is = nxg;
    if (js <= POP_NGHOST + 1 && js >= 1 && is >= 1 && is <= nxg)
      a[z * n2 + (size_t)(je0 + j - 1) * nxb + i] =
          (T)(isign * buf[(z * 3 + (js - 1)) * nxg + (is - 1)]);
  }
}

// Centre-located scalar fields (every solver vector, TRACER, RHO, PSURF): the fold needs neither the
// symmetrisation of the top row nor a sign, so the east-west wrap and the tripole copy only read
// physical columns and only write ghost cells; both run in ONE launch, straight from the source rows
// (ghost row je+1 <- mirrored row je, ghost row je+2 <- mirrored row je-1; same cells and values as
// halo_ew_wrap + halo_tripole_in + halo_tripole_out with ioffset = joffset = 0).
template <typename T>
__global__ void halo_center_scalar_local(T* __restrict__ a, int nz, int nxb, int nyb, size_t n2, int nxg,
                                         int do_ew, int do_tripole, int je0,
                                         const int* __restrict__ iglob, const int* __restrict__ jglob) {
  pdl_wait();
  pdl_trigger();
  const size_t n_ew = do_ew ? (size_t)nz * nyb * 4 : 0;
  const size_t n_tp = do_tripole ? (size_t)nz * 2 * nxb : 0;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_ew + n_tp;
       p += (size_t)gridDim.x * blockDim.x) {
    if (p < n_ew) {
      const int c = (int)(p % 4), j = (int)((p / 4) % nyb);
      const size_t z = p / ((size_t)4 * nyb);
      if (jglob[j] <= 0) continue;
      T* row = a + z * n2 + (size_t)j * nxb;
      if (c < 2) row[c] = row[nxb - 4 + c];
      else row[nxb - 4 + c] = row[c];
    } else {
      const size_t t = p - n_ew;
      const int i = (int)(t % nxb), r = (int)((t / nxb) % 2);  // r = 0: ghost row je+1, r = 1: je+2
      const size_t z = t / ((size_t)2 * nxb);
      int is = nxg - iglob[i] + 1;
      if (is == 0) is = nxg;
      if (is >= 1 && is <= nxg) {
        T* az = a + z * n2;
        az[(size_t)(je0 + 1 + r) * nxb + i] = az[(size_t)(je0 - r) * nxb + (POP_NGHOST - 1 + is)];
      }
    }
  }
}

// ------------------------------------------------------------------ peer-memory strip exchange
// One fused kernel per halo update replaces pack + ncclSend/ncclRecv + unpack: every CTA stores its
// share of the two boundary rows of every level straight into the neighbour's mailbox (peer stores over
// NVLink), the last CTA to finish publishes the sequence number in the neighbours' flag words, and the
// CTAs then wait for their own flags and copy the received rows into the ghost rows.  A push never
// waits, so the wait of any rank depends only on kernels its neighbours have already been able to
// start; the grid is kept within one resident wave so that waiting CTAs cannot starve CTAs that still
// have to push.  Mailbox slots alternate with the parity of the sequence number: slot p is rewritten
// by the neighbour's update n+2, which it can only start after this rank pushed n+1, i.e. after this
// rank finished reading update n.
#ifndef POP_EMUL
// The wait for a neighbour's flag is bounded so that a DEAD peer cannot hang the GPU, but ordinary skew between ranks
// (a first-step module load, a host stall, a cudaMalloc on one rank) must never trip it: the bound is 30 s by default
// (POP_B200_P2P_TIMEOUT_S), i.e. far beyond anything but a lost process.  On a timeout the kernel does NOT pull
// (nothing stale reaches the ghost rows), sets the error word, and every ABI entry reports POP_FAIL (p2p_check).
static long long p2p_spin_cycles() {
  static long long v = 0;
  if (v == 0) {
    double sec = 30.0;
    if (const char* e = getenv("POP_B200_P2P_TIMEOUT_S")) { const double x = atof(e); if (x > 0.0) sec = x; }
    v = (long long)(sec * 2.0e9);  // clock64 ticks at ~2 GHz
  }
  return v;
}
__global__ void __launch_bounds__(POP_EW_THREADS)
halo_p2p_kernel(double* __restrict__ a, int nz, int nxb, size_t n2, int nxg, int rowS, int rowN, int ghostS,
                int ghostN, double* __restrict__ toS, double* __restrict__ toN,
                unsigned long long* flagAtS, unsigned long long* flagAtN, const double* __restrict__ fromS,
                const double* __restrict__ fromN, volatile unsigned long long* myFlags,
                unsigned long long seq, unsigned int* counter, int* err, int fuse_ew, int fuse_tripole, int nyb,
                int je0, const int* __restrict__ iglob, const int* __restrict__ jglob, long long spin_limit, int nr) {
  // nr: rows per side and level (2 for a halo update; the ghost depth of the solver's deep strips otherwise)
  __shared__ int s_ok;
  pdl_wait();
  pdl_trigger();
  const size_t n = (size_t)nz * nr * nxg;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // centre scalars: the east-west wrap of the rows this rank owns and the tripole fold only read physical
  // cells and only write ghost cells, so they ride along before the wait (halo_center_scalar_local)
  if ((fuse_ew & 1) || fuse_tripole) {  // fuse_ew bit 0: wrap of the rows this rank owns; any bit: wrap of the pulled rows
    const size_t n_ew = (fuse_ew & 1) ? (size_t)nz * nyb * 4 : 0;
    const size_t n_tp = fuse_tripole ? (size_t)nz * 2 * nxb : 0;
    for (size_t p = t0; p < n_ew + n_tp; p += stride) {
      if (p < n_ew) {
        const int c = (int)(p % 4), j = (int)((p / 4) % nyb);
        const size_t z = p / ((size_t)4 * nyb);
        if (j < POP_NGHOST || j >= nyb - POP_NGHOST || jglob[j] <= 0) continue;  // ghost rows: see the pull
        double* row = a + z * n2 + (size_t)j * nxb;
        if (c < 2) row[c] = row[nxb - 4 + c];
        else row[nxb - 4 + c] = row[c];
      } else {
        const size_t t = p - n_ew;
        const int i = (int)(t % nxb), r = (int)((t / nxb) % 2);
        const size_t z = t / ((size_t)2 * nxb);
        int is = nxg - iglob[i] + 1;
        if (is == 0) is = nxg;
        if (is >= 1 && is <= nxg) {
          double* az = a + z * n2;
          az[(size_t)(je0 + 1 + r) * nxb + i] = az[(size_t)(je0 - r) * nxb + (POP_NGHOST - 1 + is)];
        }
      }
    }
  }
  // ---- push
  for (size_t p = t0; p < n; p += stride) {
    const int ig = (int)(p % nxg), r = (int)((p / nxg) % nr);
    const size_t z = p / ((size_t)nr * nxg);
    const double* az = a + z * n2 + POP_NGHOST + ig;
    if (toS) toS[p] = az[(size_t)(rowS + r) * nxb];
    if (toN) toN[p] = az[(size_t)(rowN + r) * nxb];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counter, 1u) + 1u;
    if (done == gridDim.x) {  // every CTA of this rank has pushed and fenced
      *counter = 0u;
      __threadfence_system();
      if (flagAtS) *(volatile unsigned long long*)flagAtS = seq;
      if (flagAtN) *(volatile unsigned long long*)flagAtN = seq;
      __threadfence_system();
    }
    // ---- wait for the neighbours' rows (flag 0: from the south, flag 1: from the north)
    const long long t_start = clock64();
    bool ok = true;
    while ((fromS && myFlags[0] < seq) || (fromN && myFlags[1] < seq)) {
      if (clock64() - t_start > spin_limit) { ok = false; break; }
    }
    if (!ok) *err = 1;
    s_ok = ok ? 1 : 0;
    __threadfence_system();
  }
  __syncthreads();
  if (!s_ok) return;  // timed out: leave the ghost rows alone rather than consume a stale mailbox
  // ---- pull
  for (size_t p = t0; p < n; p += stride) {
    const int ig = (int)(p % nxg), r = (int)((p / nxg) % nr);
    const size_t z = p / ((size_t)nr * nxg);
    double* az = a + z * n2 + POP_NGHOST + ig;
    if (fromS) {
      const double v = __ldcv(fromS + p);
      double* row = az + (size_t)(ghostS + r) * nxb;
      row[0] = v;
      if (fuse_ew) {  // east-west ghost columns of the received row
        if (ig < POP_NGHOST) row[nxg] = v;
        else if (ig >= nxg - POP_NGHOST) row[-nxg] = v;
      }
    }
    if (fromN) {
      const double v = __ldcv(fromN + p);
      double* row = az + (size_t)(ghostN + r) * nxb;
      row[0] = v;
      if (fuse_ew) {
        if (ig < POP_NGHOST) row[nxg] = v;
        else if (ig >= nxg - POP_NGHOST) row[-nxg] = v;
      }
    }
  }
}
#endif

// ------------------------------------------------------------------ communicator
#ifndef POP_EMUL
int comm_unique_id(char* id128) {
  ncclUniqueId id;
  ncclResult_t r = ncclGetUniqueId(&id);
  POP_REQUIRE(r == ncclSuccess, "ncclGetUniqueId: %s", ncclGetErrorString(r));
  static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId larger than 128 bytes");
  memset(id128, 0, 128);
  memcpy(id128, &id, sizeof(id));
  return POP_SUCCESS;
}
int comm_init(int rank, int nranks, const char* id128) {
  POP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "pop_comm_init: bad rank %d/%d", rank, nranks);
  if (G.nccl_comm) comm_finalize();
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  ncclResult_t r = ncclCommInitRank(&comm, nranks, id, rank);
  POP_REQUIRE(r == ncclSuccess, "ncclCommInitRank: %s", ncclGetErrorString(r));
  G.nccl_comm = (void*)comm;
  return POP_SUCCESS;
}
int comm_finalize() {
  if (G.nccl_comm) ncclCommDestroy((ncclComm_t)G.nccl_comm);
  G.nccl_comm = nullptr;
  return POP_SUCCESS;
}
int p2p_teardown() {
  if (G.p2p_peerS) cudaIpcCloseMemHandle(G.p2p_peerS);
  if (G.p2p_peerN && G.p2p_peerN != G.p2p_peerS) cudaIpcCloseMemHandle(G.p2p_peerN);
  cudaFree(G.p2p_mbox);
  cudaFree(G.p2p_counter);
  cudaFree(G.p2p_err);
  G.p2p_peerS = G.p2p_peerN = G.p2p_mbox = nullptr;
  G.p2p_counter = nullptr;
  G.p2p_err = nullptr;
  G.p2p_on = false;
  G.p2p_cap = 0;
  G.p2p_seq = 0;
  return POP_SUCCESS;
}
int p2p_setup() {
  p2p_teardown();
  if (G.nranks <= 1 || !G.nccl_comm) return POP_SUCCESS;
  if (getenv("POP_B200_NO_P2P") && getenv("POP_B200_NO_P2P")[0] == '1') return POP_SUCCESS;
  const int ns = G.cfg.ns_boundary_type;
  const int south = (G.rank > 0) ? G.rank - 1 : (ns == POP_BNDY_CYCLIC ? G.nranks - 1 : -1);
  const int north = (G.rank < G.nranks - 1) ? G.rank + 1 : (ns == POP_BNDY_CYCLIC ? 0 : -1);
  const int nzmax = G.km * G.nt > G.km + 2 ? G.km * G.nt : G.km + 2;
  G.p2p_cap = (size_t)nzmax * 2 * G.nxg;
  // the solver's deep strips exchange eight levels of deep_halo rows at a time (pop_barotropic.cu)
  if (G.deep_halo > 0 && (size_t)8 * G.deep_halo * G.nxg > G.p2p_cap) G.p2p_cap = (size_t)8 * G.deep_halo * G.nxg;
  const size_t bytes = (4 * G.p2p_cap + 2) * sizeof(double);
  // every step below is collective; a rank that fails still takes part and the verdict is agreed with a
  // min-reduction so that all ranks choose the same path
  double ok = 1.0;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&G.p2p_mbox, bytes) != cudaSuccess || cudaMemset(G.p2p_mbox, 0, bytes) != cudaSuccess ||
      cudaMalloc(&G.p2p_counter, sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(G.p2p_counter, 0, sizeof(unsigned int)) != cudaSuccess ||
      cudaMalloc(&G.p2p_err, sizeof(int)) != cudaSuccess || cudaMemset(G.p2p_err, 0, sizeof(int)) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine, G.p2p_mbox) != cudaSuccess)
    ok = 0.0;
  cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  // The handles travel through G.d_gather / G.d_local-sized scratch of our own; a local failure anywhere below only
  // clears `ok` -- this rank still enters the all-gather and the min-reduction, so no peer is left waiting in a
  // collective, and the scratch is freed on every path.
  char *d_send = nullptr, *d_all = nullptr;
  std::vector<cudaIpcMemHandle_t> all(G.nranks);
  memset(all.data(), 0, sizeof(cudaIpcMemHandle_t) * all.size());
  bool have = cudaMalloc(&d_send, 64) == cudaSuccess && cudaMalloc(&d_all, 64 * (size_t)G.nranks) == cudaSuccess &&
              cudaMemcpyAsync(d_send, &mine, 64, cudaMemcpyHostToDevice, G.stream) == cudaSuccess;
  if (!have) {
    // the collective still needs valid buffers: fall back to the reduction scratch (allocated by reduce_alloc)
    ok = 0.0;
    cudaGetLastError();
  }
  ncclResult_t r = ncclSuccess;
  if (have) r = ncclAllGather(d_send, d_all, 64, ncclChar, (ncclComm_t)G.nccl_comm, G.stream);
  else r = ncclAllGather(G.d_local, G.d_gather, sizeof(double) * 2 * POP_RED_NF, ncclChar, (ncclComm_t)G.nccl_comm, G.stream);
  if (r != ncclSuccess) ok = 0.0;
  if (have && r == ncclSuccess &&
      (cudaMemcpyAsync(all.data(), d_all, 64 * (size_t)G.nranks, cudaMemcpyDeviceToHost, G.stream) != cudaSuccess ||
       cudaStreamSynchronize(G.stream) != cudaSuccess))
    ok = 0.0;
  cudaGetLastError();
  cudaFree(d_send);
  cudaFree(d_all);
  if (ok == 1.0) {
    void* ps = nullptr;
    void* pn = nullptr;
    if (south >= 0 && cudaIpcOpenMemHandle(&ps, all[south], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0.0;
    if (ok == 1.0 && north >= 0) {
      if (north == south) pn = ps;
      else if (cudaIpcOpenMemHandle(&pn, all[north], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0.0;
    }
    cudaGetLastError();
    G.p2p_peerS = (double*)ps;
    G.p2p_peerN = (double*)pn;
  }
  POP_TRY(comm_allreduce_min(&ok));
  if (ok != 1.0) {
    p2p_teardown();
    return POP_SUCCESS;  // NCCL send/recv path
  }
  G.p2p_on = true;
  return POP_SUCCESS;
}
int p2p_check() {
  if (!G.p2p_on) return POP_SUCCESS;
  int e = 0;
  POP_CHECK_CUDA(cudaMemcpyAsync(&e, G.p2p_err, sizeof(int), cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  if (e != 0) {
    // report once: the word is cleared so that a later call is judged on its own exchange.  The mailbox sequence of the
    // ranks is no longer aligned after a lost exchange; the caller is expected to stop (as the reference does when a
    // message is lost) or to re-initialise the library.
    cudaMemsetAsync(G.p2p_err, 0, sizeof(int), G.stream);
    POP_REQUIRE(false, "POP_HaloUpdate: timed out waiting for a neighbouring rank's strip rows (POP_B200_P2P_TIMEOUT_S)");
  }
  return POP_SUCCESS;
}
int comm_allreduce_min(double* v) {
  if (G.nranks == 1) return POP_SUCCESS;
  double* d = G.d_local;
  POP_CHECK_CUDA(cudaMemcpyAsync(d, v, sizeof(double), cudaMemcpyHostToDevice, G.stream));
  ncclResult_t r = ncclAllReduce(d, d, 1, ncclDouble, ncclMin, (ncclComm_t)G.nccl_comm, G.stream);
  POP_REQUIRE(r == ncclSuccess, "ncclAllReduce(min): %s", ncclGetErrorString(r));
  POP_CHECK_CUDA(cudaMemcpyAsync(v, d, sizeof(double), cudaMemcpyDeviceToHost, G.stream));
  POP_CHECK_CUDA(cudaStreamSynchronize(G.stream));
  return POP_SUCCESS;
}
#else
int comm_unique_id(char* id128) { memset(id128, 0, 128); return POP_SUCCESS; }
int comm_init(int, int, const char*) { return POP_SUCCESS; }
int comm_finalize() { return POP_SUCCESS; }
int comm_allreduce_min(double*) { return POP_SUCCESS; }
int p2p_setup() { return POP_SUCCESS; }
int p2p_teardown() { return POP_SUCCESS; }
int p2p_check() { return POP_SUCCESS; }
#endif

static int ensure_halo_buffers(size_t elems) {  // elems: doubles per message
  if (elems <= G.halo_buf_elems) return POP_SUCCESS;
  cudaFree(G.d_sendS); cudaFree(G.d_sendN); cudaFree(G.d_recvS); cudaFree(G.d_recvN);
  G.d_sendS = G.d_sendN = G.d_recvS = G.d_recvN = nullptr;
  G.halo_buf_elems = 0;
  POP_CHECK_CUDA(cudaMalloc(&G.d_sendS, elems * sizeof(double)));
  POP_CHECK_CUDA(cudaMalloc(&G.d_sendN, elems * sizeof(double)));
  POP_CHECK_CUDA(cudaMalloc(&G.d_recvS, elems * sizeof(double)));
  POP_CHECK_CUDA(cudaMalloc(&G.d_recvN, elems * sizeof(double)));
  G.halo_buf_elems = elems;
  return POP_SUCCESS;
}

// xmode 1: the north-south exchange only (peer-memory path, centre scalars): neighbours' rows + their east-west ghost
// columns + the tripole fold, but NOT the east-west wrap of the rows this rank owns (halo_ew_own_rows does that later)
template <typename T>
static int halo_update_t(T* a, int nz, int loc, int kind, T fill, bool rows_only = false, int xmode = 0) {
  (void)fill;  // no eliminated land blocks in the strip decomposition: nothing receives fillValue
  POP_REQUIRE(G.initialized, "POP_HaloUpdate: library not initialized");
  POP_REQUIRE(a != nullptr && nz >= 1, "POP_HaloUpdate: bad array");
  POP_REQUIRE(loc >= POP_LOC_CENTER && loc <= POP_LOC_EFACE, "POP_HaloUpdate: unknown field location %d", loc);
  POP_REQUIRE(kind >= POP_KIND_SCALAR && kind <= POP_KIND_ANGLE, "POP_HaloUpdate: unknown field kind %d", kind);
  ScopedTimer tm("HALO");
  const int nxb = G.nxb, nyb = G.nyb, nxg = G.nxg;
  const size_t n2 = G.n2;
  const int ns = G.cfg.ns_boundary_type, ew = G.cfg.ew_boundary_type;
  const size_t msg = (size_t)nz * 2 * nxg;
  const unsigned gmsg = ew_grid(msg) < 4096 ? ew_grid(msg) : 4096;
  // ---- north/south
  if (G.nranks > 1) {
#ifndef POP_EMUL
    const int south = (G.rank > 0) ? G.rank - 1 : (ns == POP_BNDY_CYCLIC ? G.nranks - 1 : -1);
    const int north = (G.rank < G.nranks - 1) ? G.rank + 1 : (ns == POP_BNDY_CYCLIC ? 0 : -1);
    if (G.p2p_on && sizeof(T) == sizeof(double) && msg <= G.p2p_cap) {
      const unsigned long long seq = ++G.p2p_seq;
      const size_t par = (size_t)(seq & 1ull);
      double* mb = G.p2p_mbox;
      unsigned long long* myFlags = (unsigned long long*)(mb + 4 * G.p2p_cap);
      // slot(parity, side): side 0 holds rows from the south neighbour, side 1 rows from the north neighbour
      double* toS = south >= 0 ? G.p2p_peerS + (par * 2 + 1) * G.p2p_cap : nullptr;  // I am its north neighbour
      double* toN = north >= 0 ? G.p2p_peerN + (par * 2 + 0) * G.p2p_cap : nullptr;
      unsigned long long* fS = south >= 0 ? (unsigned long long*)(G.p2p_peerS + 4 * G.p2p_cap) + 1 : nullptr;
      unsigned long long* fN = north >= 0 ? (unsigned long long*)(G.p2p_peerN + 4 * G.p2p_cap) + 0 : nullptr;
      const double* fromS = south >= 0 ? mb + (par * 2 + 0) * G.p2p_cap : nullptr;
      const double* fromN = north >= 0 ? mb + (par * 2 + 1) * G.p2p_cap : nullptr;
      unsigned grid = ew_grid(msg);
      const unsigned wave = (unsigned)G.sm_count * 2;  // one resident wave (<= 8 CTAs of 256 threads fit per SM)
      if (grid > wave) grid = wave;
      const bool cs = (loc == POP_LOC_CENTER && kind == POP_KIND_SCALAR);
      const int f_ew = (cs && ew == POP_BNDY_CYCLIC) ? (xmode == 1 ? 2 : 1) : 0;
      const int f_tp = (cs && ns == POP_BNDY_TRIPOLE && G.rank == G.nranks - 1 && !rows_only) ? 1 : 0;
      POP_LAUNCH_PDL(halo_p2p_kernel, grid, POP_EW_THREADS, 0, (double*)a, nz, nxb, n2, nxg, G.jb - 1, G.je - 2, 0, G.je,
                 toS, toN, fS, fN, fromS, fromN, (volatile unsigned long long*)myFlags, seq, G.p2p_counter,
                 G.p2p_err, f_ew, f_tp, nyb, G.je - 1, G.d_iglob, G.d_jglob, p2p_spin_cycles(), 2);
      if (cs) return pop_post_launch("halo_update");  // wrap and fold were fused into the exchange
    } else {
    const size_t msg_d = (msg * sizeof(T) + sizeof(double) - 1) / sizeof(double);
    POP_TRY(ensure_halo_buffers(msg_d));
    T *sS = (T*)G.d_sendS, *sN = (T*)G.d_sendN, *rS = (T*)G.d_recvS, *rN = (T*)G.d_recvN;
    if (south >= 0) POP_LAUNCH(halo_pack_rows<T>, gmsg, POP_EW_THREADS, 0, a, sS, nz, nxb, n2, nxg, G.jb - 1);
    if (north >= 0) POP_LAUNCH(halo_pack_rows<T>, gmsg, POP_EW_THREADS, 0, a, sN, nz, nxb, n2, nxg, G.je - 2);
    ncclComm_t comm = (ncclComm_t)G.nccl_comm;
    const size_t bytes = msg * sizeof(T);
    // Two ranks on a cyclic north-south boundary have the same peer on both sides: NCCL pairs the
    // sends and receives of a peer in issue order, so my message for the peer's SOUTH ghost rows (my
    // north rows) must be my first send, matching the peer's first receive (from its south).
    ncclGroupStart();
    if (north >= 0) ncclSend(sN, bytes, ncclChar, north, comm, G.stream);
    if (south >= 0) ncclSend(sS, bytes, ncclChar, south, comm, G.stream);
    if (south >= 0) ncclRecv(rS, bytes, ncclChar, south, comm, G.stream);
    if (north >= 0) ncclRecv(rN, bytes, ncclChar, north, comm, G.stream);
    ncclResult_t r = ncclGroupEnd();
    POP_REQUIRE(r == ncclSuccess, "halo ncclGroupEnd: %s", ncclGetErrorString(r));
    if (south >= 0) POP_LAUNCH(halo_unpack_rows<T>, gmsg, POP_EW_THREADS, 0, a, rS, nz, nxb, n2, nxg, 0);
    if (north >= 0) POP_LAUNCH(halo_unpack_rows<T>, gmsg, POP_EW_THREADS, 0, a, rN, nz, nxb, n2, nxg, G.je);
    }
#endif
  } else if (ns == POP_BNDY_CYCLIC) {
    const size_t n = (size_t)nz * 4 * nxg;
    POP_LAUNCH(halo_ns_cyclic_local<T>, ew_grid(n) < 4096 ? ew_grid(n) : 4096, POP_EW_THREADS, 0, a, nz, nxb, nyb, n2, nxg);
  }
  const bool tripole_here = (ns == POP_BNDY_TRIPOLE && G.rank == G.nranks - 1) && !rows_only;
  if (loc == POP_LOC_CENTER && kind == POP_KIND_SCALAR) {  // fused east-west wrap + tripole copy
    const int do_ew = (ew == POP_BNDY_CYCLIC) ? 1 : 0, do_tp = tripole_here ? 1 : 0;
    const size_t n = (do_ew ? (size_t)nz * nyb * 4 : 0) + (do_tp ? (size_t)nz * 2 * nxb : 0);
    if (n > 0)
      POP_LAUNCH_PDL(halo_center_scalar_local<T>, ew_grid(n) < 4096 ? ew_grid(n) : 4096, POP_EW_THREADS, 0, a, nz, nxb,
                 nyb, n2, nxg, do_ew, do_tp, G.je - 1, G.d_iglob, G.d_jglob);
    return pop_post_launch("halo_update");
  }
  // ---- east/west
  if (ew == POP_BNDY_CYCLIC) {
    const size_t n = (size_t)nz * nyb * 4;
    POP_LAUNCH(halo_ew_wrap<T>, ew_grid(n) < 4096 ? ew_grid(n) : 4096, POP_EW_THREADS, 0, a, nz, nxb, nyb, n2, G.d_jglob);
  }
  // ---- tripole fold on the last rank
  if (tripole_here) {
    const size_t need = ((size_t)nz * 3 * nxg * sizeof(T) + sizeof(double) - 1) / sizeof(double);
    if (need > G.tripole_elems) {
      cudaFree(G.d_tripole);
      G.d_tripole = nullptr;
      G.tripole_elems = 0;
      POP_CHECK_CUDA(cudaMalloc(&G.d_tripole, need * sizeof(double)));
      G.tripole_elems = need;
    }
    T* buf = (T*)G.d_tripole;
    const int isign = (kind == POP_KIND_SCALAR) ? 1 : -1;
    int ioff = 0, joff = 0;
    if (loc == POP_LOC_NECORNER) { ioff = 1; joff = 1; }
    else if (loc == POP_LOC_EFACE) { ioff = 1; joff = 0; }
    else if (loc == POP_LOC_NFACE) { ioff = 0; joff = 1; }
    const size_t nin = (size_t)nz * 3 * nxg, nout = (size_t)nz * 3 * nxb;
    POP_LAUNCH(halo_tripole_in<T>, ew_grid(nin) < 4096 ? ew_grid(nin) : 4096, POP_EW_THREADS, 0, a, buf, nz, nxb, n2, nxg, G.je - 1, loc, isign);
    POP_LAUNCH(halo_tripole_out<T>, ew_grid(nout) < 4096 ? ew_grid(nout) : 4096, POP_EW_THREADS, 0, a, buf, nz, nxb, n2, nxg, G.je - 1, ioff, joff, isign, G.d_iglob);
  }
  return pop_post_launch("halo_update");
}

int halo_update(double* a, int nz, int loc, int kind, double fill) {
  return halo_update_t<double>(a, nz, loc, kind, fill);
}
// strip rows from the neighbouring ranks + east-west wrap only (no fold): makes k-invariant coefficient
// arrays that were derived locally exact copies of the owner's values two ghost rows deep
int halo_rows_only(double* a, int nz) {
  return halo_update_t<double>(a, nz, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0, true);
}
int halo_exchange_rows(double* a, int nz) {
  POP_REQUIRE(G.nranks > 1 && G.p2p_on, "halo_exchange_rows: needs the peer-memory path");
  return halo_update_t<double>(a, nz, POP_LOC_CENTER, POP_KIND_SCALAR, 0.0, false, 1);
}
// Strip exchange of the solver's deep strips (pop_barotropic.cu): `a` holds nz levels of nyd = ny_local + 2*gd rows
// (level stride n2d); the gd ghost rows on either side receive the neighbour's gd outermost physical rows, east-west
// ghost columns included.  Same mailboxes, flags and sequence numbers as the two-row exchange.
int halo_exchange_deep(double* a, int nz, int gd, size_t n2d) {
  if (G.nranks <= 1) return POP_SUCCESS;  // no neighbour strips (the single-rank test mode of the deep layout)
#ifndef POP_EMUL
  const int ns = G.cfg.ns_boundary_type;
  POP_REQUIRE(G.p2p_on && ns != POP_BNDY_CYCLIC, "halo_exchange_deep: needs the peer-memory path");
  const size_t msg = (size_t)nz * gd * G.nxg;
  POP_REQUIRE(msg <= G.p2p_cap, "halo_exchange_deep: message exceeds the mailbox");
  ScopedTimer tm("HALO");
  const int south = (G.rank > 0) ? G.rank - 1 : -1, north = (G.rank < G.nranks - 1) ? G.rank + 1 : -1;
  const unsigned long long seq = ++G.p2p_seq;
  const size_t par = (size_t)(seq & 1ull);
  double* mb = G.p2p_mbox;
  unsigned long long* myFlags = (unsigned long long*)(mb + 4 * G.p2p_cap);
  double* toS = south >= 0 ? G.p2p_peerS + (par * 2 + 1) * G.p2p_cap : nullptr;
  double* toN = north >= 0 ? G.p2p_peerN + (par * 2 + 0) * G.p2p_cap : nullptr;
  unsigned long long* fS = south >= 0 ? (unsigned long long*)(G.p2p_peerS + 4 * G.p2p_cap) + 1 : nullptr;
  unsigned long long* fN = north >= 0 ? (unsigned long long*)(G.p2p_peerN + 4 * G.p2p_cap) + 0 : nullptr;
  const double* fromS = south >= 0 ? mb + (par * 2 + 0) * G.p2p_cap : nullptr;
  const double* fromN = north >= 0 ? mb + (par * 2 + 1) * G.p2p_cap : nullptr;
  unsigned grid = ew_grid(msg);
  const unsigned wave = (unsigned)G.sm_count * 2;
  if (grid > wave) grid = wave;
  const int f_ew = (G.cfg.ew_boundary_type == POP_BNDY_CYCLIC) ? 2 : 0;  // wrap of the pulled rows only
  const int nyd = G.ny_local + 2 * gd;
  POP_LAUNCH_PDL(halo_p2p_kernel, grid, POP_EW_THREADS, 0, a, nz, G.nxb, n2d, G.nxg, gd, G.ny_local, 0, gd + G.ny_local,
                 toS, toN, fS, fN, fromS, fromN, (volatile unsigned long long*)myFlags, seq, G.p2p_counter, G.p2p_err,
                 f_ew, 0, nyd, gd + G.ny_local - 1, G.d_iglob, G.d_jglob, p2p_spin_cycles(), gd);
  return pop_post_launch("halo_exchange_deep");
#else
  (void)a; (void)nz; (void)gd; (void)n2d;
  POP_REQUIRE(false, "halo_exchange_deep: more than one rank in the emulator build");
#endif
}
int halo_ew_own_rows(double* a, int nz) {
  const size_t n = (size_t)nz * G.nyb * 4;
  POP_LAUNCH(halo_center_scalar_local<double>, ew_grid(n) < 4096 ? ew_grid(n) : 4096, POP_EW_THREADS, 0, a, nz, G.nxb,
             G.nyb, G.n2, G.nxg, 1, 0, G.je - 1, G.d_iglob, G.d_jglob);
  return pop_post_launch("halo_ew_own_rows");
}
int halo_update_i4(int* a, int nz, int loc, int kind, int fill) {
  return halo_update_t<int>(a, nz, loc, kind, fill);
}
int halo_update_r4(float* a, int nz, int loc, int kind, float fill) {
  return halo_update_t<float>(a, nz, loc, kind, fill);
}
