// p2p.cu -- ghost exchange and small all-reduces over NVLink peer memory (one process per GPU).
//
// NCCL costs 20-30 us per small operation; a BiCGStab iteration on a partitioned 10 M-tet mesh needs two ghost
// exchanges and three 1-2 value all-reduces around two ~50 us SpMVs, so NCCL latency, not bandwidth, bounds the
// strong scaling.  Here every rank exports one device arena (cudaIpc) that holds the vectors whose ghosts are
// exchanged plus a small header of flags; peers map it at set-up.  Then
//   * ghost exchange = ONE kernel: the owner packs its boundary values and stores them straight into the ghost
//     tail of the neighbour's vector (remote stores over NVLink), fences, raises a sequence flag in the
//     neighbour's header, and the last block per neighbour waits for that neighbour's flag in the local header;
//   * all-reduce (<= 8 doubles) = ONE single-block kernel: every rank stores its partials into its slot on every
//     peer, raises a flag, waits for all slots and adds them in rank order (bit-identical on every rank).
// Sequence numbers come from the host and are identical on all ranks because every rank issues the same
// operations; kernels return at once when the solver's device-side "done" flag is set (all ranks decide alike).
// A wait that exceeds P2P_TIMEOUT_NS records an error instead of hanging the GPU.
//
// Replaces the MPI traffic of the reference's KSP (VecScatter ghost update, MPI_Allreduce of VecDot/VecNorm;
// SURVEY.md section 2.2).  NCCL stays as the fallback and for large or rare collectives.
#include <stdio.h>

#include <algorithm>

#include "p2p_dev.cuh"
#include "rdc_internal.h"

namespace rdc {

// grid (BLK, n_nbr): blocks of column k serve neighbour k
__global__ void __launch_bounds__(256) k_p2p_halo(const HaloArgs A, int nv, const int32_t* __restrict__ send_idx,
                                                  const double* __restrict__ x, unsigned* counter, P2PHeader* hdr,
                                                  unsigned long long seq, const int* __restrict__ done) {
  if (done && *done) return;
  const int k = blockIdx.y;
  const int nb = A.nblk[k];
  if ((int)blockIdx.x >= nb) return;
  const int s0 = A.send_ptr[k];
  const int n = (A.send_ptr[k + 1] - s0) * nv;
  double* dst = A.dst[k];
  // one or two entries per thread: the index load, the gather and the remote store of an entry are a dependent
  // chain, so the exchange is latency-bound unless it is spread over many threads
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nb * blockDim.x) {
    const int node = i / nv, a = i - node * nv;
    dst[i] = x[(size_t)send_idx[s0 + node] * nv + a];
  }
  halo_publish_and_wait(A, k, nb, counter, hdr, seq);
}

// one block, >= nranks threads.  in/out: n <= 8 doubles in device memory (reduced in place)
__global__ void k_p2p_allreduce(int n, double* buf, int me, int nranks, P2PHeader* const* peer_hdr, P2PHeader* hdr,
                                unsigned long long seq, const int* __restrict__ done) {
  if (done && *done) return;
  const int par = (int)(seq & 1ull);
  const int q = threadIdx.x;
  if (q < nranks) {
    P2PHeader* ph = peer_hdr[q];
    for (int k = 0; k < n; k++) ph->ar_val[par][me][k] = buf[k];
    st_release_sys(&ph->ar_flag[par][me], seq);   // release orders this thread's own stores above: no extra fence
    wait_flag(&hdr->ar_flag[par][q], seq, hdr);
  }
  __syncthreads();
  if (q < n) {
    double s = 0.0;
    for (int r = 0; r < nranks; r++) s += __ldcg(&hdr->ar_val[par][r][q]);
    buf[q] = s;
  }
}

int p2p_launch_allreduce(rdc_ctx* c, double* d_buf, int n, bool check_done) {
  P2P* P = c->p2p;
  P->ar_seq++;
  const int* done = (check_done && c->work_state) ? c->work_state : nullptr;
  k_p2p_allreduce<<<1, 32, 0, c->stream>>>(n, d_buf, c->S.rank, c->S.nranks, (P2PHeader* const*)P->d_peer, (P2PHeader*)P->arena,
                                           P->ar_seq, done);
  c->st.kernel_launches++;
  P->dirty = false;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// per-call description of one ghost exchange of the arena vector x (also used by the fused vector+exchange kernels)
void p2p_fill_halo_args(rdc_ctx* c, const double* x, HaloArgs* A, int* max_blk, int* total_blk) {
  P2P* P = c->p2p;
  const int nn = (int)c->S.nbr_rank.size();
  const size_t off = (size_t)((const unsigned char*)x - P->arena);
  *max_blk = 1;
  *total_blk = 0;
  A->n_nbr = nn;
  A->blk_ptr[0] = 0;
  for (int k = 0; k < nn; k++) {
    const int q = c->S.nbr_rank[k];
    A->dst[k] = (double*)((unsigned char*)P->peer[q] + off) + (size_t)P->dst_node_off[k] * c->nv;
    A->flag[k] = &((P2PHeader*)P->peer[q])->halo_flag[c->S.rank];
    A->nbr_rank[k] = q;
    A->send_ptr[k] = c->S.send_ptr[k];
    const int cnt = (c->S.send_ptr[k + 1] - c->S.send_ptr[k]) * c->nv;
    A->nblk[k] = std::max(1, std::min(HALO_MAX_BLK, (cnt + 255) / 256));
    *max_blk = std::max(*max_blk, A->nblk[k]);
    A->blk_ptr[k + 1] = A->blk_ptr[k] + A->nblk[k];
  }
  A->send_ptr[nn] = c->S.send_ptr[nn];
  *total_blk = A->blk_ptr[nn];
}

// bookkeeping of an exchange that a fused kernel is about to perform: the guard all-reduce when needed, the
// sequence number.  Returns the sequence number to pass to the kernel.
int p2p_halo_begin(rdc_ctx* c, unsigned long long* seq) {
  P2P* P = c->p2p;
  if (P->dirty) {
    int rc = p2p_launch_allreduce(c, P->d_scratch, 0, false);
    if (rc) return rc;
  }
  *seq = ++P->halo_seq;
  P->dirty = true;
  return 0;
}

int p2p_launch_halo(rdc_ctx* c, double* x, bool check_done) {
  P2P* P = c->p2p;
  const int nn = (int)c->S.nbr_rank.size();
  if (nn == 0) return 0;
  if (P->dirty) {  // no all-reduce since the last exchange: make sure every peer has consumed its ghosts
    int rc = p2p_launch_allreduce(c, P->d_scratch, 0, false);
    if (rc) return rc;
  }
  P->halo_seq++;
  HaloArgs A;
  int max_blk = 1, total = 0;
  p2p_fill_halo_args(c, x, &A, &max_blk, &total);
  const int* done = (check_done && c->work_state) ? c->work_state : nullptr;
  k_p2p_halo<<<dim3(max_blk, nn), 256, 0, c->stream>>>(A, c->nv, c->d_send_idx, x, P->d_counter, (P2PHeader*)P->arena, P->halo_seq, done);
  c->st.kernel_launches++;
  P->dirty = true;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

int p2p_check_error(rdc_ctx* c) {
  P2P* P = c->p2p;
  if (!P || !P->on) return 0;
  int e = 0;
  RDC_CUDA(cudaMemcpyAsync(&e, &((P2PHeader*)P->arena)->error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  if (e) { c->err = "peer-memory exchange timed out (a rank did not arrive)"; return RDC_E_COMM; }
  return 0;
}

}  // namespace rdc
