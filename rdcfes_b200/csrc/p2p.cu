// p2p.cu -- ghost exchange and small all-reduces over NVLink peer memory (one process per GPU).
//
// NCCL costs 20-30 us per small operation; a BiCGStab iteration on a partitioned 10 M-tet mesh needs two ghost
// exchanges and three 1-2 value all-reduces around two ~50 us SpMVs, so NCCL latency, not bandwidth, bounds the
// strong scaling.  Here every rank exports one device arena (cudaIpc) that holds the vectors whose ghosts are
// exchanged plus a small header of flags; peers map it at set-up.  Then
//   * ghost exchange = part of the kernel that produces the vector (or ONE stand-alone kernel): the owner stores
//     every boundary value as a 16-byte word {value, sequence number} straight into the neighbour's staging area
//     (remote stores over NVLink); the receiver polls the words and moves the values into the ghost tail of its
//     vector -- data and flag travel together, no fence, no flag round trip (p2p_dev.cuh);
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

// stand-alone exchange of vector x; grid (max blocks, n_nbr): blocks of column k serve neighbour k
__global__ void __launch_bounds__(256) k_p2p_halo(const HaloArgs A, int nv, const int32_t* __restrict__ send_idx,
                                                  double* __restrict__ x, P2PHeader* hdr, unsigned long long seq,
                                                  const int* __restrict__ done) {
  if (done && *done) return;
  const int k = blockIdx.y;
  const int nb = A.nblk[k];
  if ((int)blockIdx.x >= nb) return;
  const double* xr = x;
  halo_exchange_block(A, k, (int)blockIdx.x, nb, nv, send_idx, x, seq, hdr, [xr](size_t j) { return xr[j]; });
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
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// per-call description of one ghost exchange (also used by the fused vector+exchange kernels); consumes one
// sequence number
void p2p_fill_halo_args(rdc_ctx* c, HaloArgs* A, int* max_blk, int* total_blk, unsigned long long* seq) {
  P2P* P = c->p2p;
  const int nn = (int)c->S.nbr_rank.size();
  *seq = ++P->halo_seq;
  const int par = (int)(*seq & 1ull);
  *max_blk = 1;
  *total_blk = 0;
  A->n_nbr = nn;
  A->n_owned = c->S.n_owned;
  A->blk_ptr[0] = 0;
  A->src = (const ulonglong2*)(P->arena + P->stage_off[par]);
  for (int k = 0; k < nn; k++) {
    const int q = c->S.nbr_rank[k];
    A->dst[k] = (ulonglong2*)((unsigned char*)P->peer[q] + P->stage_off[par]) + (size_t)P->dst_node_off[k] * c->nv;
    A->send_ptr[k] = c->S.send_ptr[k];
    A->recv_ptr[k] = c->S.recv_ptr[k];
    const int cnt = std::max(c->S.send_ptr[k + 1] - c->S.send_ptr[k], c->S.recv_ptr[k + 1] - c->S.recv_ptr[k]) * c->nv;
    A->nblk[k] = std::max(1, std::min(HALO_MAX_BLK, (cnt + 255) / 256));
    *max_blk = std::max(*max_blk, A->nblk[k]);
    A->blk_ptr[k + 1] = A->blk_ptr[k] + A->nblk[k];
  }
  A->send_ptr[nn] = c->S.send_ptr[nn];
  A->recv_ptr[nn] = c->S.recv_ptr[nn];
  *total_blk = A->blk_ptr[nn];
}

// the same description for a given parity, without consuming a sequence number (persistent solver: the kernel counts)
void p2p_fill_halo_args_parity(rdc_ctx* c, HaloArgs* A, int par) {
  P2P* P = c->p2p;
  const int nn = (int)c->S.nbr_rank.size();
  A->n_nbr = nn;
  A->n_owned = c->S.n_owned;
  A->blk_ptr[0] = 0;
  A->src = (const ulonglong2*)(P->arena + P->stage_off[par]);
  for (int k = 0; k < nn; k++) {
    const int q = c->S.nbr_rank[k];
    A->dst[k] = (ulonglong2*)((unsigned char*)P->peer[q] + P->stage_off[par]) + (size_t)P->dst_node_off[k] * c->nv;
    A->send_ptr[k] = c->S.send_ptr[k];
    A->recv_ptr[k] = c->S.recv_ptr[k];
    const int cnt = std::max(c->S.send_ptr[k + 1] - c->S.send_ptr[k], c->S.recv_ptr[k + 1] - c->S.recv_ptr[k]) * c->nv;
    A->nblk[k] = std::max(1, std::min(HALO_MAX_BLK, (cnt + 255) / 256));
    A->blk_ptr[k + 1] = A->blk_ptr[k] + A->nblk[k];
  }
  A->send_ptr[nn] = c->S.send_ptr[nn];
  A->recv_ptr[nn] = c->S.recv_ptr[nn];
}

int p2p_launch_halo(rdc_ctx* c, double* x, bool check_done) {
  P2P* P = c->p2p;
  const int nn = (int)c->S.nbr_rank.size();
  if (nn == 0) return 0;
  HaloArgs A;
  int max_blk = 1, total = 0;
  unsigned long long seq = 0;
  p2p_fill_halo_args(c, &A, &max_blk, &total, &seq);
  const int* done = (check_done && c->work_state) ? c->work_state : nullptr;
  k_p2p_halo<<<dim3(max_blk, nn), 256, 0, c->stream>>>(A, c->nv, c->d_send_idx, x, (P2PHeader*)P->arena, seq, done);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

int p2p_check_error(rdc_ctx* c) {
  P2P* P = c->p2p;
  if (!P || !P->on) return 0;
  int e = 0;
  RDC_CUDA(cudaMemcpyAsync(&e, &((P2PHeader*)P->arena)->error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  if (e) {   // report once: the flag is cleared so that a later, healthy exchange is not blamed for it
    RDC_CUDA(cudaMemsetAsync(&((P2PHeader*)P->arena)->error, 0, sizeof(int), c->stream));
    c->err = "peer-memory exchange timed out (a rank did not arrive); ghost values are stale";
    return RDC_E_COMM;
  }
  return 0;
}

}  // namespace rdc
