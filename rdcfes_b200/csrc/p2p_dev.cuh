// p2p_dev.cuh -- device helpers of the NVLink peer-memory exchange (p2p.cu, solver.cu).
#pragma once
#include <cuda_runtime.h>

#include "rdc_internal.h"

namespace rdc {

// generous: ranks may reach their first exchange seconds apart (caller-side work between rdc_create and the first step)
static constexpr unsigned long long P2P_TIMEOUT_NS = 60000000000ull;
static constexpr int HALO_MAX_BLK = 256;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= seq ; false (and the error mark) after the timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long seq, P2PHeader* hdr) {
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < seq) {
    if (global_ns() - t0 > P2P_TIMEOUT_NS) { hdr->error = 1; return false; }
  }
  return true;
}

// One ghost exchange with the tag-in-word protocol: every value travels as a 16-byte word {bits of the double,
// sequence number} written by ONE vector store straight into the receiver's staging area (NVLink peer memory); the
// receiver polls each word until it carries the current sequence number and moves the value into the ghost tail of
// its vector.  Data and "flag" arrive together: no fences, no flag round trip, no pack buffer.  Staging is
// double-buffered by the parity of the sequence number (a sender can be at most one exchange ahead of a receiver
// because it needs the receiver's data of the previous exchange to finish its own kernel).
struct HaloArgs {
  ulonglong2* dst[RDC_MAX_RANKS];             // neighbour k: where MY values go inside ITS staging area (this parity)
  const ulonglong2* src;                      // my own staging area (this parity)
  int send_ptr[RDC_MAX_RANKS + 1];            // per neighbour: range of send_idx (nodes)
  int recv_ptr[RDC_MAX_RANKS + 1];            // per neighbour: range of my ghost nodes
  int nblk[RDC_MAX_RANKS];                    // blocks that serve neighbour k
  int blk_ptr[RDC_MAX_RANKS + 1];             // prefix of nblk (fused kernels: block -> neighbour)
  int n_nbr;
  int n_owned;
};

// Block b of the nb blocks that serve neighbour k: send my boundary values val(j) (j = local dof), then receive its.
template <class F>
__device__ __forceinline__ void halo_exchange_block(const HaloArgs& A, int k, int b, int nb, int nv,
                                                    const int32_t* __restrict__ send_idx, double* __restrict__ x,
                                                    unsigned long long seq, P2PHeader* hdr, F val) {
  const int s0 = A.send_ptr[k], scnt = (A.send_ptr[k + 1] - s0) * nv;
  ulonglong2* dst = A.dst[k];
  for (int i = b * blockDim.x + threadIdx.x; i < scnt; i += nb * blockDim.x) {
    const int node = i / nv, a = i - node * nv;
    const double v = val((size_t)send_idx[s0 + node] * nv + a);
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + i), "l"((unsigned long long)__double_as_longlong(v)),
                 "l"(seq)
                 : "memory");
  }
  const int r0 = A.recv_ptr[k] * nv, rcnt = (A.recv_ptr[k + 1] - A.recv_ptr[k]) * nv;
  const ulonglong2* src = A.src + r0;
  double* ghost = x + (size_t)A.n_owned * nv + r0;
  for (int i = b * blockDim.x + threadIdx.x; i < rcnt; i += nb * blockDim.x) {
    unsigned long long w0, w1;
    const unsigned long long t0 = global_ns();
    int spins = 0;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src + i) : "memory");
      if (w1 == seq) break;
      if ((++spins & 1023) == 0 && global_ns() - t0 > P2P_TIMEOUT_NS) { hdr->error = 1; break; }
    }
    ghost[i] = __longlong_as_double((long long)w0);
  }
}

}  // namespace rdc
