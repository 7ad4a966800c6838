// p2p_dev.cuh -- device helpers of the NVLink peer-memory exchange (p2p.cu, solver.cu).
#pragma once
#include <cuda_runtime.h>

#include "rdc_internal.h"

namespace rdc {

static constexpr unsigned long long P2P_TIMEOUT_NS = 4000000000ull;
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

struct HaloArgs {
  double* dst[RDC_MAX_RANKS];                 // neighbour k: ghost segment of MY values inside its copy of the vector
  unsigned long long* flag[RDC_MAX_RANKS];    // neighbour k: halo_flag[my rank] in ITS header
  int nbr_rank[RDC_MAX_RANKS];
  int send_ptr[RDC_MAX_RANKS + 1];
  int nblk[RDC_MAX_RANKS];                    // blocks that serve neighbour k
  int blk_ptr[RDC_MAX_RANKS + 1];             // prefix of nblk (fused kernels: block -> neighbour)
  int n_nbr;
};


// Tail of an exchange, called by every block that stored values for neighbour k (nb blocks do): when the last of
// them is through, one system fence orders all their remote stores (the block barrier makes it cumulative), the
// sequence flag is raised in the neighbour's header and the same thread waits for the neighbour's flag here.
__device__ __forceinline__ void halo_publish_and_wait(const HaloArgs& A, int k, int nb, unsigned* counter, P2PHeader* hdr,
                                                      unsigned long long seq) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(&counter[k], 1u) == (unsigned)nb - 1) {
      counter[k] = 0u;
      st_release_sys(A.flag[k], seq);
      wait_flag(&hdr->halo_flag[A.nbr_rank[k]], seq, hdr);
    }
  }
}

}  // namespace rdc
