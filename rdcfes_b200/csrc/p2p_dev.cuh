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

// One ghost exchange with the tag-in-word protocol: every value travels as a 16-byte slot of two 8-byte words, each
// {32 data bits | 32-bit tag = low half of the sequence number}, written straight into the receiver's staging area
// (NVLink peer memory); the receiver polls the slot until BOTH words carry the current tag and moves the value into
// the ghost tail of its vector.  PTX guarantees single-copy atomicity only per 8-byte word (a vector store may be
// performed as two scalar stores, on NVLink as on PCIe peer mappings), so every word carries its own tag: a torn or
// reordered slot can never be mistaken for a complete one.  Data and "flag" arrive together: no fences, no flag
// round trip, no pack buffer.  Staging is double-buffered by the parity of the sequence number (a sender can be at
// most one exchange ahead of a receiver because it needs the receiver's data of the previous exchange to finish its
// own kernel); a slot is reused every second exchange, 2^33 exchanges before a tag could repeat in it.
__device__ __forceinline__ void tag_store(ulonglong2* slot, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (b & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (b >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool tag_load(const ulonglong2* slot, unsigned tag, double* v) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) return false;
  *v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return true;
}
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
  const unsigned tag = (unsigned)seq;
  for (int i = b * blockDim.x + threadIdx.x; i < scnt; i += nb * blockDim.x) {
    const int node = i / nv, a = i - node * nv;
    tag_store(dst + i, val((size_t)send_idx[s0 + node] * nv + a), tag);
  }
  const int r0 = A.recv_ptr[k] * nv, rcnt = (A.recv_ptr[k + 1] - A.recv_ptr[k]) * nv;
  const ulonglong2* src = A.src + r0;
  double* ghost = x + (size_t)A.n_owned * nv + r0;
  for (int i = b * blockDim.x + threadIdx.x; i < rcnt; i += nb * blockDim.x) {
    double v = 0.0;
    const unsigned long long t0 = global_ns();
    int spins = 0;
    bool got;
    while (!(got = tag_load(src + i, tag, &v))) {
      if ((++spins & 1023) == 0 && global_ns() - t0 > P2P_TIMEOUT_NS) { hdr->error = 1; break; }
    }
    if (got) ghost[i] = v;   // after a timeout the ghost keeps its old value; the host reports RDC_E_COMM
  }
}

}  // namespace rdc
