// asm_common.cuh -- what every assembly kernel shares (assemble.cu: the five RDC models; solid.cu: the solid-mechanics
// Jacobian/residual): the argument block, the index prologue of a CTA and phase 2, the deterministic reduction of the staged
// rows into the row-local block-CSR operator.  See the header of assemble.cu for the work decomposition.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rdc {

struct AsmArgs {
  const int32_t* conn;
  const double* xyz4;
  const double* u_old;
  const double* efield;
  const double* aux0;   // RIPF: TD [n_loc*3]; PROTEAS: AUX [n_loc*2]; solid: undeformed positions [n_loc*3]
  const double* aux1;   // RIPF: RT [n_loc*3]
  const int32_t* n2e_ptr;
  const int32_t* pair;
  const int32_t* rowptr;
  const int32_t* cta_node;
  const int2* task;
  const uint16_t* clist;
  const int32_t* diag_blk;
  double* val;
  double* rhs;
  double* dinv;
};

__host__ __device__ constexpr int popc(unsigned m) { int c = 0; while (m) { c += m & 1u; m >>= 1; } return c; }
__host__ __device__ constexpr int slot_of(unsigned mask, int bitpos) { return popc(mask & ((1u << bitpos) - 1u)); }

// what a CTA learns from its 48-byte descriptor
template <int NEN, int PAIRS>
struct AsmCta {
  int node0, nnode, pair0, npairs, task0, ntask, c_base;
  int2 my_task;
};

template <int NEN, int PAIRS>
__device__ __forceinline__ void asm_prologue(const AsmArgs& A, const int tid, AsmCta<NEN, PAIRS>& cta, int* s_rowptr, int* s_n2e,
                                             int* s_diag, unsigned short* s_clist_raw) {
  // one 48-byte descriptor per CTA: a single load level instead of the chain cta_node -> n2e_ptr -> rowptr -> lists
  const int4 d0 = reinterpret_cast<const int4*>(A.cta_node)[3 * (size_t)blockIdx.x];
  const int4 d1 = reinterpret_cast<const int4*>(A.cta_node)[3 * (size_t)blockIdx.x + 1];
  const int ntask = A.cta_node[12 * (size_t)blockIdx.x + 8];
  const int node0 = d0.x, nnode = d0.y, pair0 = d0.z, npairs = d0.w;
  const int task0 = d1.w;
  // my phase-2 task: fetched now, used after phase 1 (its latency disappears behind the element work)
  int2 my_task = make_int2(0, 0);
  if (tid < ntask) my_task = __ldg(A.task + task0 + tid);
  // Index prologue: everything phase 2 needs goes to shared memory with cp.async (LDGSTS), i.e. without passing
  // through registers -- the warps do not wait for these loads before they start phase 1 (the plain load+store
  // version accounted for 18 % of the kernel's stall samples).  Raw values are stored; the CTA-relative offsets
  // are subtracted where they are used.  The uint16 contributor list is copied as 4-byte words from the
  // aligned-down address.
  const int c_base = d1.z;
  {
    auto cp4 = [](void* dst, const void* src) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    for (int r = tid; r <= nnode; r += PAIRS) { cp4(&s_rowptr[r], A.rowptr + node0 + r); cp4(&s_n2e[r], A.n2e_ptr + node0 + r); }
    for (int r = tid; r < nnode; r += PAIRS) cp4(&s_diag[r], A.diag_blk + node0 + r);
    const int n_w = (npairs * NEN + (c_base & 1) + 1) / 2;
    const unsigned* src_w = reinterpret_cast<const unsigned*>(A.clist + (c_base & ~1));
    for (int i = tid; i < n_w; i += PAIRS) cp4(reinterpret_cast<unsigned*>(s_clist_raw) + i, src_w + i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  cta.node0 = node0; cta.nnode = nnode; cta.pair0 = pair0; cta.npairs = npairs; cta.task0 = task0; cta.ntask = ntask;
  cta.c_base = c_base; cta.my_task = my_task;
}

// phase 2: stageK [NKV][NEN][PAIRS] and stageF [NV][PAIRS] hold the rows the CTA's pair threads produced
template <int NV, unsigned KMASK, int NEN, int PAIRS>
__device__ __forceinline__ void asm_phase2(const AsmArgs& A, const int tid, const AsmCta<NEN, PAIRS>& cta, const double* stageK,
                                           const double* stageF, const int* s_rowptr, const int* s_n2e, const int* s_diag,
                                           const unsigned short* s_clist_raw) {
  constexpr int NKV = popc(KMASK);
  const int node0 = cta.node0, nnode = cta.nnode, pair0 = cta.pair0, task0 = cta.task0, ntask = cta.ntask;
  const int2 my_task = cta.my_task;
  const unsigned short* s_clist = s_clist_raw + (cta.c_base & 1);
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // ------------------------------------------------------------------ phase 2: one task per thread
  // load vector first (one thread per owned dof, pairs of the node in ascending element order)
  for (int t = tid; t < nnode * NV; t += PAIRS) {
    const int r = t / NV, a = t - r * NV;
    const int q0 = s_n2e[r] - pair0, q1 = s_n2e[r + 1] - pair0;
    double f = 0.0;
    for (int q = q0; q < q1; q++) f += stageF[a * PAIRS + q];
    A.rhs[(size_t)(node0 + r) * NV + a] = f;
  }
  // a finished block: plane s of block kk of row r goes to val[rowptr[r]*NKV + s*L + kk] (row-local SoA); the thread
  // that finishes a diagonal block also writes the point-Jacobi scaling (every model has C[a][a] in its mask)
  auto write_block = [&](int r, int kk, const double* acc) {
    const int r0 = s_rowptr[r], L = s_rowptr[r + 1] - r0;
    double* dst = A.val + (size_t)r0 * NKV + kk;
#pragma unroll
    for (int s = 0; s < NKV; s++) dst[(size_t)s * L] = acc[s];
    if (r0 + kk == s_diag[r]) {
#pragma unroll
      for (int a = 0; a < NV; a++) A.dinv[(size_t)(node0 + r) * NV + a] = 1.0 / acc[slot_of(KMASK, a * NV + a)];
    }
  };
  // all lanes of a warp walk the (padded) task list together; the pieces of a split block sit in consecutive lanes of one
  // warp and their partial sums are added by shuffles in piece order -- no second barrier, no partial sums in memory
  const int ntask_w = (ntask + 31) & ~31;
  for (int t = tid; t < ntask_w; t += PAIRS) {
    int2 tk = make_int2(0, 0);
    if (t < ntask) tk = t == tid ? my_task : __ldg(A.task + task0 + t);
    const unsigned x = (unsigned)tk.x, y = (unsigned)tk.y;
    const int c0 = (int)(x >> 16), cnt = (int)(y & 0xffu), piece = (int)((y >> 8) & 0xffu), np = (int)((y >> 16) & 0xffu);
    double acc[NKV > 0 ? NKV : 1];
#pragma unroll
    for (int s = 0; s < NKV; s++) acc[s] = 0.0;
    for (int c = c0; c < c0 + cnt; c++) {
      const double* src = stageK + s_clist[c];
#pragma unroll
      for (int s = 0; s < NKV; s++) acc[s] += src[(size_t)s * NEN * PAIRS];
    }
    const int maxnp = __reduce_max_sync(0xffffffffu, np);
    for (int k = 1; k < maxnp; k++) {
#pragma unroll
      for (int s = 0; s < NKV; s++) {
        const double v = __shfl_down_sync(0xffffffffu, acc[s], k);
        if (piece == 0 && k < np) acc[s] += v;
      }
    }
    if (np >= 1 && piece == 0) write_block((int)(x & 0xffu), (int)((x >> 8) & 0xffu), acc);
  }
}

}  // namespace rdc
