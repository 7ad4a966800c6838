// solver.cu -- block-CSR SpMV and the device-resident Krylov solvers (GMRES(m), CG, BiCGStab) with
// Jacobi preconditioning; check_solution kernels; halo exchange + all-reduce plumbing.
//
// Replaces PetscLinearSolver::solve -> KSPSolve inside TransientLinearImplicitSystem::solve()
// (adpm.C:74, pihna.C:80, ripf.C:83, proteas.C:78, coupled_hcc.C:114) and the per-model check_solution
// (adpm.C:654-688, pihna.C:760-803, proteas.C:707-750, coupled_hcc.C:695-731, ripf.C:675-775).
// Semantics kept from the libMesh/PETSc defaults: restarted GMRES with
// classical Gram-Schmidt, LEFT preconditioning, convergence on the preconditioned residual
// ||B r|| <= max(rtol ||B b||, 1e-50), initial guess = current solution.
//
// Everything stays on the device: dot products are warp-shuffle + fixed-order block/grid reductions
// (bit-reproducible) fused into the kernels that produce their operands, the Hessenberg/Givens update is a
// one-thread kernel, and every kernel of an iteration returns immediately once the device-side "converged"
// flag is set; the host learns about convergence a few iterations late from pinned memory (BiCGStab) or by
// polling every few iterations (GMRES, CG).  A BiCGStab breakdown falls back to GMRES from the current iterate.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "p2p_dev.cuh"
#include "rdc_internal.h"

namespace rdc {

static constexpr int RED_BLOCKS = 592;   // 4 CTAs per SM on 148 SMs
static constexpr int RED_THREADS = 256;
static constexpr int SPMV_MAX_GRID = 148 * 32;

}  // namespace rdc

struct SolverWork {
  int restart_cap = 0;
  size_t vec_len = 0;        // n_loc * nv
  double* V = nullptr;       // [(restart_cap+1)] basis vectors, each vec_len (ghost space included)
  double* t0 = nullptr;      // work vectors
  double* t1 = nullptr;
  double* t2 = nullptr;
  double* t3 = nullptr;
  double* t4 = nullptr;
  double* hs = nullptr;      // BiCGStab s (exchanged)
  double* hp2 = nullptr;     // second p buffer (exchanged)
  double* partial = nullptr; // per-block partials of the fixed-order reductions
  unsigned* counter = nullptr;
  double* h = nullptr;       // [restart_cap + 2] dots of the current column (+ norm^2)
  double* H = nullptr;       // [(restart_cap+1) * restart_cap] column-major Hessenberg after rotations
  double* cs = nullptr; double* sn = nullptr; double* g = nullptr; double* y = nullptr;
  double* scal = nullptr;    // device scalars: [0] res, [1] target, [2] beta, [3] 1/hnorm, [4..] CG/BiCGStab scalars
  int* state = nullptr;      // [0] done, [1] total its, [2] j in cycle, [3] breakdown
  double* h_scal = nullptr;  // pinned mirrors
  int* h_state = nullptr;
  static constexpr int MAX_EV = 512;
  cudaEvent_t ev[2 * MAX_EV];
  int n_ev_used = 0;
  static constexpr int RING = 8;
  int* h_ring = nullptr;     // pinned copies of the convergence flag
  int4* tiles = nullptr;     // SpMV tiles {row0, nrows, first block, nblocks} (k_spmv_tma)
  int n_tiles = 0;
  int tiles_uniform = 0;                // number of rows when every tile but the last has exactly SPMV_TILE_ROWS rows, else 0
  unsigned* flag = nullptr;             // epoch flag of the persistent solver's grid barriers
  unsigned long long* t_spmv = nullptr; // [8] phase timers of the last persistent solve (see PersistArgs)
  double persist_phase_ms[4] = {0, 0, 0, 0};
  int persist_grid = 0;                 // co-resident CTAs of k_bicgstab_persist on this device (0: not queried yet)
  unsigned long long n_persist = 0;     // persistent solves so far (parity selects the epoch flag)
  unsigned char* rep = nullptr;         // device report of a persistent solve: 8 doubles (scal) | 8 ints (state) | 8 u64 (timers)
  unsigned char* h_rep = nullptr;       // its pinned copy
  bool persist_pending = false;         // a persistent solve has been launched and not yet read back
};

namespace rdc {

// ------------------------------------------------------------------------- reductions (fixed order)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Distributed runs finish the reduction inside the same kernel (ArCtx::nranks > 1): the last block stores its sums
// into a slot on every peer (NVLink peer memory, tag-in-word protocol), waits for the slots of all ranks in its own
// header and adds them in rank order, so every rank gets the bit-identical result without a separate collective.
struct ArCtx {
  P2PHeader* const* peer = nullptr;   // [nranks] headers of all ranks (device array)
  P2PHeader* mine = nullptr;
  int me = 0, nranks = 1;
  unsigned long long seq = 0;
};

__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (b & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (b >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_load(const unsigned long long* slot, unsigned tag, double* v) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  if ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag) return false;
  *v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
  return true;
}

// Block-reduce NVAL values into partial[k * gridDim.x + blockIdx.x]; the last block to finish adds the partials
// of all blocks -- thread t takes blocks t, t+256, ... in order, then a fixed shuffle/shared-memory tree -- and
// writes out[k].  One launch, no float atomics, bit-reproducible for a fixed grid size.
template <int NVAL>
__device__ __forceinline__ void grid_reduce(double (&v)[NVAL], int nval, double* partial, unsigned* counter, double* out,
                                            const ArCtx& ar = ArCtx()) {
  __shared__ double s_red[RED_THREADS / 32][NVAL];
  __shared__ double s_tot[NVAL];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NVAL; k++) {
    const double w = warp_sum(v[k]);
    if (lane == 0) s_red[wid][k] = w;
  }
  __syncthreads();
  if (threadIdx.x < NVAL && threadIdx.x < nval) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < RED_THREADS / 32; w++) s += s_red[w][threadIdx.x];
    partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < NVAL; k++) {
      if (k >= nval) break;
      double s = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += RED_THREADS) s += __ldcg(partial + (size_t)k * gridDim.x + b);
      s = warp_sum(s);
      __syncthreads();
      if (lane == 0) s_red[wid][0] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < RED_THREADS / 32; w++) t += s_red[w][0];
        s_tot[k] = t;
      }
    }
    __syncthreads();
    if (ar.nranks > 1) {
      const int par = (int)(ar.seq & 1ull);
      const unsigned tag = (unsigned)ar.seq;
      // thread (q, k): my k-th sum -> my slot on rank q ; then rank q's k-th sum <- its slot in my header
      const int q = threadIdx.x / NVAL, k = threadIdx.x % NVAL;
      double got = 0.0;
      if (q < ar.nranks && k < nval) {
        ll_store(&ar.peer[q]->ll[par][ar.me][k][0], s_tot[k], tag);
        const unsigned long long t0 = global_ns();
        int spins = 0;
        while (!ll_load(&ar.mine->ll[par][q][k][0], tag, &got)) {
          if ((++spins & 1023) == 0 && global_ns() - t0 > P2P_TIMEOUT_NS) { ar.mine->error = 1; break; }
        }
      }
      __syncthreads();
      __shared__ double s_in[RDC_MAX_RANKS][NVAL];
      if (q < ar.nranks && k < nval) s_in[q][k] = got;
      __syncthreads();
      if (threadIdx.x < NVAL && threadIdx.x < nval) {
        double t = 0.0;
        for (int r = 0; r < ar.nranks; r++) t += s_in[r][threadIdx.x];
        out[threadIdx.x] = t;
      }
    } else if (threadIdx.x < NVAL && threadIdx.x < nval) {
      out[threadIdx.x] = s_tot[threadIdx.x];
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------ SpMV
// Half-warp per block row, lane k owns blocks k, k+16, ... of the row (row-local SoA layout -> every load of a
// half-warp is one contiguous segment).  Only the NKV structurally non-zero entry planes of the model are
// stored and streamed (KMASK).  The operator is read once (read-only, no L1 allocation) while the gathered x
// stays cacheable.  A fixed-size grid walks the rows, so the fused dot products of the Krylov methods reduce
// over a bounded number of per-CTA partials in a fixed order:
//   PLAIN    y = S (A x)                       (S = rowscale or identity)
//   DOT_W    y = S (A x) ; out = { <w, y> }                                   BiCGStab  v = B A p, <r0, v>
//   DOT_SELF y = S (A x) ; out = { <x, y>, <y, y> }                           BiCGStab  t = B A s, <s,t>, <t,t>
//   RESID    y = y2 = S (w - A x) ; out = { <y, y>, <S w, S w> }              initial residual and ||B b||^2
enum { SPMV_PLAIN = 0, SPMV_DOT_W = 1, SPMV_DOT_SELF = 2, SPMV_RESID = 3 };

__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_i32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__host__ __device__ constexpr int popc_c(unsigned m) { int c = 0; while (m) { c += m & 1u; m >>= 1; } return c; }
__host__ __device__ constexpr int slot_c(unsigned mask, int bitpos) { return popc_c(mask & ((1u << bitpos) - 1u)); }

template <int NV, unsigned KMASK, int MODE, int MINB>
__global__ void __launch_bounds__(RED_THREADS, MINB)
k_spmv(int n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const double* __restrict__ val,
       const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ rowscale,
       const double* __restrict__ w, double* __restrict__ y2, double* partial, unsigned* counter, double* out,
       const int* __restrict__ done, const ArCtx ar) {
  if (done && *done) return;
  constexpr int G = 16;
  constexpr int NKV = popc_c(KMASK);
  constexpr int HW = RED_THREADS / G;     // half-warps (rows in flight) per CTA
  const int lane = threadIdx.x & (G - 1), hw = threadIdx.x / G;
  double d[2] = {0.0, 0.0};
  // Rows are dealt to the half-warps of the whole grid round-robin, so at any time the grid streams one moving
  // window of the operator (sequential DRAM pages; a contiguous chunk per CTA measured 12 % slower).
  // The kernel is bound by the latency of the per-row dependency chain, so everything that does not depend on
  // the row's dot products is issued ahead of it (EARLY): the row pointers of the next trip and the epilogue
  // operands (Jacobi scale, the dot partner).  Chain left: operator/column loads -> x gather -> FMA -> store.
  constexpr bool EARLY = MINB != 8;
  const int stride = (int)gridDim.x * HW;
  int row = (int)blockIdx.x * HW + hw;
  int r0n = 0, r1n = 0;
  if (EARLY && row < n_rows) { r0n = rowptr[row]; r1n = rowptr[row + 1]; }
#pragma unroll 1
  for (; row < n_rows; row += stride) {
    int r0, L;
    if (EARLY) {
      r0 = r0n; L = r1n - r0n;
      if (row + stride < n_rows) { r0n = rowptr[row + stride]; r1n = rowptr[row + stride + 1]; }
    } else {
      r0 = rowptr[row];
      L = rowptr[row + 1] - r0;
    }
    const size_t o = (size_t)row * NV + (lane < NV ? lane : 0);
    double sc = 1.0, wv = 0.0;
    if (EARLY && lane < NV) {
      if (rowscale) sc = rowscale[o];
      if (MODE == SPMV_DOT_W || MODE == SPMV_RESID) wv = w[o];
      if (MODE == SPMV_DOT_SELF) wv = x[o];
    }
    double acc[NV];
#pragma unroll
    for (int a = 0; a < NV; a++) acc[a] = 0.0;
    // the 16 lanes of a half-warp always run the same trips, so the half-warp mask is exact
    const unsigned hmask = 0xffffu << (threadIdx.x & 16);
#pragma unroll 1
    for (int k = lane; k < L; k += G) {
      const int c = ld_stream_i32(col + r0 + k);
      const double* v0 = val + (size_t)r0 * NKV + k;
      double a_[NKV];
#pragma unroll
      for (int e = 0; e < NKV; e++) a_[e] = ld_stream(v0 + (size_t)e * L);
      double xv[NV];
#pragma unroll
      for (int b = 0; b < NV; b++) xv[b] = x[(size_t)c * NV + b];
#pragma unroll
      for (int ab = 0; ab < NV * NV; ab++)
        if (KMASK >> ab & 1u) acc[ab / NV] = fma(a_[slot_c(KMASK, ab)], xv[ab % NV], acc[ab / NV]);
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1)
#pragma unroll
      for (int a = 0; a < NV; a++) acc[a] += __shfl_xor_sync(hmask, acc[a], off, G);
    // lanes 0..NV-1 finish one component each (consecutive addresses)
    if (lane < NV) {
      double mine = acc[0];
#pragma unroll
      for (int a = 1; a < NV; a++)
        if (lane == a) mine = acc[a];
      if (!EARLY) {
        if (rowscale) sc = rowscale[o];
        if (MODE == SPMV_DOT_W || MODE == SPMV_RESID) wv = w[o];
        if (MODE == SPMV_DOT_SELF) wv = x[o];
      }
      if (MODE == SPMV_PLAIN) {
        y[o] = mine * sc;
      } else if (MODE == SPMV_DOT_W) {
        const double yv = mine * sc;
        y[o] = yv;
        d[0] = fma(wv, yv, d[0]);
      } else if (MODE == SPMV_DOT_SELF) {
        const double yv = mine * sc;
        y[o] = yv;
        d[0] = fma(wv, yv, d[0]);
        d[1] = fma(yv, yv, d[1]);
      } else {
        const double yv = (wv - mine) * sc;
        y[o] = yv;
        if (y2) y2[o] = yv;
        d[0] = fma(yv, yv, d[0]);
        d[1] = fma(wv * sc, wv * sc, d[1]);
      }
    }
  }
  if (MODE != SPMV_PLAIN) grid_reduce<2>(d, MODE == SPMV_DOT_W ? 1 : 2, partial, counter, out, ar);
}

// ---- TMA-staged variant (the default) ------------------------------------------------------------------------
// The LDG kernel above keeps the streamed operator in registers while it is in flight, and ptxas recycles those
// registers between the FMAs, so only 2-3 of the NKV+1 loads of a thread are outstanding (SASS, ncu: long-scoreboard
// bound at ~52 % DRAM throughput with 100 % occupancy).  Here the stream never touches registers: 16 consecutive
// block rows are one contiguous record of the row-local SoA operator, so ONE elected thread moves a whole tile
// (operator values, column ids, row pointers) with three 1-D bulk copies (cp.async.bulk -> UBLKCP) into a ring of
// shared-memory stages, completion counted by an mbarrier.  STAGES-1 tiles per CTA are always in flight
// (~15 KB per CTA and stage), the half-warps compute from shared memory and only gather x through L1/L2.
// What bounds it now is the number of rows whose x gathers are in flight per SM (threads), not the stream:
// 2 stages x 6 CTAs/SM (96 rows) measured 0.334 ms, 3 stages x 4 CTAs/SM (64 rows) 0.397 ms on the headline case.
// Tried and measured without gain: tile descriptors pipelined STAGES deep in registers (0.332 ms), x gathers issued
// one tile ahead with 3 stages x 4 CTAs/SM (0.355 ms).  A read-only stream over the same array reaches 6.7 TB/s
// (rdc_bench_stream), the plain SpMV 5.2-5.5 TB/s, the SpMV inside BiCGStab (fused dots, cold x) 4.9 TB/s.
// Tiles are cut on the host (<= 16 rows, <= SPMV_CAPB blocks): {row0, nrows, first block, nblocks}.
static constexpr int SPMV_CAPB = 256;
static constexpr int SPMV_TILE_ROWS = 16;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// The same copy with the evict_first L2 priority (createpolicy): the operator is read exactly once per SpMV and is larger
// than L2 at every size that matters, so its lines should be the first to go -- the Krylov vectors (gathered x, the
// streams of the vector phases) then survive from one phase to the next.  Measured on one GPU, whole step, same box:
// 1.35 M tets (the per-rank size of an 8-GPU run) 2.264 -> 2.144 ms, 2.6 M tets 3.898 -> 3.811 ms, 5.3 M and 10.1 M
// tets unchanged (the vectors alone exceed L2 there).  Marking the leading 32-96 MB of the operator evict_last instead
// (to keep it resident across the SpMVs, with and without a persisting-L2 set-aside or an access-policy window) gained
// nothing at any size (tools/pin_sweep.sh, profiles/r2_l2_policy_sweep.log).
__device__ __forceinline__ void bulk_g2s_hint(unsigned dst, const void* src, unsigned bytes, unsigned bar, unsigned long long policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <int NKV>
struct SpmvStage {
  static constexpr int VAL_BYTES = ((SPMV_CAPB * NKV + 2) * 8 + 15) / 16 * 16;
  static constexpr int COL_BYTES = ((SPMV_CAPB + 4) * 4 + 15) / 16 * 16;
  static constexpr int RP_BYTES = ((SPMV_TILE_ROWS + 1 + 4) * 4 + 15) / 16 * 16;
  static constexpr int DESC_OFF = VAL_BYTES + COL_BYTES + RP_BYTES;   // the tile's own descriptor {row0, nrows, b0, nblk}
  static constexpr int BYTES = (DESC_OFF + 16 + 127) / 128 * 128;
};

template <int NV, unsigned KMASK, int MODE, int STAGES>
__global__ void __launch_bounds__(RED_THREADS)
k_spmv_tma(int n_tiles, int uniform16, int l2_hint, const int4* __restrict__ tiles, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
           const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
           const double* __restrict__ rowscale, const double* __restrict__ w, double* __restrict__ y2, double* partial,
           unsigned* counter, double* out, const int* __restrict__ done, const ArCtx ar) {
  if (done && *done) return;
  constexpr int G = 16;
  constexpr int NKV = popc_c(KMASK);
  typedef SpmvStage<NKV> ST;
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_bar[STAGES];
  const int tid = threadIdx.x, lane = tid & (G - 1), hw = tid / G;
  const unsigned hmask = 0xffffu << (tid & 16);
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&s_bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer side (thread 0): four bulk copies per tile -- operator values, column ids, row pointers (sources aligned
  // down to 16 B) and the 16-byte tile descriptor itself, so that the 255 consumer threads never load a descriptor from
  // global memory (that load, re-issued by every warp in every iteration, drew 17-25 % of the kernel's stall samples);
  // the producer keeps its own descriptors two tiles ahead in registers
  const unsigned long long pol_stream = l2_policy_evict_first();
  auto issue = [&](int tile, int stage, const int4 t) {
    const int skip_v = (int)(((long long)t.z * NKV) & 1), skip_c = t.z & 3, skip_r = t.x & 3;
    const unsigned vb = (unsigned)(((skip_v + t.w * NKV) * 8 + 15) & ~15);
    const unsigned cb = (unsigned)(((skip_c + t.w) * 4 + 15) & ~15);
    const unsigned rb = (unsigned)(((skip_r + t.y + 1) * 4 + 15) & ~15);
    unsigned char* base = s_raw + (size_t)stage * ST::BYTES;
    const unsigned bar = smem_u32(&s_bar[stage]);
    mbar_expect_tx(bar, vb + cb + rb + 16u);
    if (l2_hint) {
      bulk_g2s_hint(smem_u32(base), val + ((long long)t.z * NKV - skip_v), vb, bar, pol_stream);
      bulk_g2s_hint(smem_u32(base + ST::VAL_BYTES), col + (t.z - skip_c), cb, bar, pol_stream);
    } else {
      bulk_g2s(smem_u32(base), val + ((long long)t.z * NKV - skip_v), vb, bar);
      bulk_g2s(smem_u32(base + ST::VAL_BYTES), col + (t.z - skip_c), cb, bar);
    }
    bulk_g2s(smem_u32(base + ST::VAL_BYTES + ST::COL_BYTES), rowptr + (t.x - skip_r), rb, bar);
    bulk_g2s(smem_u32(base + ST::DESC_OFF), tiles + tile, 16u, bar);
  };
  const int first = (int)blockIdx.x, tstride = (int)gridDim.x;
  int4 pn = make_int4(0, 0, 0, 0);   // producer: descriptor of the next tile to issue, fetched one iteration ahead
  if (tid == 0) {
    for (int s = 0; s < STAGES - 1; s++)
      if (first + s * tstride < n_tiles) issue(first + s * tstride, s, tiles[first + s * tstride]);
    if (first + (STAGES - 1) * tstride < n_tiles) pn = tiles[first + (STAGES - 1) * tstride];
  }
  double d[2] = {0.0, 0.0};
#pragma unroll 1
  for (int i = 0, tile = first; tile < n_tiles; i++, tile += tstride) {
    const int stage = i % STAGES;
    // refill the stage that was consumed in the previous iteration (every thread has passed its barrier)
    if (tid == 0) {
      const int nt = tile + (STAGES - 1) * tstride;
      if (nt < n_tiles) issue(nt, (i + STAGES - 1) % STAGES, pn);
      if (nt + tstride < n_tiles) pn = tiles[nt + tstride];
    }
    // epilogue operands do not depend on the tile data.  When every tile has exactly SPMV_TILE_ROWS rows (regular
    // meshes; flag from the host) the row is known without the descriptor and they are fetched while the tile is still
    // landing; otherwise right after the descriptor arrived, in flight during the x gather
    double sc = 1.0, wv = 0.0;
    if (uniform16) {   // carries the number of rows when set
      const int row = tile * SPMV_TILE_ROWS + hw;
      if (row < uniform16 && lane < NV) {
        const size_t o = (size_t)row * NV + lane;
        if (rowscale) sc = rowscale[o];
        if (MODE == SPMV_DOT_W || MODE == SPMV_RESID) wv = w[o];
        if (MODE == SPMV_DOT_SELF) wv = x[o];
      }
    }
    mbar_wait(smem_u32(&s_bar[stage]), (unsigned)((i / STAGES) & 1));
    const int4 t = *reinterpret_cast<const int4*>(s_raw + (size_t)stage * ST::BYTES + ST::DESC_OFF);
    const bool live = hw < t.y;
    const int row = t.x + hw;
    const size_t o = (size_t)row * NV + (lane < NV ? lane : 0);
    if (!uniform16 && live && lane < NV) {
      if (rowscale) sc = rowscale[o];
      if (MODE == SPMV_DOT_W || MODE == SPMV_RESID) wv = w[o];
      if (MODE == SPMV_DOT_SELF) wv = x[o];
    }
    if (live) {
      const unsigned char* base = s_raw + (size_t)stage * ST::BYTES;
      const double* s_val = reinterpret_cast<const double*>(base) + (((long long)t.z * NKV) & 1);
      const int* s_col = reinterpret_cast<const int*>(base + ST::VAL_BYTES) + (t.z & 3);
      const int* s_rp = reinterpret_cast<const int*>(base + ST::VAL_BYTES + ST::COL_BYTES) + (t.x & 3);
      const int r0 = s_rp[hw] - t.z;                 // block offset inside the tile
      const int L = s_rp[hw + 1] - s_rp[hw];
      double acc[NV];
#pragma unroll
      for (int a = 0; a < NV; a++) acc[a] = 0.0;
#pragma unroll 1
      for (int k = lane; k < L; k += G) {
        const int c = s_col[r0 + k];
        double xv[NV];
#pragma unroll
        for (int b = 0; b < NV; b++) xv[b] = x[(size_t)c * NV + b];
        const double* v0 = s_val + (size_t)r0 * NKV + k;
#pragma unroll
        for (int ab = 0; ab < NV * NV; ab++)
          if (KMASK >> ab & 1u) acc[ab / NV] = fma(v0[slot_c(KMASK, ab) * L], xv[ab % NV], acc[ab / NV]);
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1)
#pragma unroll
        for (int a = 0; a < NV; a++) acc[a] += __shfl_xor_sync(hmask, acc[a], off, G);
      if (lane < NV) {
        double mine = acc[0];
#pragma unroll
        for (int a = 1; a < NV; a++)
          if (lane == a) mine = acc[a];
        if (MODE == SPMV_PLAIN) {
          y[o] = mine * sc;
        } else if (MODE == SPMV_DOT_W) {
          const double yv = mine * sc;
          y[o] = yv;
          d[0] = fma(wv, yv, d[0]);
        } else if (MODE == SPMV_DOT_SELF) {
          const double yv = mine * sc;
          y[o] = yv;
          d[0] = fma(wv, yv, d[0]);
          d[1] = fma(yv, yv, d[1]);
        } else {
          const double yv = (wv - mine) * sc;
          y[o] = yv;
          if (y2) y2[o] = yv;
          d[0] = fma(yv, yv, d[0]);
          d[1] = fma(wv * sc, wv * sc, d[1]);
        }
      }
    }
    __syncthreads();   // the stage may be refilled from the next iteration on
  }
  if (MODE != SPMV_PLAIN) grid_reduce<2>(d, MODE == SPMV_DOT_W ? 1 : 2, partial, counter, out, ar);
}

static int spmv_grid(int n_rows, int nv, int per_sm) {
  const int want = (n_rows + 15) / 16;
  int cap = 148 * (per_sm > 0 ? per_sm : (nv == 3 ? 4 : 3));
  if (cap > SPMV_MAX_GRID) cap = SPMV_MAX_GRID;   // bounded number of per-CTA partials for the fused dots
  return want < cap ? (want > 0 ? want : 1) : cap;
}

template <int NV, unsigned KMASK>
static void spmv_mode(int minb, int mode, unsigned grid, cudaStream_t st, int n, const int32_t* rowptr, const int32_t* col, const double* val,
                      const double* x, double* y, const double* rowscale, const double* w, double* y2, double* partial,
                      unsigned* counter, double* out, const int* done, const ArCtx& ar) {
#define RDC_SPMV_GO(MODE, MB) k_spmv<NV, KMASK, MODE, MB><<<grid, RED_THREADS, 0, st>>>(n, rowptr, col, val, x, y, rowscale, w, y2, partial, counter, out, done, ar)
  if (NV == 3 && minb == 4) {
    switch (mode) {
      case SPMV_PLAIN: RDC_SPMV_GO(SPMV_PLAIN, (NV == 3 ? 4 : 3)); break;
      case SPMV_DOT_W: RDC_SPMV_GO(SPMV_DOT_W, (NV == 3 ? 4 : 3)); break;
      case SPMV_DOT_SELF: RDC_SPMV_GO(SPMV_DOT_SELF, (NV == 3 ? 4 : 3)); break;
      default: RDC_SPMV_GO(SPMV_RESID, (NV == 3 ? 4 : 3)); break;
    }
    return;
  }
  if (NV == 3 && minb == 5) {
    switch (mode) {
      case SPMV_PLAIN: RDC_SPMV_GO(SPMV_PLAIN, (NV == 3 ? 5 : 3)); break;
      case SPMV_DOT_W: RDC_SPMV_GO(SPMV_DOT_W, (NV == 3 ? 5 : 3)); break;
      case SPMV_DOT_SELF: RDC_SPMV_GO(SPMV_DOT_SELF, (NV == 3 ? 5 : 3)); break;
      default: RDC_SPMV_GO(SPMV_RESID, (NV == 3 ? 5 : 3)); break;
    }
    return;
  }
  if (NV == 3 && minb == 6) {
    switch (mode) {
      case SPMV_PLAIN: RDC_SPMV_GO(SPMV_PLAIN, (NV == 3 ? 6 : 3)); break;
      case SPMV_DOT_W: RDC_SPMV_GO(SPMV_DOT_W, (NV == 3 ? 6 : 3)); break;
      case SPMV_DOT_SELF: RDC_SPMV_GO(SPMV_DOT_SELF, (NV == 3 ? 6 : 3)); break;
      default: RDC_SPMV_GO(SPMV_RESID, (NV == 3 ? 6 : 3)); break;
    }
    return;
  }
  switch (mode) {
    case SPMV_PLAIN: RDC_SPMV_GO(SPMV_PLAIN, (NV == 3 ? 8 : 3)); break;
    case SPMV_DOT_W: RDC_SPMV_GO(SPMV_DOT_W, (NV == 3 ? 8 : 3)); break;
    case SPMV_DOT_SELF: RDC_SPMV_GO(SPMV_DOT_SELF, (NV == 3 ? 8 : 3)); break;
    default: RDC_SPMV_GO(SPMV_RESID, (NV == 3 ? 8 : 3)); break;
  }
#undef RDC_SPMV_GO
}

// the five models' entry masks (models.cuh CMASK|SMASK|TMASK), spelled out here so that solver.cu does not
// depend on the model code; create_impl checks them against assemble.cu's model_kmask()
static constexpr unsigned KM_ADPM = 0x15F, KM_RIPF = 0x1F7, KM_HCC = 0x1DF;
static constexpr unsigned KM_PIHNA = 0x1EFBDEF, KM_PROTEAS = 0x127BDEF;
static constexpr unsigned KM_SOLID = 0x1FF;   // dense 3 x 3 node block of the solid-mechanics Jacobian (solid.cu)

int spmv_masks_ok() {
  return model_kmask(RDC_ADPM) == KM_ADPM && model_kmask(RDC_RIPF) == KM_RIPF && model_kmask(RDC_HCC) == KM_HCC &&
         model_kmask(RDC_PIHNA) == KM_PIHNA && model_kmask(RDC_PROTEAS) == KM_PROTEAS && model_kmask(RDC_SOLID) == KM_SOLID;
}

template <int NV, unsigned KMASK, int STAGES>
static int tma_mode(int mode, unsigned grid, cudaStream_t st, int n_tiles, int uniform16, int l2_hint, const int4* tiles, const int32_t* rowptr, const int32_t* col,
                    const double* val, const double* x, double* y, const double* rowscale, const double* w, double* y2,
                    double* partial, unsigned* counter, double* out, const int* done, const ArCtx& ar) {
  constexpr int SMEM = STAGES * SpmvStage<popc_c(KMASK)>::BYTES;
  static bool attr_done_dev[64] = {};   // kernel attributes are per device: a process may drive several GPUs
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr_done = attr_done_dev[dev & 63];
  if (!attr_done) {
    cudaError_t e = cudaSuccess;
#define RDC_TMA_ATTR(MODE) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_spmv_tma<NV, KMASK, MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)
    RDC_TMA_ATTR(SPMV_PLAIN); RDC_TMA_ATTR(SPMV_DOT_W); RDC_TMA_ATTR(SPMV_DOT_SELF); RDC_TMA_ATTR(SPMV_RESID);
#undef RDC_TMA_ATTR
    if (e != cudaSuccess) return -1;
    attr_done = true;
  }
#define RDC_TMA_GO(MODE) k_spmv_tma<NV, KMASK, MODE, STAGES><<<grid, RED_THREADS, SMEM, st>>>(n_tiles, uniform16, l2_hint, tiles, rowptr, col, val, x, y, rowscale, w, y2, partial, counter, out, done, ar)
  switch (mode) {
    case SPMV_PLAIN: RDC_TMA_GO(SPMV_PLAIN); break;
    case SPMV_DOT_W: RDC_TMA_GO(SPMV_DOT_W); break;
    case SPMV_DOT_SELF: RDC_TMA_GO(SPMV_DOT_SELF); break;
    default: RDC_TMA_GO(SPMV_RESID); break;
  }
#undef RDC_TMA_GO
  return 0;
}

// y = S (A x) [+ fused dots, see k_spmv].  `timed` brackets the launch with an event pair (summed lazily).
// Context for a reduction that is finished across the ranks by the producing kernel itself (peer-memory path).
// *fused tells the caller whether the separate all-reduce is still needed (NCCL path / single rank).
static ArCtx ar_begin(rdc_ctx* c, bool* fused) {
  ArCtx a;
  *fused = false;
  P2P* P = c->p2p;
  if (!c->opt.p2p_fused_ar || !P || !P->on) return a;
  a.peer = (P2PHeader* const*)P->d_peer;
  a.mine = (P2PHeader*)P->arena;
  a.me = c->S.rank;
  a.nranks = c->S.nranks;
  a.seq = ++P->ar_seq;
  *fused = true;
  return a;
}

// `ar` (optional): finish the fused dot products across the ranks inside the kernel (see ar_begin)
static int spmv(rdc_ctx* c, int mode, const double* x, double* y, const double* rowscale, const double* w, double* y2, double* out,
                bool check_done, bool timed, const ArCtx& ar = ArCtx()) {
  const int n = c->S.n_owned;
  SolverWork* W = c->work;
  const int* done = (check_done && W) ? W->state : nullptr;
  const unsigned grid = (unsigned)spmv_grid(n, c->nv, c->opt.spmv_ctas_per_sm);
  timed = timed && W && W->n_ev_used < SolverWork::MAX_EV;
  if (timed) cudaEventRecord(W->ev[2 * W->n_ev_used], c->stream);
  double* partial = W ? W->partial : nullptr;
  unsigned* counter = W ? W->counter : nullptr;
  if (W && W->n_tiles > 0 && c->opt.spmv_tma) {
    const int per_sm_env = c->opt.tma_ctas_per_sm, stages_env = c->opt.tma_stages;
    const int per_sm = per_sm_env > 0 ? per_sm_env : (c->nv == 3 ? 6 : 2);  // measured: 2 stages x 6 CTAs/SM beats 3 x 4 by 16 %
    unsigned tg = (unsigned)(W->n_tiles < 148 * per_sm ? W->n_tiles : 148 * per_sm);
    if (tg > (unsigned)SPMV_MAX_GRID) tg = SPMV_MAX_GRID;   // W->partial holds 2 * SPMV_MAX_GRID per-CTA partials
#define RDC_TMA_MODEL(NVV, KM, STG) tma_mode<NVV, KM, STG>(mode, tg, c->stream, W->n_tiles, W->tiles_uniform, c->opt.l2_evict_first, W->tiles, c->d_rowptr, c->d_col, c->d_val, x, y, rowscale, w, y2, partial, counter, out, done, ar)
    int trc;
    switch (c->model) {
      case RDC_ADPM: trc = stages_env == 3 ? RDC_TMA_MODEL(3, KM_ADPM, 3) : RDC_TMA_MODEL(3, KM_ADPM, 2); break;
      case RDC_RIPF: trc = RDC_TMA_MODEL(3, KM_RIPF, 2); break;
      case RDC_HCC: trc = RDC_TMA_MODEL(3, KM_HCC, 2); break;
      case RDC_PIHNA: trc = RDC_TMA_MODEL(5, KM_PIHNA, 2); break;
      case RDC_SOLID: trc = RDC_TMA_MODEL(3, KM_SOLID, 2); break;
      default: trc = RDC_TMA_MODEL(5, KM_PROTEAS, 2); break;
    }
#undef RDC_TMA_MODEL
    if (trc) { c->err = "cudaFuncSetAttribute(k_spmv_tma) failed"; return RDC_E_CUDA; }
  } else {
#define RDC_SPMV_MODEL(NVV, KM) spmv_mode<NVV, KM>(c->opt.spmv_minb, mode, grid, c->stream, n, c->d_rowptr, c->d_col, c->d_val, x, y, rowscale, w, y2, partial, counter, out, done, ar)
    switch (c->model) {
      case RDC_ADPM: RDC_SPMV_MODEL(3, KM_ADPM); break;
      case RDC_RIPF: RDC_SPMV_MODEL(3, KM_RIPF); break;
      case RDC_HCC: RDC_SPMV_MODEL(3, KM_HCC); break;
      case RDC_PIHNA: RDC_SPMV_MODEL(5, KM_PIHNA); break;
      case RDC_SOLID: RDC_SPMV_MODEL(3, KM_SOLID); break;
      default: RDC_SPMV_MODEL(5, KM_PROTEAS); break;
    }
#undef RDC_SPMV_MODEL
  }
  if (timed) { cudaEventRecord(W->ev[2 * W->n_ev_used + 1], c->stream); W->n_ev_used++; }
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

int launch_spmv(rdc_ctx* c, const double* x, double* y, const double* rowscale, bool check_done) {
  return spmv(c, SPMV_PLAIN, x, y, rowscale, nullptr, nullptr, nullptr, check_done, check_done);
}

// out[k] = <V_k, w> for k < nvec (nvec <= 8*NCH) and, when with_norm, out[nvec] = <w, w>.  One pass over w and
// the basis for the whole Gram-Schmidt column: up to 32 independent loads in flight per thread.
template <int NCH>
__global__ void __launch_bounds__(RED_THREADS) k_multidot(size_t n, int nvec, const double* __restrict__ V, size_t ldv,
                                                          const double* __restrict__ w, int with_norm, double* partial,
                                                          unsigned* counter, double* out, const int* __restrict__ done) {
  if (done && *done) return;
  constexpr int NA = 8 * NCH;
  double acc[NA + 1];
#pragma unroll
  for (int k = 0; k <= NA; k++) acc[k] = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double wi = w[i];
    double v[NA];
#pragma unroll
    for (int k = 0; k < NA; k++) v[k] = (k < nvec) ? V[(size_t)k * ldv + i] : 0.0;
#pragma unroll
    for (int k = 0; k < NA; k++) acc[k] = fma(v[k], wi, acc[k]);
    acc[NA] = fma(wi, wi, acc[NA]);
  }
  if (with_norm) {  // place the norm right after the dots
    const double nn = acc[NA];
#pragma unroll
    for (int k = 0; k < NA; k++)
      if (k == nvec) acc[k] = nn;
  }
  grid_reduce<NA + 1>(acc, nvec + (with_norm ? 1 : 0), partial, counter, out);
}

// w -= sum_k h[k] V_k (k < nvec <= 8*NCH) ; out[0] = ||w||^2
template <int NCH>
__global__ void __launch_bounds__(RED_THREADS) k_gs_update(size_t n, int nvec, const double* __restrict__ V, size_t ldv,
                                                           double* __restrict__ w, const double* __restrict__ h,
                                                           double* partial, unsigned* counter, double* out,
                                                           const int* __restrict__ done) {
  if (done && *done) return;
  constexpr int NA = 8 * NCH;
  double hk[NA];
#pragma unroll
  for (int k = 0; k < NA; k++) hk[k] = (k < nvec) ? h[k] : 0.0;
  double acc[1] = {0.0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double wi = w[i];
    double v[NA];
#pragma unroll
    for (int k = 0; k < NA; k++) v[k] = (k < nvec) ? V[(size_t)k * ldv + i] : 0.0;
#pragma unroll
    for (int k = 0; k < NA; k++) wi = fma(-hk[k], v[k], wi);
    w[i] = wi;
    acc[0] = fma(wi, wi, acc[0]);
  }
  grid_reduce<1>(acc, 1, partial, counter, out);
}

// x[i] = alpha_dev[0] * x[i]
__global__ void k_scale_dev(size_t n, double* __restrict__ x, const double* __restrict__ alpha, const int* __restrict__ done) {
  if (done && *done) return;
  const double a = *alpha;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= a;
}

// x += sum_k y[k] V_k, k < *ny
__global__ void k_update_x(size_t n, const int* __restrict__ ny, const double* __restrict__ V, size_t ldv,
                           const double* __restrict__ y, double* __restrict__ x) {
  const int m = *ny;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double xi = x[i];
    for (int k = 0; k < m; k++) xi = fma(y[k], V[(size_t)k * ldv + i], xi);
    x[i] = xi;
  }
}

// r = scale .* (b - t)   (t = A x)
__global__ void k_residual(size_t n, const double* __restrict__ b, const double* __restrict__ t, const double* __restrict__ scale,
                           double* __restrict__ r) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    r[i] = scale ? (b[i] - t[i]) * scale[i] : (b[i] - t[i]);
}
__global__ void k_mul(size_t n, const double* __restrict__ a, const double* __restrict__ s, double* __restrict__ o) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    o[i] = s ? a[i] * s[i] : a[i];
}

// ---- GMRES small kernels (one thread) ----
// start of a cycle: beta = sqrt(nrm2) ; g = (beta,0,..) ; scal[3] = 1/beta ; convergence test on beta
__global__ void k_gmres_begin(const double* nrm2, double* g, double* scal, int* state, int m) {
  const double beta = sqrt(nrm2[0]);
  for (int k = 0; k <= m; k++) g[k] = 0.0;
  g[0] = beta;
  scal[0] = beta;
  scal[2] = beta;
  scal[3] = beta != 0.0 ? 1.0 / beta : 0.0;
  state[2] = 0;
  if (beta <= scal[1]) state[0] = 1;
}
// end of inner iteration j: h[0..j] dots, hn2 = ||w||^2 -> Hessenberg column, Givens, residual estimate
__global__ void k_gmres_givens(int j, int m, const double* h, const double* hn2, double* H, double* cs, double* sn, double* g,
                               double* scal, int* state) {
  if (state[0]) return;
  const double hn = sqrt(hn2[0]);
  double* Hj = H + (size_t)j * (m + 1);
  for (int k = 0; k <= j; k++) Hj[k] = h[k];
  Hj[j + 1] = hn;
  for (int k = 0; k < j; k++) {
    const double a = Hj[k], b = Hj[k + 1];
    Hj[k] = cs[k] * a + sn[k] * b;
    Hj[k + 1] = -sn[k] * a + cs[k] * b;
  }
  const double a = Hj[j], b = Hj[j + 1];
  const double r = hypot(a, b);
  cs[j] = r == 0.0 ? 1.0 : a / r;
  sn[j] = r == 0.0 ? 0.0 : b / r;
  Hj[j] = r;
  Hj[j + 1] = 0.0;
  g[j + 1] = -sn[j] * g[j];
  g[j] = cs[j] * g[j];
  const double res = fabs(g[j + 1]);
  scal[0] = res;
  scal[3] = hn != 0.0 ? 1.0 / hn : 0.0;
  state[1] += 1;
  state[2] = j + 1;
  if (!(res == res)) { state[0] = 1; state[3] = 1; }           // NaN: breakdown
  else if (res <= scal[1] || hn == 0.0) state[0] = 1;
}
// y = H^-1 g for the first state[2] columns
__global__ void k_gmres_backsolve(int m, const double* H, const double* g, double* y, const int* state) {
  const int j = state[2];
  for (int k = j - 1; k >= 0; k--) {
    double s = g[k];
    for (int l = k + 1; l < j; l++) s -= H[(size_t)l * (m + 1) + k] * y[l];
    y[k] = s / H[(size_t)k * (m + 1) + k];
  }
}
__global__ void k_set_target(const double* nrm2, double rtol, double* scal) {
  const double bn = sqrt(nrm2[0]);
  scal[1] = fmax(rtol * bn, 1e-50);
  scal[5] = bn;
}

// ------------------------------------------------------------------------------ check_solution
// adpm.C:675-677 and siblings: if (u < 0) u = 0
__global__ void k_clamp_nonneg(size_t n, double* __restrict__ u) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = u[i];
    if (v < 0.0) u[i] = 0.0;
  }
}
// ripf.C:709-760 for every local node (ghosts included for RT_total; u/prev/TD on owned nodes)
__global__ void __launch_bounds__(RED_THREADS) k_ripf_check(int n_owned, int n_loc, double* __restrict__ u, double* __restrict__ prev,
                                                            double* __restrict__ td, double* __restrict__ rt, double HU_min,
                                                            double HU_max, double bf, double ff, int day, double DT_R,
                                                            double* partial, unsigned* counter, double* out) {
  double mx[1] = {-1.0};
  const double tf = bf + ff;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_loc; n += gridDim.x * blockDim.x) {
    const double RT_broad = rt[(size_t)n * 3], RT_focus = rt[(size_t)n * 3 + 1];
    double RT_total;
    if (day < bf) RT_total = RT_broad / bf * (day + 1);
    else if (day < tf) RT_total = RT_focus / ff * ((day + 1) - bf) + RT_broad;
    else RT_total = RT_broad + RT_focus;
    rt[(size_t)n * 3 + 2] = RT_total;
    if (n < n_owned) {
      mx[0] = fmax(mx[0], RT_total);
      const double s0 = u[(size_t)n * 3], s1 = u[(size_t)n * 3 + 1], s2 = u[(size_t)n * 3 + 2];
      double HU = s0, cc = s1, fb = s2;
      if (HU < HU_min) HU = HU_min; else if (HU > HU_max) HU = HU_max;
      if (cc < 0.0) cc = 0.0;
      if (fb < 0.0) fb = 0.0;
      td[(size_t)n * 3] = (HU - prev[(size_t)n * 3]) * DT_R;
      td[(size_t)n * 3 + 1] = (cc - prev[(size_t)n * 3 + 1]) * DT_R;
      td[(size_t)n * 3 + 2] = (fb - prev[(size_t)n * 3 + 2]) * DT_R;
      prev[(size_t)n * 3] = s0; prev[(size_t)n * 3 + 1] = s1; prev[(size_t)n * 3 + 2] = s2;
      u[(size_t)n * 3] = HU; u[(size_t)n * 3 + 1] = cc; u[(size_t)n * 3 + 2] = fb;
    }
  }
  // max-reduce (order independent)
  __shared__ double s_mx[RED_THREADS / 32];
  __shared__ bool s_last;
  double m = mx[0];
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RED_THREADS / 32; w++) m = fmax(m, s_mx[w]);
    partial[blockIdx.x] = m;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double g = -1.0;
    for (unsigned b = 0; b < gridDim.x; b++) g = fmax(g, partial[b]);
    out[0] = g;
    *counter = 0u;
  }
}

// ------------------------------------------------------------------- user vector gather / scatter
// Thread t handles dof (node_by_glob[t / nv], t % nv): consecutive threads touch ascending GLOBAL dofs, i.e. the user
// vector (possibly pinned host memory read or written over PCIe) is walked in order whatever the local numbering is.
// n_nodes_used: all local nodes (gather) or a prefix test on owned nodes (scatter writes owned dofs only).
__global__ void k_gather(size_t n, int nv, int n_owned_limit, const int32_t* __restrict__ node_by_glob, const int32_t* __restrict__ map,
                         const double* __restrict__ src, double* __restrict__ dst) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t k = t / nv;
    const int node = node_by_glob[k];
    if (node >= n_owned_limit) continue;
    const size_t i = (size_t)node * nv + (t - k * nv);
    dst[i] = src[map[i]];
  }
}
__global__ void k_scatter(size_t n, int nv, int n_owned_limit, const int32_t* __restrict__ node_by_glob, const int32_t* __restrict__ map,
                          const double* __restrict__ src, double* __restrict__ dst) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t k = t / nv;
    const int node = node_by_glob[k];
    if (node >= n_owned_limit) continue;
    const size_t i = (size_t)node * nv + (t - k * nv);
    dst[map[i]] = src[i];
  }
}
// pack nv values of each listed node
__global__ void k_pack(size_t n_nodes, int nv, const int32_t* __restrict__ idx, const double* __restrict__ x, double* __restrict__ buf) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_nodes * nv; i += (size_t)gridDim.x * blockDim.x) {
    const size_t node = i / nv, a = i - node * nv;
    buf[i] = x[(size_t)idx[node] * nv + a];
  }
}

static inline unsigned grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (unsigned)g;
}

int launch_gather(rdc_ctx* c, const double* src_glob, double* dst_loc) {
  const size_t n = (size_t)c->S.n_loc * c->nv;
  k_gather<<<grid_for(n), 256, 0, c->stream>>>(n, c->nv, c->S.n_loc, c->d_node_by_glob, c->d_dofmap, src_glob, dst_loc);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}
int launch_scatter(rdc_ctx* c, const double* src_loc, double* dst_glob) {
  const size_t n = (size_t)c->S.n_loc * c->nv;   // walks all local nodes in global order, writes the owned ones
  k_scatter<<<grid_for(n), 256, 0, c->stream>>>(n, c->nv, c->S.n_owned, c->d_node_by_glob, c->d_dofmap, src_loc, dst_glob);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack(rdc_ctx* c, const double* x, int ncomp) {
  const size_t n = c->S.send_idx.size();
  if (n == 0) return 0;
  k_pack<<<grid_for(n * ncomp), 256, 0, c->stream>>>(n, ncomp, c->d_send_idx, x, c->d_sendbuf);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// ghosts of the current solution: exchanged once after every change of the owned values, not once per consumer
int refresh_u_ghosts(rdc_ctx* c) {
  if (c->S.nranks == 1 || c->u_ghost_fresh) return 0;
  const int rc = halo_exchange(c, c->d_u);
  if (!rc) c->u_ghost_fresh = true;
  return rc;
}

int launch_clamp(rdc_ctx* c) {
  SolverWork* W = c->work;
  c->u_ghost_fresh = false;
  if (c->model == RDC_RIPF) {
    const double* p = c->params.data();
    const int day = (int)floor(c->time);  // ripf.C:705
    k_ripf_check<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->S.n_owned, c->S.n_loc, c->d_u, c->d_prev, c->d_td, c->d_rt,
                                                           p[RIPF_HU_MIN], p[RIPF_HU_MAX], p[RIPF_RT_BROAD_FRAC],
                                                           p[RIPF_RT_FOCUS_FRAC], day, 1.0 / c->dt, W->partial, W->counter, W->h);
    c->st.kernel_launches++;
    RDC_CUDA(cudaGetLastError());
    int rc = allreduce_max(c, W->h, 1);
    if (rc) return rc;
    RDC_CUDA(cudaMemcpyAsync(W->h_scal, W->h, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RDC_CUDA(cudaStreamSynchronize(c->stream));
    const double mx = W->h_scal[0];
    c->ripf_rt_max = (int)mx;  // ripf.C:772 (int truncation)
    c->st.ripf_rt_total_max = c->ripf_rt_max;
    c->ripf_primed = true;
    if (mx <= 0.0) { c->err = "RT_total_max <= 0 (ripf.C:773)"; return RDC_E_MODEL; }
    // ghosts of TD are needed by the next assembly
    if (c->S.nranks > 1) {
      if ((rc = halo_exchange(c, c->d_td))) return rc;
      if ((rc = p2p_check_error(c))) return rc;   // synchronises; the next assembly reads the TD ghosts
    }
    return 0;
  }
  const size_t n = (size_t)c->S.n_owned * c->nv;
  k_clamp_nonneg<<<grid_for(n), 256, 0, c->stream>>>(n, c->d_u);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------- solver work space
int solver_init(rdc_ctx* c) {
  SolverWork* W = new SolverWork();
  c->work = W;
  W->vec_len = (size_t)c->S.n_loc * c->nv;
  const size_t vb = W->vec_len * sizeof(double);
  RDC_CUDA(cudaMalloc(&W->t0, vb)); RDC_CUDA(cudaMalloc(&W->t1, vb));
  RDC_CUDA(cudaMemsetAsync(W->t0, 0, vb, c->stream)); RDC_CUDA(cudaMemsetAsync(W->t1, 0, vb, c->stream));
  RDC_CUDA(cudaMalloc(&W->partial, sizeof(double) * (RED_BLOCKS * 33 > 2 * SPMV_MAX_GRID ? RED_BLOCKS * 33 : 2 * SPMV_MAX_GRID)));
  RDC_CUDA(cudaMalloc(&W->counter, sizeof(unsigned)));
  RDC_CUDA(cudaMemsetAsync(W->counter, 0, sizeof(unsigned), c->stream));
  RDC_CUDA(cudaMalloc(&W->scal, sizeof(double) * 32));
  RDC_CUDA(cudaMemsetAsync(W->scal, 0, sizeof(double) * 32, c->stream));
  RDC_CUDA(cudaMalloc(&W->state, sizeof(int) * 8));
  c->work_state = W->state;
  RDC_CUDA(cudaMemsetAsync(W->state, 0, sizeof(int) * 8, c->stream));
  RDC_CUDA(cudaMallocHost(&W->h_scal, sizeof(double) * 32));
  RDC_CUDA(cudaMallocHost(&W->h_state, sizeof(int) * 8));
  RDC_CUDA(cudaMalloc(&W->h, sizeof(double) * 1024));
  for (int k = 0; k < 2 * SolverWork::MAX_EV; k++) RDC_CUDA(cudaEventCreate(&W->ev[k]));
  {  // SpMV tiles: consecutive rows, at most SPMV_TILE_ROWS rows and SPMV_CAPB blocks each (setup.cpp)
    std::vector<int32_t> tl;
    static_assert(sizeof(int4) == 4 * sizeof(int32_t), "tile descriptor layout");
    if (cut_spmv_tiles(c->S.rowptr.data(), c->S.n_owned, SPMV_TILE_ROWS, SPMV_CAPB, tl) && !tl.empty()) {
      RDC_CUDA(cudaMalloc(&W->tiles, tl.size() * sizeof(int32_t)));
      RDC_CUDA(cudaMemcpyAsync(W->tiles, tl.data(), tl.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      RDC_CUDA(cudaStreamSynchronize(c->stream));
      W->n_tiles = (int)(tl.size() / 4);
      bool uni = true;
      for (int k = 0; k + 1 < W->n_tiles && uni; k++) uni = tl[(size_t)4 * k + 1] == SPMV_TILE_ROWS && tl[(size_t)4 * k] == k * SPMV_TILE_ROWS;
      W->tiles_uniform = uni ? c->S.n_owned : 0;
    }
  }
  RDC_CUDA(cudaMalloc(&W->flag, 2 * sizeof(unsigned)));
  RDC_CUDA(cudaMemsetAsync(W->flag, 0, 2 * sizeof(unsigned), c->stream));
  RDC_CUDA(cudaMalloc(&W->t_spmv, 8 * sizeof(unsigned long long)));
  RDC_CUDA(cudaMalloc(&W->rep, 192));
  RDC_CUDA(cudaMemsetAsync(W->rep, 0, 192, c->stream));
  RDC_CUDA(cudaMallocHost(&W->h_rep, 192));
  RDC_CUDA(cudaHostAlloc(&W->h_ring, sizeof(int) * SolverWork::RING, cudaHostAllocMapped));
  memset(W->h_ring, 0, sizeof(int) * SolverWork::RING);
  return 0;
}

static int ensure_gmres(rdc_ctx* c, int restart) {
  SolverWork* W = c->work;
  if (restart <= W->restart_cap) return 0;
  if (W->V) { cudaFree(W->V); cudaFree(W->H); cudaFree(W->cs); cudaFree(W->sn); cudaFree(W->g); cudaFree(W->y); W->V = nullptr; }
  const size_t vb = W->vec_len * sizeof(double);
  RDC_CUDA(cudaMalloc(&W->V, vb * (restart + 1)));
  RDC_CUDA(cudaMemsetAsync(W->V, 0, vb * (restart + 1), c->stream));
  RDC_CUDA(cudaMalloc(&W->H, sizeof(double) * (restart + 1) * restart));
  RDC_CUDA(cudaMalloc(&W->cs, sizeof(double) * (restart + 1)));
  RDC_CUDA(cudaMalloc(&W->sn, sizeof(double) * (restart + 1)));
  RDC_CUDA(cudaMalloc(&W->g, sizeof(double) * (restart + 2)));
  RDC_CUDA(cudaMalloc(&W->y, sizeof(double) * (restart + 1)));
  W->restart_cap = restart;
  return 0;
}

static int ensure_extra_vectors(rdc_ctx* c) {
  SolverWork* W = c->work;
  if (W->t2) return 0;
  const size_t vb = W->vec_len * sizeof(double);
  RDC_CUDA(cudaMalloc(&W->t2, vb)); RDC_CUDA(cudaMalloc(&W->t4, vb));
  // the two vectors whose ghosts are exchanged every iteration (p and s) live in the peer-memory arena when there is one
  if (!(W->t3 = p2p_alloc(c, W->vec_len))) RDC_CUDA(cudaMalloc(&W->t3, vb));
  if (!(W->hs = p2p_alloc(c, W->vec_len))) RDC_CUDA(cudaMalloc(&W->hs, vb));
  if (!(W->hp2 = p2p_alloc(c, W->vec_len))) RDC_CUDA(cudaMalloc(&W->hp2, vb));
  RDC_CUDA(cudaMemsetAsync(W->hp2, 0, vb, c->stream));
  RDC_CUDA(cudaMemsetAsync(W->hs, 0, vb, c->stream));
  RDC_CUDA(cudaMemsetAsync(W->t2, 0, vb, c->stream)); RDC_CUDA(cudaMemsetAsync(W->t3, 0, vb, c->stream));
  RDC_CUDA(cudaMemsetAsync(W->t4, 0, vb, c->stream));
  return 0;
}

void solver_free(rdc_ctx* c) {
  SolverWork* W = c->work;
  if (!W) return;
  cudaFree(W->V); cudaFree(W->t0); cudaFree(W->t1); cudaFree(W->t2); cudaFree(W->t4);
  if (!p2p_owns(c, W->t3)) cudaFree(W->t3);
  if (!p2p_owns(c, W->hs)) cudaFree(W->hs);
  if (!p2p_owns(c, W->hp2)) cudaFree(W->hp2);
  cudaFree(W->partial); cudaFree(W->counter); cudaFree(W->h); cudaFree(W->H); cudaFree(W->cs); cudaFree(W->sn);
  cudaFree(W->g); cudaFree(W->y); cudaFree(W->scal); cudaFree(W->state);
  cudaFreeHost(W->h_scal); cudaFreeHost(W->h_state);
  for (int k = 0; k < 2 * SolverWork::MAX_EV; k++) cudaEventDestroy(W->ev[k]);
  cudaFreeHost(W->h_ring);
  cudaFree(W->tiles); cudaFree(W->flag); cudaFree(W->t_spmv); cudaFree(W->rep); cudaFreeHost(W->h_rep);
  delete W;
  c->work = nullptr;
}

// dots of w against V[0..nvec) (+ optional <w,w> appended); result (all-reduced) in W->h[0..]
static int multidot(rdc_ctx* c, int nvec, const double* V, const double* w, bool with_norm) {
  SolverWork* W = c->work;
  const size_t n = (size_t)c->S.n_owned * c->nv;
  int done_cols = 0;
  do {
    const int chunk = nvec - done_cols < 32 ? nvec - done_cols : 32;
    const int wn = (with_norm && done_cols + chunk == nvec) ? 1 : 0;
    const double* Vc = V + (size_t)done_cols * W->vec_len;
    double* out = W->h + done_cols;
    if (chunk <= 8) k_multidot<1><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, wn, W->partial, W->counter, out, W->state);
    else if (chunk <= 16) k_multidot<2><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, wn, W->partial, W->counter, out, W->state);
    else if (chunk <= 24) k_multidot<3><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, wn, W->partial, W->counter, out, W->state);
    else k_multidot<4><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, wn, W->partial, W->counter, out, W->state);
    c->st.kernel_launches++;
    RDC_CUDA(cudaGetLastError());
    done_cols += chunk;
  } while (done_cols < nvec);
  return allreduce_sum(c, W->h, nvec + (with_norm ? 1 : 0));
}

// w -= V[0..nvec) h ; ||w||^2 -> out
static int gs_update(rdc_ctx* c, int nvec, const double* V, double* w, const double* h, double* out) {
  SolverWork* W = c->work;
  const size_t n = (size_t)c->S.n_owned * c->nv;
  int done_cols = 0;
  do {
    const int chunk = nvec - done_cols < 32 ? nvec - done_cols : 32;
    const double* Vc = V + (size_t)done_cols * W->vec_len;
    const double* hc = h + done_cols;
    if (chunk <= 8) k_gs_update<1><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, hc, W->partial, W->counter, out, W->state);
    else if (chunk <= 16) k_gs_update<2><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, hc, W->partial, W->counter, out, W->state);
    else if (chunk <= 24) k_gs_update<3><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, hc, W->partial, W->counter, out, W->state);
    else k_gs_update<4><<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, chunk, Vc, W->vec_len, w, hc, W->partial, W->counter, out, W->state);
    c->st.kernel_launches++;
    RDC_CUDA(cudaGetLastError());
    done_cols += chunk;
  } while (done_cols < nvec);
  return 0;
}

static int poll(rdc_ctx* c) {
  SolverWork* W = c->work;
  RDC_CUDA(cudaMemcpyAsync(W->h_state, W->state, sizeof(int) * 8, cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaMemcpyAsync(W->h_scal, W->scal, sizeof(double) * 8, cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return p2p_check_error(c);   // a timed-out peer exchange surfaces at every host synchronisation of every solver
}

// GMRES(m), left Jacobi (or no) preconditioning.  x = c->d_u (initial guess and result), b = c->d_rhs.
static int gmres(rdc_ctx* c, const double* scale, double rtol, int maxits, int m, int* its_out, double* res_out) {
  SolverWork* W = c->work;
  int rc = ensure_gmres(c, m);
  if (rc) return rc;
  const size_t n = (size_t)c->S.n_owned * c->nv;
  const size_t ld = W->vec_len;
  const int sync_every = c->opt.sync_every > 0 ? c->opt.sync_every : 4;
  RDC_CUDA(cudaMemsetAsync(W->state, 0, sizeof(int) * 8, c->stream));
  // reference norm ||B b||
  k_mul<<<grid_for(n), 256, 0, c->stream>>>(n, c->d_rhs, scale, W->t0);
  c->st.kernel_launches++;
  rc = multidot(c, 0, W->t0, W->t0, true);
  if (rc) return rc;
  k_set_target<<<1, 1, 0, c->stream>>>(W->h, rtol, W->scal);
  c->st.kernel_launches++;
  int its = 0;
  while (true) {
    // v0 = B (b - A x)
    if ((rc = halo_exchange(c, c->d_u))) return rc;
    if ((rc = launch_spmv(c, c->d_u, W->t0, nullptr, false))) return rc;
    k_residual<<<grid_for(n), 256, 0, c->stream>>>(n, c->d_rhs, W->t0, scale, W->V);
    c->st.kernel_launches++;
    if ((rc = multidot(c, 0, W->V, W->V, true))) return rc;
    k_gmres_begin<<<1, 1, 0, c->stream>>>(W->h, W->g, W->scal, W->state, m);
    c->st.kernel_launches++;
    if ((rc = poll(c))) return rc;
    if (W->h_state[0]) break;
    if (its >= maxits) break;
    k_scale_dev<<<grid_for(n), 256, 0, c->stream>>>(n, W->V, W->scal + 3, W->state);
    c->st.kernel_launches++;
    int j = 0;
    for (; j < m && its < maxits; j++) {
      double* vj = W->V + (size_t)j * ld;
      double* vn = W->V + (size_t)(j + 1) * ld;
      if ((rc = halo_exchange(c, vj))) return rc;
      if ((rc = launch_spmv(c, vj, vn, scale, true))) return rc;         // vn = B A vj
      if ((rc = multidot(c, j + 1, W->V, vn, false))) return rc;    // classical Gram-Schmidt: all dots first
      if ((rc = gs_update(c, j + 1, W->V, vn, W->h, W->h + 512))) return rc;
      if ((rc = allreduce_sum(c, W->h + 512, 1))) return rc;
      k_gmres_givens<<<1, 1, 0, c->stream>>>(j, m, W->h, W->h + 512, W->H, W->cs, W->sn, W->g, W->scal, W->state);
      k_scale_dev<<<grid_for(n), 256, 0, c->stream>>>(n, vn, W->scal + 3, W->state);
      c->st.kernel_launches += 2;
      its++;
      if ((j + 1) % sync_every == 0 || j + 1 == m || its >= maxits) {
        if ((rc = poll(c))) return rc;
        if (W->h_state[0]) break;
      }
    }
    // x += V y for the columns completed on the device
    k_gmres_backsolve<<<1, 1, 0, c->stream>>>(m, W->H, W->g, W->y, W->state);
    k_update_x<<<grid_for(n), 256, 0, c->stream>>>(n, W->state + 2, W->V, ld, W->y, c->d_u);
    c->st.kernel_launches += 2;
    RDC_CUDA(cudaGetLastError());
    if ((rc = poll(c))) return rc;
    its = W->h_state[1];
    if (W->h_state[3]) { c->err = "GMRES breakdown (NaN residual)"; *its_out = its; *res_out = W->h_scal[0]; return RDC_E_DIVERGED; }
    if (W->h_state[0]) break;
    if (its >= maxits) break;
    RDC_CUDA(cudaMemsetAsync(W->state, 0, sizeof(int), c->stream));  // keep counters, clear done
  }
  if ((rc = poll(c))) return rc;
  *its_out = W->h_state[1];
  *res_out = W->h_scal[0];
  c->st.resnorm0 = W->h_scal[5];
  return 0;
}

// ---- CG / BiCGStab building blocks: coefficients live in W->scal, computed by one-thread kernels ----
// y = a*x + b*y with a = sa * A[ia] (A == nullptr -> sa), b likewise
__global__ void k_axpby_dev(size_t n, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ S, int ia,
                            double sa, int ib, double sb, const int* __restrict__ done) {
  if (done && *done) return;
  const double a = ia >= 0 ? sa * S[ia] : sa, b = ib >= 0 ? sb * S[ib] : sb;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = fma(a, x[i], b * y[i]);
}
enum { S_RES = 0, S_TARGET = 1, S_RHO = 8, S_RHO_OLD = 9, S_ALPHA = 10, S_BETA = 11, S_OMEGA = 12, S_TMP = 13 };

// PCG scalar steps
__global__ void k_cg_alpha(const double* pAp, double* S, int* state) {  // alpha = rho / <p, Ap>
  if (state[0]) return;
  S[S_ALPHA] = S[S_RHO] / pAp[0];
}
__global__ void k_cg_beta(const double* rz_rr, double* S, int* state) {  // rho_new = <r,z>; res = sqrt(<z,z>) (precond. norm)
  if (state[0]) return;
  S[S_RHO_OLD] = S[S_RHO];
  S[S_RHO] = rz_rr[0];
  S[S_BETA] = S[S_RHO] / S[S_RHO_OLD];
  const double res = sqrt(rz_rr[1]);
  S[S_RES] = res;
  state[1] += 1;
  if (!(res == res)) { state[0] = 1; state[3] = 1; }
  else if (res <= S[S_TARGET]) state[0] = 1;
}

static int axpby(rdc_ctx* c, size_t n, const double* x, double* y, int ia, double sa, int ib, double sb) {
  SolverWork* W = c->work;
  k_axpby_dev<<<grid_for(n), 256, 0, c->stream>>>(n, x, y, W->scal, ia, sa, ib, sb, W->state);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// Jacobi-PCG (for the symmetric positive definite cases only; convergence on ||B r||)
static int pcg(rdc_ctx* c, const double* scale, double rtol, int maxits, int* its_out, double* res_out) {
  SolverWork* W = c->work;
  int rc = ensure_extra_vectors(c);
  if (rc) return rc;
  const size_t n = (size_t)c->S.n_owned * c->nv;
  const int sync_every = c->opt.sync_every > 0 ? c->opt.sync_every : 4;
  double *r = W->t1, *z = W->t2, *p = W->t3, *Ap = W->t4;
  RDC_CUDA(cudaMemsetAsync(W->state, 0, sizeof(int) * 8, c->stream));
  k_mul<<<grid_for(n), 256, 0, c->stream>>>(n, c->d_rhs, scale, W->t0);
  if ((rc = multidot(c, 0, W->t0, W->t0, true))) return rc;
  k_set_target<<<1, 1, 0, c->stream>>>(W->h, rtol, W->scal);
  if ((rc = halo_exchange(c, c->d_u))) return rc;
  if ((rc = launch_spmv(c, c->d_u, W->t0, nullptr, false))) return rc;
  k_residual<<<grid_for(n), 256, 0, c->stream>>>(n, c->d_rhs, W->t0, nullptr, r);
  k_mul<<<grid_for(n), 256, 0, c->stream>>>(n, r, scale, z);
  c->st.kernel_launches += 4;
  RDC_CUDA(cudaMemcpyAsync(p, z, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  // rho = <r,z>, res = ||z||
  if ((rc = multidot(c, 1, r, z, true))) return rc;
  RDC_CUDA(cudaMemcpyAsync(W->scal + S_RHO, W->h, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = poll(c))) return rc;
  {
    double hz[2];
    RDC_CUDA(cudaMemcpy(hz, W->h, 2 * sizeof(double), cudaMemcpyDeviceToHost));
    if (sqrt(hz[1]) <= W->h_scal[S_TARGET]) { *its_out = 0; *res_out = sqrt(hz[1]); c->st.resnorm0 = W->h_scal[5]; return 0; }
  }
  int its = 0;
  while (its < maxits) {
    if ((rc = halo_exchange(c, p))) return rc;
    if ((rc = launch_spmv(c, p, Ap, nullptr, true))) return rc;
    if ((rc = multidot(c, 1, p, Ap, false))) return rc;
    k_cg_alpha<<<1, 1, 0, c->stream>>>(W->h, W->scal, W->state);
    c->st.kernel_launches++;
    if ((rc = axpby(c, n, p, c->d_u, S_ALPHA, 1.0, -1, 1.0))) return rc;    // x += alpha p
    if ((rc = axpby(c, n, Ap, r, S_ALPHA, -1.0, -1, 1.0))) return rc;       // r -= alpha Ap
    k_mul<<<grid_for(n), 256, 0, c->stream>>>(n, r, scale, z);
    c->st.kernel_launches++;
    if ((rc = multidot(c, 1, r, z, true))) return rc;                       // <r,z>, <z,z>
    k_cg_beta<<<1, 1, 0, c->stream>>>(W->h, W->scal, W->state);
    c->st.kernel_launches++;
    if ((rc = axpby(c, n, z, p, -1, 1.0, S_BETA, 1.0))) return rc;          // p = z + beta p
    its++;
    if (its % sync_every == 0 || its >= maxits) {
      if ((rc = poll(c))) return rc;
      if (W->h_state[0]) break;
    }
  }
  if ((rc = poll(c))) return rc;
  *its_out = W->h_state[1];
  *res_out = W->h_scal[S_RES];
  c->st.resnorm0 = W->h_scal[5];
  if (W->h_state[3]) { c->err = "CG breakdown"; return RDC_E_DIVERGED; }
  return 0;
}

// ---- BiCGStab on the left-preconditioned system B A x = B b ---------------------------------------------
// Five launches per iteration: p-update, SpMV (+<r0,v>), s-update, SpMV (+<s,t>,<t,t>), x/r-update (+<r0,r>,<r,r>).
// No scalar kernels: the dot products land in the slot table D[] below (all-reduced in place when distributed) and
// every consumer derives alpha / omega / beta / the convergence decision from D in its prologue -- all threads
// compute the same values, block 0 records them.  The p-update of iteration `it` is also the convergence check of
// iteration it-1; the host reads the flag through a pipelined asynchronous copy, a few iterations late, while the
// kernels already queued return at once.
enum { D_R0V = 1, D_TS = 2, D_TT = 3, D_XR0 = 4 /* {<r0,r>,<r,r>} parity 0; parity 1 at 6,7 */, D_INIT = 8 /* {<r,r>, ||Bb||^2} */ };
enum { S_BNORM = 5 };  // next to S_RES / S_TARGET above

__device__ __forceinline__ double bi_target(const double* D, double rtol) { return fmax(rtol * sqrt(D[D_INIT + 1]), 1e-50); }

// Bundle of the ghost exchange that a vector kernel performs for the vector it produces (distributed runs): the
// first hb.total blocks of the grid recompute the boundary entries (same expression, same inputs -> same bits as the
// main blocks) and store them straight into the neighbours' ghost tails; see p2p.cu.
struct HaloBundle {
  HaloArgs A;
  const int32_t* send_idx = nullptr;
  P2PHeader* hdr = nullptr;
  unsigned long long seq = 0;
  int total = 0;   // number of exchange blocks at the front of the grid (0: no exchange)
  int nv = 1;
};

// p_new = r + beta (p_old - omega v)    (it == 0: p_new = r).  rn / ro: slots of the newest and the previous <r0,r>.
__global__ void __launch_bounds__(RED_THREADS) k_bi_p(size_t n, int it, int rn, int ro, double rtol, const double* __restrict__ r,
                                                      const double* __restrict__ v, const double* __restrict__ p_old,
                                                      double* __restrict__ p_new, const double* __restrict__ D,
                                                      double* __restrict__ S, int* state, volatile int* host_flag,
                                                      const HaloBundle hb) {
  if (state[0]) {  // converged earlier: tell the host (zero-copy pinned memory), nothing else to do
    if (blockIdx.x == 0 && threadIdx.x == 0) *host_flag = (it + 1) | (1 << 30);
    return;
  }
  const double rho = D[rn], rr = D[rn + (it == 0 ? 0 : 1)];
  const double res = sqrt(rr), target = bi_target(D, rtol);
  double beta = 0.0, omega = 0.0;
  int stop = 0, bad = 0;
  if (!(res == res)) { stop = 1; bad = 1; }
  else if (res <= target) stop = 1;
  else if (it > 0) {
    const double rho_old = D[ro];
    const double alpha = rho_old / D[D_R0V];
    omega = D[D_TT] != 0.0 ? D[D_TS] / D[D_TT] : 0.0;
    if (omega == 0.0 || rho == 0.0) { stop = 1; bad = 1; }
    beta = (rho / rho_old) * (alpha / omega);
  }
  if (blockIdx.x == (unsigned)hb.total && threadIdx.x == 0) {   // first main block keeps the books
    S[S_RES] = res; S[S_TARGET] = target; S[S_BNORM] = sqrt(D[D_INIT + 1]);
    state[1] = it;
    if (stop) { state[3] = bad; __threadfence(); state[0] = 1; }
    *host_flag = (it + 1) | (stop << 30);
  }
  if (stop) return;   // every rank takes the same decision from the same all-reduced values: nobody exchanges
  if ((int)blockIdx.x < hb.total) {
    int k = 0;
    while ((int)blockIdx.x >= hb.A.blk_ptr[k + 1]) k++;
    const int nb = hb.A.nblk[k], b = (int)blockIdx.x - hb.A.blk_ptr[k];
    halo_exchange_block(hb.A, k, b, nb, hb.nv, hb.send_idx, p_new, hb.seq, hb.hdr, [=](size_t j) {
      return it == 0 ? r[j] : fma(beta, fma(-omega, v[j], p_old[j]), r[j]);
    });
    return;
  }
  const size_t first = (blockIdx.x - hb.total) * (size_t)blockDim.x + threadIdx.x, step = (size_t)(gridDim.x - hb.total) * blockDim.x;
  if (it == 0) {
    for (size_t i = first; i < n; i += step) p_new[i] = r[i];
  } else {
    for (size_t i = first; i < n; i += step) p_new[i] = fma(beta, fma(-omega, v[i], p_old[i]), r[i]);
  }
}
// s = r - alpha v   (+ ghost exchange of s, see HaloBundle)
__global__ void __launch_bounds__(RED_THREADS) k_bi_s(size_t n, int ro, const double* __restrict__ r, const double* __restrict__ v,
                                                      double* __restrict__ s, const double* __restrict__ D,
                                                      const int* __restrict__ state, const HaloBundle hb, const int rev) {
  if (state[0]) return;
  const double alpha = D[ro] / D[D_R0V];
  if ((int)blockIdx.x < hb.total) {
    int k = 0;
    while ((int)blockIdx.x >= hb.A.blk_ptr[k + 1]) k++;
    const int nb = hb.A.nblk[k], b = (int)blockIdx.x - hb.A.blk_ptr[k];
    halo_exchange_block(hb.A, k, b, nb, hb.nv, hb.send_idx, s, hb.seq, hb.hdr, [=](size_t j) { return fma(-alpha, v[j], r[j]); });
    return;
  }
  const size_t first = (blockIdx.x - hb.total) * (size_t)blockDim.x + threadIdx.x, step = (size_t)(gridDim.x - hb.total) * blockDim.x;
  // back to front: the SpMV before this kernel wrote v in ascending row order, so its tail is what is still in L2; and the
  // head of s, written last here, is what the next SpMV's first tiles gather
  for (size_t k = first; k < n; k += step) {
    const size_t i = rev ? n - 1 - k : k;
    s[i] = fma(-alpha, v[i], r[i]);
  }
}
// x += alpha p + omega s ; r = s - omega t ; out = {<r0,r>, <r,r>}
__global__ void __launch_bounds__(RED_THREADS) k_bi_xr(size_t n, int ro, double* __restrict__ x, const double* __restrict__ p,
                                                       const double* __restrict__ s, const double* __restrict__ t,
                                                       double* __restrict__ r, const double* __restrict__ r0,
                                                       const double* __restrict__ D, double* partial, unsigned* counter,
                                                       double* out, const int* __restrict__ state, const ArCtx ar) {
  if (state[0]) return;
  const double alpha = D[ro] / D[D_R0V];
  const double omega = D[D_TT] != 0.0 ? D[D_TS] / D[D_TT] : 0.0;
  double acc[2] = {0.0, 0.0};
  // always front to back: the two dot products are summed per thread in this order, and the persistent kernel's XR phase
  // sums them the same way -- walking backwards here (measured: -0.03 ms per step) would change their last bits and
  // with them the bit-identity of the five-launch and the one-launch solver (test_solve_independent_of_spmv_variant)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double si = s[i];
    x[i] = fma(omega, si, fma(alpha, p[i], x[i]));
    const double ri = fma(-omega, t[i], si);
    r[i] = ri;
    acc[0] = fma(r0[i], ri, acc[0]);
    acc[1] = fma(ri, ri, acc[1]);
  }
  grid_reduce<2>(acc, 2, partial, counter, out, ar);
}

static bool fused_halo_ok(rdc_ctx* c, const double* x) {
  (void)x;  // any device vector can be exchanged: only the staging areas live in peer memory
  return c->opt.p2p_fused_halo && c->S.nranks > 1 && p2p_on(c) && !c->S.nbr_rank.empty();
}
// describe the exchange of arena vector x for a fused vector kernel (bookkeeping as in p2p_launch_halo)
static int halo_bundle(rdc_ctx* c, const double* x, HaloBundle* hb) {
  (void)x;
  int max_blk = 1;
  p2p_fill_halo_args(c, &hb->A, &max_blk, &hb->total, &hb->seq);
  hb->send_idx = c->d_send_idx;
  hb->hdr = (P2PHeader*)c->p2p->arena;
  hb->nv = c->nv;
  return 0;
}

// RDC_TRACE=1: event-bracket every operation of BiCGStab iteration 4 and print the device time of each (debug aid)
struct IterTrace {
  static constexpr int N = 24;
  cudaEvent_t ev[N];
  const char* name[N];
  int n = 0;
  bool on = false, made = false;
  void mark(const char* what, cudaStream_t st) {
    if (!on || n >= N) return;
    if (!made) { for (int k = 0; k < N; k++) cudaEventCreate(&ev[k]); made = true; }
    name[n] = what;
    cudaEventRecord(ev[n++], st);
  }
  void report(int rank) {
    if (!on || n < 2) return;
    cudaEventSynchronize(ev[n - 1]);
    char line[1024];
    int o = snprintf(line, sizeof(line), "[rdc trace rank %d]", rank);
    for (int k = 1; k < n; k++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[k - 1], ev[k]);
      o += snprintf(line + o, sizeof(line) - o, " %s=%.1fus", name[k], ms * 1e3f);
    }
    fprintf(stderr, "%s\n", line);
    n = 0;
  }
};

static int bicgstab(rdc_ctx* c, const double* scale, double rtol, int maxits, int* its_out, double* res_out) {
  SolverWork* W = c->work;
  static IterTrace TR;
  const int trace_env = c->opt.trace;
  int rc = ensure_extra_vectors(c);
  if (rc) return rc;
  rc = ensure_gmres(c, 1);  // borrow V for one more vector
  if (rc) return rc;
  const size_t n = (size_t)c->S.n_owned * c->nv;
  const int depth_env = c->opt.sync_every;  // iterations queued ahead of the convergence flag the host has seen
  // collectives cannot return early once convergence is flagged, so fewer iterations are queued ahead when distributed
  int depth = depth_env > 0 ? depth_env : (c->S.nranks > 1 ? 2 : 4);
  if (depth > SolverWork::RING - 1) depth = SolverWork::RING - 1;
  double *r = W->t1, *r0 = W->t2, *v = W->t4, *s = W->hs, *t = W->V;
  double* pbuf[2] = {W->t3, W->hp2};   // p is double-buffered: the exchange blocks of the p-update read the old p
  double* D = W->h;
  const unsigned vg = grid_for(n);
  RDC_CUDA(cudaMemsetAsync(W->state, 0, sizeof(int) * 8, c->stream));
  memset(W->h_ring, 0, sizeof(int) * SolverWork::RING);  // the previous solve ended with a stream synchronisation
  // r = r0 = B (b - A x), <r,r>, ||B b||^2 in one pass over the operator
  if ((rc = refresh_u_ghosts(c))) return rc;
  bool fused;
  c->st.bicg_persistent = 0;
  {
    const ArCtx ar = ar_begin(c, &fused);
    if ((rc = spmv(c, SPMV_RESID, c->d_u, r, scale, c->d_rhs, r0, D + D_INIT, false, false, ar))) return rc;
    if (!fused && (rc = allreduce_sum(c, D + D_INIT, 2))) return rc;
  }
  // pipelined polls without stream work: the p-update of iteration `it` stores its decision in a ring of pinned,
  // device-mapped host words; the host looks at the word of iteration it-depth, which has long been written
  for (int it = 0;; it++) {
    const int rn = it == 0 ? D_INIT : D_XR0 + 2 * ((it - 1) & 1);
    const int ro = it <= 1 ? D_INIT : D_XR0 + 2 * ((it - 2) & 1);
    TR.on = trace_env && it == 4;
    TR.mark("start", c->stream);
    const int slot = it % SolverWork::RING;
    double *p = pbuf[it & 1], *p_old = pbuf[(it + 1) & 1];
    HaloBundle hb;
    const bool fuse_halo = fused_halo_ok(c, p) && fused_halo_ok(c, s);
    if (fuse_halo && it < maxits) { if ((rc = halo_bundle(c, p, &hb))) return rc; }
    k_bi_p<<<vg + hb.total, RED_THREADS, 0, c->stream>>>(n, it, rn, ro, rtol, r, v, p_old, p, D, W->scal, W->state, W->h_ring + slot, hb);
    TR.mark("p_update", c->stream);
    c->st.kernel_launches++;
    if (it >= maxits) break;
    if (it >= depth) {  // the p-update of iteration it-depth wrote (it-depth+1) | stop<<30 straight into pinned host memory
      volatile int* f = W->h_ring + (it - depth) % SolverWork::RING;
      int spins = 0;
      while ((*f & 0x3fffffff) != it - depth + 1) {
        if (++spins > 2000) {  // it should be there already; fall back to waiting for the stream (also surfaces errors)
          RDC_CUDA(cudaStreamSynchronize(c->stream));
          if ((*f & 0x3fffffff) != it - depth + 1) { c->err = "BiCGStab: convergence flag never arrived"; return RDC_E_CUDA; }
        }
      }
      if (*f >> 30) break;
    }
    TR.mark("poll", c->stream);
    if (!fuse_halo && (rc = halo_exchange(c, p, true))) return rc;
    TR.mark("halo_p", c->stream);
    ArCtx ar = ar_begin(c, &fused);
    if ((rc = spmv(c, SPMV_DOT_W, p, v, scale, r0, nullptr, D + D_R0V, true, true, ar))) return rc;     // v = B A p, <r0,v>
    TR.mark("spmv1", c->stream);
    if (!fused && (rc = allreduce_sum(c, D + D_R0V, 1, true))) return rc;
    TR.mark("ar1", c->stream);
    hb = HaloBundle();
    if (fuse_halo && (rc = halo_bundle(c, s, &hb))) return rc;
    k_bi_s<<<vg + hb.total, RED_THREADS, 0, c->stream>>>(n, rn, r, v, s, D, W->state, hb, c->opt.vec_reverse);
    TR.mark("s_update", c->stream);
    if (!fuse_halo && (rc = halo_exchange(c, s, true))) return rc;
    TR.mark("halo_s", c->stream);
    ar = ar_begin(c, &fused);
    if ((rc = spmv(c, SPMV_DOT_SELF, s, t, scale, nullptr, nullptr, D + D_TS, true, true, ar))) return rc;  // t = B A s, <s,t>, <t,t>
    TR.mark("spmv2", c->stream);
    if (!fused && (rc = allreduce_sum(c, D + D_TS, 2, true))) return rc;
    TR.mark("ar2", c->stream);
    double* xr_out = D + D_XR0 + 2 * (it & 1);
    ar = ar_begin(c, &fused);
    k_bi_xr<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, rn, c->d_u, p, s, t, r, r0, D, W->partial, W->counter, xr_out, W->state, ar);
    TR.mark("xr_update", c->stream);
    if (!fused && (rc = allreduce_sum(c, xr_out, 2, true))) return rc;
    TR.mark("ar3", c->stream);
    TR.report(c->S.rank);
    c->st.kernel_launches += 2;
    RDC_CUDA(cudaGetLastError());
  }
  if ((rc = poll(c))) return rc;
  *its_out = W->h_state[1];
  *res_out = W->h_scal[S_RES];
  c->st.resnorm0 = W->h_scal[S_BNORM];
  if (W->h_state[3]) { c->err = "BiCGStab breakdown"; return RDC_E_DIVERGED; }
  return 0;
}


// ============================================================================================================
// Persistent BiCGStab: the whole solve is ONE cooperative launch.
//
// The five-launch iteration above pays, per kernel, the launch, the ramp of 888 CTAs and the tail of the slowest one
// (~10 us each at 1.3 M tets per GPU -- a third of the iteration at 8 GPUs, 12 % at one GPU).  Here a resident grid
// (the TMA SpMV's own: 148 x CTAs/SM, operator tiles dealt round-robin) walks through the phases of every iteration
//   P: p = r + beta (p - omega v) [+ghosts] | S1: v = B A p, <r0,v> | S: s = r - alpha v [+ghosts] |
//   S2: t = B A s, <s,t>, <t,t> | XR: x += alpha p + omega s, r = s - omega t, <r0,r>, <r,r>
// separated by grid barriers.  The three reductions ARE barriers: every CTA adds its partial sums to a list, the last
// one to arrive adds the list in a fixed order (and, distributed, finishes the sum across the ranks over NVLink peer
// memory with the tag-in-word slots), publishes the result and releases the epoch flag all CTAs spin on.  Scalars
// (alpha, omega, beta) and the convergence decision are derived by every thread from the same published sums, so all
// CTAs -- and all ranks -- leave the loop in the same iteration without any host involvement.
// Coherence: vectors written in one phase are read in the next by other CTAs; the barrier's release/acquire pair plus a
// gpu-scope fence in every thread (which drops the SM's L1 lines) makes the plain, L1-cached loads of the next phase see
// them -- the x gather of the SpMV keeps its L1 reuse inside a phase.
// Same arithmetic, same summation orders as the five-launch version: the two produce bit-identical iterates (tested).
struct PersistHalo {
  HaloArgs A[2];                 // by parity of the exchange sequence number (staging areas are double-buffered)
  const int32_t* send_idx = nullptr;
  P2PHeader* hdr = nullptr;
  unsigned long long seq0 = 0;   // exchanges done before this solve
  int on = 0, nv = 1;
};
struct PersistArgs {
  int n_tiles; int uniform16; int l2_hint; const int4* tiles; const int32_t* rowptr; const int32_t* col; const double* val;
  const double* scale; const double* b;
  size_t n;
  double *x, *r, *r0, *v, *s, *t, *p0, *p1;
  double* partial; unsigned* counter; unsigned* flag; unsigned* flag_other; double* D; double* S; int* state;
  double rtol; int maxits;
  int timing;          // block 0 times the phases with %globaltimer (each read costs ~1 us of the critical path)
  ArCtx ar;            // ar.seq = all-reduces done before this solve
  PersistHalo halo;
  unsigned long long* t_spmv;   // [8] globaltimer ns of block 0: [0] SpMV phases, [1] their number, [2] P, [3] S, [4] XR phases (each with its barrier), [5] set-up
};

// arrive at barrier `epoch` (1, 2, ...) and wait for it.  Release: every thread's writes -> __syncthreads -> thread 0
// fence + atomic; acquire: thread 0 spins on the flag and fences (L1 invalidation), __syncthreads.
__device__ __forceinline__ void grid_wait(unsigned* flag, unsigned epoch) {
  if (threadIdx.x == 0) {
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v < epoch);
    __threadfence();   // one gpu-scope fence per CTA: orders the CTA behind the release and drops the SM's stale L1 lines
  }
  __syncthreads();     // the other threads are ordered behind thread 0's acquire by the CTA barrier (cumulativity)
}
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned* flag, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {
      *counter = 0u;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
  }
  grid_wait(flag, epoch);
}
// reduction + barrier: out[k] = sum over all threads of the grid (and all ranks) of v[k], k < nval <= 2
__device__ __forceinline__ void grid_reduce_barrier(double (&v)[2], int nval, double* partial, unsigned* counter, unsigned* flag,
                                                    unsigned epoch, double* out, const ArCtx& ar) {
  __shared__ double s_red[RED_THREADS / 32][2];
  __shared__ double s_tot[2];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const double w = warp_sum(v[k]);
    if (lane == 0) s_red[wid][k] = w;
  }
  __syncthreads();
  if (threadIdx.x < 2 && threadIdx.x < nval) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < RED_THREADS / 32; w++) s += s_red[w][threadIdx.x];
    partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    __threadfence();   // the two writers publish their partial sums ...
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(counter, 1u) == gridDim.x - 1); }   // ... and the CTA's vector writes
  __syncthreads();
  if (s_last) {   // same order of additions as grid_reduce: thread t takes blocks t, t+256, ..., then the fixed tree --
    // but both values in one pass and all loads of a thread in flight together (the partial list sits in L2: four
    // dependent round trips per value were 3 us of a 7.7 us reduction-barrier)
    __threadfence();
    constexpr int NB = (SPMV_MAX_GRID + RED_THREADS - 1) / RED_THREADS;
    double pv[2][NB];
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const unsigned b = threadIdx.x + (unsigned)j * RED_THREADS;
#pragma unroll
      for (int k = 0; k < 2; k++) pv[k][j] = (b < gridDim.x && k < nval) ? __ldcg(partial + (size_t)k * gridDim.x + b) : 0.0;
    }
    double sk[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 2; k++) {
#pragma unroll
      for (int j = 0; j < NB; j++)
        if (threadIdx.x + (unsigned)j * RED_THREADS < gridDim.x) sk[k] += pv[k][j];
      sk[k] = warp_sum(sk[k]);
    }
    __syncthreads();
    if (lane == 0) { s_red[wid][0] = sk[0]; s_red[wid][1] = sk[1]; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < RED_THREADS / 32; w++) t += s_red[w][threadIdx.x];
      s_tot[threadIdx.x] = t;
    }
    __syncthreads();
    if (ar.nranks > 1) {
      const int par = (int)(ar.seq & 1ull);
      const unsigned tag = (unsigned)ar.seq;
      const int q = threadIdx.x / 2, k = threadIdx.x % 2;
      double got = 0.0;
      if (q < ar.nranks && k < nval) {
        ll_store(&ar.peer[q]->ll[par][ar.me][k][0], s_tot[k], tag);
        const unsigned long long t0 = global_ns();
        int spins = 0;
        while (!ll_load(&ar.mine->ll[par][q][k][0], tag, &got)) {
          if ((++spins & 1023) == 0 && global_ns() - t0 > P2P_TIMEOUT_NS) { ar.mine->error = 1; break; }
        }
      }
      __syncthreads();
      __shared__ double s_in[RDC_MAX_RANKS][2];
      if (q < ar.nranks && k < nval) s_in[q][k] = got;
      __syncthreads();
      if (threadIdx.x < 2 && threadIdx.x < nval) {
        double t = 0.0;
        for (int r = 0; r < ar.nranks; r++) t += s_in[r][threadIdx.x];
        out[threadIdx.x] = t;
      }
    } else if (threadIdx.x < 2 && threadIdx.x < nval) {
      out[threadIdx.x] = s_tot[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *counter = 0u;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
  }
  grid_wait(flag, epoch);
}

// Reduction + barrier of the persistent solver, second form: every CTA leaves its partial sums in the list, passes ONE plain
// barrier, then reads the whole list itself (<= 1024 entries from L2, all loads in flight) and adds it in the fixed order
// -- every CTA gets the bit-identical total without waiting for a designated block to compute and publish it (7.3 -> ~4 us
// per reduction on one GPU).  Distributed: block 0 sends the rank's total to every peer (tag-in-word slots) and EVERY CTA
// polls the slots of all ranks in its own rank's header and adds them in rank order.  The results go to the CTA's own
// copy of the slot table in shared memory (s_out), so the scalars of the recurrences are read from shared memory.
static constexpr int PERSIST_MAX_GRID = 1024;
__device__ __noinline__ void grid_allreduce(double v0, double v1, int nval, double* partial, unsigned* counter, unsigned* flag,
                                            unsigned epoch, const ArCtx ar, double* s_out) {
  __shared__ double s_red[RED_THREADS / 32][2];
  __shared__ double s_tot[2];
  __shared__ double s_in[RDC_MAX_RANKS][2];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  {
    const double w0 = warp_sum(v0), w1 = warp_sum(v1);
    if (lane == 0) { s_red[wid][0] = w0; s_red[wid][1] = w1; }
  }
  __syncthreads();
  if (threadIdx.x < 2 && threadIdx.x < nval) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < RED_THREADS / 32; w++) s += s_red[w][threadIdx.x];
    partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    __threadfence();
  }
  grid_barrier(counter, flag, epoch);
  constexpr int NB = PERSIST_MAX_GRID / RED_THREADS;
  double pv[2][NB];
#pragma unroll
  for (int j = 0; j < NB; j++) {
    const unsigned b = threadIdx.x + (unsigned)j * RED_THREADS;
#pragma unroll
    for (int k = 0; k < 2; k++) pv[k][j] = (b < gridDim.x && k < nval) ? __ldcg(partial + (size_t)k * gridDim.x + b) : 0.0;
  }
  double sk[2] = {0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 2; k++) {
#pragma unroll
    for (int j = 0; j < NB; j++)
      if (threadIdx.x + (unsigned)j * RED_THREADS < gridDim.x) sk[k] += pv[k][j];
    sk[k] = warp_sum(sk[k]);
  }
  if (lane == 0) { s_red[wid][0] = sk[0]; s_red[wid][1] = sk[1]; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < RED_THREADS / 32; w++) t += s_red[w][threadIdx.x];
    s_tot[threadIdx.x] = t;
  }
  __syncthreads();
  if (ar.nranks > 1) {
    const int par = (int)(ar.seq & 1ull);
    const unsigned tag = (unsigned)ar.seq;
    const int q = threadIdx.x / 2, k = threadIdx.x % 2;
    if (q < ar.nranks && k < nval) {
      if (blockIdx.x == 0) ll_store(&ar.peer[q]->ll[par][ar.me][k][0], s_tot[k], tag);
      double got = 0.0;
      const unsigned long long t0 = global_ns();
      int spins = 0;
      while (!ll_load(&ar.mine->ll[par][q][k][0], tag, &got)) {
        if ((++spins & 1023) == 0 && global_ns() - t0 > P2P_TIMEOUT_NS) { ar.mine->error = 1; break; }
      }
      s_in[q][k] = got;
    }
    __syncthreads();
    if (threadIdx.x < 2 && threadIdx.x < nval) {
      double t = 0.0;
      for (int r = 0; r < ar.nranks; r++) t += s_in[r][threadIdx.x];
      s_out[threadIdx.x] = t;
    }
  } else if (threadIdx.x < 2 && threadIdx.x < nval) {
    s_out[threadIdx.x] = s_tot[threadIdx.x];
  }
  __syncthreads();
}

// one SpMV phase of the persistent kernel: the tile loop of k_spmv_tma with a tile counter `gi` that keeps running over
// the phases (stage index and mbarrier parity follow it)
// Not inlined on purpose: the phases of the persistent kernel are separate functions so that each gets the whole register
// budget of the kernel (40 registers at 6 CTAs/SM) for its own loop; inlined, the loop-carried state of the solver was
// spilled INSIDE the tile loop (800 bytes of spill traffic per thread, +12 % per iteration).
struct SpmvOp {
  int n_tiles; int uniform16; int l2_hint; const int4* tiles; const int32_t* rowptr; const int32_t* col; const double* val; const double* scale;
};
struct SpmvRes { double d0, d1; unsigned gi; };
template <int NV, unsigned KMASK, int MODE>
__device__ __noinline__ SpmvRes persist_spmv(const SpmvOp A, const double* __restrict__ x, double* __restrict__ y,
                                             const double* __restrict__ w, double* __restrict__ y2, unsigned char* s_raw,
                                             unsigned long long* s_bar, unsigned gi) {
  constexpr int G = 16, STAGES = 2;
  constexpr int NKV = popc_c(KMASK);
  typedef SpmvStage<NKV> ST;
  const int tid = threadIdx.x, lane = tid & (G - 1), hw = tid / G;
  const unsigned hmask = 0xffffu << (tid & 16);
  const int4* __restrict__ tiles = A.tiles;
  const int n_tiles = A.n_tiles;
  double d[2] = {0.0, 0.0};
  const unsigned long long pol_stream = l2_policy_evict_first();
  auto issue = [&](int tile, unsigned slot, const int4 t) {
    const int skip_v = (int)(((long long)t.z * NKV) & 1), skip_c = t.z & 3, skip_r = t.x & 3;
    const unsigned vb = (unsigned)(((skip_v + t.w * NKV) * 8 + 15) & ~15);
    const unsigned cb = (unsigned)(((skip_c + t.w) * 4 + 15) & ~15);
    const unsigned rb = (unsigned)(((skip_r + t.y + 1) * 4 + 15) & ~15);
    unsigned char* base = s_raw + (size_t)(slot % STAGES) * ST::BYTES;
    const unsigned bar = smem_u32(&s_bar[slot % STAGES]);
    mbar_expect_tx(bar, vb + cb + rb + 16u);
    if (A.l2_hint) {
      bulk_g2s_hint(smem_u32(base), A.val + ((long long)t.z * NKV - skip_v), vb, bar, pol_stream);
      bulk_g2s_hint(smem_u32(base + ST::VAL_BYTES), A.col + (t.z - skip_c), cb, bar, pol_stream);
    } else {
      bulk_g2s(smem_u32(base), A.val + ((long long)t.z * NKV - skip_v), vb, bar);
      bulk_g2s(smem_u32(base + ST::VAL_BYTES), A.col + (t.z - skip_c), cb, bar);
    }
    bulk_g2s(smem_u32(base + ST::VAL_BYTES + ST::COL_BYTES), A.rowptr + (t.x - skip_r), rb, bar);
    bulk_g2s(smem_u32(base + ST::DESC_OFF), tiles + tile, 16u, bar);   // the consumers read the descriptor from the stage
  };
  const int first = (int)blockIdx.x, tstride = (int)gridDim.x;
  int4 pn = make_int4(0, 0, 0, 0);
  if (tid == 0) {
    if (first < n_tiles) issue(first, gi, tiles[first]);
    if (first + tstride < n_tiles) pn = tiles[first + tstride];
  }
#pragma unroll 1
  for (int tile = first; tile < n_tiles; tile += tstride, gi++) {
    const unsigned stage = gi % STAGES;
    if (tid == 0) {
      if (tile + tstride < n_tiles) issue(tile + tstride, gi + 1, pn);
      if (tile + 2 * tstride < n_tiles) pn = tiles[tile + 2 * tstride];
    }
    double sc = 1.0, wv = 0.0;
    if (A.uniform16) {   // regular tiles: the epilogue operands are fetched while the tile is still landing (see k_spmv_tma)
      const int row = tile * SPMV_TILE_ROWS + hw;
      if (row < A.uniform16 && lane < NV) {
        const size_t o = (size_t)row * NV + lane;
        if (A.scale) sc = A.scale[o];
        if (MODE == SPMV_DOT_W || MODE == SPMV_RESID) wv = w[o];
        if (MODE == SPMV_DOT_SELF) wv = x[o];
      }
    }
    mbar_wait(smem_u32(&s_bar[stage]), (unsigned)((gi / STAGES) & 1u));
    const int4 t = *reinterpret_cast<const int4*>(s_raw + (size_t)stage * ST::BYTES + ST::DESC_OFF);
    const bool live = hw < t.y;
    const int row = t.x + hw;
    const size_t o = (size_t)row * NV + (lane < NV ? lane : 0);
    if (!A.uniform16 && live && lane < NV) {
      if (A.scale) sc = A.scale[o];
      if (MODE == SPMV_DOT_W || MODE == SPMV_RESID) wv = w[o];
      if (MODE == SPMV_DOT_SELF) wv = x[o];
    }
    if (live) {
      const unsigned char* base = s_raw + (size_t)stage * ST::BYTES;
      const double* s_val = reinterpret_cast<const double*>(base) + (((long long)t.z * NKV) & 1);
      const int* s_col = reinterpret_cast<const int*>(base + ST::VAL_BYTES) + (t.z & 3);
      const int* s_rp = reinterpret_cast<const int*>(base + ST::VAL_BYTES + ST::COL_BYTES) + (t.x & 3);
      const int r0 = s_rp[hw] - t.z;
      const int L = s_rp[hw + 1] - s_rp[hw];
      double acc[NV];
#pragma unroll
      for (int a = 0; a < NV; a++) acc[a] = 0.0;
#pragma unroll 1
      for (int k = lane; k < L; k += G) {
        const int c = s_col[r0 + k];
        double xv[NV];
#pragma unroll
        for (int b = 0; b < NV; b++) xv[b] = x[(size_t)c * NV + b];
        const double* v0 = s_val + (size_t)r0 * NKV + k;
#pragma unroll
        for (int ab = 0; ab < NV * NV; ab++)
          if (KMASK >> ab & 1u) acc[ab / NV] = fma(v0[slot_c(KMASK, ab) * L], xv[ab % NV], acc[ab / NV]);
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1)
#pragma unroll
        for (int a = 0; a < NV; a++) acc[a] += __shfl_xor_sync(hmask, acc[a], off, G);
      if (lane < NV) {
        double mine = acc[0];
#pragma unroll
        for (int a = 1; a < NV; a++)
          if (lane == a) mine = acc[a];
        if (MODE == SPMV_DOT_W) {
          const double yv = mine * sc;
          y[o] = yv;
          d[0] = fma(wv, yv, d[0]);
        } else if (MODE == SPMV_DOT_SELF) {
          const double yv = mine * sc;
          y[o] = yv;
          d[0] = fma(wv, yv, d[0]);
          d[1] = fma(yv, yv, d[1]);
        } else {
          const double yv = (wv - mine) * sc;
          y[o] = yv;
          if (y2) y2[o] = yv;
          d[0] = fma(yv, yv, d[0]);
          d[1] = fma(wv * sc, wv * sc, d[1]);
        }
      }
    }
    __syncthreads();
  }
  SpmvRes out;
  out.d0 = d[0]; out.d1 = d[1]; out.gi = gi;
  return out;
}

// the three vector phases (own entries only; the ghost exchange is done by the caller around them)
__device__ __noinline__ void persist_p(size_t n, int it, double beta, double omega, const double* __restrict__ r,
                                       const double* __restrict__ v, const double* __restrict__ po, double* __restrict__ pn) {
  const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  if (it == 0) {
#pragma unroll 4
    for (size_t i = first; i < n; i += step) pn[i] = r[i];
  } else {
#pragma unroll 4
    for (size_t i = first; i < n; i += step) pn[i] = fma(beta, fma(-omega, v[i], po[i]), r[i]);
  }
}
__device__ __noinline__ void persist_s(size_t n, double alpha, const double* __restrict__ r, const double* __restrict__ v,
                                       double* __restrict__ sn) {
  const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
#pragma unroll 4
  for (size_t i = first; i < n; i += step) sn[i] = fma(-alpha, v[i], r[i]);
}
__device__ __noinline__ double2 persist_xr(size_t n, double alpha, double om, double* __restrict__ x, const double* __restrict__ pp,
                                          const double* __restrict__ s, const double* __restrict__ t, double* __restrict__ r,
                                          const double* __restrict__ r0) {
  const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll 2
  for (size_t i = first; i < n; i += step) {
    const double si = s[i];
    x[i] = fma(om, si, fma(alpha, pp[i], x[i]));
    const double ri = fma(-om, t[i], si);
    r[i] = ri;
    a0 = fma(r0[i], ri, a0);
    a1 = fma(ri, ri, a1);
  }
  return make_double2(a0, a1);
}

// ghost exchange of a vector inside a phase: virtual exchange blocks are dealt to the CTAs round-robin; all sends of a
// CTA go out before it polls, so no CTA ever waits for a peer before the peer's data can be on its way
template <class F>
__device__ __forceinline__ void persist_halo_send(const PersistHalo& H, int par, unsigned long long seq, F val) {
  const HaloArgs& A = H.A[par];
  const unsigned tag = (unsigned)seq;
  const int total = A.blk_ptr[A.n_nbr];
  for (int vb = (int)blockIdx.x; vb < total; vb += (int)gridDim.x) {
    int k = 0;
    while (vb >= A.blk_ptr[k + 1]) k++;
    const int nb = A.nblk[k], b = vb - A.blk_ptr[k];
    const int s0 = A.send_ptr[k], scnt = (A.send_ptr[k + 1] - s0) * H.nv;
    ulonglong2* dst = A.dst[k];
    for (int i = b * blockDim.x + threadIdx.x; i < scnt; i += nb * blockDim.x) {
      const int node = i / H.nv, a = i - node * H.nv;
      tag_store(dst + i, val((size_t)H.send_idx[s0 + node] * H.nv + a), tag);
    }
  }
}
__device__ __forceinline__ void persist_halo_recv(const PersistHalo& H, int par, unsigned long long seq, double* __restrict__ x) {
  const HaloArgs& A = H.A[par];
  const unsigned tag = (unsigned)seq;
  const int total = A.blk_ptr[A.n_nbr];
  for (int vb = (int)blockIdx.x; vb < total; vb += (int)gridDim.x) {
    int k = 0;
    while (vb >= A.blk_ptr[k + 1]) k++;
    const int nb = A.nblk[k], b = vb - A.blk_ptr[k];
    const int r0 = A.recv_ptr[k] * H.nv, rcnt = (A.recv_ptr[k + 1] - A.recv_ptr[k]) * H.nv;
    const ulonglong2* src = A.src + r0;
    double* ghost = x + (size_t)A.n_owned * H.nv + r0;
    for (int i = b * blockDim.x + threadIdx.x; i < rcnt; i += nb * blockDim.x) {
      double v = 0.0;
      const unsigned long long t0 = global_ns();
      int spins = 0;
      bool got;
      while (!(got = tag_load(src + i, tag, &v))) {
        if ((++spins & 1023) == 0 && global_ns() - t0 > P2P_TIMEOUT_NS) { H.hdr->error = 1; break; }
      }
      if (got) ghost[i] = v;
    }
  }
}

template <int NV, unsigned KMASK>
__global__ void __launch_bounds__(RED_THREADS, (NV == 3 ? 6 : 2)) k_bicgstab_persist(const PersistArgs A) {
  constexpr int STAGES = 2;
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(smem_u32(&s_bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // the epoch flag of the NEXT solve is cleared here (every CTA of the previous solve left its last barrier long ago);
  // this solve's flag was cleared by the previous one (both start at zero): no memset between the solves
  if (blockIdx.x == 0 && threadIdx.x == 0) { *A.flag_other = 0u; A.state[0] = 0; A.state[3] = 0; }
  unsigned epoch = 0, gi = 0;
  const unsigned long long t_begin = A.timing ? global_ns() : 0ull;
  unsigned long long nar = 0, nhalo = 0;          // reductions / exchanges of this solve so far
  unsigned long long t_acc = 0, t_cnt = 0, t_p = 0, t_s = 0, t_xr = 0, t_mark = 0;
  const size_t n = A.n;
  const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  __shared__ double s_D[16];             // this CTA's copy of the dot-product slot table (every CTA computes the same sums)
  double* const D = s_D;
  auto Dl = [&](int k) { return s_D[k]; };
  auto ar_next = [&]() { ArCtx a = A.ar; a.seq = A.ar.seq + (++nar); return a; };
  const bool timer = A.timing && blockIdx.x == 0 && threadIdx.x == 0;

  // r = r0 = B (b - A x), <r,r>, ||B b||^2 (the ghosts of x were exchanged by the host-side launch before)
  unsigned long long t_resid = 0;
  SpmvOp op;
  op.n_tiles = A.n_tiles; op.uniform16 = A.uniform16; op.l2_hint = A.l2_hint; op.tiles = A.tiles; op.rowptr = A.rowptr; op.col = A.col; op.val = A.val; op.scale = A.scale;
  {
    const SpmvRes sr = persist_spmv<NV, KMASK, SPMV_RESID>(op, A.x, A.r, A.b, A.r0, s_raw, s_bar, gi);
    gi = sr.gi;
    double d[2] = {sr.d0, sr.d1};
    grid_allreduce(d[0], d[1], 2, A.partial, A.counter, A.flag, ++epoch, ar_next(), D + D_INIT);
    if (timer) t_resid = global_ns() - t_begin;
  }
  int it = 0;
  for (;; it++) {
    const int rn = it == 0 ? D_INIT : D_XR0 + 2 * ((it - 1) & 1);
    const int ro = it <= 1 ? D_INIT : D_XR0 + 2 * ((it - 2) & 1);
    const double rho = Dl(rn), rr = Dl(rn + (it == 0 ? 0 : 1));
    const double bnorm = sqrt(Dl(D_INIT + 1));
    const double res = sqrt(rr), target = fmax(A.rtol * bnorm, 1e-50);
    double beta = 0.0, omega = 0.0;
    int stop = 0, bad = 0;
    if (!(res == res)) { stop = 1; bad = 1; }
    else if (res <= target) stop = 1;
    else if (it > 0) {
      const double rho_old = Dl(ro);
      const double alpha = rho_old / Dl(D_R0V);
      omega = Dl(D_TT) != 0.0 ? Dl(D_TS) / Dl(D_TT) : 0.0;
      if (omega == 0.0 || rho == 0.0) { stop = 1; bad = 1; }
      beta = (rho / rho_old) * (alpha / omega);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      A.S[S_RES] = res; A.S[S_TARGET] = target; A.S[S_BNORM] = bnorm;
      A.state[1] = it;
      if (stop) { A.state[3] = bad; A.state[0] = 1; }
    }
    if (stop || it >= A.maxits) break;      // same published sums on every CTA and every rank: everybody leaves together
    double* p = (it & 1) ? A.p1 : A.p0;
    const double* p_old = (it & 1) ? A.p0 : A.p1;
    // ---- P
    if (timer) t_mark = global_ns();
    {
      const double* __restrict__ r = A.r;
      const double* __restrict__ v = A.v;
      const double* __restrict__ po = p_old;
      double* __restrict__ pn = p;
      auto pval = [=](size_t j) { return it == 0 ? r[j] : fma(beta, fma(-omega, v[j], po[j]), r[j]); };
      unsigned long long hs = 0;
      if (A.halo.on) { hs = A.halo.seq0 + (++nhalo); persist_halo_send(A.halo, (int)(hs & 1ull), hs, pval); }
      persist_p(n, it, beta, omega, r, v, po, pn);
      if (A.halo.on) persist_halo_recv(A.halo, (int)(hs & 1ull), hs, p);
    }
    grid_barrier(A.counter, A.flag, ++epoch);
    if (timer) t_p += global_ns() - t_mark;
    // ---- S1: v = B A p, <r0, v>
    {
      unsigned long long t0 = 0;
      if (timer) t0 = global_ns();
      const SpmvRes sr = persist_spmv<NV, KMASK, SPMV_DOT_W>(op, p, A.v, A.r0, nullptr, s_raw, s_bar, gi);
      gi = sr.gi;
      double d[2] = {sr.d0, sr.d1};
      grid_allreduce(d[0], d[1], 1, A.partial, A.counter, A.flag, ++epoch, ar_next(), D + D_R0V);
      if (timer) { t_acc += global_ns() - t0; t_cnt++; }
    }
    const double alpha = Dl(rn) / Dl(D_R0V);
    // ---- S
    if (timer) t_mark = global_ns();
    {
      const double* __restrict__ r = A.r;
      const double* __restrict__ v = A.v;
      double* __restrict__ sn = A.s;
      auto sval = [=](size_t j) { return fma(-alpha, v[j], r[j]); };
      unsigned long long hs = 0;
      if (A.halo.on) { hs = A.halo.seq0 + (++nhalo); persist_halo_send(A.halo, (int)(hs & 1ull), hs, sval); }
      persist_s(n, alpha, r, v, sn);
      if (A.halo.on) persist_halo_recv(A.halo, (int)(hs & 1ull), hs, A.s);
    }
    grid_barrier(A.counter, A.flag, ++epoch);
    if (timer) t_s += global_ns() - t_mark;
    // ---- S2: t = B A s, <s,t>, <t,t>
    {
      unsigned long long t0 = 0;
      if (timer) t0 = global_ns();
      const SpmvRes sr = persist_spmv<NV, KMASK, SPMV_DOT_SELF>(op, A.s, A.t, nullptr, nullptr, s_raw, s_bar, gi);
      gi = sr.gi;
      double d[2] = {sr.d0, sr.d1};
      grid_allreduce(d[0], d[1], 2, A.partial, A.counter, A.flag, ++epoch, ar_next(), D + D_TS);
      if (timer) { t_acc += global_ns() - t0; t_cnt++; }
    }
    // ---- XR
    if (timer) t_mark = global_ns();
    {
      const double om = Dl(D_TT) != 0.0 ? Dl(D_TS) / Dl(D_TT) : 0.0;
      const double2 a2 = persist_xr(n, alpha, om, A.x, p, A.s, A.t, A.r, A.r0);
      double acc[2] = {a2.x, a2.y};
      grid_allreduce(acc[0], acc[1], 2, A.partial, A.counter, A.flag, ++epoch, ar_next(), D + D_XR0 + 2 * (it & 1));
    }
    if (timer) t_xr += global_ns() - t_mark;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    A.state[4] = (int)nar;
    A.state[5] = (int)nhalo;
    A.state[6] = 0;
    if (A.ar.mine) { A.state[6] = A.ar.mine->error; A.ar.mine->error = 0; }   // a peer wait that timed out: reported once
    if (A.t_spmv) { A.t_spmv[0] = t_acc; A.t_spmv[1] = t_cnt; A.t_spmv[2] = t_p; A.t_spmv[3] = t_s; A.t_spmv[4] = t_xr; A.t_spmv[5] = A.timing ? global_ns() - t_begin : 0ull; A.t_spmv[6] = t_resid; }
  }
}


template <int NV, unsigned KMASK>
static int persist_launch(rdc_ctx* c, const PersistArgs& A) {
  SolverWork* W = c->work;
  constexpr int SMEM = 2 * SpmvStage<popc_c(KMASK)>::BYTES;
  auto kern = k_bicgstab_persist<NV, KMASK>;
  static int grid_dev[64] = {};   // per device: attributes set, co-resident grid size
  int& grid = grid_dev[c->device & 63];
  if (grid == 0) {
    RDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int per_sm = 0, sms = 0, coop = 0;
    RDC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    RDC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    RDC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RED_THREADS, SMEM));
    if (!coop || per_sm < 1) { c->err = "cooperative launch is not available for the persistent solver"; return RDC_E_CUDA; }
    const int want = c->opt.tma_ctas_per_sm > 0 ? c->opt.tma_ctas_per_sm : (NV == 3 ? 6 : 2);
    grid = sms * (per_sm < want ? per_sm : want);
    if (grid > PERSIST_MAX_GRID) grid = PERSIST_MAX_GRID / sms * sms;
  }
  W->persist_grid = grid;
  void* args[] = {(void*)&A};
  RDC_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)grid), dim3(RED_THREADS), args, (size_t)SMEM, c->stream));
  c->st.kernel_launches++;
  return 0;
}

// describe the exchange for one parity of the sequence number without consuming a sequence number (p2p.cu)
void p2p_fill_halo_args_parity(rdc_ctx* c, HaloArgs* A, int par);

// BiCGStab as ONE cooperative launch (see k_bicgstab_persist).  Needs the TMA tiles; distributed runs need the peer-memory
// transport (the exchanges happen inside the kernel).  Returns 1 when the caller has to use the five-launch version.
// Which BiCGStab runs.  Measured on B200 (tools/iter_probe.py, ADPM, us per iteration persistent / five-launch): 1.3 M tets
// per GPU (the per-rank size of an 8-GPU run of the 10 M-tet mesh, 16 operator tiles per CTA and SpMV) 132 / 145; 2.5 M tets
// (31 tiles) 225 / 225; 5.1 M tets (62 tiles) 420 / 401; 10.1 M tets (122 tiles) 821 / 774 -- the SpMV phases of the
// cooperative kernel run ~6 % below the stand-alone kernel, its barriers cost less than launches only when the phases are
// short.  The persistent kernel also stops in the iteration that converges instead of a few queued launches later.
static bool want_persistent(const rdc_ctx* c) {
  if (c->opt.bicg_persist >= 0) return c->opt.bicg_persist != 0;
  const SolverWork* W = c->work;
  const int grid = 148 * (c->nv == 3 ? 6 : 2);
  return W->n_tiles <= 40 * grid;
}

static int bicgstab_persist_begin(rdc_ctx* c, const double* scale, double rtol, int maxits) {
  SolverWork* W = c->work;
  if (!(W->n_tiles > 0 && c->opt.spmv_tma)) return 1;
  if (c->S.nranks > 1 && !(p2p_on(c) && c->opt.p2p_fused_ar && c->opt.p2p_fused_halo)) return 1;
  int rc = ensure_extra_vectors(c);
  if (rc) return rc;
  if ((rc = ensure_gmres(c, 1))) return rc;   // borrow V for one more vector
  if ((rc = refresh_u_ghosts(c))) return rc;
  PersistArgs A;
  A.n_tiles = W->n_tiles; A.uniform16 = W->tiles_uniform; A.l2_hint = c->opt.l2_evict_first; A.tiles = W->tiles; A.rowptr = c->d_rowptr; A.col = c->d_col; A.val = c->d_val;
  A.scale = scale; A.b = c->d_rhs;
  A.n = (size_t)c->S.n_owned * c->nv;
  A.x = c->d_u; A.r = W->t1; A.r0 = W->t2; A.v = W->t4; A.s = W->hs; A.t = W->V; A.p0 = W->t3; A.p1 = W->hp2;
  A.partial = W->partial; A.counter = W->counter; A.flag = W->flag + (W->n_persist & 1); A.flag_other = W->flag + ((W->n_persist + 1) & 1);
  W->n_persist++;
  A.D = W->h;
  A.S = reinterpret_cast<double*>(W->rep); A.state = reinterpret_cast<int*>(W->rep + 64);
  A.t_spmv = reinterpret_cast<unsigned long long*>(W->rep + 96);
  A.rtol = rtol; A.maxits = maxits;
  A.timing = c->opt.persist_timing;
  A.ar = ArCtx();
  A.halo = PersistHalo();
  P2P* P = c->p2p;
  if (c->S.nranks > 1) {
    A.ar.peer = (P2PHeader* const*)P->d_peer;
    A.ar.mine = (P2PHeader*)P->arena;
    A.ar.me = c->S.rank;
    A.ar.nranks = c->S.nranks;
    A.ar.seq = P->ar_seq;
    if (!c->S.nbr_rank.empty()) {
      A.halo.on = 1;
      p2p_fill_halo_args_parity(c, &A.halo.A[0], 0);
      p2p_fill_halo_args_parity(c, &A.halo.A[1], 1);
      A.halo.send_idx = c->d_send_idx;
      A.halo.hdr = (P2PHeader*)P->arena;
      A.halo.seq0 = P->halo_seq;
      A.halo.nv = c->nv;
    }
  }
  switch (c->model) {
    case RDC_ADPM: rc = persist_launch<3, KM_ADPM>(c, A); break;
    case RDC_RIPF: rc = persist_launch<3, KM_RIPF>(c, A); break;
    case RDC_HCC: rc = persist_launch<3, KM_HCC>(c, A); break;
    case RDC_PIHNA: rc = persist_launch<5, KM_PIHNA>(c, A); break;
    case RDC_SOLID: rc = persist_launch<3, KM_SOLID>(c, A); break;
    default: rc = persist_launch<5, KM_PROTEAS>(c, A); break;
  }
  if (rc) return rc;
  c->u_ghost_fresh = false;
  W->persist_pending = true;
  c->st.bicg_persistent = 1;
  return 0;
}

// wait for the launched solve and read its 192-byte report (ONE copy, one synchronisation)
static int bicgstab_persist_end(rdc_ctx* c, int* its_out, double* res_out) {
  SolverWork* W = c->work;
  W->persist_pending = false;
  P2P* P = c->p2p;
  cudaError_t e = cudaMemcpyAsync(W->h_rep, W->rep, 192, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  const double* scal = reinterpret_cast<const double*>(W->h_rep);
  const int* state = reinterpret_cast<const int*>(W->h_rep + 64);
  const unsigned long long* ts = reinterpret_cast<const unsigned long long*>(W->h_rep + 96);
  if (e != cudaSuccess) {
    if (P && c->S.nranks > 1) { P->ar_seq += 1ull << 20; P->halo_seq += 1ull << 20; }   // a failed solve: never reuse its tags
    c->err = std::string("persistent BiCGStab: ") + cudaGetErrorString(e);
    return RDC_E_CUDA;
  }
  if (P && c->S.nranks > 1) { P->ar_seq += (unsigned long long)state[4]; P->halo_seq += (unsigned long long)state[5]; }
  memcpy(W->h_state, state, sizeof(int) * 8);
  memcpy(W->h_scal, scal, sizeof(double) * 8);
  *its_out = state[1];
  *res_out = scal[S_RES];
  c->st.resnorm0 = scal[S_BNORM];
  c->st.ms_spmv_total = (double)ts[0] * 1e-6;
  c->st.n_spmv = (int)ts[1];
  if (c->opt.trace)
    fprintf(stderr, "[rdc persist rank %d] its %d grid %d: spmv+reduce %.1f us each, P+barrier %.1f, S+barrier %.1f, XR+reduce %.1f us per iteration; "
            "kernel %.1f us (block 0), initial residual phase %.1f us\n",
            c->S.rank, state[1], W->persist_grid, ts[1] ? ts[0] * 1e-3 / ts[1] : 0.0, state[1] ? ts[2] * 1e-3 / state[1] : 0.0,
            state[1] ? ts[3] * 1e-3 / state[1] : 0.0, state[1] ? ts[4] * 1e-3 / state[1] : 0.0, ts[5] * 1e-3, ts[6] * 1e-3);
  if (state[6]) { c->err = "peer-memory exchange timed out (a rank did not arrive); ghost values are stale"; return RDC_E_COMM; }
  if (state[3]) { c->err = "BiCGStab breakdown"; return RDC_E_DIVERGED; }
  return 0;
}

// the two halves for rdc_step (api.cu): between them the caller may queue work that does not need the host (the clamp)
int solver_persist_begin(rdc_ctx* c, int pc, double rtol, int maxits) {
  if (!want_persistent(c) || (pc != RDC_PC_JACOBI && pc != RDC_PC_NONE) || maxits < 0) return 1;
  c->work->n_ev_used = 0;
  return bicgstab_persist_begin(c, pc == RDC_PC_JACOBI ? c->d_dinv : nullptr, rtol, maxits);
}
int solver_persist_end(rdc_ctx* c, int* its, double* res) { return bicgstab_persist_end(c, its, res); }

int solver_solve(rdc_ctx* c, int ksp, int pc, double rtol, int maxits, int restart, int* its, double* res) {
  const double* scale = nullptr;
  if (pc == RDC_PC_JACOBI) {
    scale = c->d_dinv;  // written by the assembly kernel next to the diagonal blocks
  } else if (pc != RDC_PC_NONE) {
    c->err = "preconditioner not implemented on the device path (use RDC_PC_JACOBI or RDC_PC_NONE)";
    return RDC_E_ARG;
  }
  if (restart < 1) restart = 30;
  if (restart > 500) { c->err = "rdc_solve: restart must be <= 500 (the Gram-Schmidt coefficients share a 512-entry buffer)"; return RDC_E_ARG; }
  if (maxits < 0) { c->err = "rdc_solve: maxits must be >= 0"; return RDC_E_ARG; }
  SolverWork* W = c->work;
  W->n_ev_used = 0;
  int rc;
  bool persistent = false;   // the persistent solver times its SpMV phases itself (globaltimer inside the kernel)
  if (ksp == RDC_KSP_GMRES) rc = gmres(c, scale, rtol, maxits, restart, its, res);
  else if (ksp == RDC_KSP_CG) rc = pcg(c, scale, rtol, maxits, its, res);
  else if (ksp == RDC_KSP_BICGSTAB) {
    rc = 1;
    if (want_persistent(c)) {
      rc = bicgstab_persist_begin(c, scale, rtol, maxits);
      persistent = rc != 1;
      if (rc == 0) rc = bicgstab_persist_end(c, its, res);
    }
    if (rc == 1) rc = bicgstab(c, scale, rtol, maxits, its, res);
    if (rc == RDC_E_DIVERGED && *res == *res) {
      // rho or omega vanished (not a NaN): the iterate is still valid, continue with the method that cannot break
      // down this way; all ranks take this branch alike because the flags derive from all-reduced values
      int its2 = 0;
      rc = gmres(c, scale, rtol, maxits, restart, &its2, res);
      *its += its2;
    }
  }
  else { c->err = "unknown ksp"; return RDC_E_ARG; }
  // Every solver ends with a stream synchronisation (the final poll), so the event pairs around the SpMV launches of
  // this solve are complete: sum them now.  Launches issued after convergence return at once (device-side flag) and
  // add ~0, so the mean over the REAL SpMVs is total / (its * spmv per its).
  c->u_ghost_fresh = false;   // the solution changed
  if (!persistent) {
    double tot = 0.0;
    for (int k = 0; k < W->n_ev_used; k++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, W->ev[2 * k], W->ev[2 * k + 1]) == cudaSuccess) tot += ms;
      else cudaGetLastError();
    }
    c->st.ms_spmv_total = tot;
    c->st.n_spmv = (*its) * (ksp == RDC_KSP_BICGSTAB ? 2 : 1);
  }
  return rc;
}

// Cost of the persistent solver's grid barriers on this device: `reps` barriers (mode 0) or reduction-barriers (mode 1)
// back to back in one cooperative launch of 148 x ctas_per_sm CTAs.
__global__ void __launch_bounds__(RED_THREADS) k_barrier_probe(int reps, int mode, double* partial, unsigned* counter, unsigned* flag,
                                                               double* out) {
  unsigned epoch = 0;
  double acc = 0.0;
  for (int k = 0; k < reps; k++) {
    if (mode == 0) grid_barrier(counter, flag, ++epoch);
    else {
      double d[2] = {1.0, (double)threadIdx.x};
      if (mode == 1) { grid_reduce_barrier(d, 2, partial, counter, flag, ++epoch, out, ArCtx()); acc += __ldcg(out); }
      else { __shared__ double s_o[2]; grid_allreduce(d[0], d[1], 2, partial, counter, flag, ++epoch, ArCtx(), s_o); acc += s_o[0]; }
    }
  }
  if (acc < 0.0) out[3] = acc;
}
int launch_barrier_probe(rdc_ctx* c, int reps, int ctas_per_sm, int mode) {
  SolverWork* W = c->work;
  RDC_CUDA(cudaMemsetAsync(W->flag, 0, 2 * sizeof(unsigned), c->stream));
  RDC_CUDA(cudaMemsetAsync(W->counter, 0, sizeof(unsigned), c->stream));
  double* partial = W->partial; unsigned* counter = W->counter; unsigned* flag = W->flag + 1; double* out = W->h + 900;
  void* args[] = {&reps, &mode, &partial, &counter, &flag, &out};
  RDC_CUDA(cudaLaunchCooperativeKernel((void*)k_barrier_probe, dim3(148u * ctas_per_sm), dim3(RED_THREADS), args, 0, c->stream));
  RDC_CUDA(cudaMemsetAsync(W->flag, 0, 2 * sizeof(unsigned), c->stream));   // the solver expects cleared epoch flags
  c->st.kernel_launches++;
  return 0;
}

// fp64 pipe probe: 8 independent DFMA chains per thread, full occupancy -- the rate the assembly kernel's arithmetic is
// measured against (its roofline is the fp64 pipe and the issue slots, not HBM)
__global__ void __launch_bounds__(256) k_dfma_probe(int iters, double seed, double* out) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = seed + k + threadIdx.x * 1e-3;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = fma(a[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 12345.678) out[0] = s;   // never true: keeps the chains alive
}
int launch_dfma_probe(rdc_ctx* c, int iters) {
  SolverWork* W = c->work;
  k_dfma_probe<<<148 * 8, 256, 0, c->stream>>>(iters, 1.0, W->h + 900);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// Read-only streaming probe over the operator values (16-byte loads, fixed-order reduction): the bandwidth a pure
// read of the same array reaches on this GPU, as a reference point next to the SpMV roofline.
__global__ void __launch_bounds__(RED_THREADS) k_stream_read(size_t n2, const double2* __restrict__ a, double* partial,
                                                             unsigned* counter, double* out) {
  double acc[1] = {0.0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(a + i));
    acc[0] += v.x + v.y;
  }
  grid_reduce<1>(acc, 1, partial, counter, out);
}

int launch_stream_probe(rdc_ctx* c, int ctas_per_sm) {
  SolverWork* W = c->work;
  const size_t n2 = (size_t)c->nnzb * c->nkv / 2;
  k_stream_read<<<148 * ctas_per_sm, RED_THREADS, 0, c->stream>>>(n2, (const double2*)c->d_val, W->partial, W->counter, W->h + 900);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rdc
