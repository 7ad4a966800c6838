// reduce.cu -- the per-region post-step reductions of save_solution on the device.
//
// Replaces the rank-0 serial element loops of adpm.C:690-829 (per parcellation region: element average of the LAST
// element of the region -- the reference assigns, it does not accumulate, adpm.C:780-783 -- and the volume of the
// elements whose nodes are all inside a concentration range), pihna.C:842-976 (four thresholded volumes) and
// ripf.C:777-864 (two volumes, two conditions per node), which today need the all-gathered solution
// (update_global_solution, adpm.C:700) and one dof_indices + elem->volume() call per element.
//
// Every element is counted by exactly one rank (the owner of its first node).  The rank's elements are bucketed by
// region once (stable counting sort on the host, rdc_set_subdomains) and cut into chunks of 256 that never straddle
// a region: one block reduces a chunk with a fixed tree, one block per region adds that region's chunk sums in a
// fixed pattern -- deterministic, no float atomics -- and the ranks' sums are all-reduced.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "rdc_internal.h"

struct RegionWork {
  int n_regions = 0;
  int64_t n_mine = 0;
  int n_chunks = 0;
  int32_t* d_perm = nullptr;        // [n_mine] local element ids bucketed by region, element order inside a region
  int32_t* d_chunk_ptr = nullptr;   // [n_chunks+1] into d_perm
  int32_t* d_rchunk_ptr = nullptr;  // [n_regions+1] chunk range of every region
  int32_t* d_last = nullptr;        // [n_regions] local id of the last element of the region when this rank counts it, else -1
  double* d_partial = nullptr;      // [n_chunks]
  double* d_out = nullptr;          // [n_regions]
  std::vector<double> h_out;
};

namespace rdc {

static constexpr int RCHUNK = 256;
__constant__ FeTable c_fe_red[2];  // [0] TET4, [1] HEX8
__constant__ rdc_range_cond c_cond[8];

__device__ __forceinline__ double block_sum_256(double v) {
  __shared__ double s_w[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; w++) t += s_w[w];
  __syncthreads();
  return t;  // valid in thread 0
}

// J = dx/dxi at quadrature point q of the reference element (FEMap), its determinant
__device__ __forceinline__ double jac_at(const FeTable& T, const double (*X)[3], int q) {
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int n = 0; n < T.nen; n++)
    for (int c = 0; c < 3; c++) {
      J[c][0] += X[n][c] * T.dxi[n][q];
      J[c][1] += X[n][c] * T.deta[n][q];
      J[c][2] += X[n][c] * T.dzeta[n][q];
    }
  return J[0][0] * (J[1][1] * J[2][2] - J[2][1] * J[1][2]) + J[1][0] * (J[2][1] * J[0][2] - J[0][1] * J[2][2]) +
         J[2][0] * (J[0][1] * J[1][2] - J[1][1] * J[0][2]);
}

// [upstream] Tet4::volume(): box product of the edge vectors / 6; HEX8: sum of JxW (2x2x2 Gauss is exact)
__device__ __forceinline__ double elem_volume(int nen, const double (*X)[3]) {
  if (nen == 4) {
    const double a[3] = {X[3][0] - X[0][0], X[3][1] - X[0][1], X[3][2] - X[0][2]};
    const double b[3] = {X[1][0] - X[0][0], X[1][1] - X[0][1], X[1][2] - X[0][2]};
    const double c[3] = {X[2][0] - X[0][0], X[2][1] - X[0][1], X[2][2] - X[0][2]};
    return (a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0])) / 6.;
  }
  const FeTable& T = c_fe_red[1];
  double v = 0.0;
  for (int q = 0; q < T.nqp; q++) v += jac_at(T, X, q) * T.w[q];
  return v;
}

template <int NEN>
__global__ void __launch_bounds__(RCHUNK) k_region_vol(int nv, int ncond, const int32_t* __restrict__ conn,
                                                       const double* __restrict__ xyz4, const double* __restrict__ u,
                                                       const int32_t* __restrict__ perm, const int32_t* __restrict__ chunk_ptr,
                                                       double* __restrict__ partial) {
  const int c = blockIdx.x;
  const int i = chunk_ptr[c] + threadIdx.x;
  double v = 0.0;
  if (i < chunk_ptr[c + 1]) {
    const int e = perm[i];
    int en[NEN];
#pragma unroll
    for (int l = 0; l < NEN; l++) en[l] = conn[(size_t)e * NEN + l];
    bool consider = true;
    for (int l = 0; l < NEN && consider; l++) {
      const double* un = u + (size_t)en[l] * nv;
      for (int k = 0; k < ncond && consider; k++) {
        double s = 0.0;
        bool first = true;
        for (int a = 0; a < nv; a++) {
          const double w = c_cond[k].w[a];
          if (w != 0.0) { s = first ? __dmul_rn(w, un[a]) : __dadd_rn(s, __dmul_rn(w, un[a])); first = false; }
        }
        s = s / c_cond[k].div;
        if (!(s >= c_cond[k].lo && s <= c_cond[k].hi)) consider = false;
      }
    }
    if (consider) {
      double X[NEN][3];
#pragma unroll
      for (int l = 0; l < NEN; l++) {
        X[l][0] = xyz4[(size_t)en[l] * 4]; X[l][1] = xyz4[(size_t)en[l] * 4 + 1]; X[l][2] = xyz4[(size_t)en[l] * 4 + 2];
      }
      v = elem_volume(NEN, X);
    }
  }
  const double t = block_sum_256(v);
  if (threadIdx.x == 0) partial[c] = t;
}

// out[r] = sum of the chunk sums of region r (thread t takes chunks t, t+256, ... in order, then the fixed tree)
__global__ void __launch_bounds__(RCHUNK) k_region_sum(const int32_t* __restrict__ rchunk_ptr, const double* __restrict__ partial,
                                                       double* __restrict__ out) {
  const int r = blockIdx.x;
  double s = 0.0;
  for (int c = rchunk_ptr[r] + threadIdx.x; c < rchunk_ptr[r + 1]; c += RCHUNK) s += partial[c];
  const double t = block_sum_256(s);
  if (threadIdx.x == 0) out[r] = t;
}

// one thread per region: element average of variable `var` in the region's last element (0 when another rank counts it)
__global__ void k_region_last_mean(int n_regions, int etype, int nv, int var, const int32_t* __restrict__ conn,
                                   const double* __restrict__ xyz4, const double* __restrict__ u,
                                   const int32_t* __restrict__ last, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_regions) return;
  const int e = last[r];
  if (e < 0) { out[r] = 0.0; return; }
  const FeTable& T = c_fe_red[etype == RDC_TET4 ? 0 : 1];
  double X[RDC_MAX_NEN][3], U[RDC_MAX_NEN];
  for (int l = 0; l < T.nen; l++) {
    const int n = conn[(size_t)e * T.nen + l];
    X[l][0] = xyz4[(size_t)n * 4]; X[l][1] = xyz4[(size_t)n * 4 + 1]; X[l][2] = xyz4[(size_t)n * 4 + 2];
    U[l] = u[(size_t)n * nv + var];
  }
  double avg = 0.0;
  for (int q = 0; q < T.nqp; q++) {
    double val = 0.0;
    for (int l = 0; l < T.nen; l++) val += T.phi[l][q] * U[l];
    avg += jac_at(T, X, q) * T.w[q] * val;
  }
  out[r] = avg / elem_volume(T.nen, X);
}

void region_free(rdc_ctx* c) {
  RegionWork* R = c->region;
  if (!R) return;
  cudaFree(R->d_perm); cudaFree(R->d_chunk_ptr); cudaFree(R->d_rchunk_ptr); cudaFree(R->d_last); cudaFree(R->d_partial);
  cudaFree(R->d_out);
  delete R;
  c->region = nullptr;
}

int region_setup(rdc_ctx* c, const int32_t* region, int n_regions) {
  region_free(c);
  const HostSetup& S = c->S;
  if (n_regions < 1) { c->err = "rdc_set_subdomains: n_regions must be positive"; return RDC_E_ARG; }
  {  // __constant__ memory is per device: upload on every set-up (a process may drive several GPUs)
    FeTable t[2];
    fe_table_fill(&t[0], RDC_TET4);
    fe_table_fill(&t[1], RDC_HEX8);
    RDC_CUDA(cudaMemcpyToSymbol(c_fe_red, t, sizeof(t)));
  }
  if (region)
    for (int64_t e = 0; e < S.E_glob; e++)
      if (region[e] < 0 || region[e] >= n_regions) { c->err = "rdc_set_subdomains: region id out of range"; return RDC_E_ARG; }
  // first node of every local element decides who counts it (the host copy of the connectivity was released after
  // rdc_create: read it back once -- set-up time, one plain copy)
  std::vector<int32_t> first((size_t)S.E_loc);
  {
    std::vector<int32_t> hconn((size_t)S.E_loc * c->nen);
    RDC_CUDA(cudaMemcpyAsync(hconn.data(), c->d_conn, hconn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    RDC_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t le = 0; le < S.E_loc; le++) first[le] = hconn[(size_t)le * c->nen];
  }
  RegionWork* R = new RegionWork();
  c->region = R;
  R->n_regions = n_regions;
  auto reg_of = [&](int64_t le) { return region ? region[S.elem_glob[le]] : 0; };
  std::vector<uint8_t> counted((size_t)S.E_loc);
  std::vector<int32_t> reg_loc((size_t)S.E_loc);
  for (int64_t le = 0; le < S.E_loc; le++) { counted[le] = first[le] < S.n_owned; reg_loc[le] = reg_of(le); }
  std::vector<int32_t> perm, chunk_ptr, rchunk_ptr;
  bucket_regions(S.E_loc, counted.data(), reg_loc.data(), n_regions, RCHUNK, perm, chunk_ptr, rchunk_ptr);
  R->n_mine = (int64_t)perm.size();
  R->n_chunks = (int)chunk_ptr.size() - 1;
  // the last element (global order) of every region, when it is counted here
  std::vector<int32_t> last((size_t)n_regions, -1);
  {
    std::vector<int64_t> last_glob((size_t)n_regions, -1);
    for (int64_t e = 0; e < S.E_glob; e++) last_glob[region ? region[e] : 0] = e;
    for (int64_t le = 0; le < S.E_loc; le++) {
      const int r = reg_of(le);
      if (S.elem_glob[le] == last_glob[r] && first[le] < S.n_owned) last[r] = (int32_t)le;
    }
  }
  auto up = [&](int32_t** d, const std::vector<int32_t>& h) -> int {
    RDC_CUDA(cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(int32_t)));
    if (!h.empty()) RDC_CUDA(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    return 0;
  };
  int rc;
  if ((rc = up(&R->d_perm, perm)) || (rc = up(&R->d_chunk_ptr, chunk_ptr)) || (rc = up(&R->d_rchunk_ptr, rchunk_ptr)) ||
      (rc = up(&R->d_last, last)))
    return rc;
  RDC_CUDA(cudaMalloc((void**)&R->d_partial, std::max(R->n_chunks, 1) * sizeof(double)));
  RDC_CUDA(cudaMalloc((void**)&R->d_out, (size_t)n_regions * sizeof(double)));
  R->h_out.resize((size_t)n_regions);
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

static int finish(rdc_ctx* c, double* host_out) {
  RegionWork* R = c->region;
  int rc = allreduce_sum(c, R->d_out, R->n_regions);
  if (rc) return rc;
  RDC_CUDA(cudaMemcpyAsync(host_out, R->d_out, (size_t)R->n_regions * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return p2p_check_error(c);   // the ghost exchange / all-reduce above may have timed out on a missing peer
}

int region_volumes(rdc_ctx* c, int ncond, const rdc_range_cond* cond, double* vol) {
  RegionWork* R = c->region;
  if (ncond < 1 || ncond > 8) { c->err = "rdc_region_volumes: between 1 and 8 conditions"; return RDC_E_ARG; }
  for (int k = 0; k < ncond; k++)
    if (cond[k].div == 0.0) { c->err = "rdc_region_volumes: div must not be zero"; return RDC_E_ARG; }
  RDC_CUDA(cudaMemcpyToSymbolAsync(c_cond, cond, sizeof(rdc_range_cond) * ncond, 0, cudaMemcpyHostToDevice, c->stream));
  int rc = halo_exchange(c, c->d_u);   // the conditions look at every node of an element, ghosts included
  if (rc) return rc;
  if (R->n_chunks > 0) {
    if (c->nen == 4) k_region_vol<4><<<R->n_chunks, RCHUNK, 0, c->stream>>>(c->nv, ncond, c->d_conn, c->d_xyz, c->d_u, R->d_perm, R->d_chunk_ptr, R->d_partial);
    else k_region_vol<8><<<R->n_chunks, RCHUNK, 0, c->stream>>>(c->nv, ncond, c->d_conn, c->d_xyz, c->d_u, R->d_perm, R->d_chunk_ptr, R->d_partial);
  }
  k_region_sum<<<R->n_regions, RCHUNK, 0, c->stream>>>(R->d_rchunk_ptr, R->d_partial, R->d_out);
  c->st.kernel_launches += 2;
  RDC_CUDA(cudaGetLastError());
  return finish(c, vol);
}

int region_last_mean(rdc_ctx* c, int var, double* mean) {
  RegionWork* R = c->region;
  if (var < 0 || var >= c->nv) { c->err = "rdc_region_last_mean: no such variable"; return RDC_E_ARG; }
  int rc = halo_exchange(c, c->d_u);
  if (rc) return rc;
  k_region_last_mean<<<(R->n_regions + 127) / 128, 128, 0, c->stream>>>(R->n_regions, c->etype, c->nv, var, c->d_conn, c->d_xyz, c->d_u,
                                                                      R->d_last, R->d_out);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return finish(c, mean);
}

}  // namespace rdc
