// pms_kernels.cu — hand-written sm_100a kernels of the mesh-smoothing hot path.
//
// Compiled twice (see pms_kernels.h): -DPMS_NS=pms_fast (FMA contraction) and
// -DPMS_NS=pms_strict -fmad=false (bit-comparable with the CPU oracle).
//
// Data layout: every nodal field is SoA, component c of node n at f[c*stride + n]; structured blocks are
// concatenated (node_off) with i fastest inside a block (the reference's LayoutLeft View<double****>,
// MeshType.hpp:48-82).  Reductions are deterministic: warp shuffle tree -> shared-memory tree -> per-block
// partial -> the last block to finish (atomic ticket) sums the partials in index order.
// No floating-point atomics anywhere: the nodal gradient is a gather in ascending element order.
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "pms_device.cuh"
#include "pms_kernels.h"

#ifndef PMS_NS
#define PMS_NS pms_fast
#endif

namespace PMS_NS {

using namespace pmsk;
using namespace pmsdev;

constexpr int NT = 256;  // threads per block for all kernels
constexpr int TBX = 32, TBY = 4, TBZ = 2;  // structured cell tile per block

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
struct OpSum {
  __device__ static double f(double a, double b) { return a + b; }
  __device__ static double id() { return 0.0; }
};
struct OpMin {
  __device__ static double f(double a, double b) { return a < b ? a : b; }
  __device__ static double id() { return DBL_MAX; }
};
struct OpMax {
  __device__ static double f(double a, double b) { return a > b ? a : b; }
  __device__ static double id() { return 0.0; }
};

__device__ __forceinline__ int lin_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }
__device__ __forceinline__ unsigned lin_bid() { return blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); }
__device__ __forceinline__ unsigned n_blocks() { return gridDim.x * gridDim.y * gridDim.z; }

template <class OP>
__device__ __forceinline__ double block_reduce(double v, double* sm /*[NT/32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = OP::f(v, __shfl_down_sync(0xffffffffu, v, o));
  const int t = lin_tid();
  __syncthreads();
  if ((t & 31) == 0) sm[t >> 5] = v;
  __syncthreads();
  if (t < 32) {
    v = t < NT / 32 ? sm[t] : OP::id();
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v = OP::f(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  return v;  // valid in thread 0
}
__device__ __forceinline__ unsigned long long block_reduce_u(unsigned long long v, unsigned long long* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int t = lin_tid();
  __syncthreads();
  if ((t & 31) == 0) sm[t >> 5] = v;
  __syncthreads();
  if (t < 32) {
    v = t < NT / 32 ? sm[t] : 0ull;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;
}

// Block-level values -> global result.  NV double values reduced with OP, optionally one counter.
template <class OP, int NV, bool WITH_U>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], unsigned long long u, const Reduce& R, int slot) {
  __shared__ double smd[NT / 32];
  __shared__ unsigned long long smu[NT / 32];
  __shared__ bool is_last;
  const int t = lin_tid();
  const unsigned nb = n_blocks(), bid = lin_bid();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double r = block_reduce<OP>(v[k], smd);
    if (t == 0) R.partial_d[(size_t)k * R.max_blocks + bid] = r;
  }
  if (WITH_U) {
    unsigned long long r = block_reduce_u(u, smu);
    if (t == 0) R.partial_u[bid] = r;
  }
  if (t == 0) {
    __threadfence();
    unsigned int tk = atomicAdd(R.ticket, 1u);
    is_last = (tk == nb - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double a = OP::id();
    for (unsigned i = t; i < nb; i += NT) a = OP::f(a, __ldcg(&R.partial_d[(size_t)k * R.max_blocks + i]));
    a = block_reduce<OP>(a, smd);
    if (t == 0) R.out_d[slot + k] = a;
  }
  if (WITH_U) {
    unsigned long long a = 0;
    for (unsigned i = t; i < nb; i += NT) a += __ldcg(&R.partial_u[i]);
    a = block_reduce_u(a, smu);
    if (t == 0) R.out_u[slot] = a;
  }
  if (t == 0) *R.ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// element addressing
// ------------------------------------------------------------------------------------------------
template <int FL>
__device__ __forceinline__ bool element_of_thread(const MeshDev& m, long long& e, long long (&nid)[8]) {
  if (FL == FL_SGRID) {
    const int i = blockIdx.x * TBX + threadIdx.x, j = blockIdx.y * TBY + threadIdx.y, k = blockIdx.z * TBZ + threadIdx.z;
    const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
    if (i >= nci || j >= ncj || k >= nck) return false;
    e = i + (long long)nci * (j + (long long)ncj * k);
    const long long n0 = m.node_off + i + (long long)m.ni * (j + (long long)m.nj * k);
    const long long sj = m.ni, sk = (long long)m.ni * m.nj;
#pragma unroll
    for (int a = 0; a < 8; a++) nid[a] = n0 + (a & 1) + ((a >> 1) & 1) * sj + (a >> 2) * sk;  // total_element_metric.cpp:286-302
    return true;
  } else {
    e = (long long)blockIdx.x * NT + threadIdx.x;
    if (e >= m.nelems) return false;
    constexpr int NN = npe<FL>();
#pragma unroll
    for (int a = 0; a < NN; a++) nid[a] = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
    return true;
  }
}

template <int FL>
static dim3 elem_grid(const MeshDev& m, dim3& block) {
  if (FL == FL_SGRID) {
    block = dim3(TBX, TBY, TBZ);
    return dim3((m.ni - 1 + TBX - 1) / TBX, (m.nj - 1 + TBY - 1) / TBY, (m.nk - 1 + TBZ - 1) / TBZ);
  }
  block = dim3(NT, 1, 1);
  return dim3((unsigned)((m.nelems + NT - 1) / NT), 1, 1);
}

// ------------------------------------------------------------------------------------------------
// K2: total element metric at x + alpha*s          (total_element_metric.cpp:271-307, Def.hpp:330-388)
// ------------------------------------------------------------------------------------------------
template <int FL, int STAGE, bool USE_REF, bool WITH_S>
__global__ void __launch_bounds__(NT) k_metric(const MeshDev m, const double* __restrict__ x, const double* __restrict__ s,
                                               const double* __restrict__ w, const long long stride, const double alpha,
                                               const Reduce R, const int slot, double* __restrict__ em,
                                               int32_t* __restrict__ ef) {
  long long e, nid[8];
  double val[1] = {0.0};
  unsigned long long inv = 0;
  if (element_of_thread<FL>(m, e, nid)) {
    constexpr int NN = npe<FL>();
    double vc[8][3], vo[8][3];
    unsigned bits = 0;
    if (WITH_S) bits = m.elem_fixed_bits[m.elem_off + e];
#pragma unroll
    for (int a = 0; a < NN; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double xv = __ldg(&x[r * stride + nid[a]]);
        if (WITH_S) {
          if (!((bits >> a) & 1u)) {
            const double dt = alpha * __ldg(&s[r * stride + nid[a]]);  // update_coordinates.cpp:255-256
            xv += dt;
          }
        }
        vc[a][r] = xv;
        vo[a][r] = __ldg(&w[r * stride + nid[a]]);
      }
    bool valid;
    double mm = element_metric<FL, STAGE, USE_REF>(vc, vo, valid);
    bool counted = true;
    if (FL != FL_SGRID && m.elem_owned) counted = m.elem_owned[e] != 0;
    if (counted) {
      val[0] = mm;
      inv = valid ? 0 : 1;
    }
    if (em) em[m.elem_off + e] = counted ? mm : 0.0;
    if (ef) ef[m.elem_off + e] = counted ? (valid ? 0 : 1) : 0;
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

template <int FL, int STAGE, bool USE_REF>
static void launch_metric2(cudaStream_t st, const MeshDev& m, const double* x, const double* s, const double* w,
                           long long stride, double alpha, const Reduce& R, int slot, double* em, int32_t* ef) {
  dim3 block, grid = elem_grid<FL>(m, block);
  if (s && alpha != 0.0)
    k_metric<FL, STAGE, USE_REF, true><<<grid, block, 0, st>>>(m, x, s, w, stride, alpha, R, slot, em, ef);
  else
    k_metric<FL, STAGE, USE_REF, false><<<grid, block, 0, st>>>(m, x, s, w, stride, alpha, R, slot, em, ef);
}
template <int FL>
static void launch_metric1(cudaStream_t st, const MeshDev& m, int stage, int use_ref, const double* x, const double* s,
                           const double* w, long long stride, double alpha, const Reduce& R, int slot, double* em,
                           int32_t* ef) {
  if (stage == 0)
    launch_metric2<FL, 0, true>(st, m, x, s, w, stride, alpha, R, slot, em, ef);
  else if (use_ref)
    launch_metric2<FL, 1, true>(st, m, x, s, w, stride, alpha, R, slot, em, ef);
  else
    launch_metric2<FL, 1, false>(st, m, x, s, w, stride, alpha, R, slot, em, ef);
}
static void launch_metric(cudaStream_t st, const MeshDev& m, int stage, int use_ref, const double* x, const double* s,
                          const double* w, long long stride, double alpha, const Reduce& R, int slot, double* em,
                          int32_t* ef) {
  if (m.kind == KIND_SGRID)
    launch_metric1<FL_SGRID>(st, m, stage, use_ref, x, s, w, stride, alpha, R, slot, em, ef);
  else if (m.kind == KIND_HEX8)
    launch_metric1<FL_HEX8>(st, m, stage, use_ref, x, s, w, stride, alpha, R, slot, em, ef);
  else
    launch_metric1<FL_TET4>(st, m, stage, use_ref, x, s, w, stride, alpha, R, slot, em, ef);
}

// ------------------------------------------------------------------------------------------------
// K1a: per-element gradient (+ fused metric) -> scratch eg[(a*3+r)*eg_stride + e]
//      (gradient_functors.hpp:349-407 / Def.hpp:1195-1250 without the scatter)
// ------------------------------------------------------------------------------------------------
template <int FL, int STAGE, bool USE_REF>
__global__ void __launch_bounds__(NT) k_grad_elements(const MeshDev m, const double* __restrict__ x,
                                                      const double* __restrict__ w, const long long stride,
                                                      double* __restrict__ eg, const long long eg_stride, const Reduce R,
                                                      const int slot, double* __restrict__ em, int32_t* __restrict__ ef) {
  long long e, nid[8];
  double val[1] = {0.0};
  unsigned long long inv = 0;
  if (element_of_thread<FL>(m, e, nid)) {
    constexpr int NN = npe<FL>();
    double vc[8][3], vo[8][3], grad[8][3];
#pragma unroll
    for (int a = 0; a < NN; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        vc[a][r] = __ldg(&x[r * stride + nid[a]]);
        vo[a][r] = __ldg(&w[r * stride + nid[a]]);
      }
    bool valid;
    double mm = element_grad_metric<FL, STAGE, USE_REF>(vc, vo, valid, grad);
    const bool use = valid || STAGE == 0;  // gradient_functors.hpp:391, Def.hpp:1219
    const long long ge = m.elem_off + e;
#pragma unroll
    for (int a = 0; a < NN; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) eg[(long long)(a * 3 + r) * eg_stride + ge] = use ? grad[a][r] : 0.0;
    bool counted = true;
    if (FL != FL_SGRID && m.elem_owned) counted = m.elem_owned[e] != 0;
    if (counted) {
      val[0] = mm;
      inv = valid ? 0 : 1;
    }
    if (em) em[ge] = counted ? mm : 0.0;
    if (ef) ef[ge] = counted ? (valid ? 0 : 1) : 0;
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

template <int FL>
static void launch_grad_elements1(cudaStream_t st, const MeshDev& m, int stage, int use_ref, const double* x,
                                  const double* w, long long stride, double* eg, long long egs, const Reduce& R, int slot,
                                  double* em, int32_t* ef) {
  dim3 block, grid = elem_grid<FL>(m, block);
  if (stage == 0)
    k_grad_elements<FL, 0, true><<<grid, block, 0, st>>>(m, x, w, stride, eg, egs, R, slot, em, ef);
  else if (use_ref)
    k_grad_elements<FL, 1, true><<<grid, block, 0, st>>>(m, x, w, stride, eg, egs, R, slot, em, ef);
  else
    k_grad_elements<FL, 1, false><<<grid, block, 0, st>>>(m, x, w, stride, eg, egs, R, slot, em, ef);
}
static void launch_grad_elements(cudaStream_t st, const MeshDev& m, int stage, int use_ref, const double* x,
                                 const double* w, long long stride, double* eg, long long egs, const Reduce& R, int slot,
                                 double* em, int32_t* ef) {
  if (m.kind == KIND_SGRID)
    launch_grad_elements1<FL_SGRID>(st, m, stage, use_ref, x, w, stride, eg, egs, R, slot, em, ef);
  else if (m.kind == KIND_HEX8)
    launch_grad_elements1<FL_HEX8>(st, m, stage, use_ref, x, w, stride, eg, egs, R, slot, em, ef);
  else
    launch_grad_elements1<FL_TET4>(st, m, stage, use_ref, x, w, stride, eg, egs, R, slot, em, ef);
}

// ------------------------------------------------------------------------------------------------
// K1b: node gather in ascending element order (replaces Kokkos::atomic_add, gradient_functors.hpp:400-402)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_grad_gather_sgrid(const MeshDev m, const double* __restrict__ eg,
                                                          const long long egs, const uint8_t* __restrict__ fixed,
                                                          double* __restrict__ g, const long long stride) {
  const int i = blockIdx.x * TBX + threadIdx.x, j = blockIdx.y * TBY + threadIdx.y, k = blockIdx.z * TBZ + threadIdx.z;
  if (i >= m.ni || j >= m.nj || k >= m.nk) return;
  const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
  const long long n = m.node_off + i + (long long)m.ni * (j + (long long)m.nj * k);
  double g0 = 0.0, g1 = 0.0, g2 = 0.0;
  if (!fixed[n]) {  // fixup: gradient_functors.hpp:223-243
#pragma unroll
    for (int dk = 0; dk < 2; dk++)
#pragma unroll
      for (int dj = 0; dj < 2; dj++)
#pragma unroll
        for (int di = 0; di < 2; di++) {
          const int ci = i - 1 + di, cj = j - 1 + dj, ck = k - 1 + dk;
          if (ci < 0 || ci >= nci || cj < 0 || cj >= ncj || ck < 0 || ck >= nck) continue;
          const long long e = m.elem_off + ci + (long long)nci * (cj + (long long)ncj * ck);
          const int a = (1 - di) + 2 * (1 - dj) + 4 * (1 - dk);  // this node's local index in that cell
          g0 += __ldg(&eg[(long long)(a * 3 + 0) * egs + e]);
          g1 += __ldg(&eg[(long long)(a * 3 + 1) * egs + e]);
          g2 += __ldg(&eg[(long long)(a * 3 + 2) * egs + e]);
        }
  }
  g[n] = g0;
  g[stride + n] = g1;
  g[2 * stride + n] = g2;
}

__global__ void __launch_bounds__(NT) k_grad_gather_csr(const MeshDev m, const double* __restrict__ eg, const long long egs,
                                                        const uint8_t* __restrict__ fixed,
                                                        const uint8_t* __restrict__ owned, double* __restrict__ g,
                                                        double* __restrict__ rout, const long long stride) {
  const long long n = (long long)blockIdx.x * NT + threadIdx.x;
  if (n >= m.nnodes) return;
  double g0 = 0.0, g1 = 0.0, g2 = 0.0;
  const bool skip = fixed[n] || (owned && !owned[n]);  // Def.hpp:1226-1244
  if (!skip) {
    const long long p0 = m.csr_ptr[n], p1 = m.csr_ptr[n + 1];
    for (long long p = p0; p < p1; p++) {
      const long long ent = __ldg(&m.csr_ent[p]);
      const long long e = ent >> 3;
      const int a = (int)(ent & 7);
      g0 += __ldg(&eg[(long long)(a * 3 + 0) * egs + e]);
      g1 += __ldg(&eg[(long long)(a * 3 + 1) * egs + e]);
      g2 += __ldg(&eg[(long long)(a * 3 + 2) * egs + e]);
    }
  }
  g[n] = g0;
  g[stride + n] = g1;
  g[2 * stride + n] = g2;
  if (rout) {  // copy_field(cg_r, cg_g), Def.hpp:1508
    rout[n] = g0;
    rout[stride + n] = g1;
    rout[2 * stride + n] = g2;
  }
}

static void launch_grad_gather(cudaStream_t st, const MeshDev& m, const double* eg, long long egs, const uint8_t* fixed,
                               const uint8_t* owned, double* g, double* rout, long long stride) {
  if (m.kind == KIND_SGRID) {
    dim3 block(TBX, TBY, TBZ), grid((m.ni + TBX - 1) / TBX, (m.nj + TBY - 1) / TBY, (m.nk + TBZ - 1) / TBZ);
    k_grad_gather_sgrid<<<grid, block, 0, st>>>(m, eg, egs, fixed, g, stride);
  } else {
    k_grad_gather_csr<<<(unsigned)((m.nnodes + NT - 1) / NT), NT, 0, st>>>(m, eg, egs, fixed, owned, g, rout, stride);
  }
}

static bool launch_grad_fused_sgrid(cudaStream_t, const MeshDev&, int, const double*, const double*, long long,
                                    const uint8_t*, double*, const Reduce&, int) {
  return false;  // not in this build: the two-pass element/gather path is used
}

// ------------------------------------------------------------------------------------------------
// multi-block interface reconciliation: complete sum over all duplicates, ascending block id
// (single-block-equivalent result; cf. gradient_functors.hpp:141-220, get_edge_len_avg.hpp:146-249)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_group_sum(const long long ngroups, const int64_t* __restrict__ gptr,
                                                  const int64_t* __restrict__ gnodes, double* __restrict__ f,
                                                  const int ncomp, const long long stride) {
  const long long q = (long long)blockIdx.x * NT + threadIdx.x;
  if (q >= ngroups) return;
  const long long p0 = gptr[q], p1 = gptr[q + 1];
  for (int c = 0; c < ncomp; c++) {
    double v = f[c * stride + gnodes[p0]];
    for (long long p = p0 + 1; p < p1; p++) v += f[c * stride + gnodes[p]];
    for (long long p = p0; p < p1; p++) f[c * stride + gnodes[p]] = v;
  }
}
static void launch_group_sum(cudaStream_t st, long long ng, const int64_t* gptr, const int64_t* gnodes, double* f, int ncomp,
                             long long stride) {
  if (ng > 0) k_group_sum<<<(unsigned)((ng + NT - 1) / NT), NT, 0, st>>>(ng, gptr, gnodes, f, ncomp, stride);
}

// ------------------------------------------------------------------------------------------------
// BLAS-1   (BlockStructuredGrid.cpp:301-578, PerceptMesh.cpp:4571-4820)
// ------------------------------------------------------------------------------------------------
static inline unsigned blas_blocks(long long n) {
  long long b = (n + NT - 1) / NT;
  const long long cap = 148LL * 16;  // 148 SMs x resident CTAs; grid-stride beyond that
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

__global__ void __launch_bounds__(NT) k_axpby(const double a, const double* __restrict__ x, const double b,
                                              double* __restrict__ y, const long long n, const long long stride) {
  const long long off = (long long)blockIdx.y * stride;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT)
    y[off + i] = a * x[off + i] + b * y[off + i];
}
__global__ void __launch_bounds__(NT) k_axpbypgz(const double a, const double* __restrict__ x, const double b,
                                                 const double* __restrict__ y, const double c, double* __restrict__ z,
                                                 const long long n, const long long stride) {
  const long long off = (long long)blockIdx.y * stride;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT)
    z[off + i] = a * x[off + i] + b * y[off + i] + c * z[off + i];
}
__global__ void __launch_bounds__(NT) k_set(double* __restrict__ x, const double v, const long long n,
                                            const long long stride) {
  const long long off = (long long)blockIdx.y * stride;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) x[off + i] = v;
}
static void launch_axpby(cudaStream_t st, double a, const double* x, double b, double* y, long long n, long long stride,
                         int nc) {
  k_axpby<<<dim3(blas_blocks(n), nc), NT, 0, st>>>(a, x, b, y, n, stride);
}
static void launch_axpbypgz(cudaStream_t st, double a, const double* x, double b, const double* y, double c, double* z,
                            long long n, long long stride, int nc) {
  k_axpbypgz<<<dim3(blas_blocks(n), nc), NT, 0, st>>>(a, x, b, y, c, z, n, stride);
}
static void launch_set(cudaStream_t st, double* x, double v, long long n, long long stride, int nc) {
  k_set<<<dim3(blas_blocks(n), nc), NT, 0, st>>>(x, v, n, stride);
}

__global__ void __launch_bounds__(NT) k_dot(const double* __restrict__ x, const double* __restrict__ y,
                                            const uint8_t* __restrict__ mask, const long long n, const long long stride,
                                            const int ncomp, const Reduce R, const int slot) {
  double acc[1] = {0.0};
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    if (mask && !mask[i]) continue;
    for (int c = 0; c < ncomp; c++) acc[0] += x[c * stride + i] * y[c * stride + i];
  }
  grid_reduce<OpSum, 1, false>(acc, 0ull, R, slot);
}
static void launch_dot(cudaStream_t st, const double* x, const double* y, const uint8_t* mask, long long n, long long stride,
                       int nc, const Reduce& R, int slot) {
  k_dot<<<blas_blocks(n), NT, 0, st>>>(x, y, mask, n, stride, nc, R, slot);
}

// r = -g ; dold = d.d ; dmid = r.d ; dnew = r.r        (Def.hpp:1034-1037)
__global__ void __launch_bounds__(NT) k_cg_r_dots(const double* __restrict__ g, const double* __restrict__ d,
                                                  double* __restrict__ r, const uint8_t* __restrict__ mask,
                                                  const long long n, const long long stride, const Reduce R,
                                                  const int slot) {
  double acc[3] = {0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const bool cnt = !mask || mask[i];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double gv = g[c * stride + i], dv = d[c * stride + i];
      const double rv = -gv;
      r[c * stride + i] = rv;
      if (cnt) {
        acc[0] += dv * dv;
        acc[1] += rv * dv;
        acc[2] += rv * rv;
      }
    }
  }
  grid_reduce<OpSum, 3, false>(acc, 0ull, R, slot);
}
static void launch_cg_r_dots(cudaStream_t st, const double* g, const double* d, double* r, const uint8_t* mask, long long n,
                             long long stride, const Reduce& R, int slot) {
  k_cg_r_dots<<<blas_blocks(n), NT, 0, st>>>(g, d, r, mask, n, stride, R, slot);
}

// restart ? s = r : s = 1.0*r + beta*s ; d = r         (Def.hpp:1057-1069)
__global__ void __launch_bounds__(NT) k_cg_s_update(const double* __restrict__ r, double* __restrict__ s,
                                                    double* __restrict__ d, const double beta, const int restart,
                                                    const long long n, const long long stride) {
  const long long off = (long long)blockIdx.y * stride;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const double rv = r[off + i];
    s[off + i] = restart ? rv : 1.0 * rv + beta * s[off + i];
    d[off + i] = rv;
  }
}
static void launch_cg_s_update(cudaStream_t st, const double* r, double* s, double* d, double beta, int restart, long long n,
                               long long stride) {
  k_cg_s_update<<<dim3(blas_blocks(n), 3), NT, 0, st>>>(r, s, d, beta, restart, n, stride);
}

// alpha_0 candidates + min         (get_alpha_0_refmesh.hpp:168-193, Def.hpp:864-901)
__global__ void __launch_bounds__(NT) k_alpha0(const double* __restrict__ s, const double* __restrict__ edge,
                                               const uint8_t* __restrict__ skip, const long long n, const long long stride,
                                               const Reduce R, const int slot) {
  double acc[1] = {DBL_MAX};
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    if (skip && skip[i]) continue;
    double sn = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) sn += s[c * stride + i] * s[c * stride + i];
    sn = sqrt(sn);
    if (sn > 0.0) {
      const double cand = edge[i] / sn;
      acc[0] = cand < acc[0] ? cand : acc[0];
    }
  }
  grid_reduce<OpMin, 1, false>(acc, 0ull, R, slot);
}
static void launch_alpha0(cudaStream_t st, const double* s, const double* edge, const uint8_t* skip, long long n,
                          long long stride, const Reduce& R, int slot) {
  k_alpha0<<<blas_blocks(n), NT, 0, st>>>(s, edge, skip, n, stride, R, slot);
}

// lagged = x ; x += alpha*s on unfixed nodes ; dmax, dmax_relative   (update_coordinates.cpp:241-268, 94-118)
__global__ void __launch_bounds__(NT) k_update(double* __restrict__ x, double* __restrict__ lag, const double* __restrict__ s,
                                               const double* __restrict__ edge, const uint8_t* __restrict__ fixed,
                                               const double alpha, const long long n, const long long stride,
                                               const Reduce R, const int slot) {
  double acc[2] = {0.0, 0.0};
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const bool fx = fixed[i] != 0;
    double maxDelta = 0.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double xv = x[c * stride + i];
      lag[c * stride + i] = xv;
      if (!fx) {
        const double dt = alpha * s[c * stride + i];
        x[c * stride + i] = xv + dt;
        const double a = fabs(dt);
        maxDelta = a > maxDelta ? a : maxDelta;
      }
    }
    if (!fx) {
      acc[0] = maxDelta > acc[0] ? maxDelta : acc[0];
      const double rel = maxDelta / edge[i];
      acc[1] = rel > acc[1] ? rel : acc[1];
    }
  }
  grid_reduce<OpMax, 2, false>(acc, 0ull, R, slot);
}
static void launch_update(cudaStream_t st, double* x, double* lag, const double* s, const double* edge, const uint8_t* fixed,
                          double alpha, long long n, long long stride, const Reduce& R, int slot) {
  k_update<<<blas_blocks(n), NT, 0, st>>>(x, lag, s, edge, fixed, alpha, n, stride, R, slot);
}

// ------------------------------------------------------------------------------------------------
// nodal average edge length on the reference mesh      (get_edge_len_avg.hpp:17-144,368-382; Def.hpp:39-65)
// ------------------------------------------------------------------------------------------------
template <int NE>
__device__ __forceinline__ double mean_edge_length(const double (&v)[8][3], const int (&ed)[NE][2]) {
  double ave = 0.0;
#pragma unroll
  for (int ie = 0; ie < NE; ie++) {
    double len = 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) len += (v[ed[ie][0]][d] - v[ed[ie][1]][d]) * (v[ed[ie][0]][d] - v[ed[ie][1]][d]);
    len = sqrt(len);
    ave += len / ((double)NE);
  }
  return ave;
}

__global__ void __launch_bounds__(NT) k_edge_sgrid(const MeshDev m, const double* __restrict__ w, const long long stride,
                                                   double* __restrict__ el, double* __restrict__ na) {
  const int i = blockIdx.x * TBX + threadIdx.x, j = blockIdx.y * TBY + threadIdx.y, k = blockIdx.z * TBZ + threadIdx.z;
  if (i >= m.ni || j >= m.nj || k >= m.nk) return;
  const int edges[12][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}, {4, 5}, {5, 7}, {7, 6}, {6, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
  double nm = 0.0;
  int cnt = 0;
  // sgrid_find_connected_cells: i0 outermost, i2 innermost (get_edge_len_avg.hpp:25-41)
  for (int i0 = i - 1; i0 <= i; ++i0) {
    if (i0 < 0 || i0 > m.ni - 2) continue;
    for (int i1 = j - 1; i1 <= j; ++i1) {
      if (i1 < 0 || i1 > m.nj - 2) continue;
      for (int i2 = k - 1; i2 <= k; ++i2) {
        if (i2 < 0 || i2 > m.nk - 2) continue;
        double v[8][3];
        const long long n0 = m.node_off + i0 + (long long)m.ni * (i1 + (long long)m.nj * i2);
#pragma unroll
        for (int a = 0; a < 8; a++) {
          const long long n = n0 + (a & 1) + ((a >> 1) & 1) * (long long)m.ni + (a >> 2) * (long long)m.ni * m.nj;
#pragma unroll
          for (int d = 0; d < 3; d++) v[a][d] = __ldg(&w[d * stride + n]);
        }
        nm += mean_edge_length<12>(v, edges);
        cnt++;
      }
    }
  }
  const long long n = m.node_off + i + (long long)m.ni * (j + (long long)m.nj * k);
  el[n] = nm;
  na[n] = (double)cnt;
}

template <int FL>
__global__ void __launch_bounds__(NT) k_edge_csr(const MeshDev m, const double* __restrict__ w, const long long stride,
                                                 double* __restrict__ el, double* __restrict__ na) {
  const long long n = (long long)blockIdx.x * NT + threadIdx.x;
  if (n >= m.nnodes) return;
  const int hex_edges[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
  const int tet_edges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  constexpr int NN = npe<FL>();
  double nm = 0.0, nele = 0.0;
  for (long long p = m.csr_ptr[n]; p < m.csr_ptr[n + 1]; p++) {
    const long long e = m.csr_ent[p] >> 3;
    double v[8][3];
#pragma unroll
    for (int a = 0; a < NN; a++) {
      const long long q = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
#pragma unroll
      for (int d = 0; d < 3; d++) v[a][d] = __ldg(&w[d * stride + q]);
    }
    nm += (FL == FL_TET4) ? mean_edge_length<6>(v, tet_edges) : mean_edge_length<12>(v, hex_edges);
    nele += 1.0;
  }
  el[n] = nm / nele;
  na[n] = nele;
}
static void launch_edge_lengths(cudaStream_t st, const MeshDev& m, const double* w, long long stride, double* el, double* na) {
  if (m.kind == KIND_SGRID) {
    dim3 block(TBX, TBY, TBZ), grid((m.ni + TBX - 1) / TBX, (m.nj + TBY - 1) / TBY, (m.nk + TBZ - 1) / TBZ);
    k_edge_sgrid<<<grid, block, 0, st>>>(m, w, stride, el, na);
  } else if (m.kind == KIND_HEX8)
    k_edge_csr<FL_HEX8><<<(unsigned)((m.nnodes + NT - 1) / NT), NT, 0, st>>>(m, w, stride, el, na);
  else
    k_edge_csr<FL_TET4><<<(unsigned)((m.nnodes + NT - 1) / NT), NT, 0, st>>>(m, w, stride, el, na);
}
__global__ void __launch_bounds__(NT) k_divide(double* __restrict__ el, const double* __restrict__ na, const long long n) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) el[i] /= na[i];
}
static void launch_divide(cudaStream_t st, double* el, const double* na, long long n) {
  k_divide<<<blas_blocks(n), NT, 0, st>>>(el, na, n);
}

// ------------------------------------------------------------------------------------------------
// halo pack / unpack, fixed-vertex bit masks
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_pack(const double* __restrict__ f, const long long stride, const int ncomp,
                                             const int32_t* __restrict__ nodes, const long long count,
                                             double* __restrict__ buf) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < count; i += (long long)gridDim.x * NT) {
    const long long n = nodes[i];
    for (int c = 0; c < ncomp; c++) buf[c * count + i] = f[c * stride + n];
  }
}
__global__ void __launch_bounds__(NT) k_unpack(double* __restrict__ f, const long long stride, const int ncomp,
                                               const int32_t* __restrict__ nodes, const long long count,
                                               const double* __restrict__ buf) {
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < count; i += (long long)gridDim.x * NT) {
    const long long n = nodes[i];
    for (int c = 0; c < ncomp; c++) f[c * stride + n] = buf[c * count + i];
  }
}
static void launch_pack(cudaStream_t st, const double* f, long long stride, int nc, const int32_t* nodes, long long count,
                        double* buf) {
  if (count > 0) k_pack<<<blas_blocks(count), NT, 0, st>>>(f, stride, nc, nodes, count, buf);
}
static void launch_unpack(cudaStream_t st, double* f, long long stride, int nc, const int32_t* nodes, long long count,
                          const double* buf) {
  if (count > 0) k_unpack<<<blas_blocks(count), NT, 0, st>>>(f, stride, nc, nodes, count, buf);
}

template <int FL>
__global__ void __launch_bounds__(NT) k_elem_fixed_bits(const MeshDev m, const uint8_t* __restrict__ fixed,
                                                        uint8_t* __restrict__ bits) {
  long long e, nid[8];
  if (!element_of_thread<FL>(m, e, nid)) return;
  unsigned b = 0;
  constexpr int NN = npe<FL>();
#pragma unroll
  for (int a = 0; a < NN; a++) b |= (fixed[nid[a]] ? 1u : 0u) << a;
  bits[m.elem_off + e] = (uint8_t)b;
}
static void launch_elem_fixed_bits(cudaStream_t st, const MeshDev& m, const uint8_t* fixed, uint8_t* bits) {
  dim3 block, grid;
  if (m.kind == KIND_SGRID) {
    grid = elem_grid<FL_SGRID>(m, block);
    k_elem_fixed_bits<FL_SGRID><<<grid, block, 0, st>>>(m, fixed, bits);
  } else if (m.kind == KIND_HEX8) {
    grid = elem_grid<FL_HEX8>(m, block);
    k_elem_fixed_bits<FL_HEX8><<<grid, block, 0, st>>>(m, fixed, bits);
  } else {
    grid = elem_grid<FL_TET4>(m, block);
    k_elem_fixed_bits<FL_TET4><<<grid, block, 0, st>>>(m, fixed, bits);
  }
}

static const Launchers g_launchers = {launch_metric,    launch_grad_elements, launch_grad_gather, launch_grad_fused_sgrid,
                                      launch_group_sum, launch_axpby,         launch_axpbypgz,    launch_set,
                                      launch_dot,       launch_cg_r_dots,     launch_cg_s_update, launch_alpha0,
                                      launch_update,    launch_edge_lengths,  launch_divide,      launch_pack,
                                      launch_unpack,    launch_elem_fixed_bits};

const Launchers* launchers() { return &g_launchers; }

}  // namespace PMS_NS
