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
#include "pms_device_fast.cuh"
#include "pms_kernels.h"

#ifndef PMS_NS
#define PMS_NS pms_fast
#endif

namespace PMS_NS {

using namespace pmsk;
using namespace pmsdev;

#ifdef PMS_STRICT
constexpr bool STRICT = true;   // reference expression order, -fmad=false: bit-comparable with the CPU oracle
#else
constexpr bool STRICT = false;  // flop-reduced hex formulation (pms_device_fast.cuh) + FMA contraction
#endif
// hex elements take the flop-reduced path unless strict, or ShapeB1 without a reference mesh (T = A, rare option)
template <int FL, int STAGE, bool USE_REF>
__host__ __device__ constexpr bool fast_hex() { return !STRICT && FL != FL_TET4 && (STAGE == 0 || USE_REF); }

constexpr int NT = 256;   // threads per block of the simple kernels
constexpr int NTP = 128;  // threads per block of the pipelined / quad-lane element kernels
constexpr int TBX = 32, TBY = 4, TBZ = 2;  // structured cell tile per block

static int sm_count() {  // SMs of the current device (persistent grids are sized from it)
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
  }
  return n;
}

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

// BAR == 0: all threads of the block (__syncthreads).  BAR > 0: the first NTH threads only, on named barrier BAR
// (warp-specialised kernels whose helper warps have already exited).
template <int BAR, int NTH>
__device__ __forceinline__ void red_sync() {
  if (BAR == 0)
    __syncthreads();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(BAR), "r"(NTH) : "memory");
}
template <int BAR, int NTH>
__device__ __forceinline__ int red_threads() { return BAR == 0 ? (int)(blockDim.x * blockDim.y * blockDim.z) : NTH; }

template <class OP, int BAR = 0, int NTH = 0>
__device__ __forceinline__ double block_reduce(double v, double* sm /*[NT/32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = OP::f(v, __shfl_down_sync(0xffffffffu, v, o));
  const int t = lin_tid();
  red_sync<BAR, NTH>();
  if ((t & 31) == 0) sm[t >> 5] = v;
  red_sync<BAR, NTH>();
  if (t < 32) {
    const int nw = red_threads<BAR, NTH>() >> 5;
    v = t < nw ? sm[t] : OP::id();
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v = OP::f(v, __shfl_down_sync(0xffffffffu, v, o));
  }
  return v;  // valid in thread 0
}
template <int BAR = 0, int NTH = 0>
__device__ __forceinline__ unsigned long long block_reduce_u(unsigned long long v, unsigned long long* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int t = lin_tid();
  red_sync<BAR, NTH>();
  if ((t & 31) == 0) sm[t >> 5] = v;
  red_sync<BAR, NTH>();
  if (t < 32) {
    const int nw = red_threads<BAR, NTH>() >> 5;
    v = t < nw ? sm[t] : 0ull;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;
}

// Block-level values -> global result.  NV double values reduced with OP, optionally one counter.
template <class OP, int NV, bool WITH_U, int BAR = 0, int NTH = 0>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], unsigned long long u, const Reduce& R, int slot) {
  __shared__ double smd[NT / 32];
  __shared__ unsigned long long smu[NT / 32];
  __shared__ bool is_last;
  const int t = lin_tid();
  const unsigned nb = n_blocks(), bid = lin_bid();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double r = block_reduce<OP, BAR, NTH>(v[k], smd);
    if (t == 0) R.partial_d[(size_t)k * R.max_blocks + bid] = r;
  }
  if (WITH_U) {
    unsigned long long r = block_reduce_u<BAR, NTH>(u, smu);
    if (t == 0) R.partial_u[bid] = r;
  }
  if (t == 0) {
    __threadfence();
    unsigned int tk = atomicAdd(R.ticket, 1u);
    is_last = (tk == nb - 1);
  }
  red_sync<BAR, NTH>();
  if (!is_last) return;
  __threadfence();
  const int nth = red_threads<BAR, NTH>();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double a = OP::id();
    for (unsigned i = t; i < nb; i += nth) a = OP::f(a, __ldcg(&R.partial_d[(size_t)k * R.max_blocks + i]));
    a = block_reduce<OP, BAR, NTH>(a, smd);
    if (t == 0) R.out_d[slot + k] = a;
  }
  if (WITH_U) {
    unsigned long long a = 0;
    for (unsigned i = t; i < nb; i += nth) a += __ldcg(&R.partial_u[i]);
    a = block_reduce_u<BAR, NTH>(a, smu);
    if (t == 0) R.out_u[slot] = a;
  }
  if (t == 0) *R.ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// element addressing
// ------------------------------------------------------------------------------------------------
// Work is cut into tiles of NT elements (structured: TBX x TBY x TBZ cells); blocks are persistent and stride over
// the tiles, so each thread accumulates many elements and the block reduces once at the end.
template <int FL>
__host__ __device__ inline long long num_tiles(const MeshDev& m) {
  if (FL == FL_SGRID)
    return (long long)((m.ni - 1 + TBX - 1) / TBX) * ((m.nj - 1 + TBY - 1) / TBY) * ((m.nk - 1 + TBZ - 1) / TBZ);
  return (m.nelems + NT - 1) / NT;
}

template <int FL>
__device__ __forceinline__ bool element_of_tile(const MeshDev& m, const long long tile, long long& e, long long (&nid)[8]) {
  if (FL == FL_SGRID) {
    const int nbx = (m.ni - 1 + TBX - 1) / TBX, nby = (m.nj - 1 + TBY - 1) / TBY;
    const int bx = (int)(tile % nbx), by = (int)((tile / nbx) % nby), bz = (int)(tile / ((long long)nbx * nby));
    const int t = threadIdx.x;
    const int i = bx * TBX + (t & (TBX - 1)), j = by * TBY + ((t / TBX) & (TBY - 1)), k = bz * TBZ + t / (TBX * TBY);
    const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
    if (i >= nci || j >= ncj || k >= nck) return false;
    e = i + (long long)nci * (j + (long long)ncj * k);
    const long long n0 = m.node_off + i + (long long)m.ni * (j + (long long)m.nj * k);
    const long long sj = m.ni, sk = (long long)m.ni * m.nj;
#pragma unroll
    for (int a = 0; a < 8; a++) nid[a] = n0 + (a & 1) + ((a >> 1) & 1) * sj + (a >> 2) * sk;  // total_element_metric.cpp:286-302
    return true;
  } else {
    e = tile * NT + threadIdx.x;
    if (e >= m.nelems) return false;
    constexpr int NN = npe<FL>();
#pragma unroll
    for (int a = 0; a < NN; a++) nid[a] = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
    return true;
  }
}

template <int FL>
static unsigned elem_grid(const MeshDev& m) {
  const long long nt = num_tiles<FL>(m), cap = (long long)sm_count() * 8;  // SMs x up to 8 CTAs; tiles beyond that are strided
  return (unsigned)(nt < cap ? (nt < 1 ? 1 : nt) : cap);
}

// position of local vertex a in the register array: Exodus hex8 order is permuted to structured numbering on the
// flop-reduced path (same corner set, BlockStructuredGrid.cpp:20)
template <int FL, bool FAST>
__device__ __forceinline__ constexpr int vpos(int a) { return (FAST && FL == FL_HEX8) ? hex_perm(a) : a; }

// ------------------------------------------------------------------------------------------------
// Corner-parallel hex kernels (default build): 8 lanes per hex, lane c evaluates the corner Jacobian at local node c
// (structured numbering), neighbour vertices and gradient columns travel by warp shuffle (xor 1,2,4).  ~60-100
// registers per thread instead of ~250, so several CTAs per SM keep the FP64 pipe fed.  A CTA handles 32 hexes at a
// time (4 per warp).  Structured blocks: a CTA owns an 8 x 4 column of cells and marches a chunk of k-layers, so the
// per-hex index arithmetic is one pointer increment.
//
// Edges are taken "neighbour minus self" (e_s = v[c^(1<<s)] - v[c]); versus the canonical +axis orientation this flips
// the columns s with bit s of c set, in A and W alike: T = A W^-1 and y are unchanged, both determinants pick up
// p = (-1)^popc(c), and d(mu)/d(e_s) comes out directly in the flipped frame (see pms_device_fast.cuh).
// ------------------------------------------------------------------------------------------------
constexpr int CT_I = 8, CT_J = 4;   // structured tile: 8 x 4 cells per k-layer
constexpr int CT_KCHUNK = 64;       // k-layers marched by one CTA

__device__ __forceinline__ double shfl_xor_d(double v, int lanemask) { return __shfl_xor_sync(0xffffffffu, v, lanemask); }

// One hex corner per lane.  Accumulates the metric into val/inv, writes the per-element gradient to eg.
template <int FL, int STAGE, bool GRAD, bool WITH_S>
__device__ __forceinline__ void corner_body(const MeshDev& m, const bool active, const long long e, const long long node,
                                            const int lane, const int c, const double* __restrict__ x,
                                            const double* __restrict__ s, const double* __restrict__ w,
                                            const long long stride, const double alpha, double* __restrict__ eg,
                                            const long long eg_stride, const uint8_t* __restrict__ fixed,
                                            double* __restrict__ em, int32_t* __restrict__ ef, double& val,
                                            unsigned long long& inv) {
  double xc[3] = {0.0, 0.0, 0.0}, wc[3] = {0.0, 0.0, 0.0};
  if (active) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      double xv = __ldg(&x[r * stride + node]);
      if (WITH_S) {
        if (!fixed[node]) xv += alpha * __ldg(&s[r * stride + node]);  // update_coordinates.cpp:255-256
      }
      xc[r] = xv;
      wc[r] = __ldg(&w[r * stride + node]);
    }
  }
  double a[3][3], ww[3][3];
#pragma unroll
  for (int ax = 0; ax < 3; ax++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      a[ax][r] = shfl_xor_d(xc[r], 1 << ax) - xc[r];
      ww[ax][r] = shfl_xor_d(wc[r], 1 << ax) - wc[r];
    }
  const double p = (__popc(c) & 1) ? -1.0 : 1.0;
  double detA, dM[3][3];
  bool have;
  const double mm = hex_corner_fast<STAGE, GRAD>(a[0], a[1], a[2], ww[0], ww[1], ww[2], p, detA, dM, have);
  const unsigned bal = __ballot_sync(0xffffffffu, detA > 0.0);
  const bool valid = ((bal >> (lane & 24)) & 0xffu) == 0xffu;  // all 8 corners of this hex
  bool counted = active;
  if (FL != FL_SGRID && m.elem_owned && active) counted = m.elem_owned[e] != 0;
  if (counted) {
    val += mm;
    if (c == 0) inv += valid ? 0 : 1;
  }
  if (em) {  // per-element outputs (debug/parity aid): sum the 8 corner terms of the hex
    double t = mm;
    t += shfl_xor_d(t, 1);
    t += shfl_xor_d(t, 2);
    t += shfl_xor_d(t, 4);
    if (active && c == 0) {
      em[m.elem_off + e] = counted ? t : 0.0;
      if (ef) ef[m.elem_off + e] = counted ? (valid ? 0 : 1) : 0;
    }
  }
  if (GRAD) {
    const bool use = have && (valid || STAGE == 0);  // gradient_functors.hpp:391, Def.hpp:1219
    double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int ax = 0; ax < 3; ax++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double mine = use ? dM[r][ax] : 0.0;
        g[r] += shfl_xor_d(mine, 1 << ax) - mine;  // +column from the neighbour's corner, -column of my own
      }
    if (active) {
      const int a_out = (FL == FL_HEX8) ? (c ^ ((c >> 1) & 1)) : c;  // structured corner -> Exodus local node
      const long long ge = m.elem_off + e;
#pragma unroll
      for (int r = 0; r < 3; r++) eg[(long long)(a_out * 3 + r) * eg_stride + ge] = g[r];
    }
  }
}

template <int FL, int STAGE, bool GRAD, bool WITH_S>
__global__ void __launch_bounds__(NT) k_corners(const MeshDev m, const double* __restrict__ x, const double* __restrict__ s,
                                                const double* __restrict__ w, const long long stride, const double alpha,
                                                double* __restrict__ eg, const long long eg_stride,
                                                const uint8_t* __restrict__ fixed, const Reduce R, const int slot,
                                                double* __restrict__ em, int32_t* __restrict__ ef) {
  const int lane = threadIdx.x & 31, c = lane & 7, warp = threadIdx.x >> 5, el = lane >> 3;
  double val[1] = {0.0};
  unsigned long long inv = 0;
  if (FL == FL_SGRID) {
    const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
    const int nbx = (nci + CT_I - 1) / CT_I;
    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    const int i = bx * CT_I + (warp & 1) * 4 + el, j = by * CT_J + (warp >> 1);
    const bool active = i < nci && j < ncj;
    const int k0 = blockIdx.y * CT_KCHUNK, k1 = min(k0 + CT_KCHUNK, nck);
    const long long sk = (long long)m.ni * m.nj, se = (long long)nci * ncj;
    long long node = m.node_off + (i + (c & 1)) + (long long)m.ni * (j + ((c >> 1) & 1)) + sk * (k0 + (c >> 2));
    long long e = i + (long long)nci * j + se * k0;
    for (int k = k0; k < k1; k++, node += sk, e += se)
      corner_body<FL, STAGE, GRAD, WITH_S>(m, active, e, node, lane, c, x, s, w, stride, alpha, eg, eg_stride, fixed, em, ef,
                                           val[0], inv);
  } else {
    const long long ntiles = (m.nelems + 31) / 32;
    const int a = c ^ ((c >> 1) & 1);  // Exodus local node of structured corner c (swap 2<->3, 6<->7)
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long e = tile * 32 + warp * 4 + el;
      const bool active = e < m.nelems;
      long long node = 0;
      if (active) node = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
      corner_body<FL, STAGE, GRAD, WITH_S>(m, active, e, node, lane, c, x, s, w, stride, alpha, eg, eg_stride, fixed, em, ef,
                                           val[0], inv);
    }
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

template <int FL>
static dim3 corner_grid(const MeshDev& m) {
  if (FL == FL_SGRID)
    return dim3((unsigned)(((m.ni - 1 + CT_I - 1) / CT_I) * ((m.nj - 1 + CT_J - 1) / CT_J)),
                (unsigned)((m.nk - 1 + CT_KCHUNK - 1) / CT_KCHUNK), 1);
  const long long nt = (m.nelems + 31) / 32;
  const long long cap = (long long)sm_count() * 16;
  return dim3((unsigned)(nt < cap ? (nt < 1 ? 1 : nt) : cap), 1, 1);
}

// ------------------------------------------------------------------------------------------------
// Quad-lane hex kernels (default build): 4 lanes per hex; lane q = (di,dj) owns the vertical edge of the hex at that
// (i,j) corner, i.e. the two nodes (di,dj,0) and (di,dj,1), and evaluates the two corner Jacobians at those nodes.
// The i- and j-neighbours' vertices come from lanes q^1 and q^2 by warp shuffle, the k-neighbour is the lane's own
// other node.  Compared with one hex per thread this needs ~130 instead of ~250 registers (3-4 CTAs per SM instead
// of 2) and still gives two independent corner chains per thread; compared with 8 lanes per hex it halves the
// shuffle traffic per FP64 instruction.  Structured blocks march in k: the top nodes of one step are the bottom
// nodes of the next, so each step loads one node per lane and shuffles one level.
// ------------------------------------------------------------------------------------------------
constexpr int QT_I = 8, QT_J = 4;  // structured tile per CTA and k-layer: 8 x 4 hexes (one j-row per warp)
constexpr int QT_KCHUNK = 64;

struct QNode {  // one node's current (x) and reference (w) position
  double x[3], w[3];
};

template <bool WITH_S>
__device__ __forceinline__ void qload(QNode& n, const bool active, const long long node, const double* __restrict__ x,
                                      const double* __restrict__ s, const double* __restrict__ w, const long long stride,
                                      const double alpha, const uint8_t* __restrict__ fixed) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
    n.x[r] = 0.0;
    n.w[r] = 0.0;
  }
  if (active) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      double xv = __ldg(&x[r * stride + node]);
      if (WITH_S) {
        if (!fixed[node]) xv += alpha * __ldg(&s[r * stride + node]);  // update_coordinates.cpp:255-256
      }
      n.x[r] = xv;
      n.w[r] = __ldg(&w[r * stride + node]);
    }
  }
}
__device__ __forceinline__ void qshfl(const QNode& me, QNode& out, const int lanemask) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
    out.x[r] = __shfl_xor_sync(0xffffffffu, me.x[r], lanemask);
    out.w[r] = __shfl_xor_sync(0xffffffffu, me.w[r], lanemask);
  }
}

// the two corners of this lane's vertical edge; n0/n1 own nodes, i0/j0/i1/j1 the shuffled neighbours per level
template <int FL, int STAGE, bool GRAD>
__device__ __forceinline__ void quad_body(const MeshDev& m, const bool active, const long long e, const int lane, const int q,
                                          const QNode& n0, const QNode& n1, const QNode& i0, const QNode& j0, const QNode& i1,
                                          const QNode& j1, double* __restrict__ eg, const long long eg_stride,
                                          double* __restrict__ em, int32_t* __restrict__ ef, double& val,
                                          unsigned long long& inv) {
  double a[3][3], ww[3][3], dM0[3][3], dM1[3][3], detA0, detA1;
  bool have0, have1;
  const double p0 = (__popc(q) & 1) ? -1.0 : 1.0;
#pragma unroll
  for (int r = 0; r < 3; r++) {
    a[0][r] = i0.x[r] - n0.x[r];
    a[1][r] = j0.x[r] - n0.x[r];
    a[2][r] = n1.x[r] - n0.x[r];
    ww[0][r] = i0.w[r] - n0.w[r];
    ww[1][r] = j0.w[r] - n0.w[r];
    ww[2][r] = n1.w[r] - n0.w[r];
  }
  const double mm0 = hex_corner_fast<STAGE, GRAD>(a[0], a[1], a[2], ww[0], ww[1], ww[2], p0, detA0, dM0, have0);
#pragma unroll
  for (int r = 0; r < 3; r++) {
    a[0][r] = i1.x[r] - n1.x[r];
    a[1][r] = j1.x[r] - n1.x[r];
    a[2][r] = n0.x[r] - n1.x[r];
    ww[0][r] = i1.w[r] - n1.w[r];
    ww[1][r] = j1.w[r] - n1.w[r];
    ww[2][r] = n0.w[r] - n1.w[r];
  }
  const double mm1 = hex_corner_fast<STAGE, GRAD>(a[0], a[1], a[2], ww[0], ww[1], ww[2], -p0, detA1, dM1, have1);
  const unsigned bal = __ballot_sync(0xffffffffu, detA0 > 0.0 && detA1 > 0.0);
  const bool valid = ((bal >> (lane & 28)) & 0xfu) == 0xfu;  // all 8 corners of this hex
  bool counted = active;
  if (FL != FL_SGRID && m.elem_owned && active) counted = m.elem_owned[e] != 0;
  if (counted) {
    val += mm0 + mm1;
    if (q == 0) inv += valid ? 0 : 1;
  }
  if (em) {
    double t = mm0 + mm1;
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    if (active && q == 0) {
      em[m.elem_off + e] = counted ? t : 0.0;
      if (ef) ef[m.elem_off + e] = counted ? (valid ? 0 : 1) : 0;
    }
  }
  if (GRAD) {
    const bool u0 = have0 && (valid || STAGE == 0), u1 = have1 && (valid || STAGE == 0);  // gradient_functors.hpp:391
    double g0[3], g1[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const double m00 = u0 ? dM0[r][0] : 0.0, m01 = u0 ? dM0[r][1] : 0.0, m02 = u0 ? dM0[r][2] : 0.0;
      const double m10 = u1 ? dM1[r][0] : 0.0, m11 = u1 ? dM1[r][1] : 0.0, m12 = u1 ? dM1[r][2] : 0.0;
      const double ri0 = __shfl_xor_sync(0xffffffffu, m00, 1), rj0 = __shfl_xor_sync(0xffffffffu, m01, 2);
      const double ri1 = __shfl_xor_sync(0xffffffffu, m10, 1), rj1 = __shfl_xor_sync(0xffffffffu, m11, 2);
      g0[r] = ((ri0 + rj0) + m12) - ((m00 + m01) + m02);
      g1[r] = ((ri1 + rj1) + m02) - ((m10 + m11) + m12);
    }
    if (active) {
      const int a0 = (FL == FL_HEX8) ? (q ^ ((q >> 1) & 1)) : q;  // structured corner -> Exodus local node
      const long long ge = m.elem_off + e;
#pragma unroll
      for (int r = 0; r < 3; r++) {
        eg[(long long)(a0 * 3 + r) * eg_stride + ge] = g0[r];
        eg[(long long)((a0 + 4) * 3 + r) * eg_stride + ge] = g1[r];
      }
    }
  }
}

#ifndef QUAD_MINB
#define QUAD_MINB 3
#endif
template <int FL, int STAGE, bool GRAD, bool WITH_S>
__global__ void __launch_bounds__(NTP, QUAD_MINB) k_quads(const MeshDev m, const double* __restrict__ x,
                                                          const double* __restrict__ s, const double* __restrict__ w,
                                                          const long long stride, const double alpha, double* __restrict__ eg,
                                                          const long long eg_stride, const uint8_t* __restrict__ fixed,
                                                          const Reduce R, const int slot, double* __restrict__ em,
                                                          int32_t* __restrict__ ef) {
  const int lane = threadIdx.x & 31, q = lane & 3, warp = threadIdx.x >> 5, hx = lane >> 2;
  double val[1] = {0.0};
  unsigned long long inv = 0;
  if (FL == FL_SGRID) {
    const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
    const int nbx = (nci + QT_I - 1) / QT_I;
    const int bx = blockIdx.x % nbx, by = blockIdx.x / nbx;
    const int i = bx * QT_I + hx, j = by * QT_J + warp;
    const bool active = i < nci && j < ncj;
    const int k0 = blockIdx.y * QT_KCHUNK, k1 = min(k0 + QT_KCHUNK, nck);
    const long long sk = (long long)m.ni * m.nj, se = (long long)nci * ncj;
    long long node = m.node_off + (i + (q & 1)) + (long long)m.ni * (j + (q >> 1)) + sk * k0;
    long long e = i + (long long)nci * j + se * k0;
    QNode n0, n1, i0, j0, i1, j1;
    qload<WITH_S>(n0, active, node, x, s, w, stride, alpha, fixed);
    qshfl(n0, i0, 1);
    qshfl(n0, j0, 2);
    for (int k = k0; k < k1; k++, e += se) {
      node += sk;
      qload<WITH_S>(n1, active, node, x, s, w, stride, alpha, fixed);
      qshfl(n1, i1, 1);
      qshfl(n1, j1, 2);
      quad_body<FL, STAGE, GRAD>(m, active, e, lane, q, n0, n1, i0, j0, i1, j1, eg, eg_stride, em, ef, val[0], inv);
      n0 = n1;
      i0 = i1;
      j0 = j1;
    }
  } else {
    const long long ntiles = (m.nelems + 31) / 32;
    const int a0 = q ^ ((q >> 1) & 1);  // Exodus local node of structured corner q
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long e = tile * 32 + warp * 8 + hx;
      const bool active = e < m.nelems;
      long long nd0 = 0, nd1 = 0;
      if (active) {
        nd0 = __ldg(&m.conn[(long long)a0 * m.conn_stride + e]);
        nd1 = __ldg(&m.conn[(long long)(a0 + 4) * m.conn_stride + e]);
      }
      QNode n0, n1, i0, j0, i1, j1;
      qload<WITH_S>(n0, active, nd0, x, s, w, stride, alpha, fixed);
      qload<WITH_S>(n1, active, nd1, x, s, w, stride, alpha, fixed);
      qshfl(n0, i0, 1);
      qshfl(n0, j0, 2);
      qshfl(n1, i1, 1);
      qshfl(n1, j1, 2);
      quad_body<FL, STAGE, GRAD>(m, active, e, lane, q, n0, n1, i0, j0, i1, j1, eg, eg_stride, em, ef, val[0], inv);
    }
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

template <int FL>
static dim3 quad_grid(const MeshDev& m) {
  if (FL == FL_SGRID)
    return dim3((unsigned)(((m.ni - 1 + QT_I - 1) / QT_I) * ((m.nj - 1 + QT_J - 1) / QT_J)),
                (unsigned)((m.nk - 1 + QT_KCHUNK - 1) / QT_KCHUNK), 1);
  const long long nt = (m.nelems + 31) / 32;
  const long long cap = (long long)sm_count() * 16;
  return dim3((unsigned)(nt < cap ? (nt < 1 ? 1 : nt) : cap), 1, 1);
}

// ------------------------------------------------------------------------------------------------
// Software-pipelined element kernel (default build, all element kinds): one element per thread, persistent CTAs of
// NTP threads.  While a thread evaluates element t, the vertex data of its next element (x, [s], w for 8 or 4 nodes)
// is already in flight: cp.async (LDGSTS) gathers it into a THREAD-PRIVATE column of shared memory
// sm[value][thread] (bank-conflict free, no __syncthreads), and connectivity is fetched one element further ahead.
// Memory latency is thus hidden behind ~2k instructions of FP64 work instead of by occupancy, which the ~170-250
// registers per thread of the hex metric do not allow.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int FL>
__host__ __device__ inline long long num_ptiles(const MeshDev& m) {
  if (FL == FL_SGRID) return (long long)((m.ni - 1 + 31) / 32) * ((m.nj - 1 + 3) / 4) * (m.nk - 1);
  return (m.nelems + NTP - 1) / NTP;
}

// element of this thread in pipelined tile `tile` (structured tile: 32 x 4 cells of one k-layer); structured node ids
template <int FL>
__device__ __forceinline__ bool ptile_element(const MeshDev& m, const long long tile, long long& e, long long& n0) {
  if (FL == FL_SGRID) {
    const int nci = m.ni - 1, ncj = m.nj - 1;
    const int nbx = (nci + 31) / 32, nby = (ncj + 3) / 4;
    const int bx = (int)(tile % nbx);
    const long long rest = tile / nbx;
    const int by = (int)(rest % nby), k = (int)(rest / nby);
    const int i = bx * 32 + (threadIdx.x & 31), j = by * 4 + (threadIdx.x >> 5);
    if (i >= nci || j >= ncj) return false;
    e = i + (long long)nci * (j + (long long)ncj * k);
    n0 = m.node_off + i + (long long)m.ni * (j + (long long)m.nj * k);
    return true;
  } else {
    e = tile * NTP + threadIdx.x;
    n0 = 0;
    return e < m.nelems;
  }
}

#ifndef PIPE_MINB_METRIC
#define PIPE_MINB_METRIC 2
#endif
#ifndef PIPE_MINB_GRAD
#define PIPE_MINB_GRAD 2
#endif
template <int FL, int STAGE, bool USE_REF, bool GRAD, bool WITH_S>
__global__ void __launch_bounds__(NTP, (GRAD ? PIPE_MINB_GRAD : PIPE_MINB_METRIC)) k_elem_pipe(const MeshDev m, const double* __restrict__ x,
                                                   const double* __restrict__ s, const double* __restrict__ w,
                                                   const long long stride, const double alpha, double* __restrict__ eg,
                                                   const long long eg_stride, const Reduce R, const int slot,
                                                   double* __restrict__ em, int32_t* __restrict__ ef) {
  constexpr bool FAST = fast_hex<FL, STAGE, USE_REF>();
  constexpr int NN = npe<FL>();
  // gradient of a hex on the fast path: vertices are read IN PLACE from shared memory (frees 96 registers for the
  // 24 gradient accumulators + interleaved corner chains), so the prefetch targets the other of two buffers
  constexpr bool INPLACE = FAST && GRAD && !WITH_S;
  constexpr bool TETREF = FL == FL_TET4 && USE_REF;  // reference part cached per element (when the mesh carries it)
  constexpr int NVAL = NN * 3 * (WITH_S ? 3 : 2);
  extern __shared__ double sm[];  // [INPLACE ? 2 : 1][NVAL][NTP]
  const int tid = threadIdx.x;
  int buf = 0;
  auto slot_in = [&](int b, int arr, int a, int r) -> double* { return sm + (size_t)b * NVAL * NTP + ((arr * NN + a) * 3 + r) * NTP + tid; };

  double val[1] = {0.0};
  unsigned long long inv = 0;
  const long long ntiles = num_ptiles<FL>(m);

  // --- pipeline state: element being computed (cur), element whose vertex data is in flight (nxt),
  //     element whose connectivity is in flight (unstructured only: nn_*)
  long long e_cur = 0, e_nxt = 0, n0 = 0;
  bool act_cur = false, act_nxt = false;
  unsigned bits_cur = 0, bits_nxt = 0;
  int conn_nn[8];
  bool act_nn = false;
  long long e_nn = 0;
  // ownership flag of the element (multi-GPU: ghost elements are evaluated but not counted), fetched with the
  // connectivity two elements ahead: read at the point of use it cost a full memory latency per element (+11 % at N > 1)
  uint8_t own_nn = 1, own_nxt = 1, own_cur = 1;

  auto load_conn = [&](long long tile) {  // connectivity of the element after next
    act_nn = false;
    if (tile < ntiles) {
      long long dummy;
      act_nn = ptile_element<FL>(m, tile, e_nn, dummy);
      if (act_nn) {
#pragma unroll
        for (int a = 0; a < NN; a++) conn_nn[a] = __ldg(&m.conn[(long long)a * m.conn_stride + e_nn]);
        if (m.elem_owned) own_nn = __ldg(&m.elem_owned[e_nn]);
      }
    }
  };
  auto issue = [&](long long tile) {  // start the gather of the next element's vertex data
    act_nxt = false;
    if (FL == FL_SGRID) {
      if (tile < ntiles) act_nxt = ptile_element<FL>(m, tile, e_nxt, n0);
    } else {
      act_nxt = act_nn;
      e_nxt = e_nn;
      own_nxt = own_nn;
    }
    if (act_nxt) {
      const long long sj = m.ni, sk = (long long)m.ni * m.nj;
#pragma unroll
      for (int a = 0; a < NN; a++) {
        long long nid;
        if (FL == FL_SGRID)
          nid = n0 + (a & 1) + ((a >> 1) & 1) * sj + (a >> 2) * sk;  // total_element_metric.cpp:286-302
        else
          nid = conn_nn[a];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          cp_async8(slot_in(buf, 0, a, r), &x[r * stride + nid]);
          if (!TETREF || !m.tetref) cp_async8(slot_in(buf, 1, a, r), &w[r * stride + nid]);
          if (WITH_S) cp_async8(slot_in(buf, 2, a, r), &s[r * stride + nid]);
        }
      }
      if (TETREF && m.tetref) {  // the element's cached reference part: 10 coalesced values instead of 12 scattered ones
#pragma unroll
        for (int v = 0; v < 10; v++) cp_async8(slot_in(buf, 1, v / 3, v % 3), &m.tetref[(long long)v * m.tetref_stride + m.elem_off + e_nxt]);
      }
      if (WITH_S) bits_nxt = m.elem_fixed_bits[m.elem_off + e_nxt];
    }
    cp_async_commit();
  };

  long long tile = blockIdx.x;
  if (FL != FL_SGRID) load_conn(tile);
  issue(tile);
  if (FL != FL_SGRID) load_conn(tile + gridDim.x);

  for (; tile < ntiles; tile += gridDim.x) {
    e_cur = e_nxt;
    act_cur = act_nxt;
    bits_cur = bits_nxt;
    own_cur = own_nxt;
    cp_async_wait_all();
    const int cur = buf;  // buffer holding the current element
    double vc[8][3], vo[8][3];
    if (!INPLACE && act_cur) {
#pragma unroll
      for (int a = 0; a < NN; a++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
          double xv = *slot_in(cur, 0, a, r);
          if (WITH_S) {
            if (!((bits_cur >> a) & 1u)) xv += alpha * *slot_in(cur, 2, a, r);  // update_coordinates.cpp:255-256
          }
          vc[vpos<FL, FAST>(a)][r] = xv;
          vo[vpos<FL, FAST>(a)][r] = *slot_in(cur, 1, a, r);
        }
    }
    // single buffer: the values are in registers before the slots are re-armed (same-thread program order on the
    // LSU); in-place mode: the next element goes to the other buffer
    asm volatile("" ::: "memory");
    if (INPLACE) buf ^= 1;
    issue(tile + gridDim.x);
    if (FL != FL_SGRID) load_conn(tile + 2 * (long long)gridDim.x);
    if (!act_cur) continue;

    bool valid;
    double mm;
    double grad[8][3];
    if (INPLACE) {
      const SmemVerts<FL == FL_HEX8, NTP> VC(slot_in(cur, 0, 0, 0)), VO(slot_in(cur, 1, 0, 0));
      mm = hex_metric_fast_v<STAGE, GRAD>(VC, VO, valid, grad);
    } else if (FAST && !GRAD && STAGE == 1) {
      mm = hex_shapeb1_metric_ilp(vc, vo, valid);
    } else if (FAST) {
      mm = hex_metric_fast<STAGE, GRAD>(vc, vo, valid, grad);
    } else if (TETREF && m.tetref) {
      mm = tet_metric_cached<STAGE, GRAD>(vc, vo, valid, grad);
    } else if (GRAD) {
      mm = element_grad_metric<FL, STAGE, USE_REF>(vc, vo, valid, grad);
    } else {
      mm = element_metric<FL, STAGE, USE_REF>(vc, vo, valid);
    }
    const long long ge = m.elem_off + e_cur;
    if (GRAD) {
      const bool use = valid || STAGE == 0;  // gradient_functors.hpp:391, Def.hpp:1219
      if (use || !PMS_USE_BRANCH) {  // a branch, not 24 selects: invalid elements are rare where they matter (ShapeB1 stage)
#pragma unroll
        for (int a = 0; a < NN; a++)
#pragma unroll
          for (int r = 0; r < 3; r++) eg[(long long)(a * 3 + r) * eg_stride + ge] = (PMS_USE_BRANCH || use) ? grad[vpos<FL, FAST>(a)][r] : 0.0;
      } else {
#pragma unroll
        for (int a = 0; a < NN; a++)
#pragma unroll
          for (int r = 0; r < 3; r++) eg[(long long)(a * 3 + r) * eg_stride + ge] = 0.0;
      }
    }
    const bool counted = FL == FL_SGRID || own_cur != 0;
    if (counted) {
      val[0] += mm;
      inv += valid ? 0 : 1;
    }
    if (em) em[ge] = counted ? mm : 0.0;
    if (ef) ef[ge] = counted ? (valid ? 0 : 1) : 0;
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

template <int FL, int STAGE, bool USE_REF, bool GRAD, bool WITH_S>
static void launch_elem_pipe(cudaStream_t st, const MeshDev& m, const double* x, const double* s, const double* w,
                             long long stride, double alpha, double* eg, long long egs, const Reduce& R, int slot,
                             double* em, int32_t* ef) {
  constexpr int NV = npe<FL>() * 3 * (WITH_S ? 3 : 2);
  constexpr bool INPLACE = fast_hex<FL, STAGE, USE_REF>() && GRAD && !WITH_S;
  const size_t smem = (size_t)NV * NTP * sizeof(double) * (INPLACE ? 2 : 1);
  auto kern = k_elem_pipe<FL, STAGE, USE_REF, GRAD, WITH_S>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, NTP, smem);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  const long long nt = num_ptiles<FL>(m);
  const long long cap = (long long)sm_count() * blocks_per_sm;
  const unsigned grid = (unsigned)(nt < cap ? (nt < 1 ? 1 : nt) : cap);
  kern<<<grid, NTP, smem, st>>>(m, x, s, w, stride, alpha, eg, egs, R, slot, em, ef);
}

// ------------------------------------------------------------------------------------------------
// Scratch-free hex8 gradient over bricks (default build).  The 192 B/hex element->node scratch of the two-pass scheme
// costs 5x the algorithmic HBM traffic and a 0.6 ms gather.  Here elements are visited in Morton order of their
// reference-mesh centroids, 128 per CTA iteration ("brick", a compact 8x4x4-like clump of ~225 distinct nodes):
//   stage   the brick's DISTINCT nodes (x and reference w) are cp.async'ed once into shared memory, one brick ahead
//           (~11 8-byte copies per thread instead of 48), together with the brick's tables;
//   compute one thread per hex reads its 8 vertices through brick-local indices and evaluates metric + 24 derivatives
//           exactly as k_elem_pipe does, then drops the 24 contributions into a shared-memory exchange column;
//   gather  one barrier later (at the top of the next iteration) thread j sums node j's contributions in ascending
//           (element id, local node) order from an 8-wide ELL row.  Nodes whose adjacent elements all lie in the
//           brick go straight to cg_g / cg_r; brick-surface nodes write one partial, and k_brick_combine adds a node's
//           partials in ascending brick id.
// Deterministic, no atomics, no redundant corner evaluations, one __syncthreads per brick.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}

constexpr int BR_NN = BRICK_MAX_NODES;
constexpr int BR_XS = 6 * BR_NN;            // doubles per vertex staging buffer: [x|w][3][BR_NN]
constexpr int BR_EX = BRICK_EX_ZERO + 4;    // doubles per exchange buffer (+ the zero slot, 16-byte multiple)
constexpr int BR_NGT = 3;                   // gather-table buffers (the warp-specialised kernel stages two bricks ahead)
constexpr size_t BRICK_SMEM = (size_t)2 * (BR_XS + BR_EX) * sizeof(double) + 2 * BRICK_CTAB_BYTES + BR_NGT * BRICK_GTAB_BYTES;

// nodes of one brick <- exchange column (shared memory), thread j takes nodes j, j+128, ...
__device__ __forceinline__ void brick_gather(const double* __restrict__ ex, const char* __restrict__ gt, const int tid,
                                             const BrickDev& B, double* __restrict__ g, double* __restrict__ rout,
                                             const long long stride) {
  const int* T = reinterpret_cast<const int*>(gt);
  const int nn = T[0], novf = T[1], nn4 = (nn + 3) & ~3;
  const int* tnode = T + 4;
  const int* tdst = tnode + nn4;
  const uint4* tell = reinterpret_cast<const uint4*>(tdst + nn4);
  for (int j = tid; j < nn; j += NTP) {
    const uint4 q = tell[j];
    const unsigned qq[4] = {q.x, q.y, q.z, q.w};
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {  // ascending (element id, local node) order; unused entries read the zero slot
      const double* col = ex + ((qq[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
      s0 += col[0];
      s1 += col[1];
      s2 += col[2];
    }
    if (novf > 0) {  // node shared by more than 8 elements of the brick (irregular meshes)
      const uint16_t* ov = reinterpret_cast<const uint16_t*>(tell + nn);
      for (int i = 0; i < novf;) {
        const int jj = ov[i], cnt = ov[i + 1];
        if (jj == j)
          for (int k = 0; k < cnt; k++) {
            const double* col = ex + ov[i + 2 + k];
            s0 += col[0];
            s1 += col[1];
            s2 += col[2];
          }
        i += 2 + cnt;
      }
    }
    const int node = tnode[j], dst = tdst[j];
    if (dst < 0) {
      if (dst == -2) s0 = s1 = s2 = 0.0;  // fixed or not owned (Def.hpp:1226-1244)
      g[node] = s0;
      g[stride + node] = s1;
      g[2 * stride + node] = s2;
      if (rout) {
        rout[node] = s0;
        rout[stride + node] = s1;
        rout[2 * stride + node] = s2;
      }
    } else {
      B.partial[dst] = s0;
      B.partial[B.npartial + dst] = s1;
      B.partial[2 * B.npartial + dst] = s2;
    }
  }
}

// one hex of the brick: vertices through brick-local indices, contributions into the exchange column
template <int STAGE>
__device__ __forceinline__ void brick_element(const MeshDev& m, const double* __restrict__ xb, const char* __restrict__ cb,
                                              double* __restrict__ exw, const int tid, double& val, unsigned long long& inv) {
  const uint4 lc = *reinterpret_cast<const uint4*>(cb + tid * 16);
  const int e_cur = reinterpret_cast<const int*>(cb + 2048)[tid];
  const int off[8] = {(int)(lc.x & 0xffffu), (int)(lc.x >> 16), (int)(lc.y & 0xffffu), (int)(lc.y >> 16),
                      (int)(lc.z & 0xffffu), (int)(lc.z >> 16), (int)(lc.w & 0xffffu), (int)(lc.w >> 16)};
  double grad[8][3];
  bool use = false;
  if (e_cur >= 0) {
    const IdxVerts<BR_NN> VC(xb, off), VO(xb + 3 * BR_NN, off);
    bool valid;
    const double mm = hex_metric_fast_v<STAGE, true>(VC, VO, valid, grad);
    use = valid || STAGE == 0;  // Def.hpp:1219
    bool counted = true;
    if (m.elem_owned) counted = m.elem_owned[e_cur] != 0;
    if (counted) {
      val += mm;
      inv += valid ? 0 : 1;
    }
  }
#pragma unroll
  for (int p = 0; p < 8; p++)
#pragma unroll
    for (int r = 0; r < 3; r++) exw[p * 3 + r] = use ? grad[p][r] : 0.0;
}

template <int STAGE>
__global__ void __launch_bounds__(NTP, 2) k_brick_grad(const MeshDev m, const BrickDev B, const double* __restrict__ x,
                                                       const double* __restrict__ w, const long long stride,
                                                       double* __restrict__ g, double* __restrict__ rout, const Reduce R,
                                                       const int slot) {
  static_assert(NTP == BRICK_ELEMS, "one thread per brick element");
  extern __shared__ double sm[];
  double* const xs = sm;                    // [2][BR_XS]
  double* const exb = sm + 2 * BR_XS;       // [2][BR_EX]
  char* const ctb = reinterpret_cast<char*>(exb + 2 * BR_EX);  // [2][BRICK_CTAB_BYTES]
  char* const gtb = ctb + 2 * BRICK_CTAB_BYTES;                // [BR_NGT][BRICK_GTAB_BYTES] (two used here)
  const int tid = threadIdx.x;
  double val[1] = {0.0};
  unsigned long long inv = 0;
  const long long nbricks = B.nbricks, gstep = gridDim.x;
  if (tid < 8) exb[(tid >> 2) * BR_EX + BRICK_EX_ZERO + (tid & 3)] = 0.0;

  // metadata registers: node ids this thread stages for the next brick, gather-table range of the current one
  int nid[3] = {0, 0, 0}, nn_s = 0, go0 = 0, go1 = 0;
  auto load_meta = [&](long long b_stage, long long b_gt) {
    nn_s = 0;
    go0 = go1 = 0;
    if (b_gt < nbricks) {
      go0 = __ldg(&B.goff[b_gt]);
      go1 = __ldg(&B.goff[b_gt + 1]);
    }
    if (b_stage < nbricks) {
      nn_s = __ldg(&B.bnn[b_stage]);
      const int* tn = reinterpret_cast<const int*>(B.gtab + __ldg(&B.goff[b_stage])) + 4;
#pragma unroll
      for (int q = 0; q < 3; q++) nid[q] = (tid + q * NTP < nn_s) ? __ldg(&tn[tid + q * NTP]) : 0;
    }
  };
  auto issue = [&](long long b_stage, int sbuf, int gbuf) {
    if (b_stage < nbricks) {
      double* xb = xs + (size_t)sbuf * BR_XS;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int j = tid + q * NTP;
        if (j < nn_s) {
#pragma unroll
          for (int r = 0; r < 3; r++) {
            cp_async8(xb + r * BR_NN + j, &x[r * stride + nid[q]]);
            cp_async8(xb + (3 + r) * BR_NN + j, &w[r * stride + nid[q]]);
          }
        }
      }
      char* cb = ctb + (size_t)sbuf * BRICK_CTAB_BYTES;
      const uint4* src = B.ctab + b_stage * (BRICK_CTAB_BYTES / 16);
      cp_async16(cb + tid * 16, src + tid);
      if (tid < 32) cp_async16(cb + 2048 + tid * 16, src + 128 + tid);
    }
    char* gb = gtb + (size_t)gbuf * BRICK_GTAB_BYTES;
    for (int i = go0 + tid; i < go1; i += NTP) cp_async16(gb + (size_t)(i - go0) * 16, &B.gtab[i]);
    cp_async_commit();
  };

  long long brick = blockIdx.x;
  load_meta(brick, nbricks);
  issue(brick, 0, 0);  // vertices + compute table of the first brick (no gather table yet: go0 == go1)
  load_meta(brick + gstep, brick);
  int it = 0;
  for (; brick < nbricks; brick += gstep, it++) {
    const int cur = it & 1;
    cp_async_wait_all();
    __syncthreads();  // staged vertices/tables of this brick and the exchange column of the previous one are complete
    if (it > 0) brick_gather(exb + (size_t)(cur ^ 1) * BR_EX, gtb + (size_t)(cur ^ 1) * BRICK_GTAB_BYTES, tid, B, g, rout, stride);
    issue(brick + gstep, cur ^ 1, cur);  // next brick's vertices; this brick's gather table (needed next iteration)
    load_meta(brick + 2 * gstep, brick + gstep);
    brick_element<STAGE>(m, xs + (size_t)cur * BR_XS, ctb + (size_t)cur * BRICK_CTAB_BYTES,
                         exb + (size_t)cur * BR_EX + tid * BRICK_EX_STRIDE, tid, val[0], inv);
  }
  cp_async_wait_all();
  __syncthreads();
  if (it > 0) brick_gather(exb + (size_t)((it - 1) & 1) * BR_EX, gtb + (size_t)((it - 1) & 1) * BRICK_GTAB_BYTES, tid, B, g, rout, stride);
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

// Warp-specialised form of k_brick_grad (variant 5).  256 threads: warpgroup 0 (128 threads, 216 registers via
// setmaxnreg) only evaluates elements; warpgroup 1 (40 registers) stages the next-but-one brick with cp.async and
// gathers / stores the previous brick, so the FP64 pipe is not left idle during table walks, address arithmetic and
// stores.  Hand-off through named barriers (ids 1-6); two vertex / exchange buffers, three gather-table buffers:
//   STAGED[b&1]  helper -> compute   vertices + compute table of brick b are in shared memory
//   READY[b&1]   compute -> helper   contributions of brick b written, its vertex buffer is free
//   EXFREE[b&1]  helper -> compute   contributions of brick b-2 gathered, exchange buffer reusable
__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
constexpr int WS_REG_COMPUTE = 216, WS_REG_HELPER = 40;  // (216 + 40) * 128 threads = half the register file

template <int STAGE>
__global__ void __launch_bounds__(2 * NTP, 2) k_brick_ws(const MeshDev m, const BrickDev B, const double* __restrict__ x,
                                                         const double* __restrict__ w, const long long stride,
                                                         double* __restrict__ g, double* __restrict__ rout, const Reduce R,
                                                         const int slot) {
  extern __shared__ double sm[];
  double* const xs = sm;                    // [2][BR_XS]
  double* const exb = sm + 2 * BR_XS;       // [2][BR_EX]
  char* const ctb = reinterpret_cast<char*>(exb + 2 * BR_EX);  // [2][BRICK_CTAB_BYTES]
  char* const gtb = ctb + 2 * BRICK_CTAB_BYTES;                // [BR_NGT][BRICK_GTAB_BYTES]
  const int tid = threadIdx.x & (NTP - 1);
  const long long nbricks = B.nbricks, gstep = gridDim.x, first = blockIdx.x;
  const int n = first < nbricks ? (int)((nbricks - first + gstep - 1) / gstep) : 0;  // bricks of this CTA
  enum { STAGED = 1, READY = 3, EXFREE = 5, COMPUTE = 8 };

  if (threadIdx.x >= NTP) {
    // ---------------------------------------------------------------- helper warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_HELPER));
    if (tid < 8) exb[(tid >> 2) * BR_EX + BRICK_EX_ZERO + (tid & 3)] = 0.0;
    // metadata of the brick staged next, fetched one iteration early
    int nn_s = 0, nid[3] = {0, 0, 0}, go0 = 0, go1 = 0;
    auto load_meta = [&](int k) {
      nn_s = 0;
      go0 = go1 = 0;
      if (k < n) {
        const long long b = first + (long long)k * gstep;
        nn_s = __ldg(&B.bnn[b]);
        go0 = __ldg(&B.goff[b]);
        go1 = __ldg(&B.goff[b + 1]);
        const int* tn = reinterpret_cast<const int*>(B.gtab + go0) + 4;
#pragma unroll
        for (int q = 0; q < 3; q++) nid[q] = (tid + q * NTP < nn_s) ? __ldg(&tn[tid + q * NTP]) : 0;
      }
    };
    auto stage = [&](int k, int gbuf) {  // vertices, compute table and gather table of this CTA's k-th brick
      const long long b = first + (long long)k * gstep;
      double* xb = xs + (size_t)(k & 1) * BR_XS;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const int j = tid + q * NTP;
        if (j < nn_s) {
#pragma unroll
          for (int r = 0; r < 3; r++) {
            cp_async8(xb + r * BR_NN + j, &x[r * stride + nid[q]]);
            cp_async8(xb + (3 + r) * BR_NN + j, &w[r * stride + nid[q]]);
          }
        }
      }
      char* cb = ctb + (size_t)(k & 1) * BRICK_CTAB_BYTES;
      const uint4* src = B.ctab + b * (BRICK_CTAB_BYTES / 16);
      cp_async16(cb + tid * 16, src + tid);
      if (tid < 32) cp_async16(cb + 2048 + tid * 16, src + 128 + tid);
      char* gb = gtb + (size_t)gbuf * BRICK_GTAB_BYTES;
      for (int i = go0 + tid; i < go1; i += NTP) cp_async16(gb + (size_t)(i - go0) * 16, &B.gtab[i]);
      cp_async_commit();
    };
    load_meta(0);
    for (int k = 0; k < 2 && k < n; k++) {
      stage(k, k);
      load_meta(k + 1);
      cp_async_wait_all();
      nbar_arrive(STAGED + (k & 1), 2 * NTP);
    }
    int gcur = 0;  // k % 3
    for (int k = 0; k < n; k++) {
      const int cur = k & 1;
      nbar_sync(READY + cur, 2 * NTP);  // contributions of brick k written, its vertex buffer free; gather k-1 done
      const bool more = k + 2 < n;
      if (more) {
        stage(k + 2, gcur == 0 ? 2 : gcur - 1);  // (k + 2) % 3
        load_meta(k + 3);
      }
      brick_gather(exb + (size_t)cur * BR_EX, gtb + (size_t)gcur * BRICK_GTAB_BYTES, tid, B, g, rout, stride);
      if (more) {
        nbar_arrive(EXFREE + cur, 2 * NTP);
        cp_async_wait_all();
        nbar_arrive(STAGED + cur, 2 * NTP);
      }
      gcur = gcur == 2 ? 0 : gcur + 1;
    }
    return;
  }
  // ------------------------------------------------------------------ compute warpgroup
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REG_COMPUTE));
  double val[1] = {0.0};
  unsigned long long inv = 0;
  for (int k = 0; k < n; k++) {
    const int cur = k & 1;
    nbar_sync(STAGED + cur, 2 * NTP);
    // (the exchange column of brick k-2 must have been gathered before it is overwritten: the wait sits inside
    //  brick_element's store phase in program order, i.e. after the FP64 work)
    double cval = 0.0;
    unsigned long long cinv = 0;
    double stash[24];
    brick_element<STAGE>(m, xs + (size_t)cur * BR_XS, ctb + (size_t)cur * BRICK_CTAB_BYTES, stash, tid, cval, cinv);
    if (k >= 2) nbar_sync(EXFREE + cur, 2 * NTP);
    double* exw = exb + (size_t)cur * BR_EX + tid * BRICK_EX_STRIDE;
#pragma unroll
    for (int q = 0; q < 24; q++) exw[q] = stash[q];
    val[0] += cval;
    inv += cinv;
    nbar_arrive(READY + cur, 2 * NTP);
  }
  grid_reduce<OpSum, 1, true, COMPUTE, NTP>(val, inv, R, slot);
}

__global__ void __launch_bounds__(NT) k_brick_combine(const BrickDev B, const uint8_t* __restrict__ fixed,
                                                      const uint8_t* __restrict__ owned, double* __restrict__ g,
                                                      double* __restrict__ rout, const long long stride) {
  const long long i = (long long)blockIdx.x * NT + threadIdx.x;
  if (i >= B.ncomb) return;
  const int node = B.comb_node[i];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  if (!(fixed[node] || (owned && !owned[node]))) {
    for (int p = B.comb_ptr[i]; p < B.comb_ptr[i + 1]; p++) {
      const long long q = B.comb_ent[p];
      s0 += B.partial[q];
      s1 += B.partial[B.npartial + q];
      s2 += B.partial[2 * B.npartial + q];
    }
  }
  g[node] = s0;
  g[stride + node] = s1;
  g[2 * stride + node] = s2;
  if (rout) {
    rout[node] = s0;
    rout[stride + node] = s1;
    rout[2 * stride + node] = s2;
  }
}

int g_grad_corners = 2;  // default-build gradient: 2 = pipelined element pass + gather, 1 = corner-parallel hex,
                         // 0 = simple, 3 = quad-lane, 4 = bricks, 5 = warp-specialised bricks
static bool launch_grad_bricks(cudaStream_t st, const MeshDev& m, const BrickDev& B, int stage, const double* x, const double* w,
                               long long stride, const uint8_t* fixed, const uint8_t* owned, double* g, double* rout,
                               const Reduce& R, int slot, double* em, int32_t* ef) {
  if (STRICT || m.kind != KIND_HEX8 || em || ef) return false;
  const bool ws = g_grad_corners == 5;
  void (*kern)(MeshDev, BrickDev, const double*, const double*, long long, double*, double*, Reduce, int) =
      ws ? (stage == 0 ? k_brick_ws<0> : k_brick_ws<1>) : (stage == 0 ? k_brick_grad<0> : k_brick_grad<1>);
  const int nthreads = ws ? 2 * NTP : NTP;
  static int blocks_per_sm[2][2] = {{0, 0}, {0, 0}};
  int& bps = blocks_per_sm[ws][stage];
  if (!bps) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BRICK_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, nthreads, BRICK_SMEM);
    if (bps < 1) bps = 1;
  }
  const long long cap = (long long)sm_count() * bps;
  const unsigned grid = (unsigned)(B.nbricks < cap ? (B.nbricks < 1 ? 1 : B.nbricks) : cap);
  kern<<<grid, nthreads, BRICK_SMEM, st>>>(m, B, x, w, stride, g, rout, R, slot);
  if (B.ncomb > 0) k_brick_combine<<<(unsigned)((B.ncomb + NT - 1) / NT), NT, 0, st>>>(B, fixed, owned, g, rout, stride);
  return true;
}

// ------------------------------------------------------------------------------------------------
// Occupancy-oriented hex metric (metric variant 4).  k_elem_pipe hides memory latency by software pipelining at
// 2 warps per scheduler and keeps all 48 vertex doubles in registers (~250).  Here the vertices stay in the thread's
// shared-memory column and are read in place (immediate-offset LDS), x + alpha*s is formed in place, and the register
// budget is capped so that OCC_MINB CTAs (OCC_MINB warps per scheduler) are resident: latency is hidden by other CTAs
// instead of by a prefetch, and the FP64 pipe sees more independent instruction streams.
// ------------------------------------------------------------------------------------------------
#ifndef OCC_MINB
#define OCC_MINB 3
#endif
template <int FL, int STAGE, bool WITH_S>
__global__ void __launch_bounds__(NTP, OCC_MINB) k_metric_occ(const MeshDev m, const double* __restrict__ x, const double* __restrict__ s,
                                                              const double* __restrict__ w, const long long stride,
                                                              const double alpha, const Reduce R, const int slot) {
  extern __shared__ double sm[];  // [8 * 3 * (WITH_S ? 3 : 2)][NTP]
  const int tid = threadIdx.x;
  auto slot_in = [&](int arr, int a, int r) -> double* { return sm + ((arr * 8 + a) * 3 + r) * NTP + tid; };
  double val[1] = {0.0};
  unsigned long long inv = 0;
  const long long ntiles = num_ptiles<FL>(m);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long e, n0;
    if (!ptile_element<FL>(m, tile, e, n0)) continue;
    const long long sj = m.ni, sk = (long long)m.ni * m.nj;
#pragma unroll
    for (int a = 0; a < 8; a++) {
      long long nid;
      if (FL == FL_SGRID)
        nid = n0 + (a & 1) + ((a >> 1) & 1) * sj + (a >> 2) * sk;
      else
        nid = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        cp_async8(slot_in(0, a, r), &x[r * stride + nid]);
        cp_async8(slot_in(1, a, r), &w[r * stride + nid]);
        if (WITH_S) cp_async8(slot_in(2, a, r), &s[r * stride + nid]);
      }
    }
    unsigned bits = 0;
    if (WITH_S) bits = m.elem_fixed_bits[m.elem_off + e];
    cp_async_commit();
    cp_async_wait_all();
    if (WITH_S) {
#pragma unroll
      for (int a = 0; a < 8; a++)
        if (!((bits >> a) & 1u)) {
#pragma unroll
          for (int r = 0; r < 3; r++) *slot_in(0, a, r) += alpha * *slot_in(2, a, r);  // update_coordinates.cpp:255-256
        }
    }
    const SmemVerts<FL == FL_HEX8, NTP> VC(slot_in(0, 0, 0)), VO(slot_in(1, 0, 0));
    bool valid;
    double grad[8][3];
    const double mm = hex_metric_fast_v<STAGE, false>(VC, VO, valid, grad);
    bool counted = true;
    if (FL != FL_SGRID && m.elem_owned) counted = m.elem_owned[e] != 0;
    if (counted) {
      val[0] += mm;
      inv += valid ? 0 : 1;
    }
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

template <int FL, int STAGE, bool WITH_S>
static void launch_metric_occ(cudaStream_t st, const MeshDev& m, const double* x, const double* s, const double* w, long long stride,
                              double alpha, const Reduce& R, int slot) {
  const size_t smem = (size_t)8 * 3 * (WITH_S ? 3 : 2) * NTP * sizeof(double);
  auto kern = k_metric_occ<FL, STAGE, WITH_S>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, NTP, smem);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  const long long nt = num_ptiles<FL>(m);
  const long long cap = (long long)sm_count() * blocks_per_sm;
  kern<<<(unsigned)(nt < cap ? (nt < 1 ? 1 : nt) : cap), NTP, smem, st>>>(m, x, s, w, stride, alpha, R, slot);
}

int g_sgrid_tile = 1;  // structured blocks: 1 = TMA-staged tile kernels for the metric passes, 0 = per-thread pipelined kernels
#include "pms_sgrid_tile.cuh"

// ------------------------------------------------------------------------------------------------
// Multi-alpha hex metric: the line search's trial sequence alpha_init * tau^k is known before its outcomes are, so one
// pass evaluates NB of them (hex_metric_multi: per-corner polynomial coefficients once, ~20 FP64 per extra alpha).
// Vertices of x, s, w stay in the thread's shared-memory column (in place, single buffer, 2 CTAs/SM interleave).
// ------------------------------------------------------------------------------------------------
#ifndef MULTI_MINB
#define MULTI_MINB 2
#endif
#ifndef MULTI_CORNER_UNROLL_BIG
#define MULTI_CORNER_UNROLL_BIG 8  // corner-loop unroll of the 12- and 16-alpha kernels (fewer: no spills, run-time vertex addressing)
#endif
template <int NB>
struct AlphaPack {
  double a[NB];
};
template <int FL, int STAGE, int NB, int NF>
__global__ void __launch_bounds__(NTP, MULTI_MINB) k_metric_multi(const MeshDev m, const double* __restrict__ x, const double* __restrict__ s,
                                                         const double* __restrict__ w, const long long stride,
                                                         const AlphaPack<NB> al, const Reduce R, const int slot) {
  extern __shared__ double sm[];  // [72][NTP]
  const int tid = threadIdx.x;
  auto slot_in = [&](int arr, int a, int r) -> double* { return sm + ((arr * 8 + a) * 3 + r) * NTP + tid; };
  double sums[NB];
  unsigned cnt[NB];
#pragma unroll
  for (int k = 0; k < NB; k++) {
    sums[k] = 0.0;
    cnt[k] = 0;
  }
  const long long ntiles = num_ptiles<FL>(m);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long e, n0;
    if (!ptile_element<FL>(m, tile, e, n0)) continue;
    const long long sj = m.ni, sk = (long long)m.ni * m.nj;
#pragma unroll
    for (int a = 0; a < 8; a++) {
      long long nid;
      if (FL == FL_SGRID)
        nid = n0 + (a & 1) + ((a >> 1) & 1) * sj + (a >> 2) * sk;
      else
        nid = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        cp_async8(slot_in(0, a, r), &x[r * stride + nid]);
        cp_async8(slot_in(1, a, r), &w[r * stride + nid]);
        cp_async8(slot_in(2, a, r), &s[r * stride + nid]);
      }
    }
    const unsigned bits = m.elem_fixed_bits[m.elem_off + e];
    bool counted = true;  // (loaded here, beside the vertex gather, not at its use after the evaluation)
    if (FL != FL_SGRID && m.elem_owned) counted = __ldg(&m.elem_owned[e]) != 0;
    cp_async_commit();
    cp_async_wait_all();
#pragma unroll
    for (int a = 0; a < 8; a++)
      if ((bits >> a) & 1u) {  // fixed nodes do not move (update_coordinates.cpp:255-256)
#pragma unroll
        for (int r = 0; r < 3; r++) *slot_in(2, a, r) = 0.0;
      }
    // (measured alternatives, both slower: reloading accessors SmemVertsReload without spills at 2 CTAs/SM 4.30 ms, or at
    //  168 registers / 3 CTAs/SM 4.61 ms, vs 3.63 ms for 8 alphas at 256^3 with the values kept in registers)
    const SmemVerts<FL == FL_HEX8, NTP> VC(slot_in(0, 0, 0)), VO(slot_in(1, 0, 0)), VS(slot_in(2, 0, 0));
    // accumulated straight into the thread's running sums (a separate per-element array costs 2 NB more live registers in
    // a kernel that already spills: 416 -> 16 B at 12 alphas, 4.79 -> 4.24 ms at 256^3; the structured tile form measured
    // equal either way and keeps the per-element array); ghost elements (multi-GPU aura) are not counted, so they are not evaluated either
    if (counted) {
      unsigned invmask;
      hex_metric_multi<STAGE, NB, (NB >= 12 ? MULTI_CORNER_UNROLL_BIG : 8), NF>(VC, VS, VO, al.a, sums, invmask);
#pragma unroll
      for (int k = 0; k < NB; k++) cnt[k] += (invmask >> k) & 1u;
    }
  }
  double v[2 * NB];
#pragma unroll
  for (int k = 0; k < NB; k++) {
    v[k] = sums[k];
    v[NB + k] = (double)cnt[k];
  }
  grid_reduce<OpSum, 2 * NB, false>(v, 0ull, R, slot);
}

template <int FL, int STAGE, int NB, int NF>
static void launch_metric_multi2(cudaStream_t st, const MeshDev& m, const double* x, const double* s, const double* w, long long stride,
                                 const double* alphas, const Reduce& R, int slot) {
  static_assert(2 * NB <= REDUCE_MAX_NV, "reduction scratch too small");
  const size_t smem = (size_t)72 * NTP * sizeof(double);
  auto kern = k_metric_multi<FL, STAGE, NB, NF>;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, NTP, smem);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  if (!x) return;  // prepare only (prepare_line_search_kernels): the attribute / occupancy calls have loaded the kernel
  AlphaPack<NB> al;
  for (int k = 0; k < NB; k++) al.a[k] = alphas[k];
  const long long nt = num_ptiles<FL>(m);
  long long cap = (long long)sm_count() * blocks_per_sm;
  if (cap > R.max_blocks) cap = R.max_blocks;
  kern<<<(unsigned)(nt < cap ? (nt < 1 ? 1 : nt) : cap), NTP, smem, st>>>(m, x, s, w, stride, al, R, slot);
}
// nf = step lengths (the last ones of the batch) that get their metric: all of them in the untangle stage and on the
// generic structured path (the tile kernel is the structured ShapeB1 path); the smallest even count >= nf from 4 in ShapeB1
template <int FL, int NB, int NF>
static void launch_metric_multi_nf(cudaStream_t st, const MeshDev& m, int nf, const double* x, const double* s, const double* w,
                                   long long stride, const double* alphas, const Reduce& R, int slot) {
  if constexpr (NF >= NB) {
    launch_metric_multi2<FL, 1, NB, NB>(st, m, x, s, w, stride, alphas, R, slot);
  } else {
    if (nf <= NF)
      launch_metric_multi2<FL, 1, NB, NF>(st, m, x, s, w, stride, alphas, R, slot);
    else
      launch_metric_multi_nf<FL, NB, NF + 2>(st, m, nf, x, s, w, stride, alphas, R, slot);
  }
}
template <int FL, int NB>
static void launch_metric_multi1(cudaStream_t st, const MeshDev& m, int stage, int nf, const double* x, const double* s, const double* w,
                                 long long stride, const double* alphas, const Reduce& R, int slot) {
  if (stage == 0)
    launch_metric_multi2<FL, 0, NB, NB>(st, m, x, s, w, stride, alphas, R, slot);
  else if (FL == FL_SGRID)
    launch_metric_multi2<FL, 1, NB, NB>(st, m, x, s, w, stride, alphas, R, slot);
  else
    launch_metric_multi_nf<FL, NB, 4>(st, m, nf, x, s, w, stride, alphas, R, slot);
}
static bool launch_metric_multi(cudaStream_t st, const MeshDev& m, int stage, int use_ref, int nb, int nf, const double* x, const double* s,
                                const double* w, long long stride, const double* alphas, const Reduce& R, int slot,
                                const uint8_t* fixed) {
  if (STRICT || m.kind == KIND_TET4 || (stage != 0 && !use_ref) || (nb < 8 || nb > 16 || (nb & 1))) return false;  // kernels for 8, 10, 12, 14, 16
  if (nf < 1 || nf > nb) nf = nb;
  // structured blocks: TMA-staged tile kernel (pms_sgrid_tile.cuh)
  // (ShapeB1 only: the cheap untangle polynomial runs faster in the per-thread pipelined kernel, 1.73 vs 1.97 ms)
  if (m.kind == KIND_SGRID && stage == 1 && g_sgrid_tile && launch_sgrid_tile_multi(st, m, stage, nb, nf, x, s, w, stride, fixed, alphas, R, slot))
    return true;
  if (m.kind == KIND_SGRID) {
    switch (nb) {
      case 8: launch_metric_multi1<FL_SGRID, 8>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      case 10: launch_metric_multi1<FL_SGRID, 10>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      case 12: launch_metric_multi1<FL_SGRID, 12>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      case 14: launch_metric_multi1<FL_SGRID, 14>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      default: launch_metric_multi1<FL_SGRID, 16>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
    }
  } else {
    switch (nb) {
      case 8: launch_metric_multi1<FL_HEX8, 8>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      case 10: launch_metric_multi1<FL_HEX8, 10>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      case 12: launch_metric_multi1<FL_HEX8, 12>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      case 14: launch_metric_multi1<FL_HEX8, 14>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
      default: launch_metric_multi1<FL_HEX8, 16>(st, m, stage, nf, x, s, w, stride, alphas, R, slot); break;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// K2: total element metric at x + alpha*s          (total_element_metric.cpp:271-307, Def.hpp:330-388)
// ------------------------------------------------------------------------------------------------
template <int FL, int STAGE, bool USE_REF, bool WITH_S>
__global__ void __launch_bounds__(NT) k_metric(const MeshDev m, const double* __restrict__ x, const double* __restrict__ s,
                                               const double* __restrict__ w, const long long stride, const double alpha,
                                               const Reduce R, const int slot, double* __restrict__ em,
                                               int32_t* __restrict__ ef) {
  constexpr bool FAST = fast_hex<FL, STAGE, USE_REF>();
  constexpr int NN = npe<FL>();
  double val[1] = {0.0};
  unsigned long long inv = 0;
  const long long ntiles = num_tiles<FL>(m);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long e, nid[8];
    if (!element_of_tile<FL>(m, tile, e, nid)) continue;
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
        vc[vpos<FL, FAST>(a)][r] = xv;
        vo[vpos<FL, FAST>(a)][r] = __ldg(&w[r * stride + nid[a]]);
      }
    bool valid;
    double mm;
    if (FAST) {
      double dummy[8][3];
      mm = hex_metric_fast<STAGE, false>(vc, vo, valid, dummy);
    } else {
      mm = element_metric<FL, STAGE, USE_REF>(vc, vo, valid);
    }
    bool counted = true;
    if (FL != FL_SGRID && m.elem_owned) counted = m.elem_owned[e] != 0;
    if (counted) {
      val[0] += mm;
      inv += valid ? 0 : 1;
    }
    if (em) em[m.elem_off + e] = counted ? mm : 0.0;
    if (ef) ef[m.elem_off + e] = counted ? (valid ? 0 : 1) : 0;
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

int g_metric_corners = 2;  // default-build metric kernel: 2 = pipelined element kernel, 1 = corner-parallel hex, 0 = simple
template <int FL, int STAGE, bool USE_REF>
static void launch_metric2(cudaStream_t st, const MeshDev& m, const double* x, const double* s, const double* w,
                           long long stride, double alpha, const Reduce& R, int slot, double* em, int32_t* ef,
                           const uint8_t* fixed) {
  if (FL == FL_SGRID && USE_REF && g_sgrid_tile && (STRICT || g_metric_corners == 2) &&
      launch_sgrid_tile_metric(st, m, STAGE, x, s, w, stride, alpha, fixed, R, slot, em, ef))
    return;
  if (!STRICT && g_metric_corners == 4 && fast_hex<FL, STAGE, USE_REF>() && FL != FL_TET4 && !em && !ef) {
    constexpr int F2 = FL == FL_TET4 ? FL_HEX8 : FL;
    if (s && alpha != 0.0)
      launch_metric_occ<F2, STAGE, true>(st, m, x, s, w, stride, alpha, R, slot);
    else
      launch_metric_occ<F2, STAGE, false>(st, m, x, s, w, stride, alpha, R, slot);
    return;
  }
  if (!STRICT && (g_metric_corners == 2 || g_metric_corners == 4)) {
    if (s && alpha != 0.0)
      launch_elem_pipe<FL, STAGE, USE_REF, false, true>(st, m, x, s, w, stride, alpha, nullptr, 0, R, slot, em, ef);
    else
      launch_elem_pipe<FL, STAGE, USE_REF, false, false>(st, m, x, s, w, stride, alpha, nullptr, 0, R, slot, em, ef);
    return;
  }
  if (fast_hex<FL, STAGE, USE_REF>() && g_metric_corners == 3 && FL != FL_TET4) {
    constexpr int F2 = FL == FL_TET4 ? FL_HEX8 : FL;
    const dim3 qg = quad_grid<F2>(m);
    if (s && alpha != 0.0)
      k_quads<F2, STAGE, false, true><<<qg, NTP, 0, st>>>(m, x, s, w, stride, alpha, nullptr, 0, fixed, R, slot, em, ef);
    else
      k_quads<F2, STAGE, false, false><<<qg, NTP, 0, st>>>(m, x, s, w, stride, alpha, nullptr, 0, fixed, R, slot, em, ef);
    return;
  }
  if (fast_hex<FL, STAGE, USE_REF>() && g_metric_corners == 1 && FL != FL_TET4) {
    constexpr int F2 = FL == FL_TET4 ? FL_HEX8 : FL;
    const dim3 cg = corner_grid<F2>(m);
    if (s && alpha != 0.0)
      k_corners<F2, STAGE, false, true><<<cg, NT, 0, st>>>(m, x, s, w, stride, alpha, nullptr, 0, fixed, R, slot, em, ef);
    else
      k_corners<F2, STAGE, false, false><<<cg, NT, 0, st>>>(m, x, s, w, stride, alpha, nullptr, 0, fixed, R, slot, em, ef);
    return;
  }
  const unsigned grid = elem_grid<FL>(m);
  if (s && alpha != 0.0)
    k_metric<FL, STAGE, USE_REF, true><<<grid, NT, 0, st>>>(m, x, s, w, stride, alpha, R, slot, em, ef);
  else
    k_metric<FL, STAGE, USE_REF, false><<<grid, NT, 0, st>>>(m, x, s, w, stride, alpha, R, slot, em, ef);
}
template <int FL>
static void launch_metric1(cudaStream_t st, const MeshDev& m, int stage, int use_ref, const double* x, const double* s,
                           const double* w, long long stride, double alpha, const Reduce& R, int slot, double* em,
                           int32_t* ef, const uint8_t* fixed) {
  if (stage == 0)
    launch_metric2<FL, 0, true>(st, m, x, s, w, stride, alpha, R, slot, em, ef, fixed);
  else if (use_ref)
    launch_metric2<FL, 1, true>(st, m, x, s, w, stride, alpha, R, slot, em, ef, fixed);
  else
    launch_metric2<FL, 1, false>(st, m, x, s, w, stride, alpha, R, slot, em, ef, fixed);
}
static void launch_metric(cudaStream_t st, const MeshDev& m, int stage, int use_ref, const double* x, const double* s,
                          const double* w, long long stride, double alpha, const Reduce& R, int slot, double* em,
                          int32_t* ef, const uint8_t* fixed) {
  if (m.kind == KIND_SGRID)
    launch_metric1<FL_SGRID>(st, m, stage, use_ref, x, s, w, stride, alpha, R, slot, em, ef, fixed);
  else if (m.kind == KIND_HEX8)
    launch_metric1<FL_HEX8>(st, m, stage, use_ref, x, s, w, stride, alpha, R, slot, em, ef, fixed);
  else
    launch_metric1<FL_TET4>(st, m, stage, use_ref, x, s, w, stride, alpha, R, slot, em, ef, fixed);
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
  constexpr bool FAST = fast_hex<FL, STAGE, USE_REF>();
  constexpr int NN = npe<FL>();
  double val[1] = {0.0};
  unsigned long long inv = 0;
  const long long ntiles = num_tiles<FL>(m);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long e, nid[8];
    if (!element_of_tile<FL>(m, tile, e, nid)) continue;
    double vc[8][3], vo[8][3], grad[8][3];
#pragma unroll
    for (int a = 0; a < NN; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        vc[vpos<FL, FAST>(a)][r] = __ldg(&x[r * stride + nid[a]]);
        vo[vpos<FL, FAST>(a)][r] = __ldg(&w[r * stride + nid[a]]);
      }
    bool valid;
    double mm;
    if (FAST)
      mm = hex_metric_fast<STAGE, true>(vc, vo, valid, grad);
    else
      mm = element_grad_metric<FL, STAGE, USE_REF>(vc, vo, valid, grad);
    const bool use = valid || STAGE == 0;  // gradient_functors.hpp:391, Def.hpp:1219
    const long long ge = m.elem_off + e;
#pragma unroll
    for (int a = 0; a < NN; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) eg[(long long)(a * 3 + r) * eg_stride + ge] = use ? grad[vpos<FL, FAST>(a)][r] : 0.0;
    bool counted = true;
    if (FL != FL_SGRID && m.elem_owned) counted = m.elem_owned[e] != 0;
    if (counted) {
      val[0] += mm;
      inv += valid ? 0 : 1;
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
  if (!STRICT && (g_grad_corners == 2 || g_grad_corners >= 4)) {
    if (stage == 0)
      launch_elem_pipe<FL, 0, true, true, false>(st, m, x, nullptr, w, stride, 0.0, eg, egs, R, slot, em, ef);
    else if (use_ref)
      launch_elem_pipe<FL, 1, true, true, false>(st, m, x, nullptr, w, stride, 0.0, eg, egs, R, slot, em, ef);
    else
      launch_elem_pipe<FL, 1, false, true, false>(st, m, x, nullptr, w, stride, 0.0, eg, egs, R, slot, em, ef);
    return;
  }
  if (!STRICT && FL != FL_TET4 && g_grad_corners == 3 && (stage == 0 || use_ref)) {
    constexpr int F2 = FL == FL_TET4 ? FL_HEX8 : FL;
    const dim3 qg = quad_grid<F2>(m);
    if (stage == 0)
      k_quads<F2, 0, true, false><<<qg, NTP, 0, st>>>(m, x, nullptr, w, stride, 0.0, eg, egs, nullptr, R, slot, em, ef);
    else
      k_quads<F2, 1, true, false><<<qg, NTP, 0, st>>>(m, x, nullptr, w, stride, 0.0, eg, egs, nullptr, R, slot, em, ef);
    return;
  }
  if (!STRICT && FL != FL_TET4 && g_grad_corners == 1 && (stage == 0 || use_ref)) {
    constexpr int F2 = FL == FL_TET4 ? FL_HEX8 : FL;
    const dim3 cg = corner_grid<F2>(m);
    if (stage == 0)
      k_corners<F2, 0, true, false><<<cg, NT, 0, st>>>(m, x, nullptr, w, stride, 0.0, eg, egs, nullptr, R, slot, em, ef);
    else
      k_corners<F2, 1, true, false><<<cg, NT, 0, st>>>(m, x, nullptr, w, stride, 0.0, eg, egs, nullptr, R, slot, em, ef);
    return;
  }
  const unsigned grid = elem_grid<FL>(m);
  if (stage == 0)
    k_grad_elements<FL, 0, true><<<grid, NT, 0, st>>>(m, x, w, stride, eg, egs, R, slot, em, ef);
  else if (use_ref)
    k_grad_elements<FL, 1, true><<<grid, NT, 0, st>>>(m, x, w, stride, eg, egs, R, slot, em, ef);
  else
    k_grad_elements<FL, 1, false><<<grid, NT, 0, st>>>(m, x, w, stride, eg, egs, R, slot, em, ef);
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
      const long long ent = m.csr_ent32 ? (long long)__ldg(&m.csr_ent32[p]) : __ldg(&m.csr_ent[p]);
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

#include "pms_elem_ws.cuh"

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
  const long long cap = (long long)sm_count() * 16;  // SMs x resident CTAs; grid-stride beyond that
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

// s update fused with s.g, s.s and the alpha_0 scan of the new s (one pass over the nodes instead of four)
__global__ void __launch_bounds__(NT) k_cg_s_update_dots(const double* __restrict__ r, double* __restrict__ s, double* __restrict__ d,
                                                         const double* __restrict__ g, const double* __restrict__ edge,
                                                         const uint8_t* __restrict__ dotmask, const uint8_t* __restrict__ skip,
                                                         const double beta, const int restart, const long long n,
                                                         const long long stride, const Reduce R, const int slot) {
  double acc[2] = {0.0, 0.0}, mn[1] = {DBL_MAX};
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    double sv[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double rv = r[c * stride + i];
      sv[c] = restart ? rv : 1.0 * rv + beta * s[c * stride + i];
      s[c * stride + i] = sv[c];
      d[c * stride + i] = rv;
    }
    if (!dotmask || dotmask[i]) {
#pragma unroll
      for (int c = 0; c < 3; c++) acc[0] += sv[c] * g[c * stride + i];
#pragma unroll
      for (int c = 0; c < 3; c++) acc[1] += sv[c] * sv[c];
    }
    if (!(skip && skip[i])) {
      double sn = 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) sn += sv[c] * sv[c];
      sn = sqrt(sn);
      if (sn > 0.0) {
        const double cand = edge[i] / sn;
        mn[0] = cand < mn[0] ? cand : mn[0];
      }
    }
  }
  grid_reduce<OpSum, 2, false>(acc, 0ull, R, slot);
  Reduce R2 = R;  // second reduction of the same kernel: its own ticket and partial area
  R2.ticket = R.ticket + 1;
  R2.partial_d = R.partial_d + (size_t)2 * R.max_blocks;
  grid_reduce<OpMin, 1, false>(mn, 0ull, R2, slot + 2);
}
static void launch_cg_s_update_dots(cudaStream_t st, const double* r, double* s, double* d, const double* g, const double* edge,
                                    const uint8_t* dotmask, const uint8_t* skip, double beta, int restart, long long n,
                                    long long stride, const Reduce& R, int slot) {
  k_cg_s_update_dots<<<blas_blocks(n), NT, 0, st>>>(r, s, d, g, edge, dotmask, skip, beta, restart, n, stride, R, slot);
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
                                               const uint8_t* __restrict__ owned, const double alpha, const long long n,
                                               const long long stride, const Reduce R, const int slot) {
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
    if (!fx && (!owned || owned[i])) {  // deltas over owned nodes only (ghost copies have incomplete edge lengths)
      acc[0] = maxDelta > acc[0] ? maxDelta : acc[0];
      const double rel = maxDelta / edge[i];
      acc[1] = rel > acc[1] ? rel : acc[1];
    }
  }
  grid_reduce<OpMax, 2, false>(acc, 0ull, R, slot);
}
static void launch_update(cudaStream_t st, double* x, double* lag, const double* s, const double* edge, const uint8_t* fixed,
                          const uint8_t* owned, double alpha, long long n, long long stride, const Reduce& R, int slot) {
  k_update<<<blas_blocks(n), NT, 0, st>>>(x, lag, s, edge, fixed, owned, alpha, n, stride, R, slot);
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
// Two-pass form of k_edge_csr: the element's mean edge length does not depend on which node asks for it, so it is
// computed once per element (instead of once per adjacent node: 8x the sqrt work and gathers for hexes) and the node
// pass sums the stored values in the same CSR order - bit-identical result, 4.6 ms -> well under 1 ms at 256^3.
template <int FL>
__global__ void __launch_bounds__(NT) k_edge_elem(const MeshDev m, const double* __restrict__ w, const long long stride,
                                                  double* __restrict__ scratch) {
  const long long e = (long long)blockIdx.x * NT + threadIdx.x;
  if (e >= m.nelems) return;
  const int hex_edges[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
  const int tet_edges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
  constexpr int NN = npe<FL>();
  double v[8][3];
#pragma unroll
  for (int a = 0; a < NN; a++) {
    const long long q = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
#pragma unroll
    for (int d = 0; d < 3; d++) v[a][d] = __ldg(&w[d * stride + q]);
  }
  scratch[e] = (FL == FL_TET4) ? mean_edge_length<6>(v, tet_edges) : mean_edge_length<12>(v, hex_edges);
}
__global__ void __launch_bounds__(NT) k_edge_node(const MeshDev m, const double* __restrict__ scratch, double* __restrict__ el,
                                                  double* __restrict__ na) {
  const long long n = (long long)blockIdx.x * NT + threadIdx.x;
  if (n >= m.nnodes) return;
  double nm = 0.0, nele = 0.0;
  for (long long p = m.csr_ptr[n]; p < m.csr_ptr[n + 1]; p++) {
    nm += scratch[m.csr_ent[p] >> 3];
    nele += 1.0;
  }
  el[n] = nm / nele;
  na[n] = nele;
}
static void launch_edge_lengths(cudaStream_t st, const MeshDev& m, const double* w, long long stride, double* el, double* na,
                                double* scratch) {
  if (m.kind != KIND_SGRID && scratch) {
    const unsigned ge = (unsigned)((m.nelems + NT - 1) / NT), gn = (unsigned)((m.nnodes + NT - 1) / NT);
    if (m.kind == KIND_HEX8)
      k_edge_elem<FL_HEX8><<<ge, NT, 0, st>>>(m, w, stride, scratch);
    else
      k_edge_elem<FL_TET4><<<ge, NT, 0, st>>>(m, w, stride, scratch);
    k_edge_node<<<gn, NT, 0, st>>>(m, scratch, el, na);
    return;
  }
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

// interface nodes duplicated on several ranks (structured block per GPU): f[node] = sum over all ranks holding a copy,
// added in ascending rank order (entry pos < 0 = this rank's own value), so every copy ends bitwise identical
__global__ void __launch_bounds__(NT) k_shared_sum(const long long nshared, const int32_t* __restrict__ nodes,
                                                   const int64_t* __restrict__ ptr, const int64_t* __restrict__ ent_pos,
                                                   const int32_t* __restrict__ ent_cnt, const double* __restrict__ recvbuf,
                                                   double* __restrict__ f, const int ncomp, const long long stride) {
  const long long j = (long long)blockIdx.x * NT + threadIdx.x;
  if (j >= nshared) return;
  const long long n = nodes[j];
  for (int c = 0; c < ncomp; c++) {
    double v = 0.0;
    bool first = true;
    for (long long p = ptr[j]; p < ptr[j + 1]; p++) {
      const long long pos = ent_pos[p];
      const double t = pos < 0 ? f[c * stride + n] : recvbuf[pos + (long long)c * ent_cnt[p]];
      v = first ? t : v + t;
      first = false;
    }
    // all copies read their own value before any write: own value is only read by this thread
    f[c * stride + n] = v;
  }
}
static void launch_shared_sum(cudaStream_t st, long long nshared, const int32_t* nodes, const int64_t* ptr, const int64_t* ent_pos,
                              const int32_t* ent_cnt, const double* recvbuf, double* f, int ncomp, long long stride) {
  if (nshared > 0)
    k_shared_sum<<<(unsigned)((nshared + NT - 1) / NT), NT, 0, st>>>(nshared, nodes, ptr, ent_pos, ent_cnt, recvbuf, f, ncomp, stride);
}

template <int FL>
__global__ void __launch_bounds__(NT) k_elem_fixed_bits(const MeshDev m, const uint8_t* __restrict__ fixed,
                                                        uint8_t* __restrict__ bits) {
  const long long ntiles = num_tiles<FL>(m);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long e, nid[8];
    if (!element_of_tile<FL>(m, tile, e, nid)) continue;
    unsigned b = 0;
    constexpr int NN = npe<FL>();
#pragma unroll
    for (int a = 0; a < NN; a++) b |= (fixed[nid[a]] ? 1u : 0u) << a;
    bits[m.elem_off + e] = (uint8_t)b;
  }
}
// ------------------------------------------------------------------------------------------------
// StructuredGridRefiner 2x prolongation (structured/StructuredGridRefinerDef.hpp:134-222).  The reference loops over
// INPUT nodes and writes the vertex, 3 edge midpoints, 3 face centres and the centroid each one owns; here one thread
// per OUTPUT node (coalesced 8-byte stores, the coarse reads hit L1/L2): its index parity says which of the four it is,
// and the coarse nodes are summed in the reference's order (i outermost, k innermost; 0.5*(a+b) for edges).  HBM-bound:
// 24 B read per coarse node + 24 B written per fine node.
// ------------------------------------------------------------------------------------------------
// fine node of parity (pi, PJ, PK) from the 2x2x2 coarse values xv[a][b][c]: i outermost / k innermost as in the
// reference's loops, 0.5*(a+b) for edge midpoints, plain copy for vertices
template <int PJ, int PK>
__device__ __forceinline__ double refine_value(const double (&xv)[2][2][2], const int pi) {
  if (PJ + PK == 0) return pi ? 0.5 * (xv[0][0][0] + xv[1][0][0]) : xv[0][0][0];
  if (PJ + PK == 1 && !pi) return 0.5 * (xv[0][0][0] + xv[0][PJ][PK]);
  const double wgt = (PJ + PK + pi) == 2 ? 0.25 : 0.125;
  double v = 0.0;
#pragma unroll
  for (int a = 0; a < 2; a++) {
    if (a > pi) break;
#pragma unroll
    for (int b = 0; b <= PJ; b++)
#pragma unroll
      for (int c = 0; c <= PK; c++) v += wgt * xv[a][b][c];
  }
  return v;
}

// One thread per (fine I, coarse j) of coarse plane k = blockIdx.y: it loads the <= 8 coarse values once and writes the
// up to four fine nodes (I, 2j + {0,1}, 2k + {0,1}); stores stay coalesced along I, 32-bit index arithmetic only.
__global__ void __launch_bounds__(NT) k_refine_sgrid(const int ni, const int nj, const int nk, const double* __restrict__ src,
                                                     const long long src_stride, double* __restrict__ dst,
                                                     const long long dst_stride) {
  const unsigned oi = 2 * ni - 1, oj = 2 * nj - 1;
  const unsigned t = blockIdx.x * NT + threadIdx.x, k = blockIdx.y;
  if (t >= oi * (unsigned)nj) return;
  const unsigned j = t / oi, I = t - j * oi;
  const int pi = I & 1;
  const bool hj = j + 1 < (unsigned)nj, hk = k + 1 < (unsigned)nk;
  const long long dj = ni, dk = (long long)ni * nj;
  const long long base = (I >> 1) + dj * j + dk * k;
  const long long o = I + (long long)oi * (2 * j + (long long)oj * 2 * k), odj = oi, odk = (long long)oi * oj;
#pragma unroll
  for (int ic = 0; ic < 3; ic++) {
    const double* x = src + ic * src_stride + base;
    double xv[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int c = 0; c < 2; c++)
          xv[a][b][c] = (a <= pi && (b == 0 || hj) && (c == 0 || hk)) ? __ldg(x + a + b * dj + c * dk) : 0.0;
    double* y = dst + ic * dst_stride + o;
    y[0] = refine_value<0, 0>(xv, pi);
    if (hj) y[odj] = refine_value<1, 0>(xv, pi);
    if (hk) y[odk] = refine_value<0, 1>(xv, pi);
    if (hj && hk) y[odj + odk] = refine_value<1, 1>(xv, pi);
  }
}

static void launch_refine_sgrid(cudaStream_t st, int ni, int nj, int nk, const double* src, long long src_off, long long src_stride,
                                double* dst, long long dst_off, long long dst_stride) {
  const long long oi = 2LL * ni - 1;
  // limits: (2ni-1)*nj below 2^32, at most 65535 coarse planes (inside the 2^31-node ctx limit for any sane block)
  k_refine_sgrid<<<dim3((unsigned)((oi * nj + NT - 1) / NT), (unsigned)nk), NT, 0, st>>>(ni, nj, nk, src + src_off, src_stride,
                                                                                       dst + dst_off, dst_stride);
}

// ------------------------------------------------------------------------------------------------
// Surface smoothing hooks: GenericAlgorithm_get_gradient_2 + project_delta_to_tangent_plane
// (Def.hpp:1401-1447, MeshSmoother.cpp:437-460), snap_nodes (Def.hpp:236-316), get_surface_normals (Def.hpp:596-706)
// with analytic evaluators in place of the CAD kernel.  One thread per listed (non-fixed, MS_SURFACE) node.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void surf_normal(const SurfaceDev& s, const double (&x)[3], double (&n)[3], double& t) {
  t = 0.0;
  if (s.kind == 0) {
#pragma unroll
    for (int i = 0; i < 3; i++) n[i] = s.u[i];
    return;
  }
  double d[3] = {x[0] - s.p[0], x[1] - s.p[1], x[2] - s.p[2]};
  if (s.kind == 2) {
    t = d[0] * s.u[0] + d[1] * s.u[1] + d[2] * s.u[2];
#pragma unroll
    for (int i = 0; i < 3; i++) d[i] = d[i] - t * s.u[i];
  }
  const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
#pragma unroll
  for (int i = 0; i < 3; i++) n[i] = r > 0.0 ? d[i] / r : 0.0;
}

__global__ void __launch_bounds__(NT) k_surf_project(const long long n, const int32_t* __restrict__ nodes,
                                                     const int32_t* __restrict__ sid, const SurfaceDev* __restrict__ surfs,
                                                     const double* __restrict__ x, double* __restrict__ g, const long long stride) {
  const long long i = (long long)blockIdx.x * NT + threadIdx.x;
  if (i >= n) return;
  const long long nd = nodes[i];
  const SurfaceDev s = surfs[sid[i]];
  const double xx[3] = {x[nd], x[stride + nd], x[2 * stride + nd]};
  double nrm[3], t;
  surf_normal(s, xx, nrm, t);
  double gg[3] = {g[nd], g[stride + nd], g[2 * stride + nd]};
  double dot = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) dot += gg[k] * nrm[k];
#pragma unroll
  for (int k = 0; k < 3; k++) g[k * stride + nd] = gg[k] - dot * nrm[k];
}

__global__ void __launch_bounds__(NT) k_surf_snap(const long long n, const int32_t* __restrict__ nodes, const int32_t* __restrict__ sid,
                                                  const SurfaceDev* __restrict__ surfs, double* __restrict__ x, const long long stride,
                                                  const Reduce R, const int slot) {
  double v[1] = {0.0};
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const long long nd = nodes[i];
    const SurfaceDev s = surfs[sid[i]];
    const double o[3] = {x[nd], x[stride + nd], x[2 * stride + nd]};
    double nrm[3], t, y[3];
    surf_normal(s, o, nrm, t);
    if (s.kind == 0) {
      const double h = (o[0] - s.p[0]) * nrm[0] + (o[1] - s.p[1]) * nrm[1] + (o[2] - s.p[2]) * nrm[2];
#pragma unroll
      for (int k = 0; k < 3; k++) y[k] = o[k] - h * nrm[k];
    } else if (s.kind == 1) {
#pragma unroll
      for (int k = 0; k < 3; k++) y[k] = s.p[k] + s.R * nrm[k];
    } else {
#pragma unroll
      for (int k = 0; k < 3; k++) y[k] = (s.p[k] + t * s.u[k]) + s.R * nrm[k];
    }
    double dm = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      dm += (o[k] - y[k]) * (o[k] - y[k]);
      x[k * stride + nd] = y[k];
    }
    v[0] = fmax(v[0], sqrt(dm));
  }
  grid_reduce<OpMax, 1, false>(v, 0ull, R, slot);
}

__global__ void __launch_bounds__(NT) k_surf_normals(const long long n, const int32_t* __restrict__ nodes,
                                                     const int32_t* __restrict__ sid, const SurfaceDev* __restrict__ surfs,
                                                     const double* __restrict__ x, double* __restrict__ out, const long long stride) {
  const long long i = (long long)blockIdx.x * NT + threadIdx.x;
  if (i >= n) return;
  const long long nd = nodes[i];
  const double xx[3] = {x[nd], x[stride + nd], x[2 * stride + nd]};
  double nrm[3], t;
  surf_normal(surfs[sid[i]], xx, nrm, t);
#pragma unroll
  for (int k = 0; k < 3; k++) out[k * stride + nd] = nrm[k];
}

static void launch_surf_project(cudaStream_t st, long long n, const int32_t* nodes, const int32_t* sid, const SurfaceDev* surfs,
                                const double* x, double* g, long long stride) {
  if (n > 0) k_surf_project<<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>(n, nodes, sid, surfs, x, g, stride);
}
static void launch_surf_snap(cudaStream_t st, long long n, const int32_t* nodes, const int32_t* sid, const SurfaceDev* surfs,
                             double* x, long long stride, const Reduce& R, int slot) {
  long long nb = (n + NT - 1) / NT;
  if (nb < 1) nb = 1;
  if (nb > R.max_blocks) nb = R.max_blocks;
  k_surf_snap<<<(unsigned)nb, NT, 0, st>>>(n, nodes, sid, surfs, x, stride, R, slot);
}
static void launch_surf_normals(cudaStream_t st, long long n, const int32_t* nodes, const int32_t* sid, const SurfaceDev* surfs,
                                const double* x, double* out, long long stride) {
  if (n > 0) k_surf_normals<<<(unsigned)((n + NT - 1) / NT), NT, 0, st>>>(n, nodes, sid, surfs, x, out, stride);
}

// tet4: W^-1 and det W of every element from the reference coordinates (once per run; tet_reference, pms_device.cuh)
__global__ void __launch_bounds__(NT) k_tet_ref(const MeshDev m, const double* __restrict__ w, const long long stride,
                                                double* __restrict__ out, const long long out_stride) {
  const long long e = (long long)blockIdx.x * NT + threadIdx.x;
  if (e >= m.nelems) return;
  double vo[8][3], ref[10];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const long long nid = __ldg(&m.conn[(long long)a * m.conn_stride + e]);
#pragma unroll
    for (int r = 0; r < 3; r++) vo[a][r] = __ldg(&w[r * stride + nid]);
  }
  tet_reference(vo, ref);
#pragma unroll
  for (int v = 0; v < 10; v++) out[(long long)v * out_stride + m.elem_off + e] = ref[v];
}
// CUDA loads a kernel when it is first used (lazy module loading, several ms for these large kernels), which would fall
// into the first line searches of a run: touch every multi-alpha instantiation a mesh of this kind can ask for at commit.
static void prepare_line_search_kernels(const MeshDev& m) {
  if (STRICT || m.kind == KIND_TET4) return;
  Reduce R{};
  for (int stage = 0; stage < 2; stage++)
    for (int nb = 8; nb <= 16; nb += 2)
      for (int nf = 4; nf <= (stage == 0 ? 4 : nb); nf += 2)
        launch_metric_multi(0, m, stage, 1, nb, nf, nullptr, nullptr, nullptr, 0, nullptr, R, 0, nullptr);
}

static void launch_tet_ref(cudaStream_t st, const MeshDev& m, const double* w, long long stride, double* out, long long out_stride) {
  if (m.kind == KIND_TET4 && m.nelems > 0)
    k_tet_ref<<<(unsigned)((m.nelems + NT - 1) / NT), NT, 0, st>>>(m, w, stride, out, out_stride);
}

static void launch_elem_fixed_bits(cudaStream_t st, const MeshDev& m, const uint8_t* fixed, uint8_t* bits) {
  if (m.kind == KIND_SGRID)
    k_elem_fixed_bits<FL_SGRID><<<elem_grid<FL_SGRID>(m), NT, 0, st>>>(m, fixed, bits);
  else if (m.kind == KIND_HEX8)
    k_elem_fixed_bits<FL_HEX8><<<elem_grid<FL_HEX8>(m), NT, 0, st>>>(m, fixed, bits);
  else
    k_elem_fixed_bits<FL_TET4><<<elem_grid<FL_TET4>(m), NT, 0, st>>>(m, fixed, bits);
}

static const Launchers g_launchers = {launch_metric,    launch_grad_elements, launch_grad_gather, launch_grad_fused_sgrid,
                                      sgrid_fused_scratch_bytes,
                                      launch_group_sum, launch_axpby,         launch_axpbypgz,    launch_set,
                                      launch_dot,       launch_cg_r_dots,     launch_cg_s_update, launch_cg_s_update_dots, launch_alpha0,
                                      launch_update,    launch_edge_lengths,  launch_divide,      launch_pack,
                                      launch_unpack,    launch_elem_fixed_bits, launch_tet_ref, prepare_line_search_kernels, launch_grad_bricks, launch_elem_ws, elem_gather_grid, launch_shared_sum,
                                      launch_metric_multi, launch_refine_sgrid, launch_surf_project, launch_surf_snap, launch_surf_normals};

const Launchers* launchers() { return &g_launchers; }
void set_variant(int metric_corners, int grad_corners, int sgrid_tile) {
  g_metric_corners = metric_corners;
  g_grad_corners = grad_corners;
  g_sgrid_tile = sgrid_tile;
}

}  // namespace PMS_NS
