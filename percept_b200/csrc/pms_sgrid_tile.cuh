// pms_sgrid_tile.cuh — structured-block tile kernels (included by pms_kernels.cu inside namespace PMS_NS).
//
// One-pass fused metric + nodal gradient of a StructuredGrid block: replaces the element pass + 192 B/cell scratch +
// node gather of the two-pass scheme (reference functor: gradient_functors.hpp:349-407, whose 24 Kokkos::atomic_add per
// cell become an in-CTA gather in a FIXED order).
//
//  * A CTA of 32 x 8 threads owns a tile of 32 x (8*TL_NSUB) cells and marches a chunk of k-layers.  The node planes
//    k and k+1 of x and of the reference mesh w (the tile's cells + a one-node halo: 33 x (8*TL_NSUB+1) nodes, 6
//    component planes) live in shared memory; plane k+2 is in flight while layer k is evaluated: 128-bit cp.async
//    (LDGSTS.128) over the 16-byte-aligned span of every node row, three rotating plane buffers.  Each node is fetched
//    once per CTA instead of 8 times, and there is no per-vertex address arithmetic.
//  * Each thread evaluates TL_NSUB cells per layer (vertices read in place from the staged planes) and writes the cell's
//    24 contributions to a shared-memory exchange array ex[a*3+r][cell].
//  * Node phase: every node of the tile's patch sums the contributions of its adjacent cells in the fixed (dk,dj,di)
//    order of k_grad_gather_sgrid (= ascending cell id = the serial reference loop): the four cells of layer k-1 were
//    summed one layer earlier into a carry held in shared memory, the four cells of layer k are added to it and the node
//    of plane k is final -> one coalesced store per component, fixed nodes zeroed (gradient_functors.hpp:223-243).
//    The sequence of floating-point additions per node is exactly that of the two-pass path, so the result is bitwise
//    identical to it (and, in the strict build, to the CPU oracle).
//  * Nodes on a line between two tiles ("seam" nodes: 1/32 + 1/(8*TL_NSUB) of all nodes) receive contributions from two
//    or four CTAs.  Adding per-tile partial sums would change the order of additions, so each CTA stores the
//    individual contributions of its own cells into a compact seam scratch (slot = the cell's position (dk,dj,di)
//    around the node) and k_sgrid_seam_sum adds the 8 slots in order.  No atomics, no index tables; the scratch is
//    (1/32 + 1/(8*TL_NSUB)) * 192 B per node instead of 192 B per cell.
//  * k is cut into chunks so that tiles x chunks fill the SMs evenly; a chunk that does not start at k = 0 first
//    re-evaluates the layer below it to build the carry (nothing of that layer is stored or counted).
//
// The same staging serves the metric-only kernels of the line search (k_sgrid_tile_metric).
#pragma once

constexpr int TL_X = 32, TL_Y = 8;  // threads of a CTA = cells of one sub-row step
#ifndef TL_NSUB
#define TL_NSUB 2
#endif
constexpr int TL_TY = TL_Y * TL_NSUB;            // cells per tile in j
constexpr int TL_PX = TL_X + 1, TL_PY = TL_TY + 1;  // node patch of a tile
constexpr int TL_ROWP = 36;                      // staged row: 18 x 16 B covering 33 nodes from an even node index
constexpr int TL_NTH = TL_X * TL_Y;
constexpr int TL_NCELL = TL_X * TL_TY;
constexpr int TL_NNODE = TL_PX * TL_PY;

__host__ __device__ inline int tl_lines(int n, int t) { return n >= 3 ? (n - 2) / t : 0; }  // interior tile lines among n nodes

struct SeamDev {
  double* J;  // [lineJ][slot 8][comp 3][nk][ni]  nodes with j on a tile line (all i)
  double* I;  // [lineI][slot 8][comp 3][nk][nj]  nodes with i on a tile line (j not on a tile line)
};

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// vertices of one cell read in place from two staged planes: V(p, r), p = di + 2 dj + 4 dk
template <int NARR>
struct TileVerts {
  const double* pl[2];  // plane k / k+1, already offset to this array's component 0
  int o[2][2];          // [dk][dj]: row * TL_ROWP + column of the cell's (di = 0) node
  PMS_DI double operator()(int p, int r) const { return pl[p >> 2][r * (TL_PY * TL_ROWP) + o[p >> 2][(p >> 1) & 1] + (p & 1)]; }
};

// stage node plane kp of NARR 3-component arrays (a0, a1[, a2]; component c at + c * stride) for the tile at (i0, j0)
template <int NARR>
__device__ __forceinline__ void tl_issue_plane(double* buf, const MeshDev& m, const double* a0, const double* a1,
                                               const double* a2, const long long stride, const int i0, const int j0,
                                               const int kp) {
  if (kp < m.nk) {
    const int rows = min(TL_PY, m.nj - j0);
    const int total = NARR * 3 * rows * (TL_ROWP / 2);
    for (int q = threadIdx.x + TL_X * threadIdx.y; q < total; q += TL_NTH) {
      const int ch = q % (TL_ROWP / 2);
      const int t = q / (TL_ROWP / 2);
      const int row = t % rows, ac = t / rows;  // ac = array * 3 + component
      const long long nrow = m.node_off + i0 + (long long)m.ni * ((j0 + row) + (long long)m.nj * kp);
      const long long a = (nrow & ~1LL) + 2 * ch;
      const double* src = ac < 3 ? a0 : (ac < 6 ? a1 : a2);
      if (a + 2 <= stride) cp_async16(buf + (ac * TL_PY + row) * TL_ROWP + 2 * ch, src + (ac % 3) * stride + a);
    }
  }
  cp_async_commit();
}

// parity of the first staged element of a node row (the row is staged from the even index at or below its first node)
__device__ __forceinline__ int tl_row_off(const MeshDev& m, const int i0, const int j, const int k) {
  return (int)((m.node_off + i0 + (long long)m.ni * (j + (long long)m.nj * k)) & 1LL);
}

template <int STAGE>
__global__ void __launch_bounds__(TL_NTH, 1) k_sgrid_tile_grad(const MeshDev m, const double* __restrict__ x,
                                                               const double* __restrict__ w, const long long stride,
                                                               const uint8_t* __restrict__ fixed, double* __restrict__ g,
                                                               const SeamDev seam, const int ntx, const int nty, const int kc,
                                                               const Reduce R, const int slot, double* __restrict__ em,
                                                               int32_t* __restrict__ ef) {
  extern __shared__ double sm[];
  constexpr int PLANE = 6 * TL_PY * TL_ROWP;
  double* const planes = sm;                    // [3][PLANE]
  double* const ex = sm + 3 * PLANE;            // [24][TL_NCELL]
  double* const carry = ex + 24 * TL_NCELL;     // [3][TL_NNODE]
  int* const info = reinterpret_cast<int*>(carry + 3 * TL_NNODE);  // [TL_NNODE]
  const int tx = threadIdx.x, ty = threadIdx.y, tid = tx + TL_X * ty;
  const int tile = blockIdx.x % (ntx * nty), chunk = blockIdx.x / (ntx * nty);
  const int i0 = (tile % ntx) * TL_X, j0 = (tile / ntx) * TL_TY;
  const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
  const int actx = min(TL_X, nci - i0), acty = min(TL_TY, ncj - j0);  // cells of this tile
  const int kb = chunk * kc, ke = min(kb + kc, nck);
  const int kstart = kb > 0 ? kb - 1 : 0;  // one warm-up layer below the chunk builds the carry

  // per patch node: bit0 exists, bit1 seam (bit2: the j line, else the i line), bits 4..7 adjacent tile cells (dj*2+di)
  for (int q = tid; q < TL_NNODE; q += TL_NTH) {
    const int pi = q % TL_PX, pj = q / TL_PX;
    const int i = i0 + pi, j = j0 + pj;
    int f = 0;
    if (i < m.ni && j < m.nj) {
      f = 1;
      const bool sj = (j % TL_TY == 0) && j > 0 && j < m.nj - 1, si = (i % TL_X == 0) && i > 0 && i < m.ni - 1;
      if (sj) f |= 2 | 4;
      else if (si) f |= 2;
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        const int cx = pi - 1 + (c4 & 1), cy = pj - 1 + (c4 >> 1);
        if (cx >= 0 && cx < actx && cy >= 0 && cy < acty) f |= 16 << c4;
      }
    }
    info[q] = f;
    carry[q] = 0.0;
    carry[TL_NNODE + q] = 0.0;
    carry[2 * TL_NNODE + q] = 0.0;
  }

  tl_issue_plane<2>(planes + (kstart % 3) * PLANE, m, x, w, nullptr, stride, i0, j0, kstart);
  tl_issue_plane<2>(planes + ((kstart + 1) % 3) * PLANE, m, x, w, nullptr, stride, i0, j0, kstart + 1);

  double val[1] = {0.0};
  unsigned long long inv = 0;
  const long long plane_nodes = (long long)m.ni * m.nj;

  for (int k = kstart; k < ke; k++) {
    const bool warm = k < kb;
    cp_async_wait<0>();
    __syncthreads();  // plane k+1 landed for everyone; ex, carry and the buffer of plane k-1 are free
    tl_issue_plane<2>(planes + ((k + 2) % 3) * PLANE, m, x, w, nullptr, stride, i0, j0, k + 2);
    const double* const p0 = planes + (k % 3) * PLANE;
    const double* const p1 = planes + ((k + 1) % 3) * PLANE;

#pragma unroll 1
    for (int sub = 0; sub < TL_NSUB; sub++) {
      const int cj = ty + TL_Y * sub;
      if (tx >= actx || cj >= acty) continue;
      TileVerts<2> VC, VO;
      VC.pl[0] = p0;
      VC.pl[1] = p1;
      VO.pl[0] = p0 + 3 * TL_PY * TL_ROWP;
      VO.pl[1] = p1 + 3 * TL_PY * TL_ROWP;
#pragma unroll
      for (int dk = 0; dk < 2; dk++)
#pragma unroll
        for (int dj = 0; dj < 2; dj++) {
          const int o = (cj + dj) * TL_ROWP + tx + tl_row_off(m, i0, j0 + cj + dj, k + dk);
          VC.o[dk][dj] = o;
          VO.o[dk][dj] = o;
        }
      bool valid;
      double mm, grad[8][3];
      if (STRICT) {
        double vc[8][3], vo[8][3];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int r = 0; r < 3; r++) {
            vc[a][r] = VC(a, r);
            vo[a][r] = VO(a, r);
          }
        mm = element_grad_metric<FL_SGRID, STAGE, true>(vc, vo, valid, grad);
      } else {
        mm = hex_metric_fast_v<STAGE, true>(VC, VO, valid, grad);
      }
      const bool use = valid || STAGE == 0;  // gradient_functors.hpp:391
      const int cell = cj * TL_X + tx;
#pragma unroll
      for (int a = 0; a < 8; a++)
#pragma unroll
        for (int r = 0; r < 3; r++) ex[(a * 3 + r) * TL_NCELL + cell] = use ? grad[a][r] : 0.0;
      if (!warm) {
        val[0] += mm;
        inv += valid ? 0 : 1;
        if (em || ef) {
          const long long e = m.elem_off + (i0 + tx) + (long long)nci * ((j0 + cj) + (long long)ncj * k);
          if (em) em[e] = mm;
          if (ef) ef[e] = valid ? 0 : 1;
        }
      }
    }
    __syncthreads();

    // node phase: plane k is completed by the bottom contributions of layer k, the top ones start plane k+1
    for (int q = tid; q < TL_NNODE; q += TL_NTH) {
      const int f = info[q];
      if (!(f & 1)) continue;
      const int pi = q % TL_PX, pj = q / TL_PX;
      const int i = i0 + pi, j = j0 + pj;
      if (f & 2) {
        if (warm) continue;
        // seam node: individual contributions, slot = dk*4 + dj*2 + di around the node
        double* base;
        long long pstride, pos0;
        if (f & 4) {
          const int line = j / TL_TY - 1;
          pstride = (long long)m.nk * m.ni;
          base = seam.J + (long long)line * 24 * pstride;
          pos0 = (long long)k * m.ni + i;
        } else {
          const int line = i / TL_X - 1;
          pstride = (long long)m.nk * m.nj;
          base = seam.I + (long long)line * 24 * pstride;
          pos0 = (long long)k * m.nj + j;
        }
        const long long pos1 = pos0 + ((f & 4) ? m.ni : m.nj);  // the same node of plane k+1
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
          if (!(f & (16 << c4))) continue;
          const int di = c4 & 1, dj = c4 >> 1;
          const int cell = (pj - 1 + dj) * TL_X + (pi - 1 + di);
          const int a = (1 - di) + 2 * (1 - dj);
#pragma unroll
          for (int r = 0; r < 3; r++) {
            base[((4 + c4) * 3 + r) * pstride + pos0] = ex[(a * 3 + r) * TL_NCELL + cell];        // dk = 1 cell of plane k
            base[(c4 * 3 + r) * pstride + pos1] = ex[((a + 4) * 3 + r) * TL_NCELL + cell];        // dk = 0 cell of plane k+1
          }
        }
        continue;
      }
      double acc[3], top[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int r = 0; r < 3; r++) acc[r] = carry[r * TL_NNODE + q];
#pragma unroll
      for (int c4 = 0; c4 < 4; c4++) {
        if (!(f & (16 << c4))) continue;
        const int di = c4 & 1, dj = c4 >> 1;
        const int cell = (pj - 1 + dj) * TL_X + (pi - 1 + di);
        const int a = (1 - di) + 2 * (1 - dj);  // this node's local index in that cell (bottom face)
#pragma unroll
        for (int r = 0; r < 3; r++) {
          acc[r] += ex[(a * 3 + r) * TL_NCELL + cell];
          top[r] += ex[((a + 4) * 3 + r) * TL_NCELL + cell];
        }
      }
#pragma unroll
      for (int r = 0; r < 3; r++) carry[r * TL_NNODE + q] = top[r];
      if (!warm) {
        const long long n = m.node_off + i + (long long)m.ni * j + plane_nodes * k;
        const bool fx = fixed[n] != 0;  // fixup: gradient_functors.hpp:223-243
#pragma unroll
        for (int r = 0; r < 3; r++) g[r * stride + n] = fx ? 0.0 : acc[r];
      }
    }
    // (the __syncthreads at the top of the next layer orders this phase before ex / carry are written again)
  }
  cp_async_wait<0>();
  if (ke == nck) {  // top plane of the block: only the cells below it exist
    __syncthreads();
    for (int q = tid; q < TL_NNODE; q += TL_NTH) {
      const int f = info[q];
      if (!(f & 1) || (f & 2)) continue;
      const int pi = q % TL_PX, pj = q / TL_PX;
      const long long n = m.node_off + (i0 + pi) + (long long)m.ni * (j0 + pj) + plane_nodes * nck;
      const bool fx = fixed[n] != 0;
#pragma unroll
      for (int r = 0; r < 3; r++) g[r * stride + n] = fx ? 0.0 : carry[r * TL_NNODE + q];
    }
  }
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

// seam nodes: the 8 adjacent cells' contributions in (dk,dj,di) order, as k_grad_gather_sgrid adds them
__global__ void __launch_bounds__(NT) k_sgrid_seam_sum(const MeshDev m, const SeamDev seam, const uint8_t* __restrict__ fixed,
                                                       double* __restrict__ g, const long long stride, const long long nJ,
                                                       const long long nI) {
  const long long t = (long long)blockIdx.x * NT + threadIdx.x;
  int i, j, k;
  const double* base;
  long long pstride, pos;
  if (t < nJ) {
    i = (int)(t % m.ni);
    const long long u = t / m.ni;
    k = (int)(u % m.nk);
    const int line = (int)(u / m.nk);
    j = (line + 1) * TL_TY;
    pstride = (long long)m.nk * m.ni;
    base = seam.J + (long long)line * 24 * pstride;
    pos = (long long)k * m.ni + i;
  } else if (t < nJ + nI) {
    const long long v = t - nJ;
    j = (int)(v % m.nj);
    const long long u = v / m.nj;
    k = (int)(u % m.nk);
    const int line = (int)(u / m.nk);
    i = (line + 1) * TL_X;
    if ((j % TL_TY == 0) && j > 0 && j < m.nj - 1) return;  // crossing: summed with the j line
    pstride = (long long)m.nk * m.nj;
    base = seam.I + (long long)line * 24 * pstride;
    pos = (long long)k * m.nj + j;
  } else {
    return;
  }
  const long long n = m.node_off + i + (long long)m.ni * (j + (long long)m.nj * k);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (!fixed[n]) {
    const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
#pragma unroll
    for (int s8 = 0; s8 < 8; s8++) {
      const int ci = i - 1 + (s8 & 1), cj = j - 1 + ((s8 >> 1) & 1), ck = k - 1 + (s8 >> 2);
      if (ci < 0 || ci >= nci || cj < 0 || cj >= ncj || ck < 0 || ck >= nck) continue;
      a0 += __ldg(&base[(s8 * 3 + 0) * pstride + pos]);
      a1 += __ldg(&base[(s8 * 3 + 1) * pstride + pos]);
      a2 += __ldg(&base[(s8 * 3 + 2) * pstride + pos]);
    }
  }
  g[n] = a0;
  g[stride + n] = a1;
  g[2 * stride + n] = a2;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
  }
  return n;
}

// k-chunk length that minimises (rounds of CTAs over the SMs) x (layers per CTA incl. the warm-up layer)
static int tl_pick_chunk(long long ntiles, int nck) {
  const int nsm = sm_count();
  long long best = -1;
  int best_kc = nck;
  for (int nch = 1; nch <= (nck + 7) / 8; nch++) {
    const int kc = (nck + nch - 1) / nch;
    const int nchunks = (nck + kc - 1) / kc;
    const long long rounds = (ntiles * nchunks + nsm - 1) / nsm;
    const long long cost = rounds * (kc + (nchunks > 1 ? 1 : 0));
    if (best < 0 || cost < best) {
      best = cost;
      best_kc = kc;
    }
  }
  return best_kc;
}

static size_t tl_seam_doubles_J(const MeshDev& m) { return (size_t)tl_lines(m.nj, TL_TY) * 24 * m.nk * m.ni; }
static size_t tl_seam_doubles_I(const MeshDev& m) { return (size_t)tl_lines(m.ni, TL_X) * 24 * m.nk * m.nj; }

// bytes of seam scratch the fused structured gradient needs for this block
static size_t sgrid_fused_scratch_bytes(const MeshDev& m) {
  return (tl_seam_doubles_J(m) + tl_seam_doubles_I(m)) * sizeof(double) + 64;
}

static bool launch_grad_fused_sgrid(cudaStream_t st, const MeshDev& m, int stage, const double* x, const double* w,
                                    long long stride, const uint8_t* fixed, double* g, const Reduce& R, int slot,
                                    double* scratch, double* em, int32_t* ef) {
  if (m.kind != KIND_SGRID || m.ni < 2 || m.nj < 2 || m.nk < 2 || !scratch) return false;
  if (((uintptr_t)x & 15) || ((uintptr_t)w & 15) || (stride & 1)) return false;  // 16-byte staging needs even plane starts
  const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
  const int ntx = (nci + TL_X - 1) / TL_X, nty = (ncj + TL_TY - 1) / TL_TY;
  const long long ntiles = (long long)ntx * nty;
  const int kc = tl_pick_chunk(ntiles, nck);
  const int nchunks = (nck + kc - 1) / kc;
  const long long grid = ntiles * nchunks;
  if (grid > R.max_blocks || grid > 0x7fffffffLL) return false;
  constexpr size_t smem = (size_t)(3 * 6 * TL_PY * TL_ROWP + 24 * TL_NCELL + 3 * TL_NNODE) * sizeof(double) + TL_NNODE * sizeof(int);
  SeamDev sd;
  sd.J = scratch;
  sd.I = scratch + tl_seam_doubles_J(m);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_sgrid_tile_grad<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_sgrid_tile_grad<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const dim3 block(TL_X, TL_Y, 1);
  if (stage == 0)
    k_sgrid_tile_grad<0><<<(unsigned)grid, block, smem, st>>>(m, x, w, stride, fixed, g, sd, ntx, nty, kc, R, slot, em, ef);
  else
    k_sgrid_tile_grad<1><<<(unsigned)grid, block, smem, st>>>(m, x, w, stride, fixed, g, sd, ntx, nty, kc, R, slot, em, ef);
  const long long nJ = (long long)tl_lines(m.nj, TL_TY) * m.nk * m.ni, nI = (long long)tl_lines(m.ni, TL_X) * m.nk * m.nj;
  if (nJ + nI > 0) k_sgrid_seam_sum<<<(unsigned)((nJ + nI + NT - 1) / NT), NT, 0, st>>>(m, sd, fixed, g, stride, nJ, nI);
  return true;
}
