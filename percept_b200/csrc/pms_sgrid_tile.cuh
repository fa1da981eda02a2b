// pms_sgrid_tile.cuh — structured-block tile kernel (included by pms_kernels.cu inside namespace PMS_NS).
//
// One-pass fused metric + nodal gradient of a StructuredGrid block: replaces the element pass + 192 B/cell scratch +
// node gather of the two-pass scheme (reference functor: gradient_functors.hpp:349-407, whose 24 Kokkos::atomic_add per
// cell become an in-CTA gather in a FIXED order).
//
//  * A CTA owns a tile of 32 x 8 cells and marches a chunk of k-layers.  The node planes of x and of the reference mesh
//    w (the tile's cells + a one-node halo: 33 x 9 nodes, 6 component planes) are staged in shared memory by TMA bulk
//    copies (cp.async.bulk -> SASS UBLKCP, one copy per node row: the 16-byte-aligned 288-byte span around its 33
//    nodes), four rotating plane buffers, completion on mbarriers.  Each node is fetched once per CTA instead of 8
//    times, and there is no per-vertex address arithmetic.
//  * Warp specialisation: 8 compute warps (setmaxnreg 232) evaluate one cell per thread per layer, vertices read in place
//    from the staged planes, and write the cell's 24 contributions to one of TWO exchange buffers in shared memory; they
//    execute no CTA barrier.  4 helper warps (setmaxnreg 40) issue the bulk copies two-three planes ahead and run the
//    node phase of layer k while the compute warps are already in layer k+1.  (A first version alternated a cell phase
//    and a node phase between __syncthreads: ncu showed the FP64 pipe idle 30 % of the time — node phase 13 %, warps
//    waiting at the barrier for the slower warp of their scheduler pair 17 % — and it was no faster than the two-pass
//    path; profiles/r02_sgrid_tile.md.)
//  * Node phase: every node of the tile's patch sums the contributions of its adjacent cells in the fixed (dk,dj,di)
//    order of k_grad_gather_sgrid (= ascending cell id = the serial reference loop): the four cells of layer k-1 were
//    summed one layer earlier into a carry held in shared memory, the four cells of layer k are added to it and the node
//    of plane k is final -> one coalesced store per component, fixed nodes zeroed (gradient_functors.hpp:223-243).
//    The sequence of floating-point additions per node is exactly that of the two-pass path, so the result is bitwise
//    identical to it (and, in the strict build, to the CPU oracle).
//  * Nodes on a line between two tiles ("seam" nodes: 1/32 + 1/8 of all nodes) receive contributions from two or four
//    CTAs.  Adding per-tile partial sums would change the order of additions, so each CTA stores the individual
//    contributions of its own cells into a compact seam scratch (slot = the cell's position (dk,dj,di) around the
//    node) and k_sgrid_seam_sum adds the 8 slots in order.  No atomics, no index tables; the scratch is
//    (1/32 + 1/8) * 192 B per node instead of 192 B per cell.
//  * k is cut into chunks so that tiles x chunks fill the SMs evenly; a chunk that does not start at k = 0 first
//    re-evaluates the layer below it to build the carry (nothing of that layer is stored or counted).
#pragma once

constexpr int TL_X = 32;               // cells of a tile in i = lanes of a warp
constexpr int TL_PX = TL_X + 1;        // nodes of its patch in i
constexpr int TL_ROWP = 36;            // staged row: 18 x 16 B covering 33 nodes from an even node index


__host__ __device__ inline int tl_lines(int n, int t) { return n >= 3 ? (n - 2) / t : 0; }  // interior tile lines among n nodes

struct SeamDev {
  double* J;  // [lineJ][slot 8][comp 3][nk][ni]  nodes with j on a tile line (all i)
  double* I;  // [lineI][slot 8][comp 3][nk][nj]  nodes with i on a tile line (j not on a tile line)
};

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier -------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one staged node plane: NARR*3 component planes of TL_PY rows of TL_ROWP doubles, then TL_PY rows of 64 mask bytes

constexpr int WS_TY = 8;                                  // cell rows of a tile
constexpr int WS_PY = WS_TY + 1;                          // node rows of its patch
constexpr int WS_NC = 256, WS_NH = 128;                   // compute / helper threads
constexpr int WS_NNODE = TL_PX * WS_PY;                   // 297
constexpr int WS_NRND = (WS_NNODE + WS_NH - 1) / WS_NH;   // node rounds per helper thread (3)
constexpr int WS_NPL = 4;                                 // staged plane buffers (planes k, k+1 in use, k+2, k+3 in flight)
constexpr int WS_PLANE = 6 * WS_PY * TL_ROWP;             // doubles per staged plane
// exchange buffer: per (local node a, component r) a grid of (WS_TY + 2) x (TL_X + 2) cells with a zero border, so the
// node phase reads the four cells around a node unconditionally (+0.0 does not change a sum that started at +0.0)
constexpr int WS_EXP = TL_X + 2;
constexpr int WS_EXPL = (WS_TY + 2) * WS_EXP;
constexpr int WS_EX = 24 * WS_EXPL;
// 256 x 232 + 128 x 40 = 64512 = the 384 x 168 registers the CTA is launched with (setmaxnreg.inc blocks for ever when the
// pool cannot supply the increase)
constexpr int WS_REG_C = 232, WS_REG_H = 40;
constexpr size_t WS_SMEM = (size_t)(WS_NPL * WS_PLANE + 2 * WS_EX + 3 * WS_NNODE + 1) * sizeof(double) + 16 * 8 +
                           WS_NRND * WS_NH * sizeof(int);

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The helper threads stage node plane kp (x, w: 3 components each) into buf, completing on bar (WS_NH arrivals): every
// node row is ONE bulk copy of the 16-byte-aligned span covering its 33 nodes (36 doubles from the even node index at
// or below the row's first node).  Helper thread h < 54 issues row h.
__device__ __forceinline__ void ws_stage_plane(double* buf, unsigned long long* bar, const MeshDev& m, const double* x,
                                               const double* w, const long long stride, const int i0, const int j0,
                                               const int kp, const int h) {
  const int ac = h / WS_PY, row = h % WS_PY;  // ac = array * 3 + component
  const int rows = min(WS_PY, m.nj - j0);
  if (kp < m.nk && ac < 6 && row < rows) {
    const long long nrow = m.node_off + i0 + (long long)m.ni * ((j0 + row) + (long long)m.nj * kp);
    const long long a = nrow & ~1LL;
    const long long n = stride - a < TL_ROWP ? stride - a : TL_ROWP;
    const unsigned nb = (unsigned)n * 8u;
    mbar_arrive_expect_tx(bar, nb);
    bulk_g2s(buf + (ac * WS_PY + row) * TL_ROWP, (ac < 3 ? x : w) + (ac % 3) * stride + a, nb, bar);
  } else {
    mbar_arrive(bar);
  }
}

struct WsVerts {  // vertices of one cell read in place from two staged planes: V(p, r), p = di + 2 dj + 4 dk
  const double* pl[2];
  int o[2][2];
  PMS_DI double operator()(int p, int r) const { return pl[p >> 2][r * (WS_PY * TL_ROWP) + o[p >> 2][(p >> 1) & 1] + (p & 1)]; }
};

template <int STAGE>
__global__ void __launch_bounds__(WS_NC + WS_NH, 1) k_sgrid_tile_ws(const MeshDev m, const double* __restrict__ x,
                                                                    const double* __restrict__ w, const long long stride,
                                                                    const uint8_t* __restrict__ fixed, double* __restrict__ g,
                                                                    const SeamDev seam, const int ntx, const int nty,
                                                                    const int kc, const Reduce R, const int slot,
                                                                    double* __restrict__ em, int32_t* __restrict__ ef) {
  extern __shared__ double sm[];
  double* const planes = sm;                        // [WS_NPL][WS_PLANE]
  double* const ex = sm + WS_NPL * WS_PLANE;        // [2][24][WS_EXPL]
  double* const carry = ex + 2 * WS_EX;             // [3][WS_NNODE]
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(carry + 3 * WS_NNODE + 1);
  unsigned long long* const pfull = bars;                 // [WS_NPL] TMA complete_tx: the plane has landed
  unsigned long long* const pempty = bars + WS_NPL;       // [WS_NPL] every compute thread is done with the plane
  unsigned long long* const xfull = bars + 2 * WS_NPL;    // [2] the layer's contributions are in the exchange buffer
  unsigned long long* const xempty = xfull + 2;           // [2] ... and have been gathered
  int* const ninfo = reinterpret_cast<int*>(bars + 16);   // [WS_NRND * WS_NH]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % (ntx * nty), chunk = blockIdx.x / (ntx * nty);
  const int i0 = (tile % ntx) * TL_X, j0 = (tile / ntx) * WS_TY;
  const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
  const int actx = min(TL_X, nci - i0), acty = min(WS_TY, ncj - j0);  // cells of this tile
  const int kb = chunk * kc, ke = min(kb + kc, nck);
  const int kstart = kb > 0 ? kb - 1 : 0;  // one warm-up layer below the chunk builds the carry
  const int nlay = ke - kstart;

  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < WS_NPL; b++) {
      mbar_init(&pfull[b], WS_NH);
      mbar_init(&pempty[b], WS_NC);
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {
      mbar_init(&xfull[b], WS_NC);
      mbar_init(&xempty[b], WS_NH);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // per patch node: bit0 exists and is summed here, bit1 seam (bit2: on a j line, else on an i line), bits 8..13 pi,
  // bits 14..17 pj
  for (int q = tid; q < WS_NRND * WS_NH; q += WS_NC + WS_NH) {
    int f = 0;
    if (q < WS_NNODE) {
      const int pi = q % TL_PX, pj = q / TL_PX;
      const int i = i0 + pi, j = j0 + pj;
      f = (pi << 8) | (pj << 14);
      // a node whose tile cells are all clipped away belongs to the neighbouring tile
      const bool mine = (pi - 1 < actx && pj - 1 < acty) && i < m.ni && j < m.nj;
      if (mine) {
        const bool sj = (j % WS_TY == 0) && j > 0 && j < m.nj - 1, si = (i % TL_X == 0) && i > 0 && i < m.ni - 1;
        f |= sj ? 7 : (si ? 3 : 1);
      }
    }
    ninfo[q] = f;
  }
  for (int q = tid; q < 2 * WS_EX + 3 * WS_NNODE; q += WS_NC + WS_NH) ex[q] = 0.0;  // both exchange buffers and the carry
  __syncthreads();
  const long long plane_nodes = (long long)m.ni * m.nj;

  if (tid >= WS_NC) {
    // ------------------------------------------------------------------------------ helper warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_H));
    const int h = tid - WS_NC;
#pragma unroll 1
    for (int p = 0; p < WS_NPL - 1; p++) ws_stage_plane(planes + p * WS_PLANE, &pfull[p], m, x, w, stride, i0, j0, kstart + p, h);
#pragma unroll 1
    for (int rel = 0; rel < nlay; rel++) {
      const int k = kstart + rel;
      const bool warm = k < kb;
      {  // plane rel+3 goes where plane rel-1 was
        const int pr = rel + WS_NPL - 1, u = pr / WS_NPL;
        if (u >= 1) mbar_wait(&pempty[pr % WS_NPL], (unsigned)((u - 1) & 1));
        ws_stage_plane(planes + (pr % WS_NPL) * WS_PLANE, &pfull[pr % WS_NPL], m, x, w, stride, i0, j0, kstart + pr, h);
      }
      const long long nbase = m.node_off + i0 + (long long)m.ni * j0 + plane_nodes * k;
      // fixed flags of this thread's nodes of plane k, in flight during the wait (gradient_functors.hpp:223-243)
      unsigned fxbits = 0;
      if (!warm) {
#pragma unroll
        for (int r = 0; r < WS_NRND; r++) {
          const int f = ninfo[h + WS_NH * r];
          if ((f & 3) == 1) fxbits |= (fixed[nbase + (long long)m.ni * ((f >> 14) & 15) + ((f >> 8) & 63)] ? 1u : 0u) << r;
        }
      }
      mbar_wait(&xfull[rel & 1], (unsigned)((rel >> 1) & 1));  // the cells of layer k
      const double* const exb = ex + (rel & 1) * WS_EX;
      // node phase: plane k is completed by the bottom contributions of layer k, the top ones start plane k+1
#pragma unroll 1
      for (int r = 0; r < WS_NRND; r++) {
        const int q = h + WS_NH * r;
        const int f = ninfo[q];
        if (!(f & 1)) continue;
        const int pi = (f >> 8) & 63, pj = (f >> 14) & 15;
        // contribution (local node a, comp c) of cell (dj, di) around this node at exn[(a * 3 + c) * WS_EXPL + dj * WS_EXP + di]
        const double* const exn = exb + pj * WS_EXP + pi;
        if (f & 2) {
          if (warm) continue;
          // seam node: individual contributions, slot = dk*4 + dj*2 + di around the node
          double* base;
          long long pstride, pos0;
          if (f & 4) {
            pstride = (long long)m.nk * m.ni;
            base = seam.J + (long long)((j0 + pj) / WS_TY - 1) * 24 * pstride;
            pos0 = (long long)k * m.ni + (i0 + pi);
          } else {
            pstride = (long long)m.nk * m.nj;
            base = seam.I + (long long)((i0 + pi) / TL_X - 1) * 24 * pstride;
            pos0 = (long long)k * m.nj + (j0 + pj);
          }
          const long long pos1 = pos0 + ((f & 4) ? m.ni : m.nj);  // the same node of plane k+1
#pragma unroll
          for (int c4 = 0; c4 < 4; c4++) {
            const int di = c4 & 1, dj = c4 >> 1;
            if (pi - 1 + di < 0 || pi - 1 + di >= actx || pj - 1 + dj < 0 || pj - 1 + dj >= acty) continue;  // not a cell of this tile
            const int a = (1 - di) + 2 * (1 - dj);
#pragma unroll
            for (int c = 0; c < 3; c++) {
              base[((4 + c4) * 3 + c) * pstride + pos0] = exn[(a * 3 + c) * WS_EXPL + dj * WS_EXP + di];      // dk = 1 cell of plane k
              base[(c4 * 3 + c) * pstride + pos1] = exn[((a + 4) * 3 + c) * WS_EXPL + dj * WS_EXP + di];      // dk = 0 cell of plane k+1
            }
          }
          continue;
        }
        const long long n = nbase + (long long)m.ni * pj + pi;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          double acc = carry[c * WS_NNODE + q], top = 0.0;
#pragma unroll
          for (int c4 = 0; c4 < 4; c4++) {
            const int di = c4 & 1, dj = c4 >> 1;
            const int a = (1 - di) + 2 * (1 - dj);  // this node's local index in that cell (bottom face)
            acc += exn[(a * 3 + c) * WS_EXPL + dj * WS_EXP + di];
            top += exn[((a + 4) * 3 + c) * WS_EXPL + dj * WS_EXP + di];
          }
          carry[c * WS_NNODE + q] = top;
          if (!warm) g[c * stride + n] = ((fxbits >> r) & 1u) ? 0.0 : acc;
        }
      }
      mbar_arrive(&xempty[rel & 1]);
    }
    if (ke == nck) {  // top plane of the block: only the cells below it exist
      const long long nbase = m.node_off + i0 + (long long)m.ni * j0 + plane_nodes * nck;
#pragma unroll 1
      for (int r = 0; r < WS_NRND; r++) {
        const int f = ninfo[h + WS_NH * r];
        if ((f & 3) != 1) continue;
        const long long n = nbase + (long long)m.ni * ((f >> 14) & 15) + ((f >> 8) & 63);
        const bool fx = fixed[n] != 0;
#pragma unroll
        for (int c = 0; c < 3; c++) g[c * stride + n] = fx ? 0.0 : carry[c * WS_NNODE + h + WS_NH * r];
      }
    }
    // no bulk copy may be in flight when the CTA exits: the last WS_NPL-2 staged planes were never consumed
    for (int pr = nlay + 1; pr < nlay + WS_NPL - 1; pr++) mbar_wait(&pfull[pr % WS_NPL], (unsigned)((pr / WS_NPL) & 1));
    return;
  }

  // ---------------------------------------------------------------------------------- compute warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REG_C));
  const int tx = tid & 31, ty = tid >> 5;
  const bool active = tx < actx && ty < acty;
  double val[1] = {0.0};
  unsigned long long inv = 0;
  const int par0 = (int)((m.node_off + i0) & 1LL), pni = m.ni & 1, pnj = m.nj & 1;
  mbar_wait(&pfull[0], 0);
#pragma unroll 1
  for (int rel = 0; rel < nlay; rel++) {
    const int k = kstart + rel;
    mbar_wait(&pfull[(rel + 1) % WS_NPL], (unsigned)(((rel + 1) / WS_NPL) & 1));
    if (rel >= 2) mbar_wait(&xempty[rel & 1], (unsigned)(((rel >> 1) - 1) & 1));
    if (active) {
      const double* const p0 = planes + (rel % WS_NPL) * WS_PLANE;
      const double* const p1 = planes + ((rel + 1) % WS_NPL) * WS_PLANE;
      WsVerts VC, VO;
      VC.pl[0] = p0;
      VC.pl[1] = p1;
      VO.pl[0] = p0 + 3 * WS_PY * TL_ROWP;
      VO.pl[1] = p1 + 3 * WS_PY * TL_ROWP;
#pragma unroll
      for (int dk = 0; dk < 2; dk++)
#pragma unroll
        for (int dj = 0; dj < 2; dj++) {
          // parity of the row's first node: the row is staged from the even node index at or below it
          const int par = (par0 + pni * ((j0 + ty + dj + pnj * (k + dk)) & 1)) & 1;
          const int o = (ty + dj) * TL_ROWP + tx + par;
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
      double* const exc = ex + (rel & 1) * WS_EX + (ty + 1) * WS_EXP + (tx + 1);
      if (use || !PMS_USE_BRANCH) {
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int r = 0; r < 3; r++) exc[(a * 3 + r) * WS_EXPL] = (PMS_USE_BRANCH || use) ? grad[a][r] : 0.0;
      } else {
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int r = 0; r < 3; r++) exc[(a * 3 + r) * WS_EXPL] = 0.0;
      }
      if (k >= kb) {
        val[0] += mm;
        inv += valid ? 0 : 1;
        if (em || ef) {
          const long long e = m.elem_off + (i0 + tx) + (long long)nci * ((j0 + ty) + (long long)ncj * k);
          if (em) em[e] = mm;
          if (ef) ef[e] = valid ? 0 : 1;
        }
      }
    }
    mbar_arrive(&xfull[rel & 1]);
    mbar_arrive(&pempty[rel % WS_NPL]);
  }
  grid_reduce<OpSum, 1, true, 1, WS_NC>(val, inv, R, slot);
}

// ------------------------------------------------------------------------------------------------------------------
// Metric-only tile kernel (the line search's passes, total_element_metric.cpp:271-307 / Def.hpp:330-388): the same TMA
// plane staging — x, the reference mesh w, the search direction s and the fixed mask, one bulk copy per node row — and
// the same compute / producer split, without exchange or node phase.  NB = 1: the metric at x + alpha*s (s may be
// absent: alpha = 0).  NB = 8 / 12 / 16: the metric at NB trial step lengths from the per-corner polynomial
// coefficients (hex_metric_multi).  The step is applied to unfixed nodes only (update_coordinates.cpp:255-256).
// ------------------------------------------------------------------------------------------------------------------
constexpr int TM_ROWS = WS_PY;                         // node rows of a tile's patch
template <int NARR>
struct TmPlane {
  static constexpr int DOUBLES = NARR * 3 * TM_ROWS * TL_ROWP + TM_ROWS * 8;  // + 64 mask bytes per row
};
#ifndef TM_CORNER_UNROLL
#define TM_CORNER_UNROLL 8  // (4: 3.87 / 5.21 ms, 1: 3.66 / 4.47 ms, 8: 3.31 / 4.46 ms for 8 / 12 alphas at 256^3)
#endif
constexpr int TM_REG_C = 240, TM_REG_H = 24;           // 256 x 240 + 128 x 24 = 64512 = 384 x 168
template <int NB>
struct TileAlphas {
  double a[NB];
};

// helper thread h issues row h of node plane kp: NARR*3 double rows (arrays a0, a1[, a2]) then the mask rows
template <int NARR>
__device__ __forceinline__ void tm_stage_plane(double* buf, unsigned long long* bar, const MeshDev& m, const double* a0,
                                               const double* a1, const double* a2, const uint8_t* fixed, const long long stride,
                                               const int i0, const int j0, const int kp, const int h) {
  const int ac = h / TM_ROWS, row = h % TM_ROWS;  // ac = array * 3 + component; ac == NARR*3: the fixed mask
  const int rows = min(TM_ROWS, m.nj - j0);
  if (kp < m.nk && ac <= NARR * 3 && row < rows) {
    const long long nrow = m.node_off + i0 + (long long)m.ni * ((j0 + row) + (long long)m.nj * kp);
    if (ac < NARR * 3) {
      const long long a = nrow & ~1LL;
      const long long n = stride - a < TL_ROWP ? stride - a : TL_ROWP;
      const double* src = ac < 3 ? a0 : (ac < 6 ? a1 : a2);
      const unsigned nb = (unsigned)n * 8u;
      mbar_arrive_expect_tx(bar, nb);
      bulk_g2s(buf + (ac * TM_ROWS + row) * TL_ROWP, src + (ac % 3) * stride + a, nb, bar);
    } else {
      const long long a = nrow & ~15LL;
      const long long n = stride - a < 64 ? stride - a : 64;
      mbar_arrive_expect_tx(bar, (unsigned)n);
      bulk_g2s(buf + NARR * 3 * TM_ROWS * TL_ROWP + row * 8, fixed + a, (unsigned)n, bar);
    }
  } else {
    mbar_arrive(bar);
  }
}

struct TmVerts {  // V(p, r) of one array from two staged planes (p may be a run-time value: selects, no indexed arrays)
  const double* pl[2];
  int o[2][2];
  PMS_DI double operator()(int p, int r) const {
    const double* b = (p & 4) ? pl[1] : pl[0];
    const int off = (p & 4) ? ((p & 2) ? o[1][1] : o[1][0]) : ((p & 2) ? o[0][1] : o[0][0]);
    return b[r * (TM_ROWS * TL_ROWP) + off + (p & 1)];
  }
};
struct TmVertsMasked {  // the search direction: zero at fixed nodes (bit p of fx)
  TmVerts v;
  unsigned fx;
  PMS_DI double operator()(int p, int r) const { return ((fx >> p) & 1u) ? 0.0 : v(p, r); }
};

template <int STAGE, int NB, bool WITH_S, int NF = NB>
__global__ void __launch_bounds__(WS_NC + WS_NH, 1) k_sgrid_tile_metric(const MeshDev m, const double* __restrict__ x,
                                                                        const double* __restrict__ s, const double* __restrict__ w,
                                                                        const long long stride, const uint8_t* __restrict__ fixed,
                                                                        const TileAlphas<NB> al, const int ntx, const int nty,
                                                                        const int kc, const Reduce R, const int slot,
                                                                        double* __restrict__ em, int32_t* __restrict__ ef) {
  constexpr int NARR = WITH_S ? 3 : 2;  // staged arrays: x, w[, s]
  constexpr int PLANE = TmPlane<NARR>::DOUBLES;
  extern __shared__ double sm[];
  double* const planes = sm;  // [WS_NPL][PLANE]
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sm + WS_NPL * PLANE);
  unsigned long long* const pfull = bars;            // [WS_NPL]
  unsigned long long* const pempty = bars + WS_NPL;  // [WS_NPL]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % (ntx * nty), chunk = blockIdx.x / (ntx * nty);
  const int i0 = (tile % ntx) * TL_X, j0 = (tile / ntx) * WS_TY;
  const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
  const int actx = min(TL_X, nci - i0), acty = min(WS_TY, ncj - j0);
  const int kb = chunk * kc, ke = min(kb + kc, nck);
  const int nlay = ke - kb;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < WS_NPL; b++) {
      mbar_init(&pfull[b], WS_NH);
      mbar_init(&pempty[b], WS_NC);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= WS_NC) {
    // ------------------------------------------------------------------------------ producer warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TM_REG_H));
    const int h = tid - WS_NC;
#pragma unroll 1
    for (int p = 0; p < WS_NPL - 1; p++)
      tm_stage_plane<NARR>(planes + p * PLANE, &pfull[p], m, x, w, s, fixed, stride, i0, j0, kb + p, h);
#pragma unroll 1
    for (int rel = 0; rel < nlay; rel++) {  // plane rel+3 goes where plane rel-1 was
      const int pr = rel + WS_NPL - 1, u = pr / WS_NPL;
      if (u >= 1) mbar_wait(&pempty[pr % WS_NPL], (unsigned)((u - 1) & 1));
      tm_stage_plane<NARR>(planes + (pr % WS_NPL) * PLANE, &pfull[pr % WS_NPL], m, x, w, s, fixed, stride, i0, j0, kb + pr, h);
    }
    // no bulk copy may be in flight when the CTA exits
    for (int pr = nlay + 1; pr < nlay + WS_NPL - 1; pr++) mbar_wait(&pfull[pr % WS_NPL], (unsigned)((pr / WS_NPL) & 1));
    return;
  }

  // ---------------------------------------------------------------------------------- compute warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TM_REG_C));
  const int tx = tid & 31, ty = tid >> 5;
  const bool active = tx < actx && ty < acty;
  double sums[NB];
  unsigned cnt[NB];
#pragma unroll
  for (int q = 0; q < NB; q++) {
    sums[q] = 0.0;
    cnt[q] = 0;
  }
  const int par0 = (int)((m.node_off + i0) & 1LL), pni = m.ni & 1, pnj = m.nj & 1;
  const long long rowbase = m.node_off + i0 + (long long)m.ni * (j0 + ty);  // first node of this thread's lower row, plane 0
  mbar_wait(&pfull[0], 0);
#pragma unroll 1
  for (int rel = 0; rel < nlay; rel++) {
    const int k = kb + rel;
    mbar_wait(&pfull[(rel + 1) % WS_NPL], (unsigned)(((rel + 1) / WS_NPL) & 1));
    if (active) {
      const double* const p0 = planes + (rel % WS_NPL) * PLANE;
      const double* const p1 = planes + ((rel + 1) % WS_NPL) * PLANE;
      TmVerts VC, VO, VS;
      VC.pl[0] = p0;
      VC.pl[1] = p1;
      VO.pl[0] = p0 + 3 * TM_ROWS * TL_ROWP;
      VO.pl[1] = p1 + 3 * TM_ROWS * TL_ROWP;
      VS.pl[0] = p0 + 6 * TM_ROWS * TL_ROWP;
      VS.pl[1] = p1 + 6 * TM_ROWS * TL_ROWP;
      unsigned fx = 0;
#pragma unroll
      for (int dk = 0; dk < 2; dk++)
#pragma unroll
        for (int dj = 0; dj < 2; dj++) {
          const int par = (par0 + pni * ((j0 + ty + dj + pnj * (k + dk)) & 1)) & 1;
          const int o = (ty + dj) * TL_ROWP + tx + par;
          VC.o[dk][dj] = o;
          VO.o[dk][dj] = o;
          VS.o[dk][dj] = o;
          if (WITH_S) {  // fixed flags of the row's two nodes: the mask row was staged from the 16-byte boundary below its first node
            const uint8_t* mrow = reinterpret_cast<const uint8_t*>((dk ? p1 : p0) + NARR * 3 * TM_ROWS * TL_ROWP) + (ty + dj) * 64 +
                                  (int)((rowbase + (long long)m.ni * dj + (long long)m.ni * m.nj * (k + dk)) & 15LL) + tx;
            fx |= (mrow[0] ? 1u : 0u) << (dk * 4 + dj * 2);
            fx |= (mrow[1] ? 1u : 0u) << (dk * 4 + dj * 2 + 1);
          }
        }
      const long long e = m.elem_off + (i0 + tx) + (long long)nci * ((j0 + ty) + (long long)ncj * k);
      if (NB == 1) {
        double vc[8][3], vo[8][3];
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
          for (int r = 0; r < 3; r++) {
            double xv = VC(a, r);
            if (WITH_S) {
              if (!((fx >> a) & 1u)) {
                const double dt = al.a[0] * VS(a, r);  // update_coordinates.cpp:255-256
                xv += dt;
              }
            }
            vc[a][r] = xv;
            vo[a][r] = VO(a, r);
          }
        bool valid;
        double mm;
        if (STRICT) {
          mm = element_metric<FL_SGRID, STAGE, true>(vc, vo, valid);
        } else if (STAGE == 1) {
          mm = hex_shapeb1_metric_ilp(vc, vo, valid);
        } else {
          double dummy[8][3];
          mm = hex_metric_fast<STAGE, false>(vc, vo, valid, dummy);
        }
        sums[0] += mm;
        cnt[0] += valid ? 0 : 1;
        if (em) em[e] = mm;
        if (ef) ef[e] = valid ? 0 : 1;
      } else {
        TmVertsMasked VSM;
        VSM.v = VS;
        VSM.fx = fx;
        double es[NB];
#pragma unroll
        for (int q = 0; q < NB; q++) es[q] = 0.0;
        unsigned invmask;
        hex_metric_multi<STAGE, NB, TM_CORNER_UNROLL, NF>(VC, VSM, VO, al.a, es, invmask);
#pragma unroll
        for (int q = 0; q < NB; q++) {
          sums[q] += es[q];
          cnt[q] += (invmask >> q) & 1u;
        }
      }
    }
    mbar_arrive(&pempty[rel % WS_NPL]);
  }
  if (NB == 1) {
    double v[1] = {sums[0]};
    grid_reduce<OpSum, 1, true, 1, WS_NC>(v, (unsigned long long)cnt[0], R, slot);
  } else {
    double v[2 * NB];
#pragma unroll
    for (int q = 0; q < NB; q++) {
      v[q] = sums[q];
      v[NB + q] = (double)cnt[q];
    }
    grid_reduce<OpSum, 2 * NB, false, 1, WS_NC>(v, 0ull, R, slot);
  }
}

// seam nodes: the 8 adjacent cells' contributions in (dk,dj,di) order, as k_grad_gather_sgrid adds them
template <int TJ>
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
    j = (line + 1) * TJ;
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
    if ((j % TJ == 0) && j > 0 && j < m.nj - 1) return;  // crossing: summed with the j line
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

static size_t tl_seam_doubles_J(const MeshDev& m) { return (size_t)tl_lines(m.nj, WS_TY) * 24 * m.nk * m.ni; }
static size_t tl_seam_doubles_I(const MeshDev& m) { return (size_t)tl_lines(m.ni, TL_X) * 24 * m.nk * m.nj; }

// bytes of seam scratch the fused structured gradient needs for this block
static size_t sgrid_fused_scratch_bytes(const MeshDev& m) {
  return (tl_seam_doubles_J(m) + tl_seam_doubles_I(m)) * sizeof(double) + 64;
}

static bool launch_grad_fused_sgrid(cudaStream_t st, const MeshDev& m, int stage, const double* x, const double* w,
                                    long long stride, const uint8_t* fixed, double* g, const Reduce& R, int slot,
                                    double* scratch, double* em, int32_t* ef) {
  if (m.kind != KIND_SGRID || m.ni < 2 || m.nj < 2 || m.nk < 2 || !scratch) return false;
  // 16-byte bulk copies need even plane starts
  if (((uintptr_t)x & 15) || ((uintptr_t)w & 15) || (stride & 1)) return false;
  const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
  const int ntx = (nci + TL_X - 1) / TL_X, nty = (ncj + WS_TY - 1) / WS_TY;
  const long long ntiles = (long long)ntx * nty;
  const int kc = tl_pick_chunk(ntiles, nck);
  const int nchunks = (nck + kc - 1) / kc;
  const long long grid = ntiles * nchunks;
  if (grid > R.max_blocks || grid > 0x7fffffffLL) return false;
  SeamDev sd;
  sd.J = scratch;
  sd.I = scratch + tl_seam_doubles_J(m);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_sgrid_tile_ws<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
    cudaFuncSetAttribute(k_sgrid_tile_ws<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
    attr_set = true;
  }
  if (stage == 0)
    k_sgrid_tile_ws<0><<<(unsigned)grid, WS_NC + WS_NH, WS_SMEM, st>>>(m, x, w, stride, fixed, g, sd, ntx, nty, kc, R, slot, em, ef);
  else
    k_sgrid_tile_ws<1><<<(unsigned)grid, WS_NC + WS_NH, WS_SMEM, st>>>(m, x, w, stride, fixed, g, sd, ntx, nty, kc, R, slot, em, ef);
  const long long nJ = (long long)tl_lines(m.nj, WS_TY) * m.nk * m.ni, nI = (long long)tl_lines(m.ni, TL_X) * m.nk * m.nj;
  if (nJ + nI > 0)
    k_sgrid_seam_sum<WS_TY><<<(unsigned)((nJ + nI + NT - 1) / NT), NT, 0, st>>>(m, sd, fixed, g, stride, nJ, nI);
  return true;
}

template <int STAGE, int NB, bool WITH_S, int NF = NB>
static bool launch_sgrid_tile_metric3(cudaStream_t st, const MeshDev& m, const double* x, const double* s, const double* w,
                                      long long stride, const uint8_t* fixed, const double* alphas, const Reduce& R, int slot,
                                      double* em, int32_t* ef) {
  constexpr size_t smem = (size_t)(WS_NPL * TmPlane<WITH_S ? 3 : 2>::DOUBLES) * sizeof(double) + 2 * WS_NPL * 8;
  auto kern = k_sgrid_tile_metric<STAGE, NB, WITH_S, NF>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  if (!x) return true;  // prepare only (prepare_line_search_kernels): the attribute call has loaded the kernel
  const int nci = m.ni - 1, ncj = m.nj - 1, nck = m.nk - 1;
  const int ntx = (nci + TL_X - 1) / TL_X, nty = (ncj + WS_TY - 1) / WS_TY;
  const long long ntiles = (long long)ntx * nty;
  // chunks of k: no warm-up layer here, but the same balance rule
  const int kc = tl_pick_chunk(ntiles, nck);
  const long long grid = ntiles * ((nck + kc - 1) / kc);
  if (grid > R.max_blocks || grid > 0x7fffffffLL) return false;
  TileAlphas<NB> al;
  for (int q = 0; q < NB; q++) al.a[q] = alphas[q];
  kern<<<(unsigned)grid, WS_NC + WS_NH, smem, st>>>(m, x, s, w, stride, fixed, al, ntx, nty, kc, R, slot, em, ef);
  return true;
}

static bool sgrid_tile_ok(const MeshDev& m, const double* x, const double* s, const double* w, long long stride, const uint8_t* fixed) {
  return !(m.kind != KIND_SGRID || m.ni < 2 || m.nj < 2 || m.nk < 2 || ((uintptr_t)x & 15) || ((uintptr_t)w & 15) ||
           ((uintptr_t)s & 15) || ((uintptr_t)fixed & 15) || (stride & 15) || !fixed);
}

// metric at x + alpha*s (s null or alpha 0: at x); returns false if this block cannot take the tile path
static bool launch_sgrid_tile_metric(cudaStream_t st, const MeshDev& m, int stage, const double* x, const double* s, const double* w,
                                     long long stride, double alpha, const uint8_t* fixed, const Reduce& R, int slot, double* em,
                                     int32_t* ef) {
  if (!sgrid_tile_ok(m, x, s, w, stride, fixed)) return false;
  const bool with_s = s && alpha != 0.0;
  if (stage == 0)
    return with_s ? launch_sgrid_tile_metric3<0, 1, true>(st, m, x, s, w, stride, fixed, &alpha, R, slot, em, ef)
                  : launch_sgrid_tile_metric3<0, 1, false>(st, m, x, nullptr, w, stride, fixed, &alpha, R, slot, em, ef);
  return with_s ? launch_sgrid_tile_metric3<1, 1, true>(st, m, x, s, w, stride, fixed, &alpha, R, slot, em, ef)
                : launch_sgrid_tile_metric3<1, 1, false>(st, m, x, nullptr, w, stride, fixed, &alpha, R, slot, em, ef);
}

template <int NB, int NF>
static bool launch_sgrid_tile_multi_nf(cudaStream_t st, const MeshDev& m, int nf, const double* x, const double* s, const double* w,
                                       long long stride, const uint8_t* fixed, const double* alphas, const Reduce& R, int slot) {
  if constexpr (NF >= NB) {
    return launch_sgrid_tile_metric3<1, NB, true>(st, m, x, s, w, stride, fixed, alphas, R, slot, nullptr, nullptr);
  } else {
    if (nf <= NF) return launch_sgrid_tile_metric3<1, NB, true, NF>(st, m, x, s, w, stride, fixed, alphas, R, slot, nullptr, nullptr);
    return launch_sgrid_tile_multi_nf<NB, NF + 2>(st, m, nf, x, s, w, stride, fixed, alphas, R, slot);
  }
}
template <int NB>
static bool launch_sgrid_tile_multi1(cudaStream_t st, const MeshDev& m, int stage, int nf, const double* x, const double* s, const double* w,
                                     long long stride, const uint8_t* fixed, const double* alphas, const Reduce& R, int slot) {
  if (stage == 0) return launch_sgrid_tile_metric3<0, NB, true>(st, m, x, s, w, stride, fixed, alphas, R, slot, nullptr, nullptr);
  return launch_sgrid_tile_multi_nf<NB, 4>(st, m, nf, x, s, w, stride, fixed, alphas, R, slot);
}
static bool launch_sgrid_tile_multi(cudaStream_t st, const MeshDev& m, int stage, int nb, int nf, const double* x, const double* s,
                                    const double* w, long long stride, const uint8_t* fixed, const double* alphas, const Reduce& R,
                                    int slot) {
  if (STRICT || (x && (!s || !sgrid_tile_ok(m, x, s, w, stride, fixed)))) return false;  // (x null: prepare only)
  if (nb == 8) return launch_sgrid_tile_multi1<8>(st, m, stage, nf, x, s, w, stride, fixed, alphas, R, slot);
  if (nb == 10) return launch_sgrid_tile_multi1<10>(st, m, stage, nf, x, s, w, stride, fixed, alphas, R, slot);
  if (nb == 12) return launch_sgrid_tile_multi1<12>(st, m, stage, nf, x, s, w, stride, fixed, alphas, R, slot);
  if (nb == 14) return launch_sgrid_tile_multi1<14>(st, m, stage, nf, x, s, w, stride, fixed, alphas, R, slot);
  if (nb == 16) return launch_sgrid_tile_multi1<16>(st, m, stage, nf, x, s, w, stride, fixed, alphas, R, slot);
  return false;
}
