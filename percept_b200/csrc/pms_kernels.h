// pms_kernels.h — host-callable launchers of the sm_100a kernels (internal; not part of the C ABI).
//
// The kernel translation unit pms_kernels.cu is compiled twice:
//   namespace pms_fast   : default nvcc FMA contraction
//   namespace pms_strict : -fmad=false, bit-comparable with the -ffp-contract=off CPU oracle
// Both export the same launcher table.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace pmsk {

enum : int { KIND_SGRID = 0, KIND_HEX8 = 1, KIND_TET4 = 2 };

// One structured block, or the whole unstructured mesh.
struct MeshDev {
  int kind;
  int ni, nj, nk;            // nodes per direction (structured)
  long long node_off;        // first node of this block in the flat node arrays
  long long elem_off;        // first element of this block in the flat element arrays
  long long nnodes, nelems;  // of this block / mesh
  const int32_t* conn;       // unstructured: SoA connectivity conn[a*conn_stride + e]
  long long conn_stride;
  const uint8_t* elem_fixed_bits;  // bit a set <=> local vertex a is a fixed node (indexed by flat element id)
  const uint8_t* elem_owned;       // unstructured: 1 = locally owned element (metric sums these); may be null
  const int64_t* csr_ptr;          // unstructured node->element CSR
  const int64_t* csr_ent;          // entries elem*8+local, ascending
  const int32_t* csr_ent32;        // same, 32-bit (when nelems*8 < 2^31): halves the gather's index traffic
  const double* tetref;            // tet4: cached reference part per element, tetref[v*tetref_stride + e], v = 0..8 W^-1, 9 det W (null: gather the reference vertices)
  long long tetref_stride;
};

constexpr int REDUCE_MAX_NV = 32;   // doubles one grid reduction can carry
constexpr int MULTI_ALPHA_NB = 16;  // most trial step lengths per multi-alpha metric pass (kernels for 8, 12 and 16)

struct Reduce {  // scratch for deterministic two-level reductions (partials + last-block ticket)
  double* partial_d;            // [REDUCE_MAX_NV * max_blocks]
  unsigned long long* partial_u;  // [max_blocks]
  unsigned int* ticket;         // [2] zero-initialised, self-resetting (the second one for a kernel's second reduction)
  double* out_d;                // result slots (device)
  unsigned long long* out_u;
  int max_blocks;
};

// Brick tables for the scratch-free unstructured hex8 gradient (built on the host from the reference-mesh centroids):
// elements are Morton-sorted and cut into bricks of 128; slot = brick*128 + t.
struct BrickDev {
  long long nbricks;
  // compute table, 2560 B per brick: uint16 lconn[128][8] (brick-local node index of structured position p of slot t)
  // then int32 elem[128] (original element id, -1 = padding)
  const uint4* ctab;
  const int32_t* bnn;        // [nbricks] distinct nodes of the brick
  // gather table, one 16-byte-aligned blob per brick:
  //   int32 nn, novf, 0, 0 | int32 node[nn4] | int32 dst[nn4] | uint16 ell[nn][8] | uint16 ovf[pad8(novf)]
  //   node: global ids ascending.  dst: -1 all adjacent elements inside the brick, free node -> cg_g; -2 same but
  //   fixed / not owned -> 0; >= 0 index of this brick's partial sum.  ell: the node's first 8 contributions (exchange
  //   column offsets, see BRICK_EX_STRIDE) in ascending (element id, local node) order; ovf: [j, count, entries...]
  //   for the 9th onward.
  const int32_t* goff;       // [nbricks+1] blob offsets in 16-byte units
  const uint4* gtab;
  double* partial;           // [3][npartial] brick-surface partial sums
  long long npartial;
  long long ncomb;           // nodes that have partials
  const int32_t* comb_node;  // [ncomb]
  const int32_t* comb_ptr;   // [ncomb+1]
  const int32_t* comb_ent;   // partial indices, ascending brick id
};
constexpr int BRICK_ELEMS = 128;      // = NTP of the pipelined kernels
constexpr int BRICK_MAX_NODES = 320;  // a brick is cut short when its distinct nodes would exceed this
constexpr int BRICK_CTAB_BYTES = 2560;
constexpr int BRICK_GTAB_BYTES = 16 + 8 * BRICK_MAX_NODES + 16 * BRICK_MAX_NODES;  // overflow shares the slack
// exchange column: contribution (p, r) of slot t at t*BRICK_EX_STRIDE + p*3 + r (odd stride: conflict-free writes);
// an ELL entry is t*BRICK_EX_STRIDE + p*3; unused ELL entries point at three zeros behind the last column.
constexpr int BRICK_EX_STRIDE = 25;
constexpr int BRICK_EX_ZERO = BRICK_ELEMS * BRICK_EX_STRIDE;

// plan of the one-kernel unstructured hex8 gradient (pms_elem_ws.cuh): nodes listed by the iteration in which their last
// adjacent element is evaluated, per-node tables of scratch offsets, cross-CTA progress words
struct WsPlanDev {
  int grid;                // CTAs the plan was built for (all co-resident)
  int D;                   // batch j (list positions [j*grid*128, (j+1)*grid*128)) is gathered during iteration j + D
  int nbatches;
  int nslabs;              // slabs of 4 element iterations
  unsigned sleep_ns;       // back-off of a CTA that has to wait for the slowest one
  long long niter;         // iterations of every CTA = max(element iterations, nbatches + D)
  const int32_t* lnode;    // [nbatches*grid*128] node id at list position q (-1: padding)
  const int4* tabA;        // scratch offsets (local node * 3 * egs + element) of the node's CSR entries 0..3, -1 = none
  const int4* tabB;        // entries 4..7
  const int32_t* need;     // [nbatches] slabs (of 4 iterations) every CTA must have published before batch j is read
  int* prog;               // [nslabs] CTAs through with slab s, zeroed before the launch
  const int32_t* ovf_nodes;  // nodes with more than 8 adjacent elements: gathered by a list kernel afterwards
  long long novf;
};

// analytic stand-in for a MeshGeometry evaluator: 0 plane (p, unit normal u), 1 sphere (centre p, R),
// 2 cylinder (axis point p, unit axis u, R)
struct SurfaceDev {
  int kind;
  double p[3], u[3], R;
};

struct Launchers {
  // metric of (x + alpha*s on unfixed nodes) vs reference w; result: out_d[slot] = sum, out_u[slot] = #invalid.
  // s may be null (alpha = 0).  elem_metric/elem_flag optional.
  void (*metric)(cudaStream_t, const MeshDev&, int stage, int use_ref, const double* x, const double* s, const double* w,
                 long long stride, double alpha, const Reduce&, int slot, double* elem_metric, int32_t* elem_flag,
                 const uint8_t* fixed);
  // element pass of the gradient: eg[(a*3+r)*eg_stride + e] (zero where the element is dropped); fused metric like above.
  void (*grad_elements)(cudaStream_t, const MeshDev&, int stage, int use_ref, const double* x, const double* w,
                        long long stride, double* eg, long long eg_stride, const Reduce&, int slot, double* elem_metric,
                        int32_t* elem_flag);
  // node pass: g[r*stride+n] = sum of adjacent elements' contributions in ascending element order; 0 at fixed /
  // non-owned nodes.  r_out (may be null) receives a copy (cg_r = cg_g on STK meshes).
  void (*grad_gather)(cudaStream_t, const MeshDev&, const double* eg, long long eg_stride, const uint8_t* fixed,
                      const uint8_t* node_owned, double* g, double* r_out, long long stride);
  // fused structured metric + gradient (pms_sgrid_tile.cuh: cell tiles staged in shared memory, in-CTA node gather in the
  // order of grad_gather, tile-seam nodes through `scratch` of sgrid_fused_scratch_bytes(block) bytes); returns false
  // if unsupported for this block.  elem_metric / elem_flag optional.
  bool (*grad_fused_sgrid)(cudaStream_t, const MeshDev&, int stage, const double* x, const double* w, long long stride,
                           const uint8_t* fixed, double* g, const Reduce&, int slot, double* scratch, double* elem_metric,
                           int32_t* elem_flag);
  size_t (*sgrid_fused_scratch_bytes)(const MeshDev&);

  // interface groups (multi-block structured): for group q, nodes grp_nodes[grp_ptr[q]..grp_ptr[q+1]) are duplicates of
  // one physical node, ascending block id.  val = sum over duplicates (in that order), written back to all.
  void (*group_sum)(cudaStream_t, long long ngroups, const int64_t* grp_ptr, const int64_t* grp_nodes, double* f, int ncomp,
                    long long stride);

  // BLAS-1 over [ncomp][stride] SoA fields, first n entries of every plane
  void (*axpby)(cudaStream_t, double a, const double* x, double b, double* y, long long n, long long stride, int ncomp);
  void (*axpbypgz)(cudaStream_t, double a, const double* x, double b, const double* y, double c, double* z, long long n,
                   long long stride, int ncomp);
  void (*set_value)(cudaStream_t, double* x, double v, long long n, long long stride, int ncomp);
  // out_d[slot] = sum_{owned n} sum_c x*y
  void (*dot)(cudaStream_t, const double* x, const double* y, const uint8_t* mask, long long n, long long stride, int ncomp,
              const Reduce&, int slot);
  // r = -g ; out_d[slot..slot+2] = d.d, r.d, r.r   (Def.hpp:1034-1037 fused)
  void (*cg_r_dots)(cudaStream_t, const double* g, const double* d, double* r, const uint8_t* mask, long long n,
                    long long stride, const Reduce&, int slot);
  // restart ? s = r : s = r + beta*s ; d = r   (Def.hpp:1057-1069 fused)
  void (*cg_s_update)(cudaStream_t, const double* r, double* s, double* d, double beta, int restart, long long n,
                      long long stride);
  // the same update fused with what the line search asks next about the new s: out_d[slot] = s.g, out_d[slot+1] = s.s
  // (both over `dotmask` nodes), out_d[slot+2] = min over non-`skip` nodes of edge/|s| (DBL_MAX if none).  Same
  // per-thread accumulation order as dot / alpha0, hence bit-identical values.
  void (*cg_s_update_dots)(cudaStream_t, const double* r, double* s, double* d, const double* g, const double* edge,
                           const uint8_t* dotmask, const uint8_t* skip, double beta, int restart, long long n, long long stride,
                           const Reduce&, int slot);
  // out_d[slot] = min over candidate nodes of edge/|s|, DBL_MAX if none (get_alpha_0_refmesh.hpp:168-193)
  void (*alpha0)(cudaStream_t, const double* s, const double* edge, const uint8_t* skip_mask, long long n, long long stride,
                 const Reduce&, int slot);
  // lagged = x ; x += alpha*s (unfixed) ; out_d[slot] = dmax, out_d[slot+1] = dmax_relative
  void (*update)(cudaStream_t, double* x, double* lagged, const double* s, const double* edge, const uint8_t* fixed,
                 const uint8_t* owned, double alpha, long long n, long long stride, const Reduce&, int slot);
  // edge lengths: el = sum over adjacent elements of mean edge length, na = adjacency count (division done by divide())
  // unstructured: `scratch` (>= nelems doubles, may be null) holds one mean edge length per element between the two passes
  void (*edge_lengths)(cudaStream_t, const MeshDev&, const double* w, long long stride, double* el, double* na, double* scratch);
  void (*divide)(cudaStream_t, double* el, const double* na, long long n);
  // halo pack / unpack: buf[c*count + i] <-> f[c*stride + nodes[i]]
  void (*pack)(cudaStream_t, const double* f, long long stride, int ncomp, const int32_t* nodes, long long count, double* buf);
  void (*unpack)(cudaStream_t, double* f, long long stride, int ncomp, const int32_t* nodes, long long count, const double* buf);
  // per-element bit mask of fixed vertices
  void (*elem_fixed_bits)(cudaStream_t, const MeshDev&, const uint8_t* fixed, uint8_t* bits);
  // tet4: reference-mesh part of every element (W^-1, det W) from the reference coordinates w -> out[v*out_stride + e]
  void (*tet_ref)(cudaStream_t, const MeshDev&, const double* w, long long stride, double* out, long long out_stride);
  // loads (CUDA lazy module loading) every multi-alpha kernel a mesh of this kind can ask for; no launch
  void (*prepare_line_search_kernels)(const MeshDev&);
  // scratch-free hex8 gradient over bricks (default build only); returns false when unsupported
  bool (*grad_bricks)(cudaStream_t, const MeshDev&, const BrickDev&, int stage, const double* x, const double* w,
                      long long stride, const uint8_t* fixed, const uint8_t* node_owned, double* g, double* r_out,
                      const Reduce&, int slot, double* elem_metric, int32_t* elem_flag);
  // one-kernel unstructured hex8 metric + gradient (element loop with the node gather riding on it through cp.async);
  // returns false when unsupported (strict build, tet4, no reference mesh) or the launch is refused
  bool (*grad_ws)(cudaStream_t, const MeshDev&, const WsPlanDev&, int stage, int use_ref, const double* x, const double* w,
                  long long stride, const uint8_t* fixed, const uint8_t* node_owned, double* g, double* r_out, double* eg,
                  long long egs, const Reduce&, int slot);
  // CTAs that kernel runs with (the plan is built for this grid); 0 = not available
  int (*grad_ws_grid)();
  // cross-rank interface sum (structured block per GPU), ascending rank order
  void (*shared_sum)(cudaStream_t, long long nshared, const int32_t* nodes, const int64_t* ptr, const int64_t* ent_pos,
                     const int32_t* ent_cnt, const double* recvbuf, double* f, int ncomp, long long stride);
  // hex metric at nb (8, 10, 12, 14 or 16) trial step lengths in one pass (line search speculation): out_d[slot + k] = metric
  // at alphas[k], out_d[slot + nb + k] = number of invalid elements at alphas[k].  nf < nb (ShapeB1 only; 4 or 6): only the
  // LAST nf step lengths get their metric, the others only their invalid count (their out_d[slot + k] is meaningless).
  // Returns false when not available (tet4, strict build, ShapeB1 without reference mesh).
  bool (*metric_multi)(cudaStream_t, const MeshDev&, int stage, int use_ref, int nb, int nf, const double* x, const double* s,
                       const double* w, long long stride, const double* alphas, const Reduce&, int slot, const uint8_t* fixed);
  // StructuredGridRefiner: 2x prolongation of a 3-component nodal field of one block (coarse ni x nj x nk nodes at
  // src + src_off, component stride src_stride) onto the (2ni-1) x (2nj-1) x (2nk-1) block at dst + dst_off
  void (*refine_sgrid)(cudaStream_t, int ni, int nj, int nk, const double* src, long long src_off, long long src_stride,
                       double* dst, long long dst_off, long long dst_stride);
  // surface smoothing hooks over the list of non-fixed surface nodes (node id, surface id):
  //   project: g -= (g.n) n with n = surface normal at x (get_gradient_2 / project_delta_to_tangent_plane)
  //   snap:    x <- closest point of the surface; out_d[slot] = max displacement (snap_nodes)
  //   normals: out <- n, component stride `stride` (get_surface_normals)
  void (*surf_project)(cudaStream_t, long long n, const int32_t* nodes, const int32_t* sid, const SurfaceDev* surfs,
                       const double* x, double* g, long long stride);
  void (*surf_snap)(cudaStream_t, long long n, const int32_t* nodes, const int32_t* sid, const SurfaceDev* surfs, double* x,
                    long long stride, const Reduce&, int slot);
  void (*surf_normals)(cudaStream_t, long long n, const int32_t* nodes, const int32_t* sid, const SurfaceDev* surfs,
                       const double* x, double* out, long long stride);
};

}  // namespace pmsk

namespace pms_fast { const pmsk::Launchers* launchers(); void set_variant(int metric_corners, int grad_corners, int sgrid_tile = 1); }
namespace pms_strict { const pmsk::Launchers* launchers(); void set_variant(int metric_corners, int grad_corners, int sgrid_tile = 1); }
