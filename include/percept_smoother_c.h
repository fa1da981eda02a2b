/* percept_smoother_c.h — C ABI of the B200-native mesh-smoothing hot path.
 *
 * This is the drop-in boundary for Percept's reference-mesh conjugate-gradient
 * smoother.  The reference has no FFI for this path: it is a C++ class template
 *   percept::ReferenceMeshSmootherConjugateGradientImpl<MeshType>
 *   (src/percept/mesh/mod/smoother/ReferenceMeshSmootherConjugateGradient.hpp:74-174)
 * whose members launch Kokkos functors.  The C++ host classes in
 * include/percept_b200/ (ReferenceMeshSmootherConjugateGradient.hpp) keep that class API and call the entry points
 * below; each entry point cites the reference member / functor it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types in signatures.
 *  - every function returns 0 (PMS_OK) or a non-zero status; the message is in
 *    pms_last_error(ctx).  The C++ wrapper turns non-zero into
 *    std::runtime_error, the reference's error convention (Util.hpp:288-314).
 *  - one ctx per GPU, one host thread per ctx, one CUDA stream per ctx
 *    (the reference is likewise single-threaded per MPI rank and not re-entrant).
 *  - caller owns host arrays; the ctx owns device memory.
 *  - fields are addressed BY NAME exactly as the reference does
 *    (PerceptMesh::add_coordinate_state_fields, PerceptMesh.cpp:4490-4539):
 *    "coordinates" "coordinates_N" "coordinates_NM1" "coordinates_lagged"
 *    "cg_g" "cg_r" "cg_d" "cg_s" (3 components), "cg_edge_length"
 *    "num_adj_elems" (1 component).
 *  - there is NO CPU fallback: every compute entry point runs CUDA kernels and
 *    fails with PMS_ERR_CUDA when no device is usable.
 */
#ifndef PERCEPT_SMOOTHER_C_H
#define PERCEPT_SMOOTHER_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pms_ctx pms_ctx;

enum pms_status {
  PMS_OK = 0,
  PMS_ERR_ARG = 1,      /* bad argument / unknown field / wrong call order        */
  PMS_ERR_CUDA = 2,     /* CUDA runtime failure or no device                      */
  PMS_ERR_NCCL = 3,     /* NCCL failure / NCCL not loadable                       */
  PMS_ERR_SMOOTHER = 4  /* algorithmic failure the reference throws on:
                           "can't reduce metric" (Def.hpp:553), "bad sDotGrad"
                           (Def.hpp:495), invalid elements at start of stage 2
                           (ReferenceMeshSmootherBase.cpp:288), "bad alpha"       */
};

/* Element topologies on the path (JacobianUtil.cpp:388-396, 448-457). */
enum pms_topology { PMS_TET4 = 4, PMS_HEX8 = 8 };

/* Host array layouts accepted by upload/download.
 *  PMS_AOS: xyz xyz ... per node        (STK field data, LayoutRight views)
 *  PMS_SOA: plane x | plane y | plane z (Kokkos LayoutLeft View<double****>,
 *           MeshType.hpp:48-82: i fastest, component slowest)               */
enum pms_layout { PMS_AOS = 0, PMS_SOA = 1 };

/* ---- life cycle ------------------------------------------------------- */
/* device: CUDA ordinal. Replaces Kokkos::initialize + view allocation.      */
int pms_create(pms_ctx** out, int device);
void pms_destroy(pms_ctx* ctx);
const char* pms_last_error(const pms_ctx* ctx);
/* Options (the reference's public members / properties):
 *  "use_ref_mesh"  (m_use_ref_mesh, ReferenceMeshSmootherBase.hpp:104) default 1
 *  "strict_fp"     1 = kernels compiled without FMA contraction, reference
 *                  expression order (bit-comparable with a -ffp-contract=off
 *                  CPU build); 0 = contracted FMA build (default)
 *  "cache_metric"  1 = driver reuses bitwise-identical metric evaluations
 *                  (SURVEY 3.3), 0 = evaluate exactly as often as the reference
 *  "speculate_line_search"  default 1 (needs cache_metric 1, hex meshes, default build): the driver evaluates the
 *                  backtracking sequence alpha_init*tau^k in multi-alpha passes (pms_total_metric_multi), takes the
 *                  metric at the new coordinates from the fused metric+gradient pass and keeps that gradient for the
 *                  next iteration, tests the quadratic step by moving there, and on an untangled mesh evaluates only the
 *                  invalid-element count of the far-too-long leading trials (a trial with invalid elements is rejected
 *                  whatever its metric).  Same decisions and iteration
 *                  counts; values agree to rounding (<= 1e-12) instead of bitwise.  Side effect: after
 *                  pms_run_one_iteration the field cg_g holds f'(x_new) (the gradient the next iteration starts
 *                  from) instead of f'(x_old).  0 = the reference's literal pass sequence.
 *  "smooth_surfaces"  (PerceptMesh::get_smooth_surfaces) default 0, see the surface smoothing hooks below
 *  "keep_element_metrics"  1 = metric passes also store per-element values for pms_get_element_metrics
 *  "variant_metric_corners" / "variant_grad_corners"  kernel variants (tuning aid)
 *  "sgrid_fused"   default 1: structured blocks use the one-pass tile kernels (TMA-staged node planes, in-CTA ordered
 *                  gather, pms_sgrid_tile.cuh); 0 = element pass + per-cell scratch + node gather.  Same bits either way.
 *  "tet_ref_cache" default 1: tet4 kernels read W^-1 and det W of every element from a per-element cache that is rebuilt
 *                  when coordinates_NM1 changes (10 coalesced values instead of gathering the 4 reference vertices in
 *                  every pass); 0 = gather them.  Agrees to rounding.
 *  "elem_ws"       default 0: 1 = unstructured hex8 gradient in one kernel (node gather riding on the element loop,
 *                  pms_elem_ws.cuh); same bits as the two passes, measured slower (profiles/r02_elem_gather.md)         */
int pms_set_option(pms_ctx* ctx, const char* key, double value);

/* ---- mesh definition ---------------------------------------------------- */
/* StructuredBlock (structured/StructuredBlock.hpp:31-101): node_size = nodes per
 * direction; loop/access ordering is the reference's fixed {0,1,2}.          */
int pms_add_structured_block(pms_ctx* ctx, const unsigned node_size[3], int* block_id);
/* Ioss::BoundaryCondition range, 1-based inclusive (StructuredBlock.cpp:158-163);
 * consumed by SGridBoundarySelector (ReferenceMeshSmootherConjugateGradient.hpp:31-69). */
int pms_block_add_boundary_condition(pms_ctx* ctx, int block, const int range_beg[3],
                                     const int range_end[3]);
/* Ioss::ZoneConnectivity (1-based ranges, transform as in MeshType.hpp:96-131). */
int pms_block_add_zone_connectivity(pms_ctx* ctx, int block, int donor_block,
                                    const int transform[3], const int owner_beg[3],
                                    const int owner_end[3], const int donor_beg[3],
                                    const int donor_end[3]);
/* Unstructured STK-like mesh: conn is nelems x npe, 0-based node indices in
 * Exodus/Shards order.  elem_owned / node_owned (may be NULL = all owned) mirror
 * locally_owned_part (total_element_metric.cpp:34, PerceptMesh.cpp:4665);
 * nodes with node_owned==0 are shared-not-owned or aura ghosts.              */
int pms_set_unstructured(pms_ctx* ctx, int topology, size_t nnodes, size_t nelems,
                         const int32_t* conn, const uint8_t* elem_owned,
                         const uint8_t* node_owned);
/* Fixed (boundary) node mask = MeshSmootherImpl::get_fixed_flag evaluated once on
 * the host (MeshSmoother.cpp:271-417).  Length = total node count (blocks
 * concatenated in block order).  NULL = nothing fixed.                        */
int pms_set_fixed_mask(pms_ctx* ctx, const uint8_t* fixed);
/* fixed = SGridBoundarySelector over the registered boundary conditions.      */
int pms_select_boundary_sgrid(pms_ctx* ctx);
/* Allocates the coordinate-state fields, builds node->element CSR / interface
 * maps / ownership masks.  Must precede any field or compute call.            */
int pms_commit(pms_ctx* ctx);
int pms_num_nodes_local(const pms_ctx* ctx, size_t* n);   /* all blocks, duplicates included */
int pms_num_elements_local(const pms_ctx* ctx, size_t* n);
/* Bit-exact connectivity artefacts for parity checks: node->element CSR
 * (row_ptr[nnodes+1], entries = elem*8+local_node), sorted by (elem, local).   */
int pms_get_node_elem_csr(const pms_ctx* ctx, int64_t* row_ptr, int64_t* entries);

/* ---- fields ------------------------------------------------------------- */
/* block = -1 addresses the whole (unstructured or single-block) mesh.          */
int pms_field_upload(pms_ctx* ctx, const char* name, int block, const double* host, int layout);
int pms_field_download(pms_ctx* ctx, const char* name, int block, double* host, int layout);
/* Device view of a field: component c of local node n is dptr[c*comp_stride+n].  The caller may read and write
 * through the pointer between calls (on the ctx stream, or after pms_synchronize).  From this call on the library
 * assumes the field may have changed before EVERY later call: cached metrics, the precomputed gradient and the
 * search-direction scalars that depend on it are not re-used across calls (results stay correct, per-call caching of
 * that field is lost) until pms_field_release_ptr(name) ends the aliasing.                                        */
int pms_field_device_ptr(pms_ctx* ctx, const char* name, double** dptr, size_t* comp_stride);
int pms_field_release_ptr(pms_ctx* ctx, const char* name);
/* Asynchronous SoA transfers (pinned host memory) for callers that evaluate many coordinate sets: the read-back of one
 * evaluation and the upload of a later one run beside the kernels (PCIe is full duplex; three streams).  Both go through
 * device staging buffers, so kernels queued afterwards never race with a copy in flight.
 *  pms_field_download_async: snapshots the field on the ctx stream and copies the snapshot to `host` on a second stream;
 *                            pms_transfer_wait blocks until the host buffer(s) are complete.
 *  pms_field_upload_async:   starts host -> staging at once (third stream); the field itself changes only when
 *                            pms_field_upload_finish is called (stream-ordered, does not block the host).            */
int pms_field_download_async(pms_ctx* ctx, const char* name, int block, double* host);
int pms_field_upload_async(pms_ctx* ctx, const char* name, int block, const double* host);
int pms_field_upload_finish(pms_ctx* ctx);
int pms_transfer_wait(pms_ctx* ctx);
/* BLAS-1: PerceptMesh::copy_field / nodal_field_* (PerceptMesh.cpp:4571-4820),
 * SB_* functors (structured/BlockStructuredGrid.cpp:301-578).                  */
int pms_copy_field(pms_ctx* ctx, const char* dst, const char* src);
int pms_nodal_field_axpby(pms_ctx* ctx, double alpha, const char* x, double beta, const char* y);
int pms_nodal_field_axpbypgz(pms_ctx* ctx, double alpha, const char* x, double beta,
                             const char* y, double gamma, const char* z);
int pms_nodal_field_dot(pms_ctx* ctx, const char* x, const char* y, double* result);
int pms_nodal_field_set_value(pms_ctx* ctx, const char* name, double value);
/* PerceptMesh::get_number_nodes (PerceptMesh.cpp:730-756): uniquely owned nodes,
 * summed over ranks.                                                           */
int pms_get_number_nodes(pms_ctx* ctx, size_t* n);

/* ---- smoother kernels ----------------------------------------------------- */
/* stage 0 = SmootherMetricUntangle, stage 1 = SmootherMetricShapeB1.           */
/* get_edge_lengths (Def.hpp:759-771; get_edge_len_avg.hpp:251-383).             */
int pms_get_edge_lengths(pms_ctx* ctx);
/* total_metric(alpha) (Def.hpp:330-388): metric of coordinates + alpha*cg_s on
 * unfixed nodes; "coordinates" is left unchanged.  n_invalid may be NULL.       */
int pms_total_metric(pms_ctx* ctx, int stage, double alpha, double* mtot, size_t* n_invalid);
/* total_metric at several step lengths of the same direction in one pass (hex meshes: the per-corner Jacobian
 * determinant and |A W^-1|^2 are polynomials in alpha, so extra alphas cost ~1/5 of a pass each).  Values agree with
 * pms_total_metric to rounding (<= 1e-12 relative); the line-search driver uses it to evaluate the known backtracking
 * sequence alpha, alpha/2, alpha/4, ... speculatively (option "speculate_line_search", default 1).                  */
int pms_total_metric_multi(pms_ctx* ctx, int stage, int nalpha, const double* alphas, double* mtot, size_t* n_invalid);

/* get_gradient (Def.hpp:1481-1520; gradient_functors.hpp:245-487): fills cg_g
 * (and cg_r = cg_g for unstructured meshes), zero at fixed nodes.               */
int pms_get_gradient(pms_ctx* ctx, int stage);
/* fused get_gradient + total_metric(0) in one pass over the elements.           */
int pms_metric_and_gradient(pms_ctx* ctx, int stage, double* mtot, size_t* n_invalid);
/* parallel_count_invalid_elements (MeshSmoother.cpp:192-234).                   */
int pms_count_invalid_elements(pms_ctx* ctx, size_t* n_invalid);
/* get_alpha_0 (Def.hpp:906-926; get_alpha_0_refmesh.hpp:55-194).                */
int pms_get_alpha_0(pms_ctx* ctx, double* alpha_0);
/* update_node_positions (Def.hpp:395-410; update_coordinates.cpp:241-323):
 * lagged <- coordinates; coordinates += alpha*cg_s on unfixed nodes; dmax,
 * dmax_relative.                                                                */
int pms_update_node_positions(pms_ctx* ctx, double alpha, double* dmax, double* dmax_relative);
/* per-element metric values and invalid flags of the last total_metric /
 * metric_and_gradient call (the reference's element_invalid_flags view,
 * total_element_metric.cpp:261,305); debug/parity aid, enabled with option
 * "keep_element_metrics".                                                       */
int pms_get_element_metrics(pms_ctx* ctx, double* metric, int32_t* invalid_flag);

/* ---- CG driver ------------------------------------------------------------ */
/* Mirror of the public state scalars (ReferenceMeshSmootherBase.hpp:71-101). */
typedef struct pms_state {
  double dmax, dmax_relative;
  double dnew, dold, d0, dmid;
  double alpha, alpha_0;
  double grad_norm_scaled;
  double total_metric, global_metric;
  int stage, iter;
  int untangled, converged;
  uint64_t num_invalid, num_nodes;
  /* bookkeeping (not in the reference): evaluations done by the last call       */
  uint64_t n_metric_evals, n_gradient_evals, n_line_search_halvings, restarted;
} pms_state;

int pms_state_init(pms_ctx* ctx, pms_state* st);
/* run_one_iteration (Def.hpp:1010-1093) incl. line_search (Def.hpp:466-591).    */
int pms_run_one_iteration(pms_ctx* ctx, pms_state* st, double grad_norm, double* metric_out);
/* line_search(bool& restarted, double mfac_mult) (Def.hpp:466-591;
 * ReferenceMeshSmootherConjugateGradient.hpp:155) on its own: Armijo back-tracking from 10*alpha_0 along cg_s with the
 * restart rules and the final quadratic fit; reads cg_s, cg_g, cg_r, cg_edge_length and st->stage/untangled/dmax_relative,
 * returns the accepted step length, does not move the nodes (the caller follows with pms_update_node_positions).     */
int pms_line_search(pms_ctx* ctx, pms_state* st, double mfac_mult, int* restarted, double* alpha);
/* check_convergence (Def.hpp:989-1008).                                         */
int pms_check_convergence(const pms_state* st, double grad_norm);
/* run_algorithm (ReferenceMeshSmootherBase.cpp:240-334): both stages, writes the
 * smootherlog rows (ReferenceMeshSmootherBase.cpp:57-81,216-238) to log_path
 * when non-NULL.  innerIter / gradNorm are the constructor arguments.           */
int pms_run_algorithm(pms_ctx* ctx, int inner_iterations, double grad_norm,
                      const char* log_path, pms_state* st);

/* ---- multi-GPU (one process per GPU, NCCL over NVLink) --------------------- */
/* 128-byte ncclUniqueId created by rank 0 and broadcast by the host program
 * (MPI_Bcast in Percept, torch.distributed in bench.py).                        */
int pms_comm_unique_id(void* id128);
int pms_comm_init(pms_ctx* ctx, const void* id128, int rank, int nranks);
/* Ghost/shared-node exchange plan replacing copy_owned_to_shared +
 * communicate_field_data(aura) (MeshType.cpp:84-88).  For neighbour q:
 * send_nodes[send_off[q] .. send_off[q+1]) are local indices of nodes I own that
 * q holds as non-owned; recv_nodes likewise the non-owned nodes q owns, in the
 * matching order.                                                               */
int pms_set_halo(pms_ctx* ctx, int n_neighbors, const int* neighbor_ranks,
                 const int64_t* send_off, const int32_t* send_nodes,
                 const int64_t* recv_off, const int32_t* recv_nodes);
/* Structured block per GPU: interface nodes are duplicated on every rank whose block contains them (as the
 * reference duplicates them per block, gradient_functors.hpp:141-220).  For neighbour q, nodes[off[q]..off[q+1]) are
 * my local copies of the nodes shared with q, in the same (global-id) order on both sides; neighbours in ascending
 * rank order.  get_gradient / get_edge_lengths then sum the copies in ascending rank order (the multi-block ==
 * single-block result).  owned (may be NULL): 1 where this rank is the lowest rank holding the node (dots, counts). */
int pms_set_interface(pms_ctx* ctx, int n_neighbors, const int* neighbor_ranks, const int64_t* off,
                      const int32_t* nodes, const uint8_t* owned);
/* owner -> ghost copy of a nodal field (used for cg_s once per iteration).      */
int pms_halo_exchange(pms_ctx* ctx, const char* name);

/* ---- instrumentation --------------------------------------------------------- */
/* ---- surface smoothing hooks ------------------------------------------------------------------------
 * The reference's `smooth_surfaces` branch (mesh_adapt --smooth_surfaces=1) lets nodes classified on a CAD surface move
 * along it: get_fixed_flag (mesh/mod/smoother/MeshSmoother.cpp:271-344) un-fixes MS_SURFACE nodes, get_gradient_2
 * projects cg_g onto the tangent plane at those nodes (ReferenceMeshSmootherConjugateGradientDef.hpp:1401-1447 +
 * MeshSmoother.cpp:437-460), and run_one_iteration re-snaps them after the update and rejects an invalid result
 * ("bad mesh after snap_node_positions...", Def.hpp:1078-1089, snap_nodes :236-316).  The CAD evaluators
 * (MeshGeometry::{classify_node, normal_at, snap_points_to_geometry}, OpenNURBS/PGEOM) are third party; here the
 * caller hands over the classification and analytic evaluators instead:
 *   pms_add_analytic_surface(kind, params): PMS_SURF_PLANE  params = point[3], normal[3]
 *                                           PMS_SURF_SPHERE params = centre[3], radius
 *                                           PMS_SURF_CYLINDER params = axis point[3], axis direction[3], radius
 *   pms_set_node_classification(dof, surface_id): per local node, dof as classify_node returns it (0 vertex, 1 curve,
 *     2 surface, 3 volume) and the surface evaluator of dof-2 nodes.  Replaces the boundary-selector fixed mask:
 *     dof 0/1 fixed, dof 2 fixed unless option "smooth_surfaces" is 1, dof 3 free.  Unstructured meshes only
 *     (the reference's StructuredGrid specialisation is "not impl").  Call after pms_commit.
 *   pms_snap_nodes(&dmax): snap_nodes() on its own; pms_get_surface_normals(out SoA (3,N)): get_surface_normals().   */
enum { PMS_SURF_PLANE = 0, PMS_SURF_SPHERE = 1, PMS_SURF_CYLINDER = 2 };
int pms_add_analytic_surface(pms_ctx* ctx, int kind, const double* params, int* surface_id);
int pms_set_node_classification(pms_ctx* ctx, const int8_t* dof, const int32_t* surface_id);
int pms_snap_nodes(pms_ctx* ctx, double* dmax);
/* snap_to(node, coordinate, reset) (MeshSmoother.hpp:79, MeshSmoother.cpp:517-563): the node is put at xyz and snapped to
 * its surface, xyz returns the snapped position; reset != 0 restores the node's coordinates afterwards.            */
int pms_snap_point(pms_ctx* ctx, size_t node, double* xyz, int reset);
int pms_get_surface_normals(pms_ctx* ctx, double* host_out);

/* ---- structured 2x refinement feeding the smoother -------------------------------------------------
 * Replaces StructuredGridRefiner::do_refine (structured/StructuredGridRefiner.cpp:40-170, .hpp:19-52) and
 * StructuredGridRefinerImpl::refine (structured/StructuredGridRefinerDef.hpp:43-226).
 * 1. pms_refine_structured_topology(coarse, fine, &ncells): `fine` is a fresh ctx (no mesh); it receives one block
 *    of 2n-1 nodes per direction for every block of `coarse`, with the 1-based boundary-condition and
 *    zone-connectivity ranges mapped r -> 2r-1.  Returns the refined cell count like do_refine().  Host only.
 * 2. caller commits `fine` (pms_commit) and chooses fixed nodes (e.g. pms_select_boundary_sgrid).
 * 3. pms_refine_structured_field(coarse, "coordinates", fine, "coordinates"): device prolongation of a 3-component
 *    nodal field: vertices copied, edge midpoints 0.5*(a+b), face centres and cell centroids summed in the
 *    reference's order (bit-identical to it).  Both ctx on the same GPU; stream-ordered, no host sync.          */
int pms_refine_structured_topology(const pms_ctx* coarse, pms_ctx* fine, unsigned long long* num_refined_cells);
int pms_refine_structured_field(pms_ctx* coarse, const char* src_field, pms_ctx* fine, const char* dst_field);

/* kernels launched by this ctx since creation (bench.py's gpu_launches).        */
int pms_launch_count(const pms_ctx* ctx, uint64_t* n);
/* CUDA stream of the ctx as an opaque handle, for event timing on that stream. */
int pms_stream(const pms_ctx* ctx, void** stream);
int pms_synchronize(pms_ctx* ctx);
/* accumulated device time (ms, CUDA events on the ctx stream) per kernel family:
 * names = "metric","gradient","blas1","update","alpha0","edge","halo","refine".          */
int pms_kernel_time_ms(pms_ctx* ctx, const char* family, double* ms, uint64_t* launches);
int pms_kernel_time_reset(pms_ctx* ctx);
int pms_set_timing(pms_ctx* ctx, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* PERCEPT_SMOOTHER_C_H */
