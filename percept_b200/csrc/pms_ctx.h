// pms_ctx.h — internal context of the C ABI (host side). Not installed.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/percept_smoother_c.h"
#include "pms_kernels.h"

struct PmsError : std::runtime_error {
  int code;
  PmsError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PMS_CUDA(call)                                                                                     \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      throw PmsError(PMS_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(e_));             \
  } while (0)

struct HBC {
  int beg[3], end[3];
};
struct HZone {
  int donor;
  int transform[3], owner_beg[3], owner_end[3], donor_beg[3], donor_end[3];
};
struct HBlock {
  unsigned n[3];
  long long off = 0, eoff = 0;
  std::vector<HBC> bcs;
  std::vector<HZone> zones;
  long long nnodes() const { return (long long)n[0] * n[1] * n[2]; }
  long long ncells() const { return (long long)(n[0] - 1) * (n[1] - 1) * (n[2] - 1); }
  long long node(long long i, long long j, long long k) const { return off + i + (long long)n[0] * (j + (long long)n[1] * k); }
};

struct DField {
  int ncomp = 0;
  double* d = nullptr;
  uint64_t version = 0;  // bumped on every write; keys the metric cache
  bool aliased = false;  // raw device pointer handed out: bumped on every API entry as well
};

enum Family { FAM_METRIC = 0, FAM_GRADIENT, FAM_BLAS1, FAM_UPDATE, FAM_ALPHA0, FAM_EDGE, FAM_HALO, FAM_REFINE, FAM_COUNT };

struct pms_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  const pmsk::Launchers* L = nullptr;
  bool strict_fp = false, use_ref_mesh = true, cache_metric = true, keep_elem = false, timing = false;
  bool structured = false, committed = false;
  int var_metric = 2, var_grad = 2;  // kernel variant switches (tuning aid)

  // mesh (host description)
  std::vector<HBlock> blocks;
  int topo = 0;
  long long N = 0, Npad = 0, E = 0, Epad = 0;
  std::vector<int32_t> h_conn;  // AoS as given
  std::vector<uint8_t> h_fixed, h_node_owned, h_elem_owned, h_dotmask;
  std::vector<int64_t> h_csr_ptr, h_csr_ent;
  bool fixed_set = false, dotmask_trivial = true;
  uint64_t num_nodes_owned = 0;

  // device mesh
  std::vector<pmsk::MeshDev> dev_blocks;  // one per structured block, or a single entry
  int32_t* d_conn = nullptr;
  uint8_t *d_fixed = nullptr, *d_node_owned = nullptr, *d_elem_owned = nullptr, *d_dotmask = nullptr, *d_elem_bits = nullptr,
          *d_skip = nullptr;
  int64_t *d_csr_ptr = nullptr, *d_csr_ent = nullptr;
  int32_t* d_csr_ent32 = nullptr;
  int64_t *d_grp_ptr = nullptr, *d_grp_nodes = nullptr;
  long long n_groups = 0;
  double* d_eg = nullptr;  // element-gradient scratch [3*npe][Epad]
  // tet4: reference part per element (W^-1, det W), rebuilt when coordinates_NM1 changes (pms_device.cuh tet_metric_cached)
  double* d_tetref = nullptr;
  uint64_t tetref_wv = ~0ull;
  bool tet_ref_cache = true;
  // one-kernel unstructured gradient (pms_elem_ws.cuh)
  pmsk::WsPlanDev ws{};
  bool ws_built = false, ws_failed = false, elem_ws = false;  // (off by default: measured slower than the two passes)
  int ws_grid = 0;
  double* d_seam = nullptr;  // tile-seam scratch of the fused structured gradient (pms_sgrid_tile.cuh)
  bool sgrid_fused = true;
  // brick tables of the scratch-free hex8 gradient (built lazily from the reference-mesh centroids)
  pmsk::BrickDev bricks{};
  bool bricks_built = false, bricks_failed = false;
  uint64_t bricks_wv = 0;
  std::vector<void*> brick_allocs;
  double* d_elem_metric = nullptr;
  int32_t* d_elem_flag = nullptr;

  // surface smoothing hooks: classification as MeshGeometry::classify_node gives it + analytic evaluators
  bool smooth_surfaces = false;
  std::vector<pmsk::SurfaceDev> h_surfaces;
  std::vector<int8_t> h_node_dof;
  std::vector<int32_t> h_node_surf;
  pmsk::SurfaceDev* d_surfaces = nullptr;
  int32_t *d_surf_nodes = nullptr, *d_surf_sid = nullptr;  // non-fixed MS_SURFACE nodes and their evaluator
  long long n_surf_nodes = 0;

  std::map<std::string, DField> fields;
  int n_aliased = 0;

  // reductions
  pmsk::Reduce R{};
  double* h_out_d = nullptr;              // pinned mirrors of R.out_d / R.out_u
  unsigned long long* h_out_u = nullptr;
  static constexpr int NSLOT = 64;

  // metric cache (identical evaluations reuse the previous result)
  struct CacheEnt {
    uint64_t xv, sv, wv;
    int stage;
    double alpha, mtot;
    uint64_t ninv;
    bool use_ref;
    bool value_known = true;  // false: only ninv is known (validity-only step length of a multi-alpha pass)
  };
  std::vector<CacheEnt> cache;
  static constexpr size_t CACHE_CAP = 32;
  // gradient evaluated ahead of time at the end of an iteration (together with the metric the iteration returns)
  bool pre_valid = false;
  uint64_t pre_xv = 0, pre_wv = 0, pre_gv = 0;
  int pre_stage = -1;
  // what the line search asks about a freshly updated search direction, computed by the update kernel itself
  struct SInfo {
    bool valid = false;
    uint64_t sv = 0, gv = 0;
    double alpha0 = 0, sg = 0, ss = 0;
  } sinfo;
  int spec_prev = 0;  // trial step lengths the previous line search visited (sizes the next speculative batch)
  int spec_kv_prev = -1;  // index of the first trial of the previous batch at which the whole mesh was valid (-1: unknown)
  bool speculate_line_search = true;  // evaluate the line search's alpha_init * tau^k trials in multi-alpha passes
  bool adj_fixed = false;

  // multi-GPU
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  std::vector<int> nb_ranks;
  std::vector<int64_t> send_off, recv_off;
  int32_t *d_send_nodes = nullptr, *d_recv_nodes = nullptr;
  double *d_send_buf = nullptr, *d_recv_buf = nullptr;
  double* d_coll = nullptr;  // small device buffer for scalar allreduces
  // cross-rank interface sums (structured block per GPU)
  std::vector<int> if_ranks;
  std::vector<int64_t> if_off;
  int32_t *d_if_nodes = nullptr, *d_sh_nodes = nullptr, *d_sh_cnt = nullptr;
  int64_t *d_sh_ptr = nullptr, *d_sh_pos = nullptr;
  double *d_if_send = nullptr, *d_if_recv = nullptr;
  long long n_shared = 0;

  // asynchronous transfers through device staging buffers (so no kernel ever races with a copy in flight):
  //   download: field -> stage_out (D2D, ctx stream) -> host (D2H, d2h_stream)
  //   upload:   host -> stage_in (H2D, h2d_stream), later stage_in -> field (D2D, ctx stream) at upload_finish
  cudaStream_t d2h_stream = nullptr, h2d_stream = nullptr;
  cudaEvent_t ev_out_ready = nullptr, ev_out_done = nullptr, ev_in_done = nullptr, ev_in_free = nullptr;
  double *stage_out = nullptr, *stage_in = nullptr;
  bool out_pending = false, in_pending = false;
  std::string in_field;
  int in_block = -1;

  // instrumentation
  uint64_t launches = 0;
  double fam_ms[FAM_COUNT] = {0};
  uint64_t fam_launches[FAM_COUNT] = {0};
  struct Pending {
    int fam;
    cudaEvent_t a, b;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;

  std::string err;

  DField& field(const char* name);
  void bump(const char* name);
};
