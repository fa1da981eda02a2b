// pms_api.cpp — implementation of the C ABI in include/percept_smoother_c.h.
//
// Host-side restatement of the control flow of
//   ReferenceMeshSmootherConjugateGradientImpl (SM/ReferenceMeshSmootherConjugateGradientDef.hpp:330-591, 906-1093, 1481-1520)
//   ReferenceMeshSmootherBaseImpl::run_algorithm (SM/ReferenceMeshSmootherBase.cpp:216-334)
// with every data-parallel step dispatched to the sm_100a kernels of pms_kernels.cu.  Scalars are
// `double`, i.e. the reference's `Double` typedef under KOKKOS_ENABLE_CUDA (MeshType.hpp:41-46).
// There is no CPU compute path in this file: all field arithmetic happens in CUDA kernels.
#include <dlfcn.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <functional>
#include <limits>
#include <numeric>
#include <thread>

#include <nvtx3/nvToolsExt.h>

#include "pms_ctx.h"

using pmsk::MeshDev;

static const char* FIELD_NAMES3[] = {"coordinates", "coordinates_N", "coordinates_NM1", "coordinates_lagged",
                                     "cg_g",        "cg_r",          "cg_d",            "cg_s"};
static const char* FIELD_NAMES1[] = {"cg_edge_length", "num_adj_elems"};

DField& pms_ctx::field(const char* name) {
  auto it = fields.find(name);
  if (it == fields.end()) throw PmsError(PMS_ERR_ARG, std::string("unknown field '") + name + "' (did you call pms_commit?)");
  return it->second;
}
void pms_ctx::bump(const char* name) { field(name).version++; }

// ---------------------------------------------------------------------------------------------
// instrumentation: launch counting and optional per-family CUDA-event timing on the ctx stream
// ---------------------------------------------------------------------------------------------
namespace {

// NVTX ranges (header-only NVTX3; no cost without a tool attached) named like the reference's stk TimeBlocks, so that a
// timeline of this library reads like the reference's printTimersTableStructured(): "GATM pr 1"
// (GenericAlgorithm_total_element_metric.cpp:190-211), "GA_gg_1 calc grad pf 1" (gradient_functors.hpp:427-448),
// "GATM1 parallel_for" (GenericAlgorithm_update_coordinates.cpp:306), "RMSCGI::run_one_iteration()" (Def.hpp:1014-1016)
const char* const FAM_NVTX[FAM_COUNT] = {"GATM pr 1",   "GA_gg_1 calc grad pf 1", "nodal_field BLAS-1",       "GATM1 parallel_for",
                                         "get_alpha_0", "get_edge_lengths",       "comm_fields / sum_fields", "StructuredGridRefiner"};
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct Timed {
  pms_ctx* c;
  int fam;
  cudaEvent_t a = nullptr, b = nullptr;
  NvtxRange range;
  Timed(pms_ctx* cc, int f, int nlaunch = 1) : c(cc), fam(f), range(FAM_NVTX[f]) {
    c->launches += nlaunch;
    c->fam_launches[fam] += nlaunch;
    if (c->timing) {
      auto get = [&]() {
        cudaEvent_t e;
        if (!c->event_pool.empty()) {
          e = c->event_pool.back();
          c->event_pool.pop_back();
        } else
          cudaEventCreate(&e);
        return e;
      };
      a = get();
      b = get();
      cudaEventRecord(a, c->stream);
    }
  }
  ~Timed() {
    if (a) {
      cudaEventRecord(b, c->stream);
      c->pending.push_back({fam, a, b});
    }
  }
};

void drain_timing(pms_ctx* c) {
  if (c->pending.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (auto& p : c->pending) {
    float ms = 0;
    cudaEventElapsedTime(&ms, p.a, p.b);
    c->fam_ms[p.fam] += ms;
    c->event_pool.push_back(p.a);
    c->event_pool.push_back(p.b);
  }
  c->pending.clear();
}

void check_launch(pms_ctx* c, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw PmsError(PMS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  (void)c;
}

// copy reduction slots [slot, slot+nd) / [slot, slot+nu) to the host and wait
void fetch(pms_ctx* c, int slot, int nd, int nu) {
  if (nd) PMS_CUDA(cudaMemcpyAsync(c->h_out_d + slot, c->R.out_d + slot, nd * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (nu)
    PMS_CUDA(cudaMemcpyAsync(c->h_out_u + slot, c->R.out_u + slot, nu * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             c->stream));
  PMS_CUDA(cudaStreamSynchronize(c->stream));
}

// ---------------------------------------------------------------------------------------------
// NCCL, resolved at run time (the process may already hold torch's libnccl.so.2)
// ---------------------------------------------------------------------------------------------
struct Uid {
  char internal[128];
};
struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Uid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
Nccl g_nccl;
constexpr int NCCL_F64 = 8, NCCL_U64 = 5, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3;

void nccl_load() {
  if (g_nccl.h) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) throw PmsError(PMS_ERR_NCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define SYM(field, name)                                                                   \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                                        \
  if (!g_nccl.field) throw PmsError(PMS_ERR_NCCL, std::string("NCCL symbol missing: ") + name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
}
#define PMS_NCCL(call)                                                                                             \
  do {                                                                                                             \
    int r_ = (call);                                                                                               \
    if (r_ != 0) throw PmsError(PMS_ERR_NCCL, std::string(#call) + " failed: " + g_nccl.GetErrorString(r_));       \
  } while (0)

// all-reduce n host doubles in place (line-search / CG scalars; replaces stk::all_reduce, Def.hpp:379-385 etc.)
void allreduce(pms_ctx* c, double* v, int n, int op) {
  if (c->nranks <= 1) return;
  PMS_CUDA(cudaMemcpyAsync(c->d_coll, v, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  {
    Timed t(c, FAM_HALO);
    PMS_NCCL(g_nccl.AllReduce(c->d_coll, c->d_coll, n, NCCL_F64, op, c->nccl_comm, c->stream));
  }
  PMS_CUDA(cudaMemcpyAsync(v, c->d_coll, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PMS_CUDA(cudaStreamSynchronize(c->stream));
}

// Results of a grid reduction, combined across ranks ON THE DEVICE (in place in the result slots) and then read back
// once: one stream synchronisation per scalar instead of D2H + H2D + allreduce + D2H.
void fetch_reduced(pms_ctx* c, int slot, int nd, int nu, int op) {
  if (c->nranks > 1) {
    Timed t(c, FAM_HALO);
    PMS_NCCL(g_nccl.GroupStart());
    if (nd) PMS_NCCL(g_nccl.AllReduce(c->R.out_d + slot, c->R.out_d + slot, nd, NCCL_F64, op, c->nccl_comm, c->stream));
    if (nu) PMS_NCCL(g_nccl.AllReduce(c->R.out_u + slot, c->R.out_u + slot, nu, NCCL_U64, NCCL_SUM, c->nccl_comm, c->stream));
    PMS_NCCL(g_nccl.GroupEnd());
  }
  fetch(c, slot, nd, nu);
}

template <class T>
T* dalloc(size_t n) {
  T* p = nullptr;
  PMS_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return p;
}
template <class T>
T* dupload(const std::vector<T>& v) {
  T* p = dalloc<T>(v.size());
  if (!v.empty()) PMS_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

void require_committed(pms_ctx* c) {
  if (!c->committed) throw PmsError(PMS_ERR_ARG, "pms_commit has not been called");
}

// device_safe_transform_block_indices, MeshType.hpp:86-131
void transform_block_indices(const int idx[3], const int tr[3], const int lbeg[3], const int dbeg[3], int donor[3]) {
  int t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = (tr[j] < 0 ? -1 : 1) * (int)(std::abs(tr[j]) == i + 1);
  const int diff[3] = {idx[0] - lbeg[0], idx[1] - lbeg[1], idx[2] - lbeg[2]};
  for (int r = 0; r < 3; r++) donor[r] = t[3 * r] * diff[0] + t[3 * r + 1] * diff[1] + t[3 * r + 2] * diff[2] + dbeg[r];
}

// interface-node ownership for dots / node counts: BlockStructuredGrid.cpp:448-458, BlockStructuredGrid.hpp:156-300
bool sgrid_node_counted(const HBlock& B, int blockID, unsigned i, unsigned j, unsigned k) {
  const bool on_end = (i == 0 || i == B.n[0] - 1 || j == 0 || j == B.n[1] - 1 || k == 0 || k == B.n[2] - 1);
  if (!on_end) return true;
  bool inAny = false;
  unsigned min_block = blockID;
  for (const HZone& z : B.zones) {
    unsigned lb[3], le[3];
    for (int d = 0; d < 3; d++) {
      lb[d] = std::min(z.owner_beg[d], z.owner_end[d]) - 1;
      le[d] = std::max(z.owner_beg[d], z.owner_end[d]) - 1;
    }
    if (lb[0] <= i && i <= le[0] && lb[1] <= j && j <= le[1] && lb[2] <= k && k <= le[2]) {
      inAny = true;
      if (min_block > (unsigned)z.donor) min_block = z.donor;
    }
  }
  if (!inAny) return true;
  return min_block == (unsigned)blockID;
}

long long find_root(std::vector<long long>& parent, long long a) {
  while (parent[a] != a) {
    parent[a] = parent[parent[a]];
    a = parent[a];
  }
  return a;
}

// ---------------------------------------------------------------------------------------------
// operations shared by the C ABI and the CG driver
// ---------------------------------------------------------------------------------------------
void invalidate_cache(pms_ctx* c) {  // options / masks changed: nothing evaluated earlier may be re-used
  c->cache.clear();
  c->pre_valid = false;
  c->sinfo.valid = false;
}

void op_copy(pms_ctx* c, const char* dst, const char* src) {
  DField& d = c->field(dst);
  DField& s = c->field(src);
  if (d.ncomp != s.ncomp) throw PmsError(PMS_ERR_ARG, "copy_field: component mismatch");
  Timed t(c, FAM_BLAS1);
  PMS_CUDA(cudaMemcpyAsync(d.d, s.d, (size_t)d.ncomp * c->Npad * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  d.version++;
}

double op_dot(pms_ctx* c, const char* x, const char* y) {
  DField& fx = c->field(x);
  DField& fy = c->field(y);
  if (fx.ncomp != fy.ncomp) throw PmsError(PMS_ERR_ARG, "nodal_field_dot: component mismatch");
  {
    Timed t(c, FAM_BLAS1);
    c->L->dot(c->stream, fx.d, fy.d, c->dotmask_trivial ? nullptr : c->d_dotmask, c->N, c->Npad, fx.ncomp, c->R, 0);
    check_launch(c, "dot");
  }
  fetch_reduced(c, 0, 1, 0, NCCL_SUM);
  return c->h_out_d[0];
}

void halo_exchange(pms_ctx* c, const char* name);
void interface_sum(pms_ctx* c, DField& F);

// tet4 meshes with the reference mesh: W^-1 and det W per element, computed once per version of coordinates_NM1; the
// element kernels then read 10 coalesced values instead of gathering 4 reference vertices (12 scattered loads)
void ensure_tetref(pms_ctx* c) {
  if (c->structured || c->topo != 4 || c->dev_blocks.empty()) return;
  pmsk::MeshDev& m = c->dev_blocks[0];
  if (c->strict_fp || !c->tet_ref_cache || !c->use_ref_mesh || c->E < 1) {
    m.tetref = nullptr;
    return;
  }
  DField& W = c->field("coordinates_NM1");
  if (!c->d_tetref) c->d_tetref = dalloc<double>((size_t)10 * c->Epad);
  if (c->tetref_wv != W.version) {
    Timed t(c, FAM_EDGE);
    m.tetref = nullptr;
    c->L->tet_ref(c->stream, m, W.d, c->Npad, c->d_tetref, c->Epad);
    check_launch(c, "tet_ref");
    c->tetref_wv = W.version;
  }
  m.tetref = c->d_tetref;
  m.tetref_stride = c->Epad;
}

// total_metric(alpha), Def.hpp:330-388, without materialising x + alpha*s
// validity_suffices: the caller only needs the metric when no element is invalid (the line search on an untangled mesh
// rejects such a trial whatever its value): a cached validity-only entry with invalid elements then answers with NaN.
void op_total_metric(pms_ctx* c, int stage, double alpha, double* mtot, uint64_t* ninv, bool count_eval = true,
                     bool validity_suffices = false) {
  (c->strict_fp ? pms_strict::set_variant : pms_fast::set_variant)(c->var_metric, c->var_grad, c->sgrid_fused ? 1 : 0);  // the switches live in the kernel TU: apply this ctx's
  DField& X = c->field("coordinates");
  DField& S = c->field("cg_s");
  DField& W = c->field("coordinates_NM1");
  const bool with_s = alpha != 0.0;
  if (c->cache_metric && !c->keep_elem) {
    for (auto& e : c->cache)
      if (e.xv == X.version && e.wv == W.version && e.stage == stage && e.alpha == alpha && e.use_ref == c->use_ref_mesh &&
          (!with_s || e.sv == S.version) && (e.value_known || (validity_suffices && e.ninv > 0))) {
        *mtot = e.value_known ? e.mtot : std::numeric_limits<double>::quiet_NaN();
        if (ninv) *ninv = e.ninv;
        return;
      }
  }
  const int nb = (int)c->dev_blocks.size();
  if (nb > pms_ctx::NSLOT) throw PmsError(PMS_ERR_ARG, "too many structured blocks in one ctx");
  ensure_tetref(c);
  for (int b = 0; b < nb; b++) {
    Timed t(c, FAM_METRIC);
    c->L->metric(c->stream, c->dev_blocks[b], stage, c->use_ref_mesh ? 1 : 0, X.d, with_s ? S.d : nullptr, W.d, c->Npad,
                 alpha, c->R, b, c->keep_elem ? c->d_elem_metric : nullptr, c->keep_elem ? c->d_elem_flag : nullptr,
                 c->d_fixed);
    check_launch(c, "metric");
  }
  fetch_reduced(c, 0, nb, nb, NCCL_SUM);
  double m = 0.0;  // per-block results added in block order (Def.hpp:355-360)
  uint64_t n = 0;
  for (int b = 0; b < nb; b++) {
    m += c->h_out_d[b];
    n += c->h_out_u[b];
  }
  *mtot = m;
  if (ninv) *ninv = n;
  (void)count_eval;
  if (c->cache_metric) {
    for (auto& e : c->cache)  // a validity-only entry of this evaluation is completed, not duplicated
      if (!e.value_known && e.xv == X.version && e.wv == W.version && e.sv == S.version && e.stage == stage && e.alpha == alpha &&
          e.use_ref == c->use_ref_mesh) {
        e.mtot = m;
        e.ninv = n;
        e.value_known = true;
        return;
      }
    if (c->cache.size() >= pms_ctx::CACHE_CAP) c->cache.erase(c->cache.begin());
    c->cache.push_back({X.version, S.version, W.version, stage, alpha, m, n, c->use_ref_mesh});
  }
}

// total_metric at nb (8, 10, 12, 14 or 16) step lengths in one pass (hex meshes, one block).  Returns false when the kernel is
// not available for this mesh / build; the caller then evaluates one alpha at a time.
// nf < nb (ShapeB1; rounded up to an even count >= 4): only the last nf step lengths get their metric, the first nb - nf only
// their invalid-element count (mtot = NaN; cached as validity-only entries).  *nf_used returns what the kernel did.
bool op_total_metric_multi(pms_ctx* c, int stage, int nb, const double* alphas, double* mtot, uint64_t* ninv, int nf = 0,
                           int* nf_used = nullptr) {
  if (c->dev_blocks.size() != 1 || c->keep_elem || c->strict_fp || nb > pmsk::MULTI_ALPHA_NB) return false;
  nf = (stage == 0 || nf <= 0 || nf >= nb) ? nb : std::max(4, (nf + 1) & ~1);  // kernels: every even count from 4
  if (nf_used) *nf_used = nf;
  DField& X = c->field("coordinates");
  DField& S = c->field("cg_s");
  DField& W = c->field("coordinates_NM1");
  {
    Timed t(c, FAM_METRIC);
    (c->strict_fp ? pms_strict::set_variant : pms_fast::set_variant)(c->var_metric, c->var_grad, c->sgrid_fused ? 1 : 0);
    if (!c->L->metric_multi(c->stream, c->dev_blocks[0], stage, c->use_ref_mesh ? 1 : 0, nb, nf, X.d, S.d, W.d, c->Npad, alphas, c->R, 0,
                            c->d_fixed)) {
      c->launches--;
      c->fam_launches[FAM_METRIC]--;
      return false;
    }
    check_launch(c, "metric_multi");
  }
  fetch_reduced(c, 0, 2 * nb, 0, NCCL_SUM);
  double v[2 * pmsk::MULTI_ALPHA_NB];
  for (int k = 0; k < 2 * nb; k++) v[k] = c->h_out_d[k];
  for (int k = 0; k < nb; k++) {
    mtot[k] = k < nb - nf ? std::numeric_limits<double>::quiet_NaN() : v[k];
    ninv[k] = (uint64_t)v[nb + k];
  }
  if (c->cache_metric)
    for (int k = 0; k < nb; k++) {
      const bool known = k >= nb - nf;
      bool have = false;
      for (auto& e : c->cache)
        if (e.xv == X.version && e.wv == W.version && e.sv == S.version && e.stage == stage && e.alpha == alphas[k] &&
            e.use_ref == c->use_ref_mesh) {
          have = true;
          if (!e.value_known && known) {
            e.mtot = mtot[k];
            e.ninv = ninv[k];
            e.value_known = true;
          }
        }
      if (have) continue;
      if (c->cache.size() >= pms_ctx::CACHE_CAP) c->cache.erase(c->cache.begin());
      c->cache.push_back({X.version, S.version, W.version, stage, alphas[k], mtot[k], ninv[k], c->use_ref_mesh, known});
    }
  return true;
}

// ---------------------------------------------------------------------------------------------
// brick tables (host): Morton order of reference-mesh centroids, 128 elements per brick, brick-local node CSR,
// brick-surface partial-sum plan.  One-time cost per mesh (a few seconds at 16.8 M hexes, multi-threaded).
// ---------------------------------------------------------------------------------------------
static inline uint64_t spread21(uint64_t v) {
  v &= 0x1fffff;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

void build_bricks(pms_ctx* c) {
  if (c->structured || c->topo != 8) {
    c->bricks_failed = true;
    return;
  }
  for (void* p : c->brick_allocs) cudaFree(p);
  c->brick_allocs.clear();
  const long long N = c->N, E = c->E;
  const int BS = 128;
  DField& W = c->field("coordinates_NM1");
  std::vector<double> w((size_t)3 * N);
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; k++) PMS_CUDA(cudaMemcpy(w.data() + (size_t)k * N, W.d + (size_t)k * c->Npad, N * sizeof(double), cudaMemcpyDeviceToHost));
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int k = 0; k < 3; k++)
    for (long long n = 0; n < N; n++) {
      lo[k] = std::min(lo[k], w[(size_t)k * N + n]);
      hi[k] = std::max(hi[k], w[(size_t)k * N + n]);
    }
  double ext = 0;
  for (int k = 0; k < 3; k++) ext = std::max(ext, hi[k] - lo[k]);
  if (!(ext > 0) || !std::isfinite(ext)) {
    c->bricks_failed = true;  // reference mesh not uploaded yet / degenerate: keep the scratch path
    return;
  }
  const unsigned nthreads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  auto parallel_for = [&](long long n, const std::function<void(long long, long long)>& fn) {
    std::vector<std::thread> th;
    const long long chunk = (n + nthreads - 1) / nthreads;
    for (unsigned t = 0; t < nthreads; t++) {
      const long long a = t * chunk, b = std::min(n, a + chunk);
      if (a < b) th.emplace_back(fn, a, b);
    }
    for (auto& x : th) x.join();
  };
  std::vector<std::pair<uint64_t, int32_t>> order(E);
  const double sc = 2097151.0 / ext;
  parallel_for(E, [&](long long a, long long b) {
    for (long long e = a; e < b; e++) {
      double cx[3] = {0, 0, 0};
      for (int v = 0; v < 8; v++) {
        const long long n = c->h_conn[(size_t)e * 8 + v];
        for (int k = 0; k < 3; k++) cx[k] += w[(size_t)k * N + n];
      }
      uint64_t key = 0;
      for (int k = 0; k < 3; k++) key |= spread21((uint64_t)std::min(2097151.0, std::max(0.0, (cx[k] * 0.125 - lo[k]) * sc))) << k;
      order[e] = {key, (int32_t)e};
    }
  });
  std::sort(order.begin(), order.end());
  // greedy cut of the Morton sequence: 128 elements per brick, fewer when the brick would touch too many nodes
  const int NODE_CAP = pmsk::BRICK_MAX_NODES - 32;  // slack of 512 B in the gather-table buffer for valence > 8
  std::vector<long long> bstart{0};
  {
    std::vector<int32_t> stamp(N, -1);
    int ne = 0, nn = 0;
    for (long long s = 0; s < E; s++) {
      const int32_t e = order[s].second;
      int fresh = 0;
      const int32_t id = (int32_t)(bstart.size() - 1);
      for (int v = 0; v < 8; v++) fresh += stamp[c->h_conn[(size_t)e * 8 + v]] != id;  // (repeated nodes over-count: harmless)
      if (ne == BS || nn + fresh > NODE_CAP) {
        bstart.push_back(s);
        ne = nn = 0;
      }
      const int32_t id2 = (int32_t)(bstart.size() - 1);
      for (int v = 0; v < 8; v++) {
        int32_t& st = stamp[c->h_conn[(size_t)e * 8 + v]];
        if (st != id2) {
          st = id2;
          nn++;
        }
      }
      ne++;
    }
    bstart.push_back(E);
  }
  const long long nb = (long long)bstart.size() - 1;
  auto pad = [](long long v, long long m) { return (v + m - 1) / m * m; };
  static const int PERM[8] = {0, 1, 3, 2, 4, 5, 7, 6};  // structured position <-> Exodus local node (an involution)
  // per brick: distinct nodes (ascending), completeness, contributions in ascending (element id, local node) order
  struct BrickTmp {
    std::vector<int32_t> node, cnt;
    std::vector<uint8_t> complete;
    std::vector<uint16_t> gent;
  };
  std::vector<BrickTmp> tmp(nb);
  std::vector<uint16_t> ctab((size_t)nb * (pmsk::BRICK_CTAB_BYTES / 2), 0);
  parallel_for(nb, [&](long long a, long long b) {
    struct Ent {
      int32_t node, elem;
      uint16_t ta;
    };
    std::vector<Ent> ents;
    for (long long br = a; br < b; br++) {
      ents.clear();
      uint16_t* lconn = ctab.data() + (size_t)br * (pmsk::BRICK_CTAB_BYTES / 2);
      int32_t* belem = reinterpret_cast<int32_t*>(lconn + BS * 8);
      for (int t = 0; t < BS; t++) {
        const long long s = bstart[br] + t;
        const int32_t e = s < bstart[br + 1] ? order[s].second : -1;
        belem[t] = e;
        if (e < 0) continue;
        for (int v = 0; v < 8; v++) ents.push_back({c->h_conn[(size_t)e * 8 + v], e, (uint16_t)(t * 8 + v)});
      }
      std::sort(ents.begin(), ents.end(), [](const Ent& x, const Ent& y) {
        if (x.node != y.node) return x.node < y.node;
        if (x.elem != y.elem) return x.elem < y.elem;
        return (x.ta & 7) < (y.ta & 7);
      });
      BrickTmp& T = tmp[br];
      for (size_t i = 0; i < ents.size();) {
        size_t j = i;
        const int lj = (int)T.node.size();
        while (j < ents.size() && ents[j].node == ents[i].node) {
          const int t = ents[j].ta >> 3, p = PERM[ents[j].ta & 7];
          T.gent.push_back((uint16_t)(t * pmsk::BRICK_EX_STRIDE + p * 3));
          lconn[t * 8 + p] = (uint16_t)lj;
          j++;
        }
        const int32_t nd = ents[i].node;
        T.node.push_back(nd);
        T.cnt.push_back((int32_t)(j - i));
        T.complete.push_back((int64_t)(j - i) == c->h_csr_ptr[nd + 1] - c->h_csr_ptr[nd]);
        i = j;
      }
    }
  });
  // blob sizes / partial numbering (serial prefix), then fill in parallel
  std::vector<int32_t> goff(nb + 1, 0), part0(nb + 1, 0), bnn(nb, 0), novf(nb, 0);
  std::vector<int32_t> pcount(N, 0);
  for (long long br = 0; br < nb; br++) {
    const BrickTmp& T = tmp[br];
    const long long nn = (long long)T.node.size();
    int ov = 0, np = 0;
    for (size_t j = 0; j < T.node.size(); j++) {
      if (T.cnt[j] > 8) ov += 2 + (T.cnt[j] - 8);
      if (!T.complete[j]) {
        np++;
        pcount[T.node[j]]++;
      }
    }
    const long long bytes = 16 + 8 * pad(nn, 4) + 16 * nn + 2 * pad(ov, 8);
    if (nn > pmsk::BRICK_MAX_NODES || bytes > pmsk::BRICK_GTAB_BYTES || (long long)goff[br] + bytes / 16 >= (1ll << 31) ||
        (long long)part0[br] + np >= (1ll << 31)) {
      c->bricks_failed = true;  // (e.g. a node shared by hundreds of elements): keep the scratch path
      return;
    }
    bnn[br] = (int32_t)nn;
    novf[br] = ov;
    goff[br + 1] = goff[br] + (int32_t)(bytes / 16);
    part0[br + 1] = part0[br] + np;
  }
  const long long npart = part0[nb];
  // combine plan: nodes with partials, partial indices in ascending brick order
  std::vector<int32_t> comb_node, comb_ptr{0}, comb_slot(N, -1);
  for (long long n = 0; n < N; n++)
    if (pcount[n] > 0) {
      comb_slot[n] = (int32_t)comb_node.size();
      comb_node.push_back((int32_t)n);
      comb_ptr.push_back(comb_ptr.back() + pcount[n]);
    }
  std::vector<int32_t> comb_ent(npart), fill(comb_ptr.begin(), comb_ptr.end() - 1);
  for (long long br = 0; br < nb; br++) {
    const BrickTmp& T = tmp[br];
    int32_t q = part0[br];
    for (size_t j = 0; j < T.node.size(); j++)
      if (!T.complete[j]) comb_ent[fill[comb_slot[T.node[j]]]++] = q++;
  }
  std::vector<int32_t> gtab((size_t)goff[nb] * 4, 0);
  const bool have_owned = !c->h_node_owned.empty();
  parallel_for(nb, [&](long long a, long long b) {
    for (long long br = a; br < b; br++) {
      const BrickTmp& T = tmp[br];
      const int nn = (int)T.node.size();
      int32_t* blob = gtab.data() + (size_t)goff[br] * 4;
      blob[0] = nn;
      blob[1] = novf[br];
      int32_t* tnode = blob + 4;
      int32_t* tdst = tnode + pad(nn, 4);
      uint16_t* tell = reinterpret_cast<uint16_t*>(tdst + pad(nn, 4));
      uint16_t* tovf = tell + (size_t)nn * 8;
      int32_t q = part0[br];
      int off = 0, o = 0;
      for (int j = 0; j < nn; j++) {
        const int32_t nd = T.node[j];
        tnode[j] = nd;
        for (int k = 0; k < 8; k++) tell[j * 8 + k] = k < T.cnt[j] ? T.gent[off + k] : (uint16_t)pmsk::BRICK_EX_ZERO;
        if (T.cnt[j] > 8) {
          tovf[o++] = (uint16_t)j;
          tovf[o++] = (uint16_t)(T.cnt[j] - 8);
          for (int k = 8; k < T.cnt[j]; k++) tovf[o++] = T.gent[off + k];
        }
        off += T.cnt[j];
        if (T.complete[j])
          tdst[j] = (c->h_fixed[nd] || (have_owned && !c->h_node_owned[nd])) ? -2 : -1;
        else
          tdst[j] = q++;
      }
    }
  });

  pmsk::BrickDev& B = c->bricks;
  auto up = [&](auto& v) {
    auto* p = dupload(v);
    c->brick_allocs.push_back((void*)p);
    return p;
  };
  B.nbricks = nb;
  B.ctab = reinterpret_cast<const uint4*>(up(ctab));
  B.bnn = up(bnn);
  B.goff = up(goff);
  B.gtab = reinterpret_cast<const uint4*>(up(gtab));
  B.npartial = std::max<long long>(npart, 1);
  B.partial = dalloc<double>((size_t)3 * B.npartial);
  c->brick_allocs.push_back((void*)B.partial);
  B.ncomb = (long long)comb_node.size();
  B.comb_node = up(comb_node);
  B.comb_ptr = up(comb_ptr);
  B.comb_ent = up(comb_ent);
  c->bricks_built = true;
  c->bricks_wv = W.version;
}

// Plan of the one-kernel unstructured hex8 gradient (pms_elem_ws.cuh): nodes listed by the iteration in which their last
// adjacent element is evaluated (tile = 128 consecutive elements, iteration = tile / grid), cut into batches of grid*128,
// the per-node CSR walk resolved into scratch offsets, the slabs each batch has to wait for, and the delay D.
void build_ws_plan(pms_ctx* c) {
  c->ws_failed = true;
  const int G = c->L->grad_ws_grid();
  if (G < 1 || c->topo != 8) return;
  const long long E = c->E, N = c->N, Epad = c->Epad;
  constexpr int K = 4, TILE = 128;  // EG_K, NTP of the kernel TU
  // iterations a CTA may run ahead of the slowest one before it has to wait for it (costs as many drain iterations at the end)
  constexpr int LAG = 8;
  if (24 * Epad >= (1ll << 31) || N >= (1ll << 31)) return;  // 32-bit scratch offsets / node ids
  const long long per = (long long)G * TILE;
  const long long ntiles = (E + TILE - 1) / TILE, niter_e = (ntiles + G - 1) / G;
  if (niter_e < 2) return;  // a single iteration: nothing to ride on
  // completion iteration per node; nodes with more than 8 elements go to the overflow list
  std::vector<int32_t> comp(N), ovf;
  std::vector<long long> goff(niter_e + 1, 0);
  for (long long n = 0; n < N; n++) {
    const int64_t p0 = c->h_csr_ptr[n], p1 = c->h_csr_ptr[n + 1];
    if (p1 - p0 > 8) {
      comp[n] = -1;
      ovf.push_back((int32_t)n);
      continue;
    }
    comp[n] = p1 > p0 ? (int32_t)(((c->h_csr_ent[p1 - 1] >> 3) / TILE) / G) : 0;
    goff[comp[n] + 1]++;
  }
  for (long long i = 0; i < niter_e; i++) goff[i + 1] += goff[i];
  const long long NL = goff[niter_e];
  const long long nb = (NL + per - 1) / per;
  if (nb < 1 || nb >= (1ll << 30)) return;
  std::vector<int32_t> lnode((size_t)nb * per, -1);
  {
    std::vector<long long> fill(goff.begin(), goff.end() - 1);
    for (long long n = 0; n < N; n++)
      if (comp[n] >= 0) lnode[fill[comp[n]]++] = (int32_t)n;
  }
  std::vector<int32_t> need(nb);
  long long D = 2;
  for (long long j = 0; j < nb; j++) {
    const long long last = std::min((j + 1) * per, NL) - 1;
    need[j] = comp[lnode[last]] / K + 1;  // (the list is sorted by completion iteration)
    D = std::max(D, (long long)need[j] * K - j + 2 + LAG);
  }
  // a mesh numbered so wildly that most nodes complete only at the end gains nothing: keep the two-pass path
  if (nb + D > niter_e + niter_e / 4 + 32) return;
  std::vector<int32_t> tabA((size_t)nb * per * 4, -1), tabB((size_t)nb * per * 4, -1);
  const unsigned nthreads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  {
    std::vector<std::thread> th;
    const long long chunk = (NL + nthreads - 1) / nthreads;
    for (unsigned t = 0; t < nthreads; t++) {
      const long long qa = t * chunk, qb = std::min(NL, qa + chunk);
      if (qa >= qb) break;
      th.emplace_back([&, qa, qb]() {
        for (long long q = qa; q < qb; q++) {
          const long long n = lnode[q];
          const int64_t p0 = c->h_csr_ptr[n], p1 = c->h_csr_ptr[n + 1];
          for (int64_t p = p0; p < p1; p++) {
            const int64_t ent = c->h_csr_ent[p];
            const int32_t off = (int32_t)((ent & 7) * 3 * Epad + (ent >> 3));
            const int k = (int)(p - p0);
            (k < 4 ? tabA : tabB)[(size_t)q * 4 + (k & 3)] = off;
          }
        }
      });
    }
    for (auto& x : th) x.join();
  }
  pmsk::WsPlanDev& P = c->ws;
  P.grid = G;
  P.D = (int)D;
  P.nbatches = (int)nb;
  P.nslabs = (int)((niter_e + K - 1) / K);
  P.sleep_ns = 1000;
  P.niter = std::max(niter_e, nb + D);
  P.lnode = dupload(lnode);
  P.tabA = reinterpret_cast<const int4*>(dupload(tabA));
  P.tabB = reinterpret_cast<const int4*>(dupload(tabB));
  P.need = dupload(need);
  P.prog = dalloc<int>((size_t)P.nslabs + 1);
  P.novf = (long long)ovf.size();
  P.ovf_nodes = ovf.empty() ? nullptr : dupload(ovf);
  c->ws_grid = G;
  c->ws_built = true;
  c->ws_failed = false;
}

void ensure_eg(pms_ctx* c) {
  if (!c->d_eg) c->d_eg = dalloc<double>((size_t)3 * (c->topo == 4 ? 4 : 8) * c->Epad);
}

// get_gradient (Def.hpp:1481-1520); optionally the fused metric of the same pass
void op_gradient(pms_ctx* c, int stage, double* mtot, uint64_t* ninv, bool write_r = true) {
  (c->strict_fp ? pms_strict::set_variant : pms_fast::set_variant)(c->var_metric, c->var_grad, c->sgrid_fused ? 1 : 0);  // the switches live in the kernel TU: apply this ctx's
  DField& X = c->field("coordinates");
  DField& W = c->field("coordinates_NM1");
  DField& G = c->field("cg_g");
  DField& Rr = c->field("cg_r");
  const int nb = (int)c->dev_blocks.size();
  // structured: get_gradient() never forwards m_use_ref_mesh to m_sgrid_gradient (Def.hpp:1485-1495) => always ref mesh
  const int use_ref = c->structured ? 1 : (c->use_ref_mesh ? 1 : 0);
  bool fused_ok = false;
  if (c->structured && c->sgrid_fused) {
    if (!c->d_seam) {  // tile-seam scratch of the fused kernel, sized for the largest block (blocks run one after another)
      size_t bytes = 0;
      for (auto& mb : c->dev_blocks) bytes = std::max(bytes, c->L->sgrid_fused_scratch_bytes(mb));
      c->d_seam = dalloc<double>(bytes / sizeof(double) + 1);
    }
    fused_ok = true;
    for (int b = 0; b < nb && fused_ok; b++) {
      Timed t(c, FAM_GRADIENT, 2);
      fused_ok = c->L->grad_fused_sgrid(c->stream, c->dev_blocks[b], stage, X.d, W.d, c->Npad, c->d_fixed, G.d, c->R, b,
                                        c->d_seam, c->keep_elem ? c->d_elem_metric : nullptr,
                                        c->keep_elem ? c->d_elem_flag : nullptr);
      check_launch(c, "grad_fused_sgrid");
      if (!fused_ok) {
        if (b > 0) throw PmsError(PMS_ERR_CUDA, "fused structured gradient refused a block after accepting another");
        c->launches -= 2;
        c->fam_launches[FAM_GRADIENT] -= 2;
      }
    }
  }
  if (!fused_ok && !c->structured && c->topo == 8 && c->elem_ws && c->var_grad == 2 && !c->strict_fp && !c->keep_elem &&
      (stage == 0 || use_ref)) {
    if (!c->ws_built && !c->ws_failed) build_ws_plan(c);
    if (c->ws_built) {
      ensure_eg(c);
      Timed t(c, FAM_GRADIENT, c->ws.novf > 0 ? 2 : 1);
      PMS_CUDA(cudaMemsetAsync(c->ws.prog, 0, ((size_t)c->ws.nslabs + 1) * sizeof(int), c->stream));
      fused_ok = c->L->grad_ws(c->stream, c->dev_blocks[0], c->ws, stage, use_ref, X.d, W.d, c->Npad, c->d_fixed, c->d_node_owned,
                               G.d, write_r ? Rr.d : nullptr, c->d_eg, c->Epad, c->R, 0);
      if (!fused_ok) {
        cudaGetLastError();
        c->ws_built = false;
        c->ws_failed = true;  // (launch refused: not tried again)
        c->launches -= c->ws.novf > 0 ? 2 : 1;
        c->fam_launches[FAM_GRADIENT] -= c->ws.novf > 0 ? 2 : 1;
      } else {
        check_launch(c, "grad_ws");
      }
    }
  }
  if (!fused_ok && !c->structured && c->topo == 8 && c->var_grad >= 4 && !c->strict_fp && !c->keep_elem && use_ref) {
    if (!c->bricks_built && !c->bricks_failed) build_bricks(c);
    if (c->bricks_built) {
      Timed t(c, FAM_GRADIENT, 2);
      fused_ok = c->L->grad_bricks(c->stream, c->dev_blocks[0], c->bricks, stage, X.d, W.d, c->Npad, c->d_fixed, c->d_node_owned,
                                   G.d, write_r ? Rr.d : nullptr, c->R, 0, nullptr, nullptr);
      check_launch(c, "grad_bricks");
      if (!fused_ok) {
        c->launches -= 2;
        c->fam_launches[FAM_GRADIENT] -= 2;
      }
    }
  }
  if (!fused_ok) {
    ensure_eg(c);
    ensure_tetref(c);
    for (int b = 0; b < nb; b++) {
      Timed t(c, FAM_GRADIENT, 2);
      c->L->grad_elements(c->stream, c->dev_blocks[b], stage, use_ref, X.d, W.d, c->Npad, c->d_eg, c->Epad, c->R, b,
                          c->keep_elem ? c->d_elem_metric : nullptr, c->keep_elem ? c->d_elem_flag : nullptr);
      check_launch(c, "grad_elements");
      c->L->grad_gather(c->stream, c->dev_blocks[b], c->d_eg, c->Epad, c->d_fixed, c->structured ? nullptr : c->d_node_owned,
                        G.d, (c->structured || !write_r) ? nullptr : Rr.d, c->Npad);
      check_launch(c, "grad_gather");
    }
  }
  if (c->n_groups > 0) {  // inter-block sum + propagate, gradient_functors.hpp:479-482
    Timed t(c, FAM_GRADIENT);
    c->L->group_sum(c->stream, c->n_groups, c->d_grp_ptr, c->d_grp_nodes, G.d, 3, c->Npad);
    check_launch(c, "group_sum");
  }
  if (c->structured) interface_sum(c, G);  // block-per-GPU: complete the sums over ranks
  G.version++;
  if (!c->structured && write_r) Rr.version++;
  if (c->smooth_surfaces && c->n_surf_nodes > 0) {  // GenericAlgorithm_get_gradient_2 (Def.hpp:1401-1447): after cg_r = cg_g
    Timed t(c, FAM_GRADIENT);
    c->L->surf_project(c->stream, c->n_surf_nodes, c->d_surf_nodes, c->d_surf_sid, c->d_surfaces, X.d, G.d, c->Npad);
    check_launch(c, "surf_project");
  }
  if (mtot) {
    fetch_reduced(c, 0, nb, nb, NCCL_SUM);
    double m = 0.0;
    uint64_t n = 0;
    for (int b = 0; b < nb; b++) {
      m += c->h_out_d[b];
      n += c->h_out_u[b];
    }
    *mtot = m;
    if (ninv) *ninv = n;
    // NOTE: the fused value is not inserted into the metric cache: it equals total_metric(0) only to rounding
    // (different instruction schedule), and cached values must be bitwise what the metric kernel returns.
  }
}

void op_edge_lengths(pms_ctx* c) {
  DField& W = c->field("coordinates_NM1");
  DField& EL = c->field("cg_edge_length");
  DField& NA = c->field("num_adj_elems");
  if (!c->structured) ensure_eg(c);  // the gradient scratch doubles as the per-element edge-length buffer
  for (auto& mb : c->dev_blocks) {
    Timed t(c, FAM_EDGE, c->structured ? 1 : 2);
    c->L->edge_lengths(c->stream, mb, W.d, c->Npad, EL.d, NA.d, c->structured ? nullptr : c->d_eg);
    check_launch(c, "edge_lengths");
  }
  if (c->structured) {
    if (c->n_groups > 0) {  // sum + propagate across interfaces, get_edge_len_avg.hpp:333-341
      Timed t(c, FAM_EDGE, 2);
      c->L->group_sum(c->stream, c->n_groups, c->d_grp_ptr, c->d_grp_nodes, EL.d, 1, c->Npad);
      c->L->group_sum(c->stream, c->n_groups, c->d_grp_ptr, c->d_grp_nodes, NA.d, 1, c->Npad);
    }
    interface_sum(c, EL);
    interface_sum(c, NA);
    Timed t(c, FAM_EDGE);
    c->L->divide(c->stream, EL.d, NA.d, c->N);
    check_launch(c, "divide");
  }
  EL.version++;
  NA.version++;
}

double op_alpha0(pms_ctx* c) {
  DField& S = c->field("cg_s");
  DField& EL = c->field("cg_edge_length");
  {
    Timed t(c, FAM_ALPHA0);
    // structured: all nodes (get_alpha_0_refmesh.hpp:168-193); STK: skip fixed nodes (Def.hpp:876-878)
    c->L->alpha0(c->stream, S.d, EL.d, c->structured ? nullptr : c->d_skip, c->N, c->Npad, c->R, 0);
    check_launch(c, "alpha0");
  }
  fetch_reduced(c, 0, 1, 0, NCCL_MIN);
  double a = c->h_out_d[0];
  if (a == DBL_MAX) a = 1.0;  // nothing moves: get_alpha_0_refmesh.hpp:164-165, Def.hpp:851-852
  if (!(a > 0.0)) throw PmsError(PMS_ERR_SMOOTHER, "bad alpha");
  return a;
}

void op_update(pms_ctx* c, double alpha, double* dmax, double* dmax_rel) {
  DField& X = c->field("coordinates");
  DField& LG = c->field("coordinates_lagged");
  DField& S = c->field("cg_s");
  DField& EL = c->field("cg_edge_length");
  {
    Timed t(c, FAM_UPDATE);
    c->L->update(c->stream, X.d, LG.d, S.d, EL.d, c->d_fixed, c->d_node_owned, alpha, c->N, c->Npad, c->R, 0);
    check_launch(c, "update");
  }
  X.version++;
  LG.version++;
  fetch_reduced(c, 0, 2, 0, NCCL_MAX);
  double v[2] = {c->h_out_d[0], c->h_out_d[1]};
  *dmax = v[0];
  *dmax_rel = v[1];
}

// snap_nodes (Def.hpp:236-316): non-fixed MS_SURFACE nodes back onto their surface; returns the largest displacement
double op_snap_nodes(pms_ctx* c) {
  if (c->n_surf_nodes == 0) return 0.0;
  DField& X = c->field("coordinates");
  {
    Timed t(c, FAM_UPDATE);
    c->L->surf_snap(c->stream, c->n_surf_nodes, c->d_surf_nodes, c->d_surf_sid, c->d_surfaces, X.d, c->Npad, c->R, 0);
    check_launch(c, "surf_snap");
  }
  X.version++;
  fetch_reduced(c, 0, 1, 0, NCCL_MAX);
  double v[1] = {c->h_out_d[0]};
  return v[0];
}

uint64_t op_count_invalid(pms_ctx* c) {  // MeshSmoother.cpp:192-234: untangle metric's valid flag
  double m;
  uint64_t n;
  if (c->cache_metric && !c->keep_elem) {
    // validity (any corner detA <= 0) is the same rule in both metrics and independent of alpha = 0 evaluations'
    // stage, so any cached evaluation at the current coordinates already holds the count
    const uint64_t xv = c->field("coordinates").version;
    for (auto& e : c->cache)
      if (e.xv == xv && e.alpha == 0.0) return e.ninv;
  }
  op_total_metric(c, 0, 0.0, &m, &n);
  return n;
}

void halo_exchange(pms_ctx* c, const char* name) {
  if (c->nranks <= 1 || c->nb_ranks.empty()) return;
  DField& F = c->field(name);
  const int nc = F.ncomp;
  const int nn = (int)c->nb_ranks.size();
  Timed t(c, FAM_HALO, 2 * nn);
  for (int q = 0; q < nn; q++) {
    const long long cnt = c->send_off[q + 1] - c->send_off[q];
    c->L->pack(c->stream, F.d, c->Npad, nc, c->d_send_nodes + c->send_off[q], cnt, c->d_send_buf + 3 * c->send_off[q]);
  }
  PMS_NCCL(g_nccl.GroupStart());
  for (int q = 0; q < nn; q++) {
    const long long sc = c->send_off[q + 1] - c->send_off[q], rc = c->recv_off[q + 1] - c->recv_off[q];
    if (sc) PMS_NCCL(g_nccl.Send(c->d_send_buf + 3 * c->send_off[q], sc * nc, NCCL_F64, c->nb_ranks[q], c->nccl_comm, c->stream));
    if (rc) PMS_NCCL(g_nccl.Recv(c->d_recv_buf + 3 * c->recv_off[q], rc * nc, NCCL_F64, c->nb_ranks[q], c->nccl_comm, c->stream));
  }
  PMS_NCCL(g_nccl.GroupEnd());
  for (int q = 0; q < nn; q++) {
    const long long cnt = c->recv_off[q + 1] - c->recv_off[q];
    c->L->unpack(c->stream, F.d, c->Npad, nc, c->d_recv_nodes + c->recv_off[q], cnt, c->d_recv_buf + 3 * c->recv_off[q]);
  }
  check_launch(c, "halo");
  F.version++;
}

// sum a nodal field over the copies of interface nodes held by different ranks (structured block per GPU)
void interface_sum(pms_ctx* c, DField& F) {
  if (c->nranks <= 1 || c->if_ranks.empty()) return;
  const int nc = F.ncomp, nn = (int)c->if_ranks.size();
  Timed t(c, FAM_HALO, nn + 1);
  for (int q = 0; q < nn; q++) {
    const long long cnt = c->if_off[q + 1] - c->if_off[q];
    c->L->pack(c->stream, F.d, c->Npad, nc, c->d_if_nodes + c->if_off[q], cnt, c->d_if_send + 3 * c->if_off[q]);
  }
  PMS_NCCL(g_nccl.GroupStart());
  for (int q = 0; q < nn; q++) {
    const long long cnt = c->if_off[q + 1] - c->if_off[q];
    if (!cnt) continue;
    PMS_NCCL(g_nccl.Send(c->d_if_send + 3 * c->if_off[q], cnt * nc, NCCL_F64, c->if_ranks[q], c->nccl_comm, c->stream));
    PMS_NCCL(g_nccl.Recv(c->d_if_recv + 3 * c->if_off[q], cnt * nc, NCCL_F64, c->if_ranks[q], c->nccl_comm, c->stream));
  }
  PMS_NCCL(g_nccl.GroupEnd());
  c->L->shared_sum(c->stream, c->n_shared, c->d_sh_nodes, c->d_sh_ptr, c->d_sh_pos, c->d_sh_cnt, c->d_if_recv, F.d, nc, c->Npad);
  check_launch(c, "interface_sum");
  F.version++;
}

// ---------------------------------------------------------------------------------------------
// CG driver (host control flow; Double = double)
// ---------------------------------------------------------------------------------------------
struct Driver {
  pms_ctx* c;
  pms_state* st;
  double gradNorm;
  Driver(pms_ctx* cc, pms_state* s, double gn) : c(cc), st(s), gradNorm(gn) {}

  // validity_suffices: see op_total_metric (the Armijo loops on an untangled mesh: `converged && total_valid`)
  double tm(double alpha, bool& valid, uint64_t* ninv = nullptr, bool validity_suffices = false) {
    double m;
    uint64_t n;
    const uint64_t before = c->fam_launches[FAM_METRIC];
    op_total_metric(c, st->stage, alpha, &m, &n, true, validity_suffices);
    if (c->fam_launches[FAM_METRIC] != before) st->n_metric_evals++;
    valid = (n == 0);
    if (ninv) *ninv = n;
    return m;
  }

  // Line-search speculation: the backtracking trials are alpha, alpha*tau, alpha*tau^2, ... (and the quadratic fit's
  // a1 = alpha/2 continues the same sequence), known before their outcomes.  One multi-alpha pass evaluates the next
  // MULTI_ALPHA_NB of them into the metric cache; tm() then finds them there.  Same decisions, fewer passes.
  void speculate(double alpha, double tau) {
    if (!c->speculate_line_search || !c->cache_metric || c->keep_elem) return;
    {
      const uint64_t xv = c->field("coordinates").version, sv = c->field("cg_s").version, wv = c->field("coordinates_NM1").version;
      for (auto& e : c->cache)
        if (e.xv == xv && e.wv == wv && e.sv == sv && e.stage == st->stage && e.alpha == alpha && e.use_ref == c->use_ref_mesh &&
            (e.value_known || (st->untangled && e.ninv > 0)))
          return;
    }
    // batch size from the previous line search: it visited spec_prev trials of the sequence (+1 for the quadratic
    // fit's a1); 8 when that is enough, else 12
    // the number of halvings changes slowly from one iteration to the next (and grows as the run converges): cover the
    // previous search's trials + 1 more + the quadratic fit's a1; a search that outruns the batch gets a batch of 8 more
    constexpr int NBMAX = pmsk::MULTI_ALPHA_NB;
    const int need = c->spec_prev > 0 ? c->spec_prev + 2 : 12, done = spec_cur;
    const int nb = done > 0 ? 8 : std::min(NBMAX, std::max(8, (need + 1) & ~1));  // kernels for 8, 10, 12, 14, 16 step lengths
    double al[NBMAX], m[NBMAX];
    uint64_t n[NBMAX];
    double a = alpha;
    for (int k = 0; k < nb; k++) {
      al[k] = a;
      a *= tau;  // the very multiplication the loop below performs, so the cache keys match bit for bit
    }
    // On an untangled mesh a trial with invalid elements is rejected whatever its metric, and the sequence starts with far
    // too long steps: values are asked only from one trial before the previous batch's first all-valid one onwards, the
    // earlier trials get their invalid count only (a wrong guess costs a single-alpha pass for the trial concerned).
    int nf = nb;
    if (st->stage == 1 && st->untangled && done == 0 && c->spec_kv_prev >= 0) nf = nb - std::max(0, c->spec_kv_prev - 1);
    if (op_total_metric_multi(c, st->stage, nb, al, m, n, nf)) {
      st->n_metric_evals++;
      if (done == 0) {
        int kv = 0;
        while (kv < nb && n[kv] != 0) kv++;
        c->spec_kv_prev = kv;
      }
    }
  }
  int spec_cur = 0;  // trials visited by the current line search (c->spec_prev: by the previous one)

  // Pass fusion (with line-search speculation on): the fused metric+gradient kernel's metric is put in the metric cache
  // as the alpha = 0 value at these coordinates, and a gradient computed at the end of the previous iteration is reused.
  bool fuse_passes() const { return c->speculate_line_search && c->cache_metric && !c->keep_elem && c->use_ref_mesh && !c->strict_fp; }
  bool pre_matches() const {
    return c->pre_valid && c->pre_xv == c->field("coordinates").version && c->pre_wv == c->field("coordinates_NM1").version &&
           c->pre_gv == c->field("cg_g").version && c->pre_stage == st->stage && fuse_passes();
  }
  void mark_pre() {
    c->pre_valid = true;
    c->pre_xv = c->field("coordinates").version;
    c->pre_wv = c->field("coordinates_NM1").version;
    c->pre_gv = c->field("cg_g").version;
    c->pre_stage = st->stage;
  }
  void gradient_here(double* m_out = nullptr, uint64_t* n_out = nullptr) {
    DField& X = c->field("coordinates");
    DField& W = c->field("coordinates_NM1");
    if (pre_matches()) {
      c->pre_valid = false;
      return;  // cg_g already holds f'(x) for these coordinates
    }
    c->pre_valid = false;
    double m;
    uint64_t n;
    op_gradient(c, st->stage, &m, &n, /*write_r=*/false);  // cg_r = cg_g (Def.hpp:1508) is overwritten by r = -g
    if (m_out) *m_out = m;
    if (n_out) *n_out = n;
    st->n_gradient_evals++;
    // (a metric of exactly 0 is the reference's early-exit test `metric_check == 0.0`, Def.hpp:1041: at that noise floor the
    //  metric kernel's own value decides, as without fusion)
    if (fuse_passes() && m != 0.0) {
      bool have = false;
      for (auto& e : c->cache)
        have = have || (e.xv == X.version && e.wv == W.version && e.stage == st->stage && e.alpha == 0.0 && e.use_ref == c->use_ref_mesh);
      if (!have) {
        if (c->cache.size() >= pms_ctx::CACHE_CAP) c->cache.erase(c->cache.begin());
        c->cache.push_back({X.version, c->field("cg_s").version, W.version, st->stage, 0.0, m, n, c->use_ref_mesh});
      }
    }
  }

  bool sinfo_ok(bool need_g) const {
    return c->sinfo.valid && c->sinfo.sv == c->field("cg_s").version && (!need_g || c->sinfo.gv == c->field("cg_g").version);
  }
  double alpha0_of_s() {
    if (!sinfo_ok(false)) return op_alpha0(c);
    double a = c->sinfo.alpha0;
    if (a == DBL_MAX) a = 1.0;  // as op_alpha0
    if (!(a > 0.0)) throw PmsError(PMS_ERR_SMOOTHER, "bad alpha");
    return a;
  }
  double s_dot_g() { return sinfo_ok(true) ? c->sinfo.sg : op_dot(c, "cg_s", "cg_g"); }
  double s_dot_s() { return sinfo_ok(false) ? c->sinfo.ss : op_dot(c, "cg_s", "cg_s"); }

  void s_equals_r() {  // copy_field(cg_s, cg_r)
    op_copy(c, "cg_s", "cg_r");
    halo_exchange(c, "cg_s");
  }

  // line_search, Def.hpp:466-591
  bool moved = false;  // line_search already applied the accepted step (update_node_positions) and holds f'(x_new)
  bool allow_move = true;  // false for the stand-alone pms_line_search, which must leave the coordinates alone
  double line_search(bool& restarted, const double mfac_mult = 1.0) {
    moved = false;
    const double alpha_fac = 10.0, tau = 0.5, c0 = 1.e-1, min_alpha_factor = 1.e-12;
    st->alpha_0 = alpha0_of_s();
    double alpha = alpha_fac * st->alpha_0;
    bool total_valid = false;
    uint64_t n_invalid = 0;
    const double metric_0 = tm(0.0, total_valid, &n_invalid);
    double metric = 0.0;
    double sDotGrad = s_dot_g();
    if (sDotGrad >= 0.0) {
      s_equals_r();
      st->alpha_0 = op_alpha0(c);
      alpha = alpha_fac * st->alpha_0;
      sDotGrad = op_dot(c, "cg_s", "cg_g");
      restarted = true;
    }
    if (!(sDotGrad < 0.0)) throw PmsError(PMS_ERR_SMOOTHER, "bad sDotGrad");
    const double armijo_offset_factor = c0 * sDotGrad;
    bool converged = false;
    total_valid = false;
    int liter = 0;
    const int niter = 1000;
    spec_cur = 0;
    while (!converged && liter < niter) {
      speculate(alpha, tau);
      spec_cur++;
      metric = tm(alpha, total_valid, &n_invalid, st->untangled != 0);
      const double mfac = alpha * armijo_offset_factor * mfac_mult;
      converged = (metric < metric_0 + mfac);
      if (st->untangled) converged = converged && total_valid;
      if (!converged) {
        alpha *= tau;
        st->n_line_search_halvings++;
      }
      if (alpha < min_alpha_factor * st->alpha_0) break;
      ++liter;
    }
    if (!converged) {
      restarted = true;
      s_equals_r();
      st->alpha_0 = op_alpha0(c);
      alpha = alpha_fac * st->alpha_0;
      sDotGrad = op_dot(c, "cg_s", "cg_g");
      if (!(sDotGrad < 0.0)) throw PmsError(PMS_ERR_SMOOTHER, "bad sDotGrad 2nd time");
    }
    liter = 0;
    while (!converged && liter < niter) {
      speculate(alpha, tau);
      spec_cur++;
      metric = tm(alpha, total_valid, nullptr, st->untangled != 0);
      const double mfac = alpha * armijo_offset_factor * mfac_mult;
      converged = (metric < metric_0 + mfac);
      if (st->untangled) {
        converged = converged && total_valid;
        if (total_valid && st->dmax_relative < gradNorm) converged = true;
      }
      if (!converged) {
        alpha *= tau;
        st->n_line_search_halvings++;
      }
      if (alpha < min_alpha_factor * st->alpha_0) break;
      ++liter;
    }
    if (!converged) throw PmsError(PMS_ERR_SMOOTHER, "can't reduce metric");
    c->spec_prev = spec_cur;
    const double a1 = alpha / 2., a2 = alpha;
    const double f0 = metric_0, f1 = tm(a1, total_valid), f2 = tm(a2, total_valid);
    const double den = 2. * (a2 * (-f0 + f1) + a1 * (f0 - f2));
    const double num = a2 * a2 * (f1 - f0) + a1 * a1 * (f0 - f2);
    if (std::fabs(den) > 1.e-10) {
      const double alpha_quadratic = num / den;
      if (alpha_quadratic > 1.e-10 && alpha_quadratic < 2 * alpha) {
        if (fuse_passes() && !c->smooth_surfaces && allow_move) {
          // The quadratic step is accepted almost always, and an accepted step is followed by the update and by the
          // fused metric+gradient pass at the new coordinates.  So move there first and let that pass provide fm;
          // only a rejected step costs extra (restore x from coordinates_lagged).
          op_update(c, alpha_quadratic, &st->dmax, &st->dmax_relative);
          double fm = 0.0;
          uint64_t nq = 0;
          c->pre_valid = false;
          gradient_here(&fm, &nq);
          if (fm < f2 && (st->stage == 0 || nq == 0)) {
            alpha = alpha_quadratic;
            mark_pre();
            moved = true;
          } else {
            op_copy(c, "coordinates", "coordinates_lagged");
          }
        } else {
          const double fm = tm(alpha_quadratic, total_valid);
          if (fm < f2 && (st->stage == 0 || total_valid)) alpha = alpha_quadratic;
        }
      }
    }
    return alpha;
  }

  bool check_convergence() const { return pms_check_convergence(st, gradNorm) != 0; }

  // run_one_iteration, Def.hpp:1010-1093
  double run_one_iteration() {
    NvtxRange range("RMSCGI::run_one_iteration()");
    bool total_valid = false;
    if (st->iter == 0) {
      op_edge_lengths(c);
      for (const char* nm : {"cg_r", "cg_d", "cg_s"}) {
        DField& f = c->field(nm);
        Timed t(c, FAM_BLAS1);
        c->L->set_value(c->stream, f.d, 0.0, c->N, c->Npad, 3);
        f.version++;
      }
    }
    // f'(x) (+ the metric at the same x from the same pass, which is what metric_check evaluates)
    gradient_here();
    // r = -g ; dold = d.d ; dmid = r.d ; dnew = r.r   (Def.hpp:1034-1037)
    {
      DField& G = c->field("cg_g");
      DField& D = c->field("cg_d");
      DField& Rr = c->field("cg_r");
      {
        Timed t(c, FAM_BLAS1);
        c->L->cg_r_dots(c->stream, G.d, D.d, Rr.d, c->dotmask_trivial ? nullptr : c->d_dotmask, c->N, c->Npad, c->R, 0);
        check_launch(c, "cg_r_dots");
      }
      Rr.version++;
      fetch_reduced(c, 0, 3, 0, NCCL_SUM);
      double v[3] = {c->h_out_d[0], c->h_out_d[1], c->h_out_d[2]};
      st->dold = v[0];
      st->dmid = v[1];
      st->dnew = v[2];
    }
    if (st->iter == 0) st->d0 = st->dnew;
    const double metric_check = tm(0.0, total_valid);
    st->total_metric = metric_check;
    if (check_convergence() || metric_check == 0.0) return tm(0.0, total_valid);

    double cg_beta = 0.0;
    if (st->dold == 0.0)
      cg_beta = 0.0;
    else if (st->iter > 0)
      cg_beta = (st->dnew - st->dmid) / st->dold;
    const uint64_t N = st->num_nodes;
    const bool restart = (N == 0 || (uint64_t)st->iter % N == 0 || cg_beta <= 0.0);
    {
      DField& Rr = c->field("cg_r");
      DField& S = c->field("cg_s");
      DField& D = c->field("cg_d");
      c->sinfo.valid = false;
      if (c->cache_metric) {
        // one pass: the update plus s.g, s.s and the alpha_0 scan the line search asks for next (bit-identical values)
        DField& G = c->field("cg_g");
        DField& EL = c->field("cg_edge_length");
        {
          Timed t(c, FAM_BLAS1);
          c->L->cg_s_update_dots(c->stream, Rr.d, S.d, D.d, G.d, EL.d, c->dotmask_trivial ? nullptr : c->d_dotmask,
                                 c->structured ? nullptr : c->d_skip, cg_beta, restart ? 1 : 0, c->N, c->Npad, c->R, 0);
          check_launch(c, "cg_s_update_dots");
        }
        S.version++;
        D.version++;
        if (c->nranks > 1) {
          Timed t(c, FAM_HALO);
          PMS_NCCL(g_nccl.GroupStart());
          PMS_NCCL(g_nccl.AllReduce(c->R.out_d, c->R.out_d, 2, NCCL_F64, NCCL_SUM, c->nccl_comm, c->stream));
          PMS_NCCL(g_nccl.AllReduce(c->R.out_d + 2, c->R.out_d + 2, 1, NCCL_F64, NCCL_MIN, c->nccl_comm, c->stream));
          PMS_NCCL(g_nccl.GroupEnd());
        }
        halo_exchange(c, "cg_s");
        fetch(c, 0, 3, 0);
        c->sinfo.valid = true;
        c->sinfo.sv = S.version;
        c->sinfo.gv = G.version;
        c->sinfo.sg = c->h_out_d[0];
        c->sinfo.ss = c->h_out_d[1];
        c->sinfo.alpha0 = c->h_out_d[2];
      } else {
        {
          Timed t(c, FAM_BLAS1);
          c->L->cg_s_update(c->stream, Rr.d, S.d, D.d, cg_beta, restart ? 1 : 0, c->N, c->Npad);
          check_launch(c, "cg_s_update");
        }
        S.version++;
        D.version++;
        halo_exchange(c, "cg_s");
      }
    }
    bool restarted = false;
    const double alpha = line_search(restarted);
    st->restarted += restarted ? 1 : 0;
    const double snorm = s_dot_s();
    st->grad_norm_scaled = st->alpha_0 * std::sqrt(snorm) / double(st->num_nodes);
    st->alpha = alpha;
    if (!moved) op_update(c, alpha, &st->dmax, &st->dmax_relative);  // update_node_positions, Def.hpp:395-410
    if (c->smooth_surfaces) {  // check if re-snapped geometry is acceptable, Def.hpp:1078-1089
      op_snap_nodes(c);
      if (st->stage != 0) {
        bool total_valid_0 = true;
        tm(0.0, total_valid_0);
        if (!total_valid_0) throw PmsError(PMS_ERR_SMOOTHER, "bad mesh after snap_node_positions...");
      }
    }
    if (fuse_passes()) {
      // the metric at the new coordinates comes from the fused metric+gradient pass whose gradient the next
      // iteration starts with (same x): one pass instead of a metric pass now and a gradient pass then
      if (!pre_matches()) {
        c->pre_valid = false;
        gradient_here();
        mark_pre();
      }
    }
    return tm(0.0, total_valid);
  }

  // run_algorithm, ReferenceMeshSmootherBase.cpp:240-334
  void run_algorithm(int innerIter, const char* log_path) {
    std::ofstream myFile;
    const unsigned width = 14;
    const bool log = log_path && c->rank == 0;
    if (log) {  // ReferenceMeshSmootherBase.cpp:57-81
      myFile.open(log_path, std::ios::out);
      myFile.precision(6);
      myFile.setf(std::ios_base::left, std::ios_base::adjustfield);
      myFile.setf(std::ios_base::scientific);
      myFile << std::setw(8) << "%iter" << std::setw(width) << "global_metric" << std::setw(width) << "invalid_elems"
             << std::setw(width) << "dmax" << std::setw(width) << "dmax_rel" << std::setw(width) << "grad/g0"
             << std::setw(width) << "gradScaled" << std::setw(width) << "alpha" << std::setw(width) << "alpha_0"
             << std::setw(6) << "stage" << std::setw(width) << "untangled_elems" << std::setw(width) << "noNodes"
             << std::endl;
    }
    pms_state_init(c, st);
    op_copy(c, "coordinates_lagged", "coordinates_NM1");
    {  // coordinates = 1.0*coordinates_N + 0.0*coordinates_NM1 + 0.0*coordinates   (:259)
      DField& X = c->field("coordinates");
      Timed t(c, FAM_BLAS1);
      c->L->axpbypgz(c->stream, 1.0, c->field("coordinates_N").d, 0.0, c->field("coordinates_NM1").d, 0.0, X.d, c->N, c->Npad, 3);
      X.version++;
    }
    st->num_invalid = op_count_invalid(c);
    st->untangled = (st->num_invalid == 0);
    for (int stage = 0; stage < 2; stage++) {
      st->stage = stage;
      if (stage == 1) {
        const uint64_t num_invalid_1 = op_count_invalid(c);
        if (num_invalid_1 != 0) throw PmsError(PMS_ERR_SMOOTHER, "Invalid elements exist for start of stage 2, quiting");
      }
      for (int iter = 0; iter < innerIter; ++iter) {
        st->iter = iter;
        st->num_invalid = op_count_invalid(c);
        const int num_invalid_0 = (int)st->num_invalid;
        if (!st->untangled && st->num_invalid == 0) st->untangled = 1;
        st->global_metric = run_one_iteration();
        const bool conv = check_convergence();
        st->converged = conv;
        if (log) {  // print_iteration_status, :216-238
          myFile << std::setw(8) << st->iter << std::setw(width) << st->global_metric << std::setw(width) << num_invalid_0
                 << std::setw(width) << st->dmax << std::setw(width) << st->dmax_relative << std::setw(width)
                 << std::sqrt(st->dnew / st->d0) << std::setw(width) << st->grad_norm_scaled << std::setw(width) << st->alpha
                 << std::setw(width) << st->alpha_0 << std::setw(6) << st->stage << std::setw(width) << (bool)st->untangled
                 << std::setw(width) << st->num_nodes << std::endl;
        }
        if (!st->untangled && st->num_invalid == 0) st->untangled = 1;
        if (conv && st->untangled) break;
      }
    }
    op_copy(c, "coordinates_lagged", "coordinates");
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
// Every entry point runs on the context's own device (a process may hold contexts on several GPUs).  A field whose raw
// device pointer has been handed out (pms_field_device_ptr) may have been written behind the library's back since the
// last call: its version is bumped on every entry, so the metric cache, the precomputed gradient and the search-
// direction scalars keyed on it are never re-used across calls (until pms_field_release_ptr).
static inline void pms_enter(pms_ctx* c) {
  if (c->stream) cudaSetDevice(c->device);
  if (c->n_aliased)
    for (auto& kv : c->fields)
      if (kv.second.aliased) kv.second.version++;
}
#define PMS_TRY \
  if (!c) return PMS_ERR_ARG; \
  try {      \
    pms_enter(c);
#define PMS_CATCH                   \
  }                                 \
  catch (const PmsError& e) {       \
    c->err = e.what();              \
    return e.code;                  \
  }                                 \
  catch (const std::exception& e) { \
    c->err = e.what();              \
    return PMS_ERR_ARG;             \
  }                                 \
  return PMS_OK;

extern "C" {

int pms_create(pms_ctx** out, int device) {
  if (!out) return PMS_ERR_ARG;
  *out = nullptr;
  pms_ctx* c = new pms_ctx();
  c->device = device;
  try {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw PmsError(PMS_ERR_CUDA, std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                                       "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) throw PmsError(PMS_ERR_ARG, "device ordinal out of range");
    PMS_CUDA(cudaSetDevice(device));
    PMS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->L = pms_fast::launchers();
  } catch (const PmsError& e) {
    c->err = e.what();
    *out = c;  // returned so the caller can read pms_last_error, then pms_destroy
    return e.code;
  }
  *out = c;
  return PMS_OK;
}

void pms_destroy(pms_ctx* c) {
  if (!c) return;
  if (c->stream) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->fields) cudaFree(kv.second.d);
    for (void* p : c->brick_allocs) cudaFree(p);
    void* ptrs[] = {c->d_conn,      c->d_fixed,     c->d_node_owned, c->d_elem_owned, c->d_dotmask,  c->d_elem_bits,
                    c->d_csr_ptr,   c->d_csr_ent,   c->d_grp_ptr,    c->d_grp_nodes,  c->d_eg,       c->d_elem_metric,
                    c->d_elem_flag, c->R.partial_d, c->R.partial_u,  c->R.ticket,     c->R.out_d,    c->R.out_u,
                    c->d_send_nodes, c->d_recv_nodes, c->d_send_buf, c->d_recv_buf,   c->d_coll, c->d_skip, c->d_csr_ent32, c->d_if_nodes, c->d_sh_nodes, c->d_sh_cnt, c->d_sh_ptr, c->d_sh_pos,
                    c->d_if_send, c->d_if_recv, c->d_seam, c->d_tetref, (void*)c->ws.lnode, (void*)c->ws.tabA, (void*)c->ws.tabB, (void*)c->ws.need, c->ws.prog, (void*)c->ws.ovf_nodes};
    for (void* p : ptrs)
      if (p) cudaFree(p);
    if (c->h_out_d) cudaFreeHost(c->h_out_d);
    if (c->h_out_u) cudaFreeHost(c->h_out_u);
    for (auto& p : c->pending) {
      cudaEventDestroy(p.a);
      cudaEventDestroy(p.b);
    }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    if (c->d2h_stream) {
      cudaStreamSynchronize(c->d2h_stream);
      cudaStreamSynchronize(c->h2d_stream);
      cudaStreamDestroy(c->d2h_stream);
      cudaStreamDestroy(c->h2d_stream);
      for (cudaEvent_t e : {c->ev_out_ready, c->ev_out_done, c->ev_in_done, c->ev_in_free}) cudaEventDestroy(e);
      if (c->stage_out) cudaFree(c->stage_out);
      if (c->stage_in) cudaFree(c->stage_in);
    }
    cudaStreamDestroy(c->stream);
  }
  delete c;
}

const char* pms_last_error(const pms_ctx* c) { return c ? c->err.c_str() : "null ctx"; }

static void apply_classification(pms_ctx* c);

int pms_set_option(pms_ctx* c, const char* key, double v) {
  PMS_TRY
  const std::string k(key);
  if (k == "use_ref_mesh") {
    c->use_ref_mesh = v != 0;
  } else if (k == "strict_fp") {
    c->strict_fp = v != 0;
    c->L = c->strict_fp ? pms_strict::launchers() : pms_fast::launchers();
    invalidate_cache(c);
  } else if (k == "cache_metric") {
    c->cache_metric = v != 0;
    invalidate_cache(c);
  } else if (k == "keep_element_metrics") {
    c->keep_elem = v != 0;
    if (c->keep_elem && c->committed && !c->d_elem_metric) {
      c->d_elem_metric = dalloc<double>(c->Epad);
      c->d_elem_flag = dalloc<int32_t>(c->Epad);
    }
  } else if (k == "speculate_line_search") {
    c->speculate_line_search = v != 0;
  } else if (k == "smooth_surfaces") {  // PerceptMesh::get_smooth_surfaces (mesh_adapt --smooth_surfaces)
    c->smooth_surfaces = v != 0;
    if (c->committed) apply_classification(c);
    invalidate_cache(c);
  } else if (k == "variant_metric_corners") {
    c->var_metric = (int)v;
    pms_fast::set_variant(c->var_metric, c->var_grad);
    invalidate_cache(c);
  } else if (k == "variant_grad_corners") {
    c->var_grad = (int)v;
    pms_fast::set_variant(c->var_metric, c->var_grad);
    invalidate_cache(c);
  } else if (k == "tet_ref_cache") {  // 1 (default): tet4 kernels read the cached reference part per element; 0: gather the reference vertices
    c->tet_ref_cache = v != 0;
    invalidate_cache(c);
  } else if (k == "elem_ws") {  // 1: one-kernel unstructured gradient (element loop + lagged gather); 0 (default): two passes
    c->elem_ws = v != 0;
    invalidate_cache(c);
  } else if (k == "sgrid_fused") {  // 1 (default): one-pass tile kernel for structured gradients; 0: element pass + gather
    c->sgrid_fused = v != 0;
    invalidate_cache(c);
  } else
    throw PmsError(PMS_ERR_ARG, "unknown option " + k);
  PMS_CATCH
}

int pms_add_structured_block(pms_ctx* c, const unsigned node_size[3], int* block_id) {
  PMS_TRY
  if (c->committed) throw PmsError(PMS_ERR_ARG, "mesh already committed");
  if (!c->structured && c->topo != 0) throw PmsError(PMS_ERR_ARG, "ctx already holds an unstructured mesh");
  for (int d = 0; d < 3; d++)
    if (node_size[d] < 2) throw PmsError(PMS_ERR_ARG, "structured block needs >= 2 nodes per direction");
  HBlock b;
  for (int d = 0; d < 3; d++) b.n[d] = node_size[d];
  c->structured = true;
  c->blocks.push_back(b);
  if (block_id) *block_id = (int)c->blocks.size() - 1;
  PMS_CATCH
}

int pms_block_add_boundary_condition(pms_ctx* c, int block, const int rb[3], const int re[3]) {
  PMS_TRY
  if (block < 0 || block >= (int)c->blocks.size()) throw PmsError(PMS_ERR_ARG, "bad block id");
  HBC bc;
  for (int d = 0; d < 3; d++) {
    bc.beg[d] = rb[d];
    bc.end[d] = re[d];
  }
  c->blocks[block].bcs.push_back(bc);
  PMS_CATCH
}

int pms_block_add_zone_connectivity(pms_ctx* c, int block, int donor, const int tr[3], const int ob[3], const int oe[3],
                                    const int db[3], const int de[3]) {
  PMS_TRY
  if (block < 0 || block >= (int)c->blocks.size() || donor < 0) throw PmsError(PMS_ERR_ARG, "bad block id");
  if (c->committed) throw PmsError(PMS_ERR_ARG, "mesh already committed");
  // 1-based inclusive ranges inside the owner block (the donor side is checked at commit, when every block is known)
  for (int d = 0; d < 3; d++) {
    const int n = (int)c->blocks[block].n[d];
    if (ob[d] < 1 || ob[d] > n || oe[d] < 1 || oe[d] > n)
      throw PmsError(PMS_ERR_ARG, "zone connectivity: owner range outside the block's nodes");
    if (std::abs(tr[d]) < 1 || std::abs(tr[d]) > 3) throw PmsError(PMS_ERR_ARG, "zone connectivity: transform entries must be +-1, +-2, +-3");
  }
  HZone z;
  z.donor = donor;
  for (int d = 0; d < 3; d++) {
    z.transform[d] = tr[d];
    z.owner_beg[d] = ob[d];
    z.owner_end[d] = oe[d];
    z.donor_beg[d] = db[d];
    z.donor_end[d] = de[d];
  }
  c->blocks[block].zones.push_back(z);
  PMS_CATCH
}

int pms_set_unstructured(pms_ctx* c, int topology, size_t nnodes, size_t nelems, const int32_t* conn,
                         const uint8_t* elem_owned, const uint8_t* node_owned) {
  PMS_TRY
  if (c->committed || c->structured) throw PmsError(PMS_ERR_ARG, "ctx already holds a mesh");
  if (topology != PMS_HEX8 && topology != PMS_TET4) throw PmsError(PMS_ERR_ARG, "topology must be PMS_HEX8 or PMS_TET4");
  if (!conn || nnodes == 0 || nelems == 0) throw PmsError(PMS_ERR_ARG, "empty mesh");
  if (nnodes >= (1ull << 31)) throw PmsError(PMS_ERR_ARG, "more than 2^31 nodes per ctx");
  c->topo = topology;
  c->N = (long long)nnodes;
  c->E = (long long)nelems;
  c->h_conn.assign(conn, conn + nelems * topology);
  for (int32_t v : c->h_conn)
    if (v < 0 || (size_t)v >= nnodes) throw PmsError(PMS_ERR_ARG, "connectivity index out of range");
  if (elem_owned) c->h_elem_owned.assign(elem_owned, elem_owned + nelems);
  if (node_owned) c->h_node_owned.assign(node_owned, node_owned + nnodes);
  PMS_CATCH
}

// nodes skipped by get_alpha_0 on STK meshes: fixed (Def.hpp:876-878) or not owned by this rank
static void rebuild_skip_mask(pms_ctx* c) {
  if (c->structured) return;
  std::vector<uint8_t> skip(c->Npad, 1);
  for (long long n = 0; n < c->N; n++)
    skip[n] = (c->h_fixed[n] || (!c->h_node_owned.empty() && !c->h_node_owned[n])) ? 1 : 0;
  if (!c->d_skip) c->d_skip = dalloc<uint8_t>(c->Npad);
  PMS_CUDA(cudaMemcpy(c->d_skip, skip.data(), c->Npad, cudaMemcpyHostToDevice));
}

// get_fixed_flag with a geometry and no boundary selector (MeshSmoother.cpp:293-340): vertex / curve nodes fixed,
// surface nodes fixed unless smooth_surfaces, volume nodes free; then the list of non-fixed surface nodes
static void apply_classification(pms_ctx* c) {
  if (c->h_node_dof.empty()) return;
  std::vector<uint8_t> fx(c->N);
  std::vector<int32_t> nodes, sid;
  for (long long n = 0; n < c->N; n++) {
    const int dof = c->h_node_dof[n];
    fx[n] = (dof == 0 || dof == 1 || (dof == 2 && !c->smooth_surfaces)) ? 1 : 0;
    if (dof == 2 && !fx[n]) {
      nodes.push_back((int32_t)n);
      sid.push_back(c->h_node_surf[n]);
    }
  }
  if (pms_set_fixed_mask(c, fx.data()) != PMS_OK) throw PmsError(PMS_ERR_ARG, c->err);
  for (void* p : {(void*)c->d_surf_nodes, (void*)c->d_surf_sid, (void*)c->d_surfaces})
    if (p) cudaFree(p);
  c->d_surf_nodes = c->d_surf_sid = nullptr;
  c->d_surfaces = nullptr;
  c->n_surf_nodes = (long long)nodes.size();
  if (c->n_surf_nodes > 0) {
    c->d_surf_nodes = dupload(nodes);
    c->d_surf_sid = dupload(sid);
    c->d_surfaces = dupload(c->h_surfaces);
  }
}

static long long total_nodes_precommit(pms_ctx* c) {
  if (!c->structured) return c->N;
  long long n = 0;
  for (auto& b : c->blocks) n += b.nnodes();
  return n;
}

int pms_set_fixed_mask(pms_ctx* c, const uint8_t* fixed) {
  PMS_TRY
  const long long n = total_nodes_precommit(c);
  if (fixed)
    c->h_fixed.assign(fixed, fixed + n);
  else
    c->h_fixed.assign(n, 0);
  c->fixed_set = true;
  if (c->committed) {
    PMS_CUDA(cudaMemcpy(c->d_fixed, c->h_fixed.data(), n, cudaMemcpyHostToDevice));
    rebuild_skip_mask(c);
    c->bricks_built = false;  // the brick tables bake the fixed mask in
    for (auto& mb : c->dev_blocks) c->L->elem_fixed_bits(c->stream, mb, c->d_fixed, c->d_elem_bits);
    check_launch(c, "elem_fixed_bits");
    invalidate_cache(c);
  }
  PMS_CATCH
}

int pms_select_boundary_sgrid(pms_ctx* c) {  // SGridBoundarySelector, ReferenceMeshSmootherConjugateGradient.hpp:31-69
  PMS_TRY
  if (!c->structured) throw PmsError(PMS_ERR_ARG, "not a structured mesh");
  long long N = 0;
  for (auto& b : c->blocks) {
    b.off = N;
    N += b.nnodes();
  }
  std::vector<uint8_t> fx(N, 0);
  for (auto& b : c->blocks)
    for (const HBC& bc : b.bcs) {
      long long beg[3], end[3];
      for (int d = 0; d < 3; d++) {
        beg[d] = std::min(bc.beg[d], bc.end[d]) - 1;
        end[d] = std::max(bc.beg[d], bc.end[d]) - 1;
        beg[d] = std::max<long long>(beg[d], 0);
        end[d] = std::min<long long>(end[d], (long long)b.n[d] - 1);
      }
      for (long long k = beg[2]; k <= end[2]; k++)
        for (long long j = beg[1]; j <= end[1]; j++)
          for (long long i = beg[0]; i <= end[0]; i++) fx[b.node(i, j, k)] = 1;
    }
  return pms_set_fixed_mask(c, fx.data());
  PMS_CATCH
}

int pms_commit(pms_ctx* c) {
  PMS_TRY
  if (c->committed) throw PmsError(PMS_ERR_ARG, "already committed");
  if (!c->structured && c->topo == 0) throw PmsError(PMS_ERR_ARG, "no mesh defined");
  PMS_CUDA(cudaSetDevice(c->device));
  if (c->structured) {
    long long off = 0, eoff = 0;
    for (auto& b : c->blocks) {
      b.off = off;
      b.eoff = eoff;
      off += b.nnodes();
      eoff += b.ncells();
    }
    c->N = off;
    c->E = eoff;
    c->topo = 8;
    for (auto& b : c->blocks)
      for (auto& z : b.zones)
      {
        if (z.donor >= (int)c->blocks.size()) throw PmsError(PMS_ERR_ARG, "zone connectivity donor block does not exist");
        for (int d = 0; d < 3; d++) {
          const int n = (int)c->blocks[z.donor].n[d];
          if (z.donor_beg[d] < 1 || z.donor_beg[d] > n || z.donor_end[d] < 1 || z.donor_end[d] > n)
            throw PmsError(PMS_ERR_ARG, "zone connectivity: donor range outside the donor block's nodes");
        }
      }
  }
  if (c->N >= (1ll << 31)) throw PmsError(PMS_ERR_ARG, "more than 2^31 nodes per ctx");
  c->Npad = (c->N + 31) / 32 * 32;
  c->Epad = (c->E + 31) / 32 * 32;
  if (!c->fixed_set) c->h_fixed.assign(c->N, 0);
  if ((long long)c->h_fixed.size() != c->N) throw PmsError(PMS_ERR_ARG, "fixed mask length != node count");

  // fields (add_coordinate_state_fields, PerceptMesh.cpp:4490-4539)
  for (const char* n : FIELD_NAMES3) {
    DField f;
    f.ncomp = 3;
    f.d = dalloc<double>((size_t)3 * c->Npad);
    PMS_CUDA(cudaMemset(f.d, 0, (size_t)3 * c->Npad * sizeof(double)));
    c->fields[n] = f;
  }
  for (const char* n : FIELD_NAMES1) {
    DField f;
    f.ncomp = 1;
    f.d = dalloc<double>((size_t)c->Npad);
    PMS_CUDA(cudaMemset(f.d, 0, (size_t)c->Npad * sizeof(double)));
    c->fields[n] = f;
  }
  c->d_fixed = dalloc<uint8_t>(c->Npad);
  PMS_CUDA(cudaMemset(c->d_fixed, 0, c->Npad));
  PMS_CUDA(cudaMemcpy(c->d_fixed, c->h_fixed.data(), c->N, cudaMemcpyHostToDevice));
  c->d_elem_bits = dalloc<uint8_t>(c->Epad);

  // ownership masks for dots / node counts
  c->h_dotmask.assign(c->N, 1);
  c->dotmask_trivial = true;
  if (c->structured) {
    for (int ib = 0; ib < (int)c->blocks.size(); ib++) {
      const HBlock& B = c->blocks[ib];
      if (B.zones.empty()) continue;
      for (unsigned k = 0; k < B.n[2]; k++)
        for (unsigned j = 0; j < B.n[1]; j++)
          for (unsigned i = 0; i < B.n[0]; i++)
            if (!sgrid_node_counted(B, ib, i, j, k)) {
              c->h_dotmask[B.node(i, j, k)] = 0;
              c->dotmask_trivial = false;
            }
    }
  } else if (!c->h_node_owned.empty()) {
    c->h_dotmask = c->h_node_owned;
    for (uint8_t v : c->h_dotmask)
      if (!v) c->dotmask_trivial = false;
  }
  c->num_nodes_owned = 0;
  for (uint8_t v : c->h_dotmask) c->num_nodes_owned += v ? 1 : 0;
  if (!c->dotmask_trivial) {
    c->d_dotmask = dalloc<uint8_t>(c->Npad);
    PMS_CUDA(cudaMemset(c->d_dotmask, 0, c->Npad));
    PMS_CUDA(cudaMemcpy(c->d_dotmask, c->h_dotmask.data(), c->N, cudaMemcpyHostToDevice));
  }

  // device mesh descriptors
  c->dev_blocks.clear();
  if (c->structured) {
    for (auto& b : c->blocks) {
      MeshDev m{};
      m.kind = pmsk::KIND_SGRID;
      m.ni = b.n[0];
      m.nj = b.n[1];
      m.nk = b.n[2];
      m.node_off = b.off;
      m.elem_off = b.eoff;
      m.nnodes = b.nnodes();
      m.nelems = b.ncells();
      m.elem_fixed_bits = c->d_elem_bits;
      c->dev_blocks.push_back(m);
    }
    // interface duplicate groups from the zone connectivities (union-find over flat node ids)
    bool any_zone = false;
    for (auto& b : c->blocks) any_zone = any_zone || !b.zones.empty();
    if (any_zone) {
      std::vector<long long> parent(c->N);
      std::iota(parent.begin(), parent.end(), 0);
      std::vector<uint8_t> touched(c->N, 0);
      for (int ib = 0; ib < (int)c->blocks.size(); ib++) {
        const HBlock& B = c->blocks[ib];
        for (const HZone& z : B.zones) {
          const HBlock& D = c->blocks[z.donor];
          int lb[3], le[3];
          for (int d = 0; d < 3; d++) {
            lb[d] = std::min(z.owner_beg[d], z.owner_end[d]);
            le[d] = std::max(z.owner_beg[d], z.owner_end[d]);
          }
          for (int k = lb[2]; k <= le[2]; k++)
            for (int j = lb[1]; j <= le[1]; j++)
              for (int i = lb[0]; i <= le[0]; i++) {
                const int li[3] = {i, j, k};
                int di[3];
                transform_block_indices(li, z.transform, z.owner_beg, z.donor_beg, di);
                for (int d = 0; d < 3; d++)
                  if (di[d] < 1 || di[d] > (int)D.n[d]) throw PmsError(PMS_ERR_ARG, "zone connectivity maps outside the donor block");
                const long long a = B.node(i - 1, j - 1, k - 1), bnode = D.node(di[0] - 1, di[1] - 1, di[2] - 1);
                touched[a] = touched[bnode] = 1;
                const long long ra = find_root(parent, a), rb = find_root(parent, bnode);
                if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
              }
        }
      }
      std::map<long long, std::vector<long long>> groups;
      for (long long n = 0; n < c->N; n++)
        if (touched[n]) groups[find_root(parent, n)].push_back(n);  // ascending n == ascending block id
      std::vector<int64_t> gptr{0}, gnodes;
      for (auto& kv : groups) {
        if (kv.second.size() < 2) continue;
        for (long long n : kv.second) gnodes.push_back(n);
        gptr.push_back((int64_t)gnodes.size());
      }
      c->n_groups = (long long)gptr.size() - 1;
      if (c->n_groups > 0) {
        c->d_grp_ptr = dupload(gptr);
        c->d_grp_nodes = dupload(gnodes);
      }
    }
  } else {
    const int npe = c->topo;
    // SoA connectivity
    std::vector<int32_t> soa((size_t)npe * c->Epad, 0);
    for (long long e = 0; e < c->E; e++)
      for (int a = 0; a < npe; a++) soa[(size_t)a * c->Epad + e] = c->h_conn[(size_t)e * npe + a];
    c->d_conn = dupload(soa);
    // node -> element CSR, entries sorted by (element, local node): deterministic gather order = ascending element id
    c->h_csr_ptr.assign(c->N + 1, 0);
    for (long long e = 0; e < c->E; e++)
      for (int a = 0; a < npe; a++) c->h_csr_ptr[c->h_conn[(size_t)e * npe + a] + 1]++;
    for (long long n = 0; n < c->N; n++) c->h_csr_ptr[n + 1] += c->h_csr_ptr[n];
    c->h_csr_ent.assign(c->h_csr_ptr[c->N], 0);
    {
      std::vector<int64_t> fill(c->h_csr_ptr.begin(), c->h_csr_ptr.end() - 1);
      for (long long e = 0; e < c->E; e++)
        for (int a = 0; a < npe; a++) c->h_csr_ent[fill[c->h_conn[(size_t)e * npe + a]]++] = (int64_t)e * 8 + a;
    }
    c->d_csr_ptr = dupload(c->h_csr_ptr);
    c->d_csr_ent = dupload(c->h_csr_ent);
    if (c->E * 8 < (1ll << 31)) {
      std::vector<int32_t> e32(c->h_csr_ent.begin(), c->h_csr_ent.end());
      c->d_csr_ent32 = dupload(e32);
    }
    if (!c->h_elem_owned.empty()) {
      std::vector<uint8_t> eo(c->Epad, 0);
      std::copy(c->h_elem_owned.begin(), c->h_elem_owned.end(), eo.begin());
      c->d_elem_owned = dupload(eo);
    }
    if (!c->h_node_owned.empty()) {
      std::vector<uint8_t> no(c->Npad, 0);
      std::copy(c->h_node_owned.begin(), c->h_node_owned.end(), no.begin());
      c->d_node_owned = dupload(no);
    }
    MeshDev m{};
    m.kind = npe == 8 ? pmsk::KIND_HEX8 : pmsk::KIND_TET4;
    m.nnodes = c->N;
    m.nelems = c->E;
    m.conn = c->d_conn;
    m.conn_stride = c->Epad;
    m.elem_fixed_bits = c->d_elem_bits;
    m.elem_owned = c->d_elem_owned;
    m.csr_ptr = c->d_csr_ptr;
    m.csr_ent = c->d_csr_ent;
    m.csr_ent32 = c->d_csr_ent32;
    c->dev_blocks.push_back(m);
  }

  if (!c->strict_fp && !c->dev_blocks.empty()) c->L->prepare_line_search_kernels(c->dev_blocks[0]);

  // reduction scratch: partial slots for the largest grid any kernel launches
  int nsm_dev = 148;
  if (cudaDeviceGetAttribute(&nsm_dev, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess || nsm_dev < 1) nsm_dev = 148;
  long long maxb = (long long)nsm_dev * 16;  // the persistent kernels launch at most SMs x 16 CTAs
  for (auto& m : c->dev_blocks) {
    long long nb = m.kind == pmsk::KIND_SGRID
                       ? (long long)((m.ni + 30) / 32) * ((m.nj + 2) / 4) * ((m.nk + 0) / 2 + 1)
                       : (m.nelems + 255) / 256;
    if (m.kind == pmsk::KIND_SGRID)  // the lane-cooperative variants (8 x 4 x 64 cell columns): larger grids on thin slabs
      nb = std::max(nb, (long long)((m.ni + 6) / 8) * ((m.nj + 2) / 4) * ((m.nk + 62) / 64));
    maxb = std::max(maxb, nb + 8);
  }
  c->R.max_blocks = (int)maxb;
  c->R.partial_d = dalloc<double>((size_t)pmsk::REDUCE_MAX_NV * maxb);
  c->R.partial_u = dalloc<unsigned long long>((size_t)maxb);
  c->R.ticket = dalloc<unsigned int>(2);
  PMS_CUDA(cudaMemset(c->R.ticket, 0, 2 * sizeof(unsigned int)));
  c->R.out_d = dalloc<double>(pms_ctx::NSLOT);
  c->R.out_u = dalloc<unsigned long long>(pms_ctx::NSLOT);
  PMS_CUDA(cudaMallocHost(&c->h_out_d, pms_ctx::NSLOT * sizeof(double)));
  PMS_CUDA(cudaMallocHost(&c->h_out_u, pms_ctx::NSLOT * sizeof(unsigned long long)));
  c->d_coll = dalloc<double>(32);
  if (c->keep_elem) {
    c->d_elem_metric = dalloc<double>(c->Epad);
    c->d_elem_flag = dalloc<int32_t>(c->Epad);
  }
  c->committed = true;
  rebuild_skip_mask(c);
  for (auto& mb : c->dev_blocks) c->L->elem_fixed_bits(c->stream, mb, c->d_fixed, c->d_elem_bits);
  check_launch(c, "elem_fixed_bits");
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  PMS_CATCH
}

int pms_num_nodes_local(const pms_ctx* c, size_t* n) {
  if (!c || !n) return PMS_ERR_ARG;
  *n = (size_t)(c->committed ? c->N : total_nodes_precommit(const_cast<pms_ctx*>(c)));
  return PMS_OK;
}
int pms_num_elements_local(const pms_ctx* c, size_t* n) {
  if (!c || !n) return PMS_ERR_ARG;
  *n = (size_t)c->E;
  return PMS_OK;
}
int pms_get_node_elem_csr(const pms_ctx* cc, int64_t* row_ptr, int64_t* entries) {
  pms_ctx* c = const_cast<pms_ctx*>(cc);
  PMS_TRY
  require_committed(c);
  if (c->structured) throw PmsError(PMS_ERR_ARG, "structured meshes use the implicit 8-cell stencil, no CSR");
  memcpy(row_ptr, c->h_csr_ptr.data(), c->h_csr_ptr.size() * sizeof(int64_t));
  memcpy(entries, c->h_csr_ent.data(), c->h_csr_ent.size() * sizeof(int64_t));
  PMS_CATCH
}

static void block_range(pms_ctx* c, int block, long long& off, long long& n) {
  if (block < 0) {
    off = 0;
    n = c->N;
    return;
  }
  if (!c->structured || block >= (int)c->blocks.size()) throw PmsError(PMS_ERR_ARG, "bad block id");
  off = c->blocks[block].off;
  n = c->blocks[block].nnodes();
}

int pms_field_upload(pms_ctx* c, const char* name, int block, const double* host, int layout) {
  PMS_TRY
  require_committed(c);
  DField& f = c->field(name);
  long long off, n;
  block_range(c, block, off, n);
  if (layout == PMS_SOA) {
    for (int k = 0; k < f.ncomp; k++)
      PMS_CUDA(cudaMemcpyAsync(f.d + (size_t)k * c->Npad + off, host + (size_t)k * n, n * sizeof(double),
                               cudaMemcpyHostToDevice, c->stream));
    PMS_CUDA(cudaStreamSynchronize(c->stream));
  } else if (layout == PMS_AOS) {
    std::vector<double> tmp((size_t)f.ncomp * n);
    for (int k = 0; k < f.ncomp; k++)
      for (long long i = 0; i < n; i++) tmp[(size_t)k * n + i] = host[(size_t)i * f.ncomp + k];
    for (int k = 0; k < f.ncomp; k++)
      PMS_CUDA(cudaMemcpyAsync(f.d + (size_t)k * c->Npad + off, tmp.data() + (size_t)k * n, n * sizeof(double),
                               cudaMemcpyHostToDevice, c->stream));
    PMS_CUDA(cudaStreamSynchronize(c->stream));
  } else
    throw PmsError(PMS_ERR_ARG, "bad layout");
  f.version++;
  PMS_CATCH
}

int pms_field_download(pms_ctx* c, const char* name, int block, double* host, int layout) {
  PMS_TRY
  require_committed(c);
  DField& f = c->field(name);
  long long off, n;
  block_range(c, block, off, n);
  if (layout == PMS_SOA) {
    for (int k = 0; k < f.ncomp; k++)
      PMS_CUDA(cudaMemcpyAsync(host + (size_t)k * n, f.d + (size_t)k * c->Npad + off, n * sizeof(double),
                               cudaMemcpyDeviceToHost, c->stream));
    PMS_CUDA(cudaStreamSynchronize(c->stream));
  } else if (layout == PMS_AOS) {
    std::vector<double> tmp((size_t)f.ncomp * n);
    for (int k = 0; k < f.ncomp; k++)
      PMS_CUDA(cudaMemcpyAsync(tmp.data() + (size_t)k * n, f.d + (size_t)k * c->Npad + off, n * sizeof(double),
                               cudaMemcpyDeviceToHost, c->stream));
    PMS_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < f.ncomp; k++)
      for (long long i = 0; i < n; i++) host[(size_t)i * f.ncomp + k] = tmp[(size_t)k * n + i];
  } else
    throw PmsError(PMS_ERR_ARG, "bad layout");
  PMS_CATCH
}

// Asynchronous transfers.  Both go through a device staging buffer, so kernels queued later never race with a copy:
// download: field -> stage_out on the ctx stream (D2D), then stage_out -> host on a second stream;
// upload:   host -> stage_in on a third stream at once, stage_in -> field on the ctx stream at pms_field_upload_finish.
static void ensure_transfer_streams(pms_ctx* c) {
  if (c->d2h_stream) return;
  PMS_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  PMS_CUDA(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&c->ev_out_ready, &c->ev_out_done, &c->ev_in_done, &c->ev_in_free})
    PMS_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
}
int pms_field_download_async(pms_ctx* c, const char* name, int block, double* host) {
  PMS_TRY
  require_committed(c);
  DField& f = c->field(name);
  long long off, n;
  block_range(c, block, off, n);
  ensure_transfer_streams(c);
  if (!c->stage_out) c->stage_out = dalloc<double>((size_t)3 * c->Npad);
  if (c->out_pending) PMS_CUDA(cudaStreamWaitEvent(c->stream, c->ev_out_done, 0));  // previous read-back still uses stage_out
  for (int k = 0; k < f.ncomp; k++)
    PMS_CUDA(cudaMemcpyAsync(c->stage_out + (size_t)k * c->Npad, f.d + (size_t)k * c->Npad + off, n * sizeof(double),
                             cudaMemcpyDeviceToDevice, c->stream));
  PMS_CUDA(cudaEventRecord(c->ev_out_ready, c->stream));
  PMS_CUDA(cudaStreamWaitEvent(c->d2h_stream, c->ev_out_ready, 0));
  for (int k = 0; k < f.ncomp; k++)
    PMS_CUDA(cudaMemcpyAsync(host + (size_t)k * n, c->stage_out + (size_t)k * c->Npad, n * sizeof(double), cudaMemcpyDeviceToHost,
                             c->d2h_stream));
  PMS_CUDA(cudaEventRecord(c->ev_out_done, c->d2h_stream));
  c->out_pending = true;
  PMS_CATCH
}
int pms_field_upload_async(pms_ctx* c, const char* name, int block, const double* host) {
  PMS_TRY
  require_committed(c);
  DField& f = c->field(name);
  long long off, n;
  block_range(c, block, off, n);
  if (c->in_pending) throw PmsError(PMS_ERR_ARG, "an asynchronous upload is already in flight: call pms_field_upload_finish first");
  ensure_transfer_streams(c);
  if (!c->stage_in) c->stage_in = dalloc<double>((size_t)3 * c->Npad);
  PMS_CUDA(cudaStreamWaitEvent(c->h2d_stream, c->ev_in_free, 0));  // (no-op until recorded) stage_in was consumed
  for (int k = 0; k < f.ncomp; k++)
    PMS_CUDA(cudaMemcpyAsync(c->stage_in + (size_t)k * c->Npad, host + (size_t)k * n, n * sizeof(double), cudaMemcpyHostToDevice,
                             c->h2d_stream));
  PMS_CUDA(cudaEventRecord(c->ev_in_done, c->h2d_stream));
  c->in_pending = true;
  c->in_field = name;
  c->in_block = block;
  PMS_CATCH
}
int pms_field_upload_finish(pms_ctx* c) {
  PMS_TRY
  if (!c->in_pending) throw PmsError(PMS_ERR_ARG, "no asynchronous upload in flight");
  DField& f = c->field(c->in_field.c_str());
  long long off, n;
  block_range(c, c->in_block, off, n);
  PMS_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in_done, 0));
  for (int k = 0; k < f.ncomp; k++)
    PMS_CUDA(cudaMemcpyAsync(f.d + (size_t)k * c->Npad + off, c->stage_in + (size_t)k * c->Npad, n * sizeof(double),
                             cudaMemcpyDeviceToDevice, c->stream));
  PMS_CUDA(cudaEventRecord(c->ev_in_free, c->stream));
  f.version++;
  c->in_pending = false;
  PMS_CATCH
}
int pms_transfer_wait(pms_ctx* c) {
  PMS_TRY
  if (c->d2h_stream) PMS_CUDA(cudaStreamSynchronize(c->d2h_stream));
  c->out_pending = false;
  PMS_CATCH
}

int pms_field_device_ptr(pms_ctx* c, const char* name, double** dptr, size_t* comp_stride) {
  PMS_TRY
  require_committed(c);
  DField& f = c->field(name);
  *dptr = f.d;
  *comp_stride = (size_t)c->Npad;
  f.version++;  // the caller may write through the pointer ...
  if (!f.aliased) {  // ... now and later: see pms_enter
    f.aliased = true;
    c->n_aliased++;
  }
  PMS_CATCH
}

int pms_field_release_ptr(pms_ctx* c, const char* name) {
  PMS_TRY
  require_committed(c);
  DField& f = c->field(name);
  if (f.aliased) {
    f.aliased = false;
    c->n_aliased--;
  }
  f.version++;
  PMS_CATCH
}

int pms_copy_field(pms_ctx* c, const char* dst, const char* src) {
  PMS_TRY
  require_committed(c);
  op_copy(c, dst, src);
  if (c->nranks > 1) halo_exchange(c, dst);
  PMS_CATCH
}

// StructuredGridRefiner::do_refine_structured, host part (structured/StructuredGridRefiner.cpp:40-170): sizes 2n-1
// (get_new_sizes :19-38), 1-based BC / zone-connectivity ranges r -> 2(r-1)+1 (m_index_base = 1, .hpp:26).
int pms_refine_structured_topology(const pms_ctx* coarse, pms_ctx* c, unsigned long long* num_refined_cells) {
  PMS_TRY
  if (!coarse || !coarse->structured) throw PmsError(PMS_ERR_ARG, "refiner input is not a structured grid");
  if (c == coarse || c->committed || c->structured || c->topo != 0) throw PmsError(PMS_ERR_ARG, "refiner output ctx must be empty");
  const int base = 1;
  unsigned long long ncell = 0;
  for (const HBlock& b : coarse->blocks) {
    HBlock nb;
    for (int d = 0; d < 3; d++) {
      if (2ull * b.n[d] - 1 > 0x7fffffffull) throw PmsError(PMS_ERR_ARG, "refined block too large");
      nb.n[d] = 2 * b.n[d] - 1;
    }
    for (const HBC& bc : b.bcs) {
      HBC r = bc;
      for (int d = 0; d < 3; d++) {
        r.beg[d] = 2 * (bc.beg[d] - base) + base;
        r.end[d] = 2 * (bc.end[d] - base) + base;
      }
      nb.bcs.push_back(r);
    }
    for (const HZone& z : b.zones) {
      HZone r = z;
      for (int d = 0; d < 3; d++) {
        r.owner_beg[d] = 2 * (z.owner_beg[d] - base) + base;
        r.owner_end[d] = 2 * (z.owner_end[d] - base) + base;
        r.donor_beg[d] = 2 * (z.donor_beg[d] - base) + base;
        r.donor_end[d] = 2 * (z.donor_end[d] - base) + base;
      }
      nb.zones.push_back(r);
    }
    ncell += (unsigned long long)nb.ncells();
    c->blocks.push_back(nb);
  }
  c->structured = true;
  if (num_refined_cells) *num_refined_cells = ncell;
  PMS_CATCH
}

// StructuredGridRefinerImpl::refine (structured/StructuredGridRefinerDef.hpp:88-222) for every block, on the device
int pms_refine_structured_field(pms_ctx* coarse, const char* src_field, pms_ctx* c, const char* dst_field) {
  PMS_TRY
  if (!coarse) throw PmsError(PMS_ERR_ARG, "null coarse ctx");
  require_committed(coarse);
  require_committed(c);
  if (!coarse->structured || !c->structured || coarse->blocks.size() != c->blocks.size())
    throw PmsError(PMS_ERR_ARG, "refiner: grids do not match");
  if (coarse->device != c->device) throw PmsError(PMS_ERR_ARG, "refiner: both grids must live on the same GPU");
  DField& fi = coarse->field(src_field);
  DField& fo = c->field(dst_field);
  if (fi.ncomp != 3 || fo.ncomp != 3) throw PmsError(PMS_ERR_ARG, "refiner prolongs 3-component nodal fields");
  for (size_t ib = 0; ib < c->blocks.size(); ib++)
    for (int d = 0; d < 3; d++)
      if (c->blocks[ib].n[d] != 2 * coarse->blocks[ib].n[d] - 1)
        throw PmsError(PMS_ERR_ARG, "refiner: output block is not the 2x refinement of the input block");
  PMS_CUDA(cudaSetDevice(c->device));
  if (coarse->stream != c->stream) {  // order after everything queued on the coarse grid's stream
    cudaEvent_t ev;
    PMS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    PMS_CUDA(cudaEventRecord(ev, coarse->stream));
    PMS_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
    PMS_CUDA(cudaEventDestroy(ev));
  }
  for (size_t ib = 0; ib < c->blocks.size(); ib++) {
    const HBlock& bi = coarse->blocks[ib];
    const HBlock& bo = c->blocks[ib];
    Timed t(c, FAM_REFINE);
    c->L->refine_sgrid(c->stream, (int)bi.n[0], (int)bi.n[1], (int)bi.n[2], fi.d, bi.off, coarse->Npad, fo.d, bo.off, c->Npad);
    check_launch(c, "refine_sgrid");
  }
  fo.version++;
  if (coarse->stream != c->stream) {  // the coarse grid may be freed or overwritten right after this call returns
    cudaEvent_t ev;
    PMS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    PMS_CUDA(cudaEventRecord(ev, c->stream));
    PMS_CUDA(cudaStreamWaitEvent(coarse->stream, ev, 0));
    PMS_CUDA(cudaEventDestroy(ev));
  }
  PMS_CATCH
}

// ---- surface smoothing hooks ----------------------------------------------------------------------
int pms_add_analytic_surface(pms_ctx* c, int kind, const double* params, int* surface_id) {
  PMS_TRY
  if (kind < PMS_SURF_PLANE || kind > PMS_SURF_CYLINDER || !params) throw PmsError(PMS_ERR_ARG, "unknown analytic surface kind");
  pmsk::SurfaceDev s{};
  s.kind = kind;
  for (int i = 0; i < 3; i++) s.p[i] = params[i];
  if (kind == PMS_SURF_SPHERE) {
    s.R = params[3];
  } else {
    const double l = std::sqrt(params[3] * params[3] + params[4] * params[4] + params[5] * params[5]);
    if (!(l > 0.0)) throw PmsError(PMS_ERR_ARG, "analytic surface: zero direction");
    for (int i = 0; i < 3; i++) s.u[i] = params[3 + i] / l;
    s.R = kind == PMS_SURF_CYLINDER ? params[6] : 0.0;
  }
  c->h_surfaces.push_back(s);
  if (surface_id) *surface_id = (int)c->h_surfaces.size() - 1;
  if (c->committed && !c->h_node_dof.empty()) apply_classification(c);
  PMS_CATCH
}

int pms_set_node_classification(pms_ctx* c, const int8_t* dof, const int32_t* surface_id) {
  PMS_TRY
  require_committed(c);
  // MeshSmootherImpl<StructuredGrid>::project_delta_to_tangent_plane is VERIFY_MSG("not impl") (MeshSmoother.cpp:462-467)
  if (c->structured) throw PmsError(PMS_ERR_ARG, "surface smoothing of structured grids: not impl (as in the reference)");
  if (!dof || !surface_id) throw PmsError(PMS_ERR_ARG, "null classification");
  for (long long n = 0; n < c->N; n++)
    if (dof[n] == 2 && (surface_id[n] < 0 || surface_id[n] >= (int)c->h_surfaces.size()))
      throw PmsError(PMS_ERR_ARG, "surface node without a surface evaluator");
  c->h_node_dof.assign(dof, dof + c->N);
  c->h_node_surf.assign(surface_id, surface_id + c->N);
  apply_classification(c);
  PMS_CATCH
}

int pms_snap_nodes(pms_ctx* c, double* dmax) {
  PMS_TRY
  require_committed(c);
  const double d = op_snap_nodes(c);
  if (dmax) *dmax = d;
  PMS_CATCH
}

// MeshSmootherImpl::snap_to (MeshSmoother.cpp:517-563): put the node at xyz, snap it to its surface, return the snapped
// position in xyz; reset != 0 restores the node's coordinates afterwards.
int pms_snap_point(pms_ctx* c, size_t node, double* xyz, int reset) {
  PMS_TRY
  require_committed(c);
  if (!xyz || (long long)node >= c->N) throw PmsError(PMS_ERR_ARG, "bad node / null coordinate");
  if (c->h_node_dof.empty() || c->h_node_dof[node] != 2 || c->h_node_surf[node] < 0) return PMS_OK;  // not on a surface: unchanged
  DField& X = c->field("coordinates");
  double save[3];
  for (int r = 0; r < 3; r++) {
    PMS_CUDA(cudaMemcpyAsync(&save[r], X.d + r * c->Npad + node, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PMS_CUDA(cudaMemcpyAsync(X.d + r * c->Npad + node, &xyz[r], sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  const int32_t hn = (int32_t)node, hs = c->h_node_surf[node];
  int32_t* dn = dalloc<int32_t>(2);
  PMS_CUDA(cudaMemcpyAsync(dn, &hn, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PMS_CUDA(cudaMemcpyAsync(dn + 1, &hs, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  {
    Timed t(c, FAM_UPDATE);
    c->L->surf_snap(c->stream, 1, dn, dn + 1, c->d_surfaces, X.d, c->Npad, c->R, 0);
    check_launch(c, "surf_snap");
  }
  for (int r = 0; r < 3; r++) PMS_CUDA(cudaMemcpyAsync(&xyz[r], X.d + r * c->Npad + node, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  if (reset)
    for (int r = 0; r < 3; r++) PMS_CUDA(cudaMemcpyAsync(X.d + r * c->Npad + node, &save[r], sizeof(double), cudaMemcpyHostToDevice, c->stream));
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(dn);
  X.version++;
  PMS_CATCH
}

int pms_get_surface_normals(pms_ctx* c, double* host_out) {
  PMS_TRY
  require_committed(c);
  if (!host_out) throw PmsError(PMS_ERR_ARG, "null output");
  double* d = dalloc<double>((size_t)3 * c->Npad);
  PMS_CUDA(cudaMemsetAsync(d, 0, (size_t)3 * c->Npad * sizeof(double), c->stream));
  if (c->n_surf_nodes > 0) {
    Timed t(c, FAM_UPDATE);
    c->L->surf_normals(c->stream, c->n_surf_nodes, c->d_surf_nodes, c->d_surf_sid, c->d_surfaces, c->field("coordinates").d, d, c->Npad);
    check_launch(c, "surf_normals");
  }
  for (int k = 0; k < 3; k++)
    PMS_CUDA(cudaMemcpyAsync(host_out + (size_t)k * c->N, d + (size_t)k * c->Npad, c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  PMS_CATCH
}

int pms_nodal_field_axpby(pms_ctx* c, double alpha, const char* x, double beta, const char* y) {
  PMS_TRY
  require_committed(c);
  DField& fx = c->field(x);
  DField& fy = c->field(y);
  if (fx.ncomp != fy.ncomp) throw PmsError(PMS_ERR_ARG, "axpby: component mismatch");
  {
    Timed t(c, FAM_BLAS1);
    c->L->axpby(c->stream, alpha, fx.d, beta, fy.d, c->N, c->Npad, fy.ncomp);
    check_launch(c, "axpby");
  }
  fy.version++;
  PMS_CATCH
}

int pms_nodal_field_axpbypgz(pms_ctx* c, double alpha, const char* x, double beta, const char* y, double gamma,
                             const char* z) {
  PMS_TRY
  require_committed(c);
  DField& fx = c->field(x);
  DField& fy = c->field(y);
  DField& fz = c->field(z);
  if (fx.ncomp != fy.ncomp || fx.ncomp != fz.ncomp) throw PmsError(PMS_ERR_ARG, "axpbypgz: component mismatch");
  {
    Timed t(c, FAM_BLAS1);
    c->L->axpbypgz(c->stream, alpha, fx.d, beta, fy.d, gamma, fz.d, c->N, c->Npad, fz.ncomp);
    check_launch(c, "axpbypgz");
  }
  fz.version++;
  PMS_CATCH
}

int pms_nodal_field_dot(pms_ctx* c, const char* x, const char* y, double* result) {
  PMS_TRY
  require_committed(c);
  *result = op_dot(c, x, y);
  PMS_CATCH
}

int pms_nodal_field_set_value(pms_ctx* c, const char* name, double value) {
  PMS_TRY
  require_committed(c);
  DField& f = c->field(name);
  {
    Timed t(c, FAM_BLAS1);
    c->L->set_value(c->stream, f.d, value, c->N, c->Npad, f.ncomp);
    check_launch(c, "set_value");
  }
  f.version++;
  PMS_CATCH
}

int pms_get_number_nodes(pms_ctx* c, size_t* n) {
  PMS_TRY
  require_committed(c);
  double v = (double)c->num_nodes_owned;
  allreduce(c, &v, 1, NCCL_SUM);
  *n = (size_t)v;
  PMS_CATCH
}

int pms_get_edge_lengths(pms_ctx* c) {
  PMS_TRY
  require_committed(c);
  op_edge_lengths(c);
  PMS_CATCH
}

int pms_total_metric(pms_ctx* c, int stage, double alpha, double* mtot, size_t* n_invalid) {
  PMS_TRY
  require_committed(c);
  if (stage != 0 && stage != 1) throw PmsError(PMS_ERR_ARG, "stage must be 0 (untangle) or 1 (ShapeB1)");
  uint64_t n;
  op_total_metric(c, stage, alpha, mtot, &n);
  if (n_invalid) *n_invalid = (size_t)n;
  PMS_CATCH
}

int pms_total_metric_multi(pms_ctx* c, int stage, int nalpha, const double* alphas, double* mtot, size_t* n_invalid) {
  PMS_TRY
  require_committed(c);
  if (nalpha < 1 || !alphas || !mtot) throw PmsError(PMS_ERR_ARG, "bad multi-alpha request");
  constexpr int NBMAX = pmsk::MULTI_ALPHA_NB;
  for (int k0 = 0; k0 < nalpha;) {
    double al[NBMAX], m[NBMAX];
    uint64_t n[NBMAX];
    const int left = nalpha - k0, nb = std::min(NBMAX, std::max(8, (left + 1) & ~1)), cnt = std::min(nb, left);
    for (int k = 0; k < nb; k++) al[k] = alphas[k0 + std::min(k, cnt - 1)];
    if (!op_total_metric_multi(c, stage, nb, al, m, n))  // tet4 / strict build / multi-block: one pass per alpha
      for (int k = 0; k < cnt; k++) op_total_metric(c, stage, al[k], &m[k], &n[k]);
    for (int k = 0; k < cnt; k++) {
      mtot[k0 + k] = m[k];
      if (n_invalid) n_invalid[k0 + k] = (size_t)n[k];
    }
    k0 += cnt;
  }
  PMS_CATCH
}

int pms_get_gradient(pms_ctx* c, int stage) {
  PMS_TRY
  require_committed(c);
  if (stage != 0 && stage != 1) throw PmsError(PMS_ERR_ARG, "stage must be 0 (untangle) or 1 (ShapeB1)");
  op_gradient(c, stage, nullptr, nullptr);
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  PMS_CATCH
}

int pms_metric_and_gradient(pms_ctx* c, int stage, double* mtot, size_t* n_invalid) {
  PMS_TRY
  require_committed(c);
  if (stage != 0 && stage != 1) throw PmsError(PMS_ERR_ARG, "stage must be 0 (untangle) or 1 (ShapeB1)");
  double m;
  uint64_t n;
  op_gradient(c, stage, &m, &n);
  if (c->structured && !c->use_ref_mesh) {
    // the structured gradient always uses the reference mesh; the metric honours use_ref_mesh
    op_total_metric(c, stage, 0.0, &m, &n);
  }
  if (mtot) *mtot = m;
  if (n_invalid) *n_invalid = (size_t)n;
  PMS_CATCH
}

int pms_count_invalid_elements(pms_ctx* c, size_t* n_invalid) {
  PMS_TRY
  require_committed(c);
  *n_invalid = (size_t)op_count_invalid(c);
  PMS_CATCH
}

int pms_get_alpha_0(pms_ctx* c, double* a) {
  PMS_TRY
  require_committed(c);
  *a = op_alpha0(c);
  PMS_CATCH
}

int pms_update_node_positions(pms_ctx* c, double alpha, double* dmax, double* dmax_rel) {
  PMS_TRY
  require_committed(c);
  double a, b;
  op_update(c, alpha, &a, &b);
  if (dmax) *dmax = a;
  if (dmax_rel) *dmax_rel = b;
  PMS_CATCH
}

int pms_get_element_metrics(pms_ctx* c, double* metric, int32_t* flag) {
  PMS_TRY
  require_committed(c);
  if (!c->d_elem_metric) throw PmsError(PMS_ERR_ARG, "option keep_element_metrics is off");
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  if (metric) PMS_CUDA(cudaMemcpy(metric, c->d_elem_metric, c->E * sizeof(double), cudaMemcpyDeviceToHost));
  if (flag) PMS_CUDA(cudaMemcpy(flag, c->d_elem_flag, c->E * sizeof(int32_t), cudaMemcpyDeviceToHost));
  PMS_CATCH
}

int pms_state_init(pms_ctx* c, pms_state* st) {
  PMS_TRY
  require_committed(c);
  memset(st, 0, sizeof(*st));
  st->d0 = 1;
  st->global_metric = DBL_MAX;
  size_t n;
  pms_get_number_nodes(c, &n);
  st->num_nodes = n;
  PMS_CATCH
}

int pms_check_convergence(const pms_state* st, double gradNorm) {  // Def.hpp:989-1008
  const double grad_check = gradNorm;
  bool retval = false;
  if (st->stage == 0 && (st->dnew == 0.0 || st->total_metric == 0.0))
    retval = true;
  else if (st->stage == 0 && st->num_invalid == 0 && (st->dmax_relative < grad_check))
    retval = true;
  else if (st->num_invalid == 0 &&
           (st->iter > 10 && (st->dmax_relative < grad_check &&
                              (st->dnew < grad_check * grad_check * st->d0 || st->grad_norm_scaled < grad_check))))
    retval = true;
  return retval ? 1 : 0;
}

int pms_run_one_iteration(pms_ctx* c, pms_state* st, double grad_norm, double* metric_out) {
  PMS_TRY
  require_committed(c);
  Driver d(c, st, grad_norm);
  const double m = d.run_one_iteration();
  st->global_metric = m;
  if (metric_out) *metric_out = m;
  PMS_CATCH
}

// line_search (Def.hpp:466-591) on its own: needs cg_s (search direction), cg_g (gradient at the current coordinates),
// cg_r (the restart direction) and cg_edge_length; leaves the coordinates where they are.
int pms_line_search(pms_ctx* c, pms_state* st, double mfac_mult, int* restarted, double* alpha) {
  PMS_TRY
  require_committed(c);
  if (!st || !alpha) throw PmsError(PMS_ERR_ARG, "null argument");
  Driver d(c, st, 0.0);
  d.allow_move = false;
  bool r = false;
  *alpha = d.line_search(r, mfac_mult);
  if (restarted) *restarted = r ? 1 : 0;
  PMS_CATCH
}

int pms_run_algorithm(pms_ctx* c, int inner_iterations, double grad_norm, const char* log_path, pms_state* st) {
  PMS_TRY
  require_committed(c);
  pms_state local;
  Driver d(c, st ? st : &local, grad_norm);
  d.run_algorithm(inner_iterations, log_path);
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  PMS_CATCH
}

// ---- multi-GPU ----
int pms_comm_unique_id(void* id128) {
  try {
    nccl_load();
    if (g_nccl.GetUniqueId(id128) != 0) return PMS_ERR_NCCL;
  } catch (...) {
    return PMS_ERR_NCCL;
  }
  return PMS_OK;
}

int pms_comm_init(pms_ctx* c, const void* id128, int rank, int nranks) {
  PMS_TRY
  nccl_load();
  PMS_CUDA(cudaSetDevice(c->device));
  Uid uid;
  memcpy(uid.internal, id128, 128);
  PMS_NCCL(g_nccl.CommInitRank(&c->nccl_comm, nranks, uid, rank));
  c->rank = rank;
  c->nranks = nranks;
  PMS_CATCH
}

int pms_set_halo(pms_ctx* c, int nn, const int* ranks, const int64_t* send_off, const int32_t* send_nodes,
                 const int64_t* recv_off, const int32_t* recv_nodes) {
  PMS_TRY
  require_committed(c);
  c->nb_ranks.assign(ranks, ranks + nn);
  c->send_off.assign(send_off, send_off + nn + 1);
  c->recv_off.assign(recv_off, recv_off + nn + 1);
  const int64_t ns = c->send_off[nn], nr = c->recv_off[nn];
  for (int64_t i = 0; i < ns; i++)
    if (send_nodes[i] < 0 || send_nodes[i] >= c->N) throw PmsError(PMS_ERR_ARG, "halo send node out of range");
  for (int64_t i = 0; i < nr; i++)
    if (recv_nodes[i] < 0 || recv_nodes[i] >= c->N) throw PmsError(PMS_ERR_ARG, "halo recv node out of range");
  std::vector<int32_t> s(send_nodes, send_nodes + ns), r(recv_nodes, recv_nodes + nr);
  c->d_send_nodes = dupload(s);
  c->d_recv_nodes = dupload(r);
  c->d_send_buf = dalloc<double>((size_t)3 * ns);
  c->d_recv_buf = dalloc<double>((size_t)3 * nr);
  PMS_CATCH
}

int pms_set_interface(pms_ctx* c, int nn, const int* ranks, const int64_t* off, const int32_t* nodes, const uint8_t* owned) {
  PMS_TRY
  require_committed(c);
  if (!c->structured) throw PmsError(PMS_ERR_ARG, "pms_set_interface is for structured blocks; unstructured meshes use pms_set_halo");
  c->if_ranks.assign(ranks, ranks + nn);
  c->if_off.assign(off, off + nn + 1);
  const int64_t tot = c->if_off[nn];
  for (int64_t i = 0; i < tot; i++)
    if (nodes[i] < 0 || nodes[i] >= c->N) throw PmsError(PMS_ERR_ARG, "interface node out of range");
  std::vector<int32_t> nd(nodes, nodes + tot);
  c->d_if_nodes = dupload(nd);
  c->d_if_send = dalloc<double>((size_t)3 * tot);
  c->d_if_recv = dalloc<double>((size_t)3 * tot);
  // per shared node: contributions in ascending rank order; neighbours are given in ascending rank order
  std::map<int32_t, std::vector<std::pair<int, std::pair<int64_t, int32_t>>>> by_node;  // node -> (rank, (pos, cnt))
  for (int q = 0; q < nn; q++) {
    const int64_t cnt = c->if_off[q + 1] - c->if_off[q];
    for (int64_t i = 0; i < cnt; i++)
      by_node[nodes[c->if_off[q] + i]].push_back({ranks[q], {3 * c->if_off[q] + i, (int32_t)cnt}});
  }
  std::vector<int32_t> sh_nodes, sh_cnt;
  std::vector<int64_t> sh_ptr{0}, sh_pos;
  for (auto& kv : by_node) {
    auto v = kv.second;
    v.push_back({c->rank, {-1, 0}});
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    sh_nodes.push_back(kv.first);
    for (auto& e : v) {
      sh_pos.push_back(e.second.first);
      sh_cnt.push_back(e.second.second);
    }
    sh_ptr.push_back((int64_t)sh_pos.size());
  }
  c->n_shared = (long long)sh_nodes.size();
  c->d_sh_nodes = dupload(sh_nodes);
  c->d_sh_ptr = dupload(sh_ptr);
  c->d_sh_pos = dupload(sh_pos);
  c->d_sh_cnt = dupload(sh_cnt);
  if (owned) {  // dot / node-count ownership: the lowest rank holding a copy counts the node
    c->h_dotmask.assign(owned, owned + c->N);
    c->dotmask_trivial = true;
    c->num_nodes_owned = 0;
    for (uint8_t v : c->h_dotmask) {
      c->num_nodes_owned += v ? 1 : 0;
      if (!v) c->dotmask_trivial = false;
    }
    if (!c->dotmask_trivial) {
      if (!c->d_dotmask) c->d_dotmask = dalloc<uint8_t>(c->Npad);
      PMS_CUDA(cudaMemset(c->d_dotmask, 0, c->Npad));
      PMS_CUDA(cudaMemcpy(c->d_dotmask, c->h_dotmask.data(), c->N, cudaMemcpyHostToDevice));
    }
  }
  PMS_CATCH
}

int pms_halo_exchange(pms_ctx* c, const char* name) {
  PMS_TRY
  require_committed(c);
  halo_exchange(c, name);
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  PMS_CATCH
}

// ---- instrumentation ----
int pms_launch_count(const pms_ctx* c, uint64_t* n) {
  if (!c || !n) return PMS_ERR_ARG;
  *n = c->launches;
  return PMS_OK;
}
int pms_stream(const pms_ctx* c, void** s) {
  if (!c || !s) return PMS_ERR_ARG;
  *s = (void*)c->stream;
  return PMS_OK;
}
int pms_synchronize(pms_ctx* c) {
  PMS_TRY
  PMS_CUDA(cudaStreamSynchronize(c->stream));
  PMS_CATCH
}
static int family_of(const std::string& f) {
  const char* names[FAM_COUNT] = {"metric", "gradient", "blas1", "update", "alpha0", "edge", "halo", "refine"};
  for (int i = 0; i < FAM_COUNT; i++)
    if (f == names[i]) return i;
  return -1;
}
int pms_kernel_time_ms(pms_ctx* c, const char* family, double* ms, uint64_t* launches) {
  PMS_TRY
  const int f = family_of(family);
  if (f < 0) throw PmsError(PMS_ERR_ARG, "unknown kernel family");
  drain_timing(c);
  if (ms) *ms = c->fam_ms[f];
  if (launches) *launches = c->fam_launches[f];
  PMS_CATCH
}
int pms_kernel_time_reset(pms_ctx* c) {
  PMS_TRY
  drain_timing(c);
  for (int i = 0; i < FAM_COUNT; i++) {
    c->fam_ms[i] = 0;
    c->fam_launches[i] = 0;
  }
  PMS_CATCH
}
int pms_set_timing(pms_ctx* c, int enabled) {
  PMS_TRY
  drain_timing(c);
  c->timing = enabled != 0;
  PMS_CATCH
}

}  // extern "C"
