// pms_elem_ws.cuh — unstructured hex8 / tet4 fused metric + nodal gradient in ONE kernel
// (included by pms_kernels.cu inside namespace PMS_NS, after the pipelined element kernel).
//
// The two-pass scheme (k_elem_pipe writes 192 B/hex of per-element contributions, k_grad_gather_csr sums them per node
// in ascending element order) leaves the FP64 pipe idle during the gather (28 % of the evaluation) and HBM idle during
// the element pass, and it moves 5.5x the algorithmic bytes through DRAM.  Here both run at the same time in one
// persistent kernel, and the scratch becomes a small ring that lives in the 126 MB L2:
//
//   * one CTA per SM (cooperative launch, so that all of them are resident): 8 compute warps (setmaxnreg 232) run the
//     software-pipelined element loop of k_elem_pipe — cp.async gathers of the next element's vertices into a
//     thread-private shared-memory column, FP64 evaluation of the current one — over warp tiles of 32 elements taken in
//     ascending element order, and write the 24 (12) contributions of every element into a RING of 4 slabs;
//   * 4 helper warps (setmaxnreg 40) follow one slab behind: for every node whose LAST adjacent element lies in slab g
//     (lists built once on the host from the node->element CSR) they add the node's contributions in ascending element
//     order — the CSR walk of k_grad_gather_csr resolved on the host into ring offsets, hence the same bits — and store it;
//   * hand-off through two arrays of per-CTA progress words in global memory, no atomics on shared addresses: prog[c] =
//     slabs CTA c's compute warps have completed (every warp owns the same number of tiles per slab and takes them in
//     order; the last of the 8 warps to finish a slab — counted in shared memory — fences and publishes), gprog[c] =
//     groups CTA c's helpers have gathered.  Helpers start group g when every prog[] >= g+1; a compute warp enters slab
//     s when every gprog[] >= s-2, i.e. when the readers of the ring slot it is about to overwrite are done.  (A first
//     version published every warp tile with __threadfence + atomicAdd on one counter per slab: 8.2 ms instead of
//     2.4 — the fence per tile and 2368 same-address atomics per slab serialised everything.)
//     Ring data is read with ld.global.cg (the L1 is not coherent and the slots are re-used).
//   * requirement: the elements around any node span at most one slab (so they sit in two consecutive slabs); the slab
//     length is chosen from the mesh's largest span, and a mesh numbered so wildly that the ring would be as large as the
//     full scratch keeps the two-pass path.
// No floating-point atomics; results are bitwise those of the two-pass path.
#pragma once


constexpr int EW_NC = 256, EW_NH = 128;  // compute / helper threads per CTA
constexpr int EW_REG_C = 224, EW_REG_H = 56;  // 256 x 224 + 128 x 56 = 64512 = 384 x 168
constexpr int EW_RING = 4;

// polling load: relaxed (an acquire load invalidates the whole L1 — SASS CCTL.IVALL — on every poll); the waiter issues
// one __threadfence() after the condition holds
__device__ __forceinline__ int ld_relaxed_i32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int FL, int STAGE, bool USE_REF>
__global__ void __launch_bounds__(EW_NC + EW_NH, 1) k_elem_ws(const MeshDev m, const WsPlanDev P, const double* __restrict__ x,
                                                              const double* __restrict__ w, const long long stride,
                                                              const uint8_t* __restrict__ fixed,
                                                              const uint8_t* __restrict__ node_owned, double* __restrict__ g,
                                                              double* __restrict__ rout, const Reduce R, const int slot) {
  constexpr bool FAST = fast_hex<FL, STAGE, USE_REF>();
  constexpr int NN = npe<FL>();
  constexpr int NVAL = NN * 3 * 2;
  constexpr bool INPLACE = FAST;  // vertices read in place from a double-buffered column (see k_elem_pipe)
  extern __shared__ double sm[];  // [INPLACE ? 2 : 1][NVAL][EW_NC]
  const int tid = threadIdx.x;
  int* const prog = P.counters;              // [G] slabs completed by each CTA's compute warps
  int* const gprog = P.counters + gridDim.x;  // [G] groups gathered by each CTA's helpers
  __shared__ int sm_cnt[8];                   // warps of this CTA that finished slab (s & 7)
  __shared__ int sm_gmin, sm_lock;            // min over CTAs of gprog[] as last seen by this CTA; its refresh lock
  if (tid < 8) sm_cnt[tid] = 0;
  if (tid == 0) sm_gmin = 0, sm_lock = 0;
  __syncthreads();
  const long long ring_elems = EW_RING * P.slab_elems;
  const int G = gridDim.x;

  if (tid >= EW_NC) {
    // ------------------------------------------------------------------------------ helper warps: lagged node gather
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(EW_REG_H));
    const int h = tid - EW_NC;
    for (int grp = 0; grp < P.nslabs; grp++) {
      // every CTA's compute warps have completed slab grp (hence all earlier ones)
      // (one warp polls, 5 words per lane: 1184 warps polling the same five cache lines saturated their L2 slice.  No
      //  fence is needed after the wait: the ring is read with ld.global.cg, issued after the flag load by control dependence.)
      if (h < 32) {
        for (;;) {
          bool ok = true;
          for (int c = h; c < G; c += 32) ok = ok && ld_relaxed_i32(&prog[c]) >= grp + 1;
          if (__all_sync(0xffffffffu, ok)) break;
          __nanosleep(400);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(EW_NH) : "memory");
      const long long g0 = P.goff[grp], cnt = P.goff[grp + 1] - g0;
      const long long chunk = (cnt + G - 1) / G;
      const long long q0 = g0 + chunk * blockIdx.x, q1 = min(q0 + chunk, g0 + cnt);
      const int ring_i = (int)ring_elems;  // (the plan guarantees 24 * ring_elems < 2^31)
      // Memory-level parallelism comes from batching and prefetching, not occupancy (128 helper threads per SM).  The
      // plan holds, per node in group order, the ring offsets of its contributions (gtab[entry][node], -1 = none: the CSR
      // walk resolved once on the host), read coalesced and ONE CHUNK AHEAD, so the only latency on the critical path is
      // that of the 12 ring loads of a chunk of 4 entries, all in flight together.  Additions keep ascending entry order.
      const int WC = P.tab_w / 4;  // chunks per node
      long long q = q0 + h;
      int c = 0, nx[4];
      long long n_nx = 0;
      bool skip_nx = false, have = q < q1;
      auto prefetch = [&]() {
#pragma unroll
        for (int i = 0; i < 4; i++) nx[i] = __ldg(&P.gtab[(long long)(c * 4 + i) * P.tab_n + q]);
        if (c == 0) {
          n_nx = P.gnodes[q];
          skip_nx = fixed[n_nx] || (node_owned && !node_owned[n_nx]);  // Def.hpp:1226-1244
        }
      };
      if (have) prefetch();
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      long long n = 0;
      bool skip = false;
      while (have) {
        int cur[4];
#pragma unroll
        for (int i = 0; i < 4; i++) cur[i] = nx[i];
        const int cc = c;
        if (cc == 0) {
          n = n_nx;
          skip = skip_nx;
        }
        if (++c == WC) {
          c = 0;
          q += EW_NH;
        }
        have = q < q1;
        if (have) prefetch();
        double v[4][3];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int r = 0; r < 3; r++) v[i][r] = cur[i] >= 0 ? __ldcg(&P.ring[cur[i] + r * ring_i]) : 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) {  // an absent entry adds +0.0, which changes nothing
          a0 += v[i][0];
          a1 += v[i][1];
          a2 += v[i][2];
        }
        if (cc == WC - 1) {
          if (skip) a0 = a1 = a2 = 0.0;
          g[n] = a0;
          g[stride + n] = a1;
          g[2 * stride + n] = a2;
          if (rout) {  // copy_field(cg_r, cg_g), Def.hpp:1508
            rout[n] = a0;
            rout[stride + n] = a1;
            rout[2 * stride + n] = a2;
          }
          a0 = a1 = a2 = 0.0;
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(EW_NH) : "memory");  // every helper thread has read its ring entries
      if (h == 0) st_release_i32(&gprog[blockIdx.x], grp + 1);  // (no global writes to order: the ring was only read)
    }
    return;
  }

  // ---------------------------------------------------------------------------------- compute warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EW_REG_C));
  const int lane = tid & 31, warp = tid >> 5;
  const long long gw = (long long)blockIdx.x * (EW_NC / 32) + warp;  // this warp among the grid's compute warps
  const long long nwarps = (long long)G * (EW_NC / 32);
  const long long ntiles = (m.nelems + 31) / 32;
  int buf = 0;
  auto slot_in = [&](int b, int arr, int a, int r) -> double* { return sm + (size_t)b * NVAL * EW_NC + ((arr * NN + a) * 3 + r) * EW_NC + tid; };

  double val[1] = {0.0};
  unsigned long long inv = 0;
  // pipeline state as in k_elem_pipe: element being evaluated, element whose vertices are in flight, element whose
  // connectivity (and ownership flag) is in flight
  long long e_cur = 0, e_nxt = 0, e_nn = 0;
  bool act_cur = false, act_nxt = false, act_nn = false;
  int conn_nn[8];
  uint8_t own_nn = 1, own_nxt = 1, own_cur = 1;
  auto load_conn = [&](long long it) {
    const long long tile = it * nwarps + gw;
    e_nn = tile * 32 + lane;
    act_nn = tile < ntiles && e_nn < m.nelems;
    if (act_nn) {
#pragma unroll
      for (int a = 0; a < NN; a++) conn_nn[a] = __ldg(&m.conn[(long long)a * m.conn_stride + e_nn]);
      if (m.elem_owned) own_nn = __ldg(&m.elem_owned[e_nn]);
    }
  };
  auto issue = [&]() {
    act_nxt = act_nn;
    e_nxt = e_nn;
    own_nxt = own_nn;
    if (act_nxt) {
#pragma unroll
      for (int a = 0; a < NN; a++) {
        const long long nid = conn_nn[a];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          cp_async8(slot_in(buf, 0, a, r), &x[r * stride + nid]);
          cp_async8(slot_in(buf, 1, a, r), &w[r * stride + nid]);
        }
      }
    }
    cp_async_commit();
  };
  const long long niter = (ntiles + nwarps - 1) / nwarps;  // the same for every warp: empty trailing iterations still count
  load_conn(0);
  issue();
  load_conn(1);
  for (long long it = 0; it < niter; it++) {
    e_cur = e_nxt;
    act_cur = act_nxt;
    own_cur = own_nxt;
    const int slab = (int)(it / P.tiles_per_warp);
    if (it % P.tiles_per_warp == 0 && slab >= EW_RING - 1) {
      // the ring slot of slab `slab` held slab-4, read by groups slab-4 and slab-3: wait for their readers in every CTA
      // (the CTA keeps the minimum it last saw in shared memory; at most one of its warps polls global memory at a time)
      const int need = slab - (EW_RING - 2);
      while (*(volatile int*)&sm_gmin < need) {
        int got = 0;
        if (lane == 0) got = atomicCAS(&sm_lock, 0, 1) == 0;
        got = __shfl_sync(0xffffffffu, got, 0);
        if (got) {
          int mn = 0x7fffffff;
          for (int c = lane; c < G; c += 32) mn = min(mn, ld_relaxed_i32(&gprog[c]));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          if (lane == 0) {
            if (mn > *(volatile int*)&sm_gmin) *(volatile int*)&sm_gmin = mn;
            __threadfence_block();
            atomicExch(&sm_lock, 0);
          }
          __syncwarp();
        }
        if (*(volatile int*)&sm_gmin < need) __nanosleep(400);
      }
    }
    cp_async_wait_all();
    const int cur = buf;
    double vc[8][3], vo[8][3];
    if (!INPLACE && act_cur) {
#pragma unroll
      for (int a = 0; a < NN; a++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
          vc[vpos<FL, FAST>(a)][r] = *slot_in(cur, 0, a, r);
          vo[vpos<FL, FAST>(a)][r] = *slot_in(cur, 1, a, r);
        }
    }
    asm volatile("" ::: "memory");
    if (INPLACE) buf ^= 1;
    issue();
    load_conn(it + 2);
    if (act_cur) {
      bool valid;
      double mm, grad[8][3];
      if (INPLACE) {
        const SmemVerts<FL == FL_HEX8, EW_NC> VC(slot_in(cur, 0, 0, 0)), VO(slot_in(cur, 1, 0, 0));
        mm = hex_metric_fast_v<STAGE, true>(VC, VO, valid, grad);
      } else {
        mm = element_grad_metric<FL, STAGE, USE_REF>(vc, vo, valid, grad);
      }
      const bool use = valid || STAGE == 0;  // gradient_functors.hpp:391, Def.hpp:1219
      const long long re = (long long)(slab % EW_RING) * P.slab_elems + (e_cur - (long long)slab * P.slab_elems);
#pragma unroll
      for (int a = 0; a < NN; a++)
#pragma unroll
        for (int r = 0; r < 3; r++) P.ring[(long long)(a * 3 + r) * ring_elems + re] = use ? grad[vpos<FL, FAST>(a)][r] : 0.0;
      if (own_cur != 0) {
        val[0] += mm;
        inv += valid ? 0 : 1;
      }
    }
    if ((it + 1) % P.tiles_per_warp == 0 || it + 1 == niter) {  // this warp is through with slab `slab`
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();  // this warp's ring stores precede its count; the publishing warp's gpu-scope release is cumulative
        if (atomicAdd(&sm_cnt[slab & 7], 1) == EW_NC / 32 - 1) {
          sm_cnt[slab & 7] = 0;
          st_release_i32(&prog[blockIdx.x], slab + 1);  // (release at gpu scope, cumulative over the other warps' stores)
        }
      }
    }
  }
  grid_reduce<OpSum, 1, true, 1, EW_NC>(val, inv, R, slot);
}

template <int FL, int STAGE, bool USE_REF>
static bool launch_elem_ws2(cudaStream_t st, const MeshDev& m, const WsPlanDev& P, int grid, const double* x, const double* w,
                            long long stride, const uint8_t* fixed, const uint8_t* node_owned, double* g, double* rout,
                            const Reduce& R, int slot) {
  constexpr bool FAST = fast_hex<FL, STAGE, USE_REF>();
  constexpr int NVAL = npe<FL>() * 3 * 2;
  const size_t smem = (size_t)NVAL * EW_NC * sizeof(double) * (FAST ? 2 : 1);
  auto kern = k_elem_ws<FL, STAGE, USE_REF>;
  static int ok = -1;
  if (ok < 0) {
    int nb = 0;
    ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                 cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, EW_NC + EW_NH, smem) == cudaSuccess && nb >= 1
             ? 1
             : 0;
  }
  if (!ok) return false;
  MeshDev mm = m;
  WsPlanDev pp = P;
  long long stride_ = stride;
  Reduce rr = R;
  void* args[] = {&mm, &pp, (void*)&x, (void*)&w, &stride_, (void*)&fixed, (void*)&node_owned, &g, &rout, &rr, &slot};
  // cooperative launch: the counters are waited on across CTAs, so every CTA of the grid must be resident
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(EW_NC + EW_NH), args, smem, st) == cudaSuccess;
}

// grid = CTAs the plan was built for (one per SM).  Returns false when unsupported (strict build, plan missing).
static bool launch_elem_ws(cudaStream_t st, const MeshDev& m, const WsPlanDev& P, int grid, int stage, int use_ref, const double* x,
                           const double* w, long long stride, const uint8_t* fixed, const uint8_t* node_owned, double* g,
                           double* rout, const Reduce& R, int slot) {
  if (STRICT || !P.ring || grid < 1 || grid > R.max_blocks) return false;
  if (m.kind == KIND_HEX8) {
    if (stage == 0) return launch_elem_ws2<FL_HEX8, 0, true>(st, m, P, grid, x, w, stride, fixed, node_owned, g, rout, R, slot);
    if (use_ref) return launch_elem_ws2<FL_HEX8, 1, true>(st, m, P, grid, x, w, stride, fixed, node_owned, g, rout, R, slot);
    return launch_elem_ws2<FL_HEX8, 1, false>(st, m, P, grid, x, w, stride, fixed, node_owned, g, rout, R, slot);
  }
  if (m.kind == KIND_TET4) {
    if (stage == 0) return launch_elem_ws2<FL_TET4, 0, true>(st, m, P, grid, x, w, stride, fixed, node_owned, g, rout, R, slot);
    if (use_ref) return launch_elem_ws2<FL_TET4, 1, true>(st, m, P, grid, x, w, stride, fixed, node_owned, g, rout, R, slot);
    return launch_elem_ws2<FL_TET4, 1, false>(st, m, P, grid, x, w, stride, fixed, node_owned, g, rout, R, slot);
  }
  return false;
}
