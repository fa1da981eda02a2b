// pms_elem_ws.cuh — unstructured hex8 fused metric + nodal gradient in ONE kernel
// (included by pms_kernels.cu inside namespace PMS_NS, after the pipelined element kernel).
//
// The two-pass scheme (k_elem_pipe writes 192 B/hex of per-element contributions, k_grad_gather_csr sums them per node
// in ascending element order) leaves the FP64 pipe idle during the gather (28 % of the evaluation) and HBM idle during
// the element pass.  Here the gather rides inside the element loop, on the SAME warps, and costs them no registers:
//
//   * the element loop is that of k_elem_pipe (persistent CTAs of 128 threads, two per SM, cp.async prefetch of the next
//     element's vertices into a thread-private shared-memory column, vertices read in place, 24 contributions per hex
//     written to the scratch eg);
//   * every iteration each thread also completes ONE node: the node's up to 8 contributions are fetched from eg with
//     cp.async (LDGSTS: global -> shared memory, no register is live while the data is in flight) into a second private
//     column, 4 entries at the top of the iteration and 4 in the middle of the hex evaluation (between corners 3 and 4),
//     because only 12 doubles per thread of shared memory are left beside the vertex buffers; each half has half an
//     iteration (~1.8 us) to land.  The sum runs over the entries in ascending CSR order, absent entries are zero-filled
//     (+0.0 changes nothing), so the bits are those of k_grad_gather_csr;
//   * nodes are taken from a list sorted by COMPLETION ITERATION (the iteration in which the last adjacent element is
//     evaluated; built once on the host from the node->element CSR), batch j = list positions [j G 128, (j+1) G 128) is
//     gathered during iteration j + D.  The per-node entry table holds scratch offsets, fetched one iteration ahead;
//   * hand-off: one counter per slab of K iterations, prog[s] = CTAs through with slab s (counted per warp in shared
//     memory, the CTA's last warp fences and adds 1), need[j] = leading slabs every CTA must be through with before batch j
//     is read; a CTA caches the number of complete leading slabs in shared memory and refreshes it with one or two
//     single-word loads.  (Per-CTA progress words polled by every CTA — 10 cache lines, millions of polls — saturated
//     their L2 slices: 3.0 ms.)  D is chosen on the host such that need[j] K <= j + D - 2, i.e. a CTA only ever waits for iterations
//     strictly before its own: the CTA furthest behind is never blocked, hence no deadlock — provided all CTAs are
//     resident (cooperative launch).  In the steady state the wait is a shared-memory compare;
//   * contributions are read at least a slab after they were written, by then they left the writer's store queue; the
//     128-byte line of a contribution holds 16 elements of the same 128-element tile, so a line is complete when any of
//     its elements is (no stale L1 sectors within the launch; L1 is invalidated between launches).
// Nodes with more than 8 adjacent elements (irregular hex meshes) are left to a list-driven CSR gather afterwards.
// No floating-point atomics; results are bitwise those of the two-pass path.
// Measured (256^3 hexes, ShapeB1, B200; profiles/r02_elem_gather.md): 2.88 ms against 2.36 ms for the two passes, so the
// option (elem_ws) stays OFF by default.  Without the 24 gather copies per thread the same kernel runs in 2.24 ms, i.e.
// the copies cost as much as the stand-alone gather (cp.async of 8 bytes per lane is an LSU/L1 instruction like any
// other: the element loop already issues 48 of them per hex), and the split of the hex evaluation, the node sums and the
// stores add 0.55 ms to the 1.69 ms element pass.  Neither a longer lag (8 -> 100 iterations) nor the back-off changes it.
// (Two earlier forms: helper warps gathering through registers from an L2 ring, 4.0-4.9 ms, bound by the helper threads'
//  load latency; per-CTA progress words polled by every CTA, 3.0 ms, L2 slices saturated by the polls.)
#pragma once

constexpr int EG_K = 4;       // iterations per published slab
constexpr int EG_HALF = 4;    // entries per half gather
constexpr size_t EG_SMEM = (size_t)(2 * 48 + 3 * EG_HALF) * NTP * sizeof(double) + (size_t)NTP * (2 * sizeof(int4) + sizeof(int)) + 16 * sizeof(int);

// polling load: relaxed (an acquire load invalidates the whole L1 — SASS CCTL.IVALL — on every poll)
__device__ __forceinline__ int ld_relaxed_i32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_i32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// 8-byte copy, zero-filled when !valid (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async8_zfill(double* smem_dst, const double* gsrc, const bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gsrc), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}

template <int STAGE>
__global__ void __launch_bounds__(NTP, 2) k_elem_gather(const MeshDev m, const WsPlanDev P, const double* __restrict__ x,
                                                        const double* __restrict__ w, const long long stride,
                                                        const uint8_t* __restrict__ fixed,
                                                        const uint8_t* __restrict__ node_owned, double* __restrict__ g,
                                                        double* __restrict__ rout, double* __restrict__ eg, const long long egs,
                                                        const Reduce R, const int slot) {
  constexpr int NN = 8, NVAL = NN * 3 * 2;
  extern __shared__ double sm[];          // [2][NVAL][NTP] vertex columns
  double* const gbuf = sm + 2 * NVAL * NTP;                       // [3 * EG_HALF][NTP] contributions of a half gather
  int4* const eA = reinterpret_cast<int4*>(gbuf + 3 * EG_HALF * NTP);  // [NTP] scratch offsets of entries 0..3 (-1: none)
  int4* const eB = eA + NTP;                                      // [NTP] entries 4..7
  int* const nd = reinterpret_cast<int*>(eB + NTP);               // [NTP] node id (-1: padding)
  int* const ctrl = nd + NTP;  // [0..7] warps through slab (s & 7); [8] min of prog[] as last seen; [9] its refresh lock; [10] own published
  const int tid = threadIdx.x, lane = tid & 31;
  const int G = gridDim.x;
  if (tid < 16) ctrl[tid] = 0;
  __syncthreads();
  int buf = 0;
  auto slot_in = [&](int b, int arr, int a, int r) -> double* { return sm + (size_t)b * NVAL * NTP + ((arr * NN + a) * 3 + r) * NTP + tid; };

  double val[1] = {0.0};
  unsigned long long inv = 0;
  const long long ntiles = (m.nelems + NTP - 1) / NTP;
  const long long niter_e = (ntiles + G - 1) / G;
  // element pipeline state as in k_elem_pipe
  long long e_cur = 0, e_nxt = 0, e_nn = 0;
  bool act_cur = false, act_nxt = false, act_nn = false;
  int conn_nn[8];
  uint8_t own_nn = 1, own_nxt = 1, own_cur = 1;
  auto load_conn = [&](long long it) {
    const long long tile = it * G + blockIdx.x;
    e_nn = tile * NTP + tid;
    act_nn = it < niter_e && e_nn < m.nelems;
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
  auto issue_half = [&](const int4 e4) {  // 12 contributions of 4 entries -> gbuf
    const int ee[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
    for (int i = 0; i < EG_HALF; i++) {
      const bool have = ee[i] >= 0;
      const double* src = eg + (have ? ee[i] : 0);
#pragma unroll
      for (int r = 0; r < 3; r++) cp_async8_zfill(gbuf + (i * 3 + r) * NTP + tid, src + r * egs, have);
    }
  };
  const long long qbase = (long long)blockIdx.x * NTP + tid, bstride = (long long)G * NTP;

  load_conn(0);
  issue();
  load_conn(1);
  int need_nxt = (1 - P.D >= 0 && 1 - P.D < P.nbatches) ? __ldg(&P.need[1 - P.D]) : 0;  // need of batch j1 of iteration 0
  unsigned fx_cur = 0, ow_cur = 1;
  for (long long it = 0; it < P.niter; it++) {
    const long long j = it - P.D, j1 = j + 1;  // batch j: second half issued now, completed in the middle; j1: first half issued in the middle
    const bool jv = j >= 0 && j < P.nbatches, j1v = j1 >= 0 && j1 < P.nbatches;
    const int slab = (int)(it / EG_K);
    if (it % EG_K == 0 && slab >= 6 && it < niter_e) {  // the shared slab counters are re-used every 8 slabs: stay within 6 of this CTA's slowest warp
      while (*(volatile int*)&ctrl[10] < slab - 5) __nanosleep(200);
    }
    {
      // every CTA has published the slabs batch j1's contributions lie in.  The minimum over the CTAs' progress words is
      // cached in shared memory; it is refreshed (one warp of the CTA at a time, 10 words per lane) one slab BEFORE it
      // would fall short, so that in the steady state nobody waits: the refreshing warp spends ~1 us every few iterations,
      // the others go on.  need_nxt was loaded an iteration ago.
      const int need = need_nxt;
      need_nxt = (j1 + 1 >= 0 && j1 + 1 < P.nbatches) ? __ldg(&P.need[j1 + 1]) : 0;
      for (;;) {
        // (lane 0's view, so that the warp takes the loop's branches together: the shuffles below name all 32 lanes)
        const int cached = __shfl_sync(0xffffffffu, *(volatile int*)&ctrl[8], 0);
        if (cached >= need + 1 || cached >= P.nslabs) break;
        int got = 0;
        if (lane == 0) got = atomicCAS(&ctrl[9], 0, 1) == 0;
        got = __shfl_sync(0xffffffffu, got, 0);
        if (got) {
          int c = cached;
          if (lane == 0) {  // P.prog[s] = CTAs through with slab s: one word per step, not one per CTA
            while (c < P.nslabs && c < need + 2 && ld_relaxed_i32(&P.prog[c]) >= G) c++;
            if (c > *(volatile int*)&ctrl[8]) *(volatile int*)&ctrl[8] = c;
            __threadfence_block();
            atomicExch(&ctrl[9], 0);
          }
          c = __shfl_sync(0xffffffffu, c, 0);
          if (c >= need) break;
        } else if (cached >= need) {
          break;  // another warp is refreshing; what this one needs is there
        }
        __nanosleep(P.sleep_ns);
      }
    }
    cp_async_wait_all();  // this element's vertices, the first half of batch j, the entries of batch j
    e_cur = e_nxt;
    act_cur = act_nxt;
    own_cur = own_nxt;
    const int cur = buf;
    // ---- node, first half: entries 0..3 have landed
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    int n_cur = -1;
    if (jv) {
      n_cur = nd[tid];
      if (n_cur >= 0) {
        // fixed / owned flags of the node (Def.hpp:1226-1244): their 4-byte words were fetched into the spent eA slot in the
        // middle of the previous iteration
        const int4 fl = eA[tid];
        fx_cur = ((unsigned)fl.x >> ((n_cur & 3) * 8)) & 0xffu;
        ow_cur = node_owned ? (((unsigned)fl.y >> ((n_cur & 3) * 8)) & 0xffu) : 1u;
#pragma unroll
        for (int i = 0; i < EG_HALF; i++) {
          a0 += gbuf[(i * 3 + 0) * NTP + tid];
          a1 += gbuf[(i * 3 + 1) * NTP + tid];
          a2 += gbuf[(i * 3 + 2) * NTP + tid];
        }
        const int4 eb = eB[tid];
        asm volatile("" ::: "memory");  // the column has been read before it is re-armed (same-thread order on the LSU)
        issue_half(eb);
      }
    }
    asm volatile("" ::: "memory");
    if (j1v) {  // entries and node id of batch j1 (used from the middle of this iteration on)
      const long long q = j1 * bstride + qbase;
      cp_async16(&eA[tid], &P.tabA[q]);
      cp_async16(&eB[tid], &P.tabB[q]);
      cp_async4(&nd[tid], &P.lnode[q]);
    }
    cp_async_commit();  // group: second half of batch j + entries of batch j1
    buf ^= 1;
    issue();            // group: vertices of the next element
    load_conn(it + 2);

    // ---- middle of the hex evaluation: second half of batch j, first half of batch j1
    auto mid = [&]() {
      cp_async_wait_1();
      if (n_cur >= 0) {
#pragma unroll
        for (int i = 0; i < EG_HALF; i++) {
          a0 += gbuf[(i * 3 + 0) * NTP + tid];
          a1 += gbuf[(i * 3 + 1) * NTP + tid];
          a2 += gbuf[(i * 3 + 2) * NTP + tid];
        }
        if (fx_cur != 0 || ow_cur == 0) a0 = a1 = a2 = 0.0;
        g[n_cur] = a0;
        g[stride + n_cur] = a1;
        g[2 * stride + n_cur] = a2;
        if (rout) {  // copy_field(cg_r, cg_g), Def.hpp:1508
          rout[n_cur] = a0;
          rout[stride + n_cur] = a1;
          rout[2 * stride + n_cur] = a2;
        }
      }
      asm volatile("" ::: "memory");
      if (j1v) {
        const int n1 = nd[tid];
        if (n1 >= 0) {
          issue_half(eA[tid]);
          asm volatile("" ::: "memory");  // eA[tid] has been read: its first two words now receive the node's flag words
          cp_async4(&eA[tid].x, fixed + (n1 & ~3));
          if (node_owned) cp_async4(&eA[tid].y, node_owned + (n1 & ~3));
        }
      }
      cp_async_commit();  // group: first half of batch j1 (+ flags)
    };
    if (act_cur) {
      bool valid;
      double grad[8][3];
      const SmemVerts<true, NTP> VC(slot_in(cur, 0, 0, 0)), VO(slot_in(cur, 1, 0, 0));
      const double mm = hex_metric_fast_v<STAGE, true>(VC, VO, valid, grad, mid);
      const bool use = valid || STAGE == 0;  // gradient_functors.hpp:391, Def.hpp:1219
      const long long ge = m.elem_off + e_cur;
#pragma unroll
      for (int a = 0; a < NN; a++)
#pragma unroll
        for (int r = 0; r < 3; r++) eg[(long long)(a * 3 + r) * egs + ge] = use ? grad[hex_perm(a)][r] : 0.0;
      if (own_cur != 0) {
        val[0] += mm;
        inv += valid ? 0 : 1;
      }
    } else {
      mid();
    }
    if (it < niter_e && ((it + 1) % EG_K == 0 || it + 1 == niter_e)) {  // this warp is through with slab `slab`
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();  // this warp's scratch stores precede its count; the publishing warp's gpu-scope release is cumulative
        if (atomicAdd(&ctrl[slab & 7], 1) == NTP / 32 - 1) {
          ctrl[slab & 7] = 0;
          __threadfence();  // release at gpu scope, cumulative over the other warps' stores
          atomicAdd(&P.prog[slab], 1);
          *(volatile int*)&ctrl[10] = slab + 1;
        }
      }
    }
  }
  cp_async_wait_all();
  grid_reduce<OpSum, 1, true>(val, inv, R, slot);
}

// nodes with more than 8 adjacent elements: the CSR walk of k_grad_gather_csr over a node list
__global__ void __launch_bounds__(NT) k_grad_gather_list(const MeshDev m, const int32_t* __restrict__ nodes, const long long nn,
                                                         const double* __restrict__ eg, const long long egs,
                                                         const uint8_t* __restrict__ fixed, const uint8_t* __restrict__ owned,
                                                         double* __restrict__ g, double* __restrict__ rout, const long long stride) {
  const long long q = (long long)blockIdx.x * NT + threadIdx.x;
  if (q >= nn) return;
  const long long n = nodes[q];
  double g0 = 0.0, g1 = 0.0, g2 = 0.0;
  if (!(fixed[n] || (owned && !owned[n]))) {
    const long long p0 = m.csr_ptr[n], p1 = m.csr_ptr[n + 1];
    for (long long p = p0; p < p1; p++) {
      const long long ent = m.csr_ent32 ? (long long)__ldg(&m.csr_ent32[p]) : __ldg(&m.csr_ent[p]);
      const long long e = ent >> 3;
      const int a = (int)(ent & 7);
      g0 += __ldcg(&eg[(long long)(a * 3 + 0) * egs + e]);
      g1 += __ldcg(&eg[(long long)(a * 3 + 1) * egs + e]);
      g2 += __ldcg(&eg[(long long)(a * 3 + 2) * egs + e]);
    }
  }
  g[n] = g0;
  g[stride + n] = g1;
  g[2 * stride + n] = g2;
  if (rout) {
    rout[n] = g0;
    rout[stride + n] = g1;
    rout[2 * stride + n] = g2;
  }
}

template <int STAGE>
static int elem_gather_grid1() {
  static int grid = -1;
  if (grid < 0) {
    int nb = 0, coop = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    auto kern = k_elem_gather<STAGE>;
    const bool ok = coop && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EG_SMEM) == cudaSuccess &&
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NTP, EG_SMEM) == cudaSuccess && nb >= 1;
    grid = ok ? sm_count() * (nb > 2 ? 2 : nb) : 0;
    if (!ok) cudaGetLastError();
  }
  return grid;
}
// CTAs of the one-kernel gradient (all co-resident); 0 = not available in this build / on this device
static int elem_gather_grid() {
  if (STRICT) return 0;
  const int g0 = elem_gather_grid1<0>(), g1 = elem_gather_grid1<1>();
  return g0 == g1 ? g0 : 0;
}

// grid = CTAs the plan was built for.  Returns false when unsupported (strict build, not hex8 with the reference mesh, plan missing).
static bool launch_elem_ws(cudaStream_t st, const MeshDev& m, const WsPlanDev& P, int stage, int use_ref, const double* x,
                           const double* w, long long stride, const uint8_t* fixed, const uint8_t* node_owned, double* g,
                           double* rout, double* eg, long long egs, const Reduce& R, int slot) {
  if (STRICT || m.kind != KIND_HEX8 || !(stage == 0 || use_ref) || !P.prog || !eg) return false;
  if (P.grid < 1 || P.grid > R.max_blocks || P.grid != elem_gather_grid()) return false;
  MeshDev mm = m;
  WsPlanDev pp = P;
  long long stride_ = stride, egs_ = egs;
  Reduce rr = R;
  void* args[] = {&mm, &pp, (void*)&x, (void*)&w, &stride_, (void*)&fixed, (void*)&node_owned, &g, &rout, &eg, &egs_, &rr, &slot};
  // cooperative launch: progress words are waited on across CTAs, so every CTA of the grid must be resident
  const void* kern = stage == 0 ? (const void*)k_elem_gather<0> : (const void*)k_elem_gather<1>;
  if (cudaLaunchCooperativeKernel(kern, dim3(P.grid), dim3(NTP), args, EG_SMEM, st) != cudaSuccess) return false;
  if (P.novf > 0)
    k_grad_gather_list<<<(unsigned)((P.novf + NT - 1) / NT), NT, 0, st>>>(m, P.ovf_nodes, P.novf, eg, egs, fixed, node_owned, g, rout, stride);
  return true;
}
