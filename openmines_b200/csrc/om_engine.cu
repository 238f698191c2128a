// Kernels + C ABI of the batched OpenMines engine (see include/openmines_b200.h).
//
// Compile for sm_100a with -fmad=false: all simulation times / tonnages are fp64 and must be
// bit-identical to CPython's IEEE-754 double arithmetic (no FMA contraction).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "om_engine.cuh"

#ifndef OM_MAX_WARPS_PER_CTA
#define OM_MAX_WARPS_PER_CTA 20
#endif
#define OM_SMEM_LIMIT (227 * 1024 - 256)   /* minus the static shared memory (mbarriers) */

// ------------------------------------------------------------------------------------------------
// TMA staging of one instance's state blob (1-D bulk copies): HBM -> shared memory completing on an mbarrier
// transaction count, shared memory -> HBM as a bulk group.  One elected lane issues, the whole warp waits.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t om_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void om_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(om_smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// called by every lane; `phase` is the parity of the barrier phase this load completes
__device__ __forceinline__ void om_stage_in(unsigned char* sm, const unsigned char* g, uint32_t bytes, uint64_t* bar,
                                            uint32_t phase, int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses to this buffer are done
  __syncwarp();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(om_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(om_smem_u32(sm)), "l"(g), "r"(bytes), "r"(om_smem_u32(bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(om_smem_u32(bar)), "r"(phase) : "memory");
  }
}

// called by every lane after the last shared-memory write of the instance; returns when the buffer may be reused
// `complete`: wait until the bytes are written (not only read out of shared memory) -- another warp is about to be told that
// the blob is in HBM (episode segments)
__device__ __forceinline__ void om_stage_out(unsigned char* g, const unsigned char* sm, uint32_t bytes, int lane, bool complete = false) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the async proxy
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(om_smem_u32(sm)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (complete) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}

// The work queue is one atomic counter; the last warp of a launch to leave it puts it (and the exit count) back to zero, so a
// launch needs no memset in front of it (one stream operation less per env step; safe under CUDA-graph replay, where the
// kernel parameters are frozen).  When the last warp retires every other warp has already seen the queue empty.
__device__ __forceinline__ void om_retire_warp(const RunParams& p, int lane) {
  if (lane == 0) {
    const unsigned long long total = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    if (atomicAdd(p.counter + 1, 1ull) == total - 1ull) {
      p.counter[0] = 0ull;
      p.counter[1] = 0ull;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void om_reset_kernel(const DevCfg* cfgs, const int32_t* cfg_id, OmStateLayout ly, unsigned char* state,
                                const uint64_t* seeds, const int32_t* dispatcher, long long n) {
  long long inst = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (inst >= n) return;
  unsigned char* blob = state + (size_t)inst * ly.total_bytes;
  const DevCfg& c = cfgs[cfg_id ? cfg_id[inst] : 0];
  om_init_state(blob, blob, ly.total_bytes, c, ly, threadIdx.x & 31, seeds[inst], dispatcher[inst]);
}

// Work-queue order of om_run: instance ids grouped by dispatcher id (counting sort, one CTA).  The episode kernel is
// instruction-fetch bound and each dispatcher has its own decision code, so SMs that run one dispatcher at a time keep
// a smaller instruction working set: +13 % on the SQ/SPTF workload against the interleaved order.
// Results are indexed by instance id, so the order is invisible outside the kernel.
// Groups are queued most expensive first (measured ms per 16384 stochastic north_pit instances: SPTF 41.9, SQ 29.6,
// FixedGroup 27.1, Random 26.7, Nearest 21.5, Naive 13.3), so the last, partially filled wave runs the cheap ones.
__device__ __forceinline__ int om_cost_rank(int d) {
  switch (d) {
    case OM_DISP_SPTF: return 0;
    case OM_DISP_SHORTESTTRIP: return 1;
    case OM_DISP_SQ: return 2;
    case OM_DISP_FIXEDGROUP: return 3;
    case OM_DISP_RANDOM: return 4;
    case OM_DISP_NEAREST: return 5;
    case OM_DISP_NAIVE: return 6;
    default: return 7;
  }
}
__global__ void om_order_kernel(const int32_t* dispatcher, long long n, int32_t* order) {
  __shared__ unsigned int start[9];
  if (threadIdx.x < 9) start[threadIdx.x] = 0u;
  __syncthreads();
  for (long long i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&start[om_cost_rank(dispatcher[i]) + 1], 1u);
  __syncthreads();
  if (threadIdx.x == 0)
    for (int d = 1; d < 9; ++d) start[d] += start[d - 1];
  __syncthreads();
  for (long long i = threadIdx.x; i < n; i += blockDim.x) order[atomicAdd(&start[om_cost_rank(dispatcher[i])], 1u)] = (int32_t)i;
}

// MULTI: mixed-config batch (om_create_multi) -- every instance names its config in p.cfg_id; the warp copies that DevCfg
// (scalars + table pointers, 320 bytes) from the device table into its shared-memory slot behind the state blob, and the
// simulation reads its config from there instead of the kernel parameters.  The layout is the same for every instance
// (sized for the largest config).
#define OM_CFG_SLOT ((uint32_t)((sizeof(DevCfg) + 15u) & ~15u))
template <int RNG, bool TRACE, bool SPLIT, bool MULTI>
__global__ void __launch_bounds__(OM_MAX_WARPS_PER_CTA * 32, 1) om_run_kernel(const __grid_constant__ RunParams p) {
  extern __shared__ __align__(16) unsigned char om_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t staged = SPLIT ? p.ly.hot_bytes : p.ly.total_bytes;   // SPLIT: truck records / jam pool stay in HBM
  unsigned char* sm = om_smem + (size_t)warp * (staged + (MULTI ? OM_CFG_SLOT : 0u));
  __shared__ __align__(8) uint64_t om_mbar[OM_MAX_WARPS_PER_CTA];
  uint32_t phase = 0;
  if (lane == 0) om_mbar_init(&om_mbar[warp]);
  __syncwarp();
  // Work items: (segment, instance), all first segments before all second ones ...  Cutting the episode into segments keeps
  // the tail of the persistent grid short (16384 instances on 2960 resident warps are 5.5 waves; when the queue runs dry
  // the last warps still have a whole episode in front of them).  A later segment of an instance waits for the flag its
  // predecessor sets after its blob is back in HBM; the predecessor was handed out earlier and waits for nothing younger.
  const int S = SPLIT ? 1 : (p.n_segments > 1 ? p.n_segments : 1);
  while (true) {
    unsigned long long inst = 0;
    int seg = 0;
    if (lane == 0) {
      const unsigned long long q = atomicAdd(p.counter, 1ull);
      inst = ~0ull;
      if (q < (unsigned long long)p.n_instances * (unsigned long long)S) {
        unsigned long long qi = q;
        #pragma unroll 1
        while (qi >= (unsigned long long)p.n_instances) { qi -= (unsigned long long)p.n_instances; ++seg; }   // S <= 16 trips
        inst = (p.order != nullptr) ? (unsigned long long)p.order[qi] : qi;
        if (seg > 0) {
          const uint32_t want = (uint32_t)seg;
          uint32_t have;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(have) : "l"(p.seg_done + inst) : "memory");
            if (have != want) __nanosleep(256);
          } while (have != want);
        }
      }
    }
    inst = __shfl_sync(OM_FULL, inst, 0);
    seg = __shfl_sync(OM_FULL, seg, 0);
    if (inst == ~0ull) break;
    // stage the instance's blob HBM -> shared memory
    unsigned char* g = p.state + (size_t)inst * p.ly.total_bytes;
    om_stage_in(sm, g, staged, &om_mbar[warp], phase, lane);   // one TMA bulk copy
    phase ^= 1u;
    const DevCfg* cfg = &p.c;
    if (MULTI) {
      DevCfg* wc = (DevCfg*)(sm + staged);
      const uint32_t* src = (const uint32_t*)(p.cfgs + p.cfg_id[inst]);
      __syncwarp();
      #pragma unroll 1
      for (uint32_t k = lane; k < sizeof(DevCfg) / 4u; k += 32) ((uint32_t*)wc)[k] = __ldg(src + k);
      __syncwarp();
      cfg = wc;
    }
    Sim<RNG, TRACE, false, SPLIT> s(sm, g, p, *cfg, lane, (long long)inst);
    s.load_hdr(p);
    // mixed-config batches: every instance stops at its own config's sim_time (or the call's `until`, if earlier)
    const double until_i = (MULTI && cfg->sim_time > 0.0 && cfg->sim_time < p.until) ? cfg->sim_time : p.until;
    const bool last_seg = seg == S - 1;
    s.run(last_seg ? until_i : until_i * (1.0 - p.seg_tail * (double)(S - 1 - seg)));
    s.store_hdr(p);
    if (last_seg && p.out.summary != nullptr) {
      double dumped = s.dumped_tons();
      if (lane == 0) {
        OmSummary r;
        r.reported_tons = s.reported_tons; r.dumped_tons = dumped; r.now = s.now;
        r.n_truck_events = s.n_events; r.n_decisions = s.n_decisions; r.n_wrapper_orders = s.n_orders;
        r.n_hops = s.n_hops; r.n_jams = s.n_jams; r.n_road_repairs = s.n_road_rep;
        r.n_truck_repairs = s.n_brk; r.n_unrepairable = s.n_unrep; r.n_ticks = s.n_ticks; r.err = s.err;
        p.out.summary[inst] = r;
      }
    }
    if (TRACE && RNG == OM_RNG_PHILOX && p.out.draw_count != nullptr && lane == 0) p.out.draw_count[inst] = (uint32_t)s.tape_pos;
#ifdef OM_BOUNDS_CHECK
    om_publish_violation(s.dbg_viol);
#endif
    om_stage_out(g, sm, staged, lane, !last_seg);
    if (S > 1 && lane == 0) {
      // the blob is in HBM: release the instance's next segment; the last segment puts the flag back to 0, so a launch
      // leaves the flags as it found them (no epoch, no memset: safe under CUDA-graph replay)
      __threadfence();
      const uint32_t v = last_seg ? 0u : (uint32_t)seg + 1u;
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.seg_done + inst), "r"(v) : "memory");
    }
  }
  om_retire_warp(p, lane);
}

// Batched gym-style step: one warp per env; the state blob round-trips HBM every step (this kernel IS
// HBM-bound: 2 * S_state + obs row per env step).
template <int RNG, bool SPLIT>
__global__ void __launch_bounds__(OM_MAX_WARPS_PER_CTA * 32, 1) om_step_kernel(const __grid_constant__ RunParams p) {
  extern __shared__ __align__(16) unsigned char om_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t staged = SPLIT ? p.ly.hot_bytes : p.ly.total_bytes;
  unsigned char* sm = om_smem + (size_t)warp * staged;
  __shared__ __align__(8) uint64_t om_mbar[OM_MAX_WARPS_PER_CTA];
  uint32_t phase = 0;
  if (lane == 0) om_mbar_init(&om_mbar[warp]);
  __syncwarp();
  while (true) {
    unsigned long long inst = 0;
    if (lane == 0) inst = atomicAdd(p.counter, 1ull);
    inst = __shfl_sync(OM_FULL, inst, 0);
    if (inst >= (unsigned long long)p.n_instances) break;
    unsigned char* g = p.state + (size_t)inst * p.ly.total_bytes;
    om_stage_in(sm, g, staged, &om_mbar[warp], phase, lane);   // one TMA bulk copy
    phase ^= 1u;
    Sim<RNG, false, true, SPLIT> s(sm, g, p, p.c, lane, (long long)inst);
    s.load_hdr(p);
    s.rl_step(p);
    s.store_hdr(p);
#ifdef OM_BOUNDS_CHECK
    om_publish_violation(s.dbg_viol);
#endif
    om_stage_out(g, sm, staged, lane);
  }
  om_retire_warp(p, lane);
}

// ------------------------------------------------------------------------------------------------
// test hooks (om_debug_*): the engine's own draw transforms and jam-position sampler, exposed so that the statistical tests
// can look at the marginal law of a single draw site (tests/test_gpu_rng_marginals.py)
// ------------------------------------------------------------------------------------------------
__global__ void om_debug_draws_kernel(int kind, uint64_t seed, uint32_t arg, long long n, double* out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)kind, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  double v;
  switch (kind) {
    case 0: v = (double)om_u01(r.x); break;                                              // uniform (0, 1]
    case 1: v = (double)om_exp1(r.x); break;                                             // exponential, mean 1
    case 2: v = (double)om_normal_fast(r.x, r.y); break;                                 // standard normal
    case 3: v = (double)sqrtf(om_exp1(r.z)); break;                                      // Weibull(shape 2), scale 1
    case 4: v = -0.1 + 0.2 * (double)om_u01(r.x); break;                                 // load jitter
    case 5: v = (double)om_randint(r.x, arg); break;                                     // randint [0, arg)
    case 6: v = (double)((float)(r.y >> 8) * 5.9604644775390625e-08f); break;            // jam coin [0, 1)
    default: v = 0.0;
  }
  out[i] = v;
}

// one warp: n_samples jam positions on road `rid` of instance `inst` as it stands in `state` (epoch = sample index)
__global__ void om_debug_jam_kernel(RunParams p, long long inst, int rid, int n_samples, double* out, int* fallbacks,
                                    double* coverage_out) {
  const int lane = threadIdx.x & 31;
  unsigned char* blob = p.state + (size_t)inst * p.ly.total_bytes;
  Sim<OM_RNG_PHILOX, false, false, false> s(blob, blob, p, p.c, lane, inst);
  s.load_hdr(p);
  const uint32_t n_on = s.debug_trucks_on_road(rid);
  if (coverage_out != nullptr)                 // Truck.get_route_coverage of every truck under way on the road, -1 for the others
    for (int i = lane; i < p.c.T; i += 32) coverage_out[i] = s.debug_coverage(i, rid);
  int fb = 0;
  for (int k = 0; k < n_samples; ++k) {
    double pos = (n_on >= 1) ? s.template sample_jam_position<true>(k, rid, n_on, &fb) : -1.0;
    if (lane == 0) out[k] = pos;
  }
  if (lane == 0 && fallbacks) *fallbacks = fb;
}

#define OM_STATS_PER_DISP 16
#define OM_STATS_DISPATCHERS 7   /* the seven device dispatchers (OM_DISP_NAIVE .. OM_DISP_SHORTESTTRIP) */
// Per-dispatcher statistics for the cross-GPU all-reduce.  Every entry is an INTEGER (held in a double on the way out):
// counts as they are, tonnages in fixed point -- reported / dumped tons in units of 2^-24 t split into a high and a low
// word at bit 25, the squared tons (of tons rounded to 2^-8 t) split at bit 32.  Integer sums are associative, so the
// totals do not depend on the order blocks, warps or ranks add them in: the same global instance ids give bit-identical
// statistics on 1 or 8 GPUs, and every partial sum stays below 2^53 (exact in the fp64 all-reduce) for up to 2^20
// instances per dispatcher.  openmines_b200/parallel.py holds the host mirror and the decoder.
__device__ __forceinline__ void om_stats_split(double tons, unsigned long long* hi, unsigned long long* lo) {
  double x = tons * 16777216.0;                       // 2^24
  x = (x < 0.0) ? 0.0 : ((x > 1.0e18) ? 1.0e18 : x);  // tonnages are >= 0; NaN -> 0
  const unsigned long long q = (unsigned long long)__double2ll_rn(x);
  *hi = q >> 25; *lo = q & 0x1FFFFFFull;
}
__global__ void __launch_bounds__(256) om_stats_kernel(const OmSummary* summary, const int32_t* dispatcher, long long n,
                                                       unsigned long long* acc, unsigned int* blocks_done, double* stats) {
  __shared__ unsigned long long sacc[OM_STATS_DISPATCHERS * OM_STATS_PER_DISP];
  __shared__ bool last;
  for (int k = threadIdx.x; k < OM_STATS_DISPATCHERS * OM_STATS_PER_DISP; k += blockDim.x) sacc[k] = 0ull;
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int d = -1;
  unsigned long long v[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) v[k] = 0ull;
  if (i < n) {
    d = dispatcher[i];
    if (d < 0 || d >= OM_STATS_DISPATCHERS) d = -1;
    const OmSummary s = summary[i];
    v[0] = 1ull;
    om_stats_split(s.reported_tons, &v[1], &v[2]);
    v[3] = s.n_truck_events; v[4] = s.n_decisions; v[5] = s.n_jams; v[6] = s.n_road_repairs; v[7] = s.n_truck_repairs;
    v[8] = s.n_unrepairable; v[9] = (s.err != 0 && s.err != OM_IERR_TRACE_OVERFLOW) ? 1ull : 0ull; v[10] = s.n_hops;
    om_stats_split(s.dumped_tons, &v[11], &v[12]);
    double t8 = s.reported_tons * 256.0;
    t8 = (t8 < 0.0) ? 0.0 : ((t8 > 4.0e9) ? 4.0e9 : t8);
    const unsigned long long q8 = (unsigned long long)__double2ll_rn(t8), sq = q8 * q8;
    v[13] = sq >> 32; v[14] = sq & 0xFFFFFFFFull;
  }
  // every value is < 2^32 (hops of one instance, 25-bit words, ...): 32 of them add up in 64 bits, word by word
#pragma unroll 1
  for (int dd = 0; dd < OM_STATS_DISPATCHERS; ++dd) {
    const bool mine = d == dd;
    if (!__any_sync(OM_FULL, mine)) continue;
#pragma unroll
    for (int k = 0; k < 15; ++k) {
      const unsigned long long x = mine ? v[k] : 0ull;
      const unsigned int lo = __reduce_add_sync(OM_FULL, (unsigned int)(x & 0xFFFFull));
      const unsigned int mid = __reduce_add_sync(OM_FULL, (unsigned int)((x >> 16) & 0xFFFFull));
      const unsigned int hi = __reduce_add_sync(OM_FULL, (unsigned int)(x >> 32));
      if ((threadIdx.x & 31) == 0)
        atomicAdd(&sacc[dd * OM_STATS_PER_DISP + k], (unsigned long long)lo + ((unsigned long long)mid << 16) + ((unsigned long long)hi << 32));
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < OM_STATS_DISPATCHERS * OM_STATS_PER_DISP; k += blockDim.x)
    if (sacc[k] != 0ull) atomicAdd(&acc[k], sacc[k]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {     // the last block to finish publishes the totals and clears the accumulators for the next call
    __threadfence();
    for (int k = threadIdx.x; k < OM_STATS_DISPATCHERS * OM_STATS_PER_DISP; k += blockDim.x) {
      stats[k] = (double)atomicExch(&acc[k], 0ull);
    }
    if (threadIdx.x == 0) *blocks_done = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct OmEngine {
  int device = 0;
  long long n = 0;
  DevCfg c;                    // the config (one-config handles) or the maxima the layout is sized for (mixed-config handles)
  std::vector<DevCfg> cfgs;    // host copies of the device config table
  const DevCfg* d_cfgs = nullptr;
  const int32_t* d_cfg_id = nullptr;   // [n] config of every instance (mixed-config handles only)
  bool multi = false;
  const void* last_obs = nullptr;      // observation buffer of the previous om_step (static columns already written there)
  uint32_t* seg_done = nullptr;        // [n] episode-segment progress flags (see om_run_kernel)
  uint32_t seg_epoch = 0;              // om_run launches so far
  int segments = 1;
  double seg_tail = 0.1;
  OmStateLayout ly;
  std::vector<void*> dev_allocs;
  unsigned long long* counter = nullptr;
  int32_t* order = nullptr;   // [n] instance ids grouped by dispatcher (written by om_reset)
  unsigned long long* stats_acc = nullptr;   // [7 * 16] integer accumulators of om_stats_reduce (zero between calls)
  unsigned int* stats_done = nullptr;
  int rng_mode = OM_RNG_DET;
  const double* tape_values = nullptr;
  const uint8_t* tape_sites = nullptr;
  const int64_t* tape_offsets = nullptr;
  int grid = 0, block = 0, smem = 0, warps = 0;
  bool split = false;   // cold part of the blob accessed in place in HBM
  bool rl = false;           // created with OM_CFG_RL_SURFACE
  bool reset_done = false;   // om_reset has been enqueued at least once (om_run / om_step on a never-written blob could spin)
  std::string err;
};

static thread_local std::string g_create_err;

static int set_err(OmEngine* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return set_err(h, OM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)

template <typename T>
static int upload(OmEngine* h, const T* src, size_t count, const T** dst) {
  void* d = nullptr;
  size_t bytes = (count ? count : 1) * sizeof(T);
  CK(cudaMalloc(&d, bytes));
  h->dev_allocs.push_back(d);
  if (count) CK(cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *dst = (const T*)d;
  return OM_OK;
}

static uint32_t take(uint32_t& off, size_t elem, size_t count) {
  off = (uint32_t)((off + elem - 1) / elem * elem);
  uint32_t r = off;
  off += (uint32_t)(elem * count);
  return r;
}

static void build_layout(const DevCfg& c, OmStateLayout& ly, bool rl, bool split) {
  memset(&ly, 0, sizeof(ly));
  uint32_t off = 0;
  const size_t T = c.T, U = c.U, R = c.R, J = c.J, NS = c.NS;
  // ---- hot part (always staged in shared memory) ----
  // Without the RL surface the header stops before the reward-window fields (InstHdr up to n_rel = 120 bytes).
  ly.hdr_bytes = rl ? OM_HDR_BYTES : 128;
  ly.hdr = take(off, 8, ly.hdr_bytes / 8);
  ly.t_time = take(off, 16, T);       // 16-byte entries {time, key, phase}, scanned on every pop
  ly.t_key = ly.t_time + 8; ly.t_phase = ly.t_time + 12; ly.e_stride = OM_ESTRIDE;
  ly.t_stride = OM_TSTRIDE; ly.u_stride = OM_USTRIDE;
  ly.u_rec = take(off, 8, (size_t)U * OM_USTRIDE / 8);
  ly.u_last_start = ly.u_rec + 0; ly.u_last_done = ly.u_rec + 8; ly.u_tons = ly.u_rec + 16;
  ly.u_svc = ly.u_rec + 24; ly.u_user = ly.u_rec + 28; ly.u_head = ly.u_rec + 30; ly.u_tail = ly.u_rec + 32;
  ly.u_qlen = ly.u_rec + 34; ly.u_repair = ly.u_rec + 36;
  ly.u_lastbrk = take(off, 8, c.S);
  ly.r_repair_end = take(off, 8, R);
  if (rl) { ly.s_tons = take(off, 4, NS); ly.s_cap = take(off, 4, NS); ly.s_ability = take(off, 4, NS); }   // float32
  // scratch: unit_est_all() results of ONE side (dumpers share the shovels' slots) or per-road counters (u32)
  (void)U;
  size_t nscr = ((size_t)c.S > (size_t)c.M ? (size_t)c.S : (size_t)c.M);
  if (nscr < (R + 1) / 2) nscr = (R + 1) / 2;
  if (nscr < 32) nscr = 32;
  ly.scratch = take(off, 8, nscr);
  // per-slot minima of the truck entries, used by the SPLIT kernels (valid within one kernel invocation only)
  if (split) ly.slot_cache = take(off, 16, (T + 31) / 32);   // (0 in unsplit layouts: their kernels scan the entries directly)
  ly.u_rel_key = take(off, 4, U);
  ly.s_qtotal = take(off, 2, NS); ly.s_coming = take(off, 2, NS);
  if (split) {   // shared-memory helpers that keep the per-minute / per-departure scans of a big mine out of HBM
    ly.t_aux = take(off, 8, T);          // {unit, next in queue, order epoch, service count} u16[4] per truck (instead of the record fields)
    ly.t_loc = take(off, 4, T);          // {status, current, target, flags} per truck (instead of the record fields)
    ly.t_lastbrk_hot = take(off, 2, T);  // minute of the last breakdown per truck, uint16 (instead of the record's f64)
    ly.r_jamcnt = take(off, 2, R);       // stored jams per road
    // trucks under way per road: count + doubly linked list, maintained at departure / arrival (jam sampling)
    ly.r_on = take(off, 2, R); ly.r_head = take(off, 2, R);
    ly.t_rnext = take(off, 2, T); ly.t_rprev = take(off, 2, T);
  }
  ly.hot_bytes = (off + 15u) & ~15u;
  off = ly.hot_bytes;
  // ---- cold part (truck records, jam pool): in shared memory too unless the engine runs split ----
  ly.t_rec = take(off, 8, (size_t)T * OM_TSTRIDE / 8);
  ly.t_load = ly.t_rec + 0; ly.t_jstart = ly.t_rec + 8; ly.t_prejam = ly.t_rec + 16; ly.t_lastbrk = ly.t_rec + 24;
  ly.t_repair = ly.t_rec + 32; ly.t_total = ly.t_rec + 40; ly.t_cycle = ly.t_rec + 48; ly.t_dist = ly.t_rec + 56;
  ly.t_rng = ly.t_rec + 64; ly.t_svc = ly.t_rec + 68; ly.t_epoch = ly.t_rec + 70;
  ly.t_next = ly.t_rec + 72; ly.t_unit = ly.t_rec + 74; ly.t_status = ly.t_rec + 77;
  ly.t_cur = ly.t_rec + 78; ly.t_tgt = ly.t_rec + 79;
  if (split) {   // published offsets follow the data (strides: t_loc_stride / 8)
    ly.t_status = ly.t_loc + 0; ly.t_cur = ly.t_loc + 1; ly.t_tgt = ly.t_loc + 2; ly.t_lastbrk = ly.t_lastbrk_hot;
    ly.t_unit = ly.t_aux + 0; ly.t_next = ly.t_aux + 2; ly.t_epoch = ly.t_aux + 4; ly.t_svc = ly.t_aux + 6;
  }
  if (rl) ly.t_since = take(off, 8, T);
  ly.j_pos = take(off, 8, J); ly.j_end = take(off, 8, J);
  if (rl) ly.j_start = take(off, 2, J);   // uint16: the minute the jam was created at
  ly.j_road = take(off, 2, J);
  ly.total_bytes = (off + 15u) & ~15u;
}

extern "C" {

const char* om_last_error(OmHandle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int om_abi_info(int32_t* out, int32_t n) {
  const int32_t v[] = {OM_ABI_VERSION, (int32_t)sizeof(OmConfig), (int32_t)sizeof(OmOutputs), (int32_t)sizeof(OmStepIO),
                       (int32_t)sizeof(OmStateLayout), (int32_t)sizeof(OmSummary), (int32_t)sizeof(OmTickRecord),
                       (int32_t)sizeof(OmTraceRecord), (int32_t)sizeof(OmUnitTick), (int32_t)sizeof(InstHdr),
                       OM_STATS_DISPATCHERS, OM_STATS_PER_DISP};
  const int32_t m = (int32_t)(sizeof(v) / sizeof(v[0]));
  if (!out || n <= 0) return m;
  for (int32_t k = 0; k < n && k < m; ++k) out[k] = v[k];
  return m;
}

// argument checks of one config (no CUDA call: these run without a GPU)
static int check_config(const OmConfig* cfg) {
  OmEngine* h = nullptr;
  if (!cfg) return set_err(h, OM_E_INVALID, "om_create: null config");
  if (cfg->abi_version != OM_ABI_VERSION) return set_err(h, OM_E_INVALID, "om_create: ABI version mismatch");
  const int T = cfg->n_trucks, L = cfg->n_load_sites, D = cfg->n_dump_sites, S = cfg->n_shovels, M = cfg->n_dumpers;
  if (T <= 0 || L <= 0 || D <= 0 || S <= 0 || M <= 0 || cfg->n_truck_types <= 0)
    return set_err(h, OM_E_INVALID, "om_create: empty mine (mine.py:346-351 asserts)");
  if (T >= 0xFFFF || L + D > 32 || L + D >= 0xFF || cfg->jam_capacity <= 0 || cfg->jam_capacity > 4096)
    return set_err(h, OM_E_UNSUPPORTED, "om_create: need T < 65535, L + D <= 32, 0 < jam_capacity <= 4096");
  if (S + M >= 0xFFFF) return set_err(h, OM_E_UNSUPPORTED, "om_create: need fewer than 65535 shovels + dumpers");
  for (int s = 0; s < L + D; ++s) {
    int nu = cfg->site_unit_begin[s + 1] - cfg->site_unit_begin[s];
    if (nu <= 0) return set_err(h, OM_E_INVALID, "om_create: every site needs at least one shovel / dumper");
  }
  // index tables: out-of-range entries would become out-of-bounds device reads
  if (cfg->site_unit_begin[0] != 0 || cfg->site_unit_begin[L] != S || cfg->site_unit_begin[L + D] != S + M)
    return set_err(h, OM_E_INVALID, "om_create: site_unit_begin must run 0 .. S .. S+M over the load sites then the dump sites");
  for (int t = 0; t < T; ++t)
    if (cfg->truck_type[t] < 0 || cfg->truck_type[t] >= cfg->n_truck_types || cfg->fixed_group[t] < 0 || cfg->fixed_group[t] >= L ||
        !(cfg->truck_speed[t] > 0.0) || !(cfg->truck_capacity[t] > 0.0))
      return set_err(h, OM_E_INVALID, "om_create: truck_type / fixed_group out of range, or non-positive truck speed / capacity");
  for (int i = 0; i < L; ++i)
    if (cfg->nearest_haul[i] < 0 || cfg->nearest_haul[i] >= D) return set_err(h, OM_E_INVALID, "om_create: nearest_haul out of range");
  for (int j = 0; j < D; ++j)
    if (cfg->nearest_back[j] < 0 || cfg->nearest_back[j] >= L) return set_err(h, OM_E_INVALID, "om_create: nearest_back out of range");
  if (cfg->nearest_init < 0 || cfg->nearest_init >= L) return set_err(h, OM_E_INVALID, "om_create: nearest_init out of range");
  if (cfg->sum_mode != 0 && cfg->sum_mode != 1) return set_err(h, OM_E_INVALID, "om_create: sum_mode must be 0 (naive) or 1 (Neumaier)");
  return OM_OK;
}

// Device copy of one packed config: scalars + uploaded tables (owned by the handle).
static int build_devcfg(OmEngine* h, const OmConfig* cfg, DevCfg& c) {
  const int T = cfg->n_trucks, L = cfg->n_load_sites, D = cfg->n_dump_sites, S = cfg->n_shovels, M = cfg->n_dumpers;
  int rc = OM_OK;
  memset(&c, 0, sizeof(c));
  c.T = T; c.L = L; c.D = D; c.S = S; c.M = M; c.U = S + M; c.R = L + 2 * L * D; c.NS = L + D;
  c.NT = cfg->n_truck_types; c.J = cfg->jam_capacity; c.sum_mode = cfg->sum_mode; c.nearest_init = cfg->nearest_init;
  c.road_lambda = cfg->road_lambda; c.road_mu = cfg->road_mu; c.road_sigma = cfg->road_sigma;
  c.detour_mu = cfg->detour_mu; c.detour_sigma = cfg->detour_sigma; c.breakdown_mean = cfg->breakdown_mean;
  c.repair_mu = cfg->repair_mu; c.repair_sigma = cfg->repair_sigma; c.unrepairable_mean = cfg->unrepairable_mean;
  c.maint_mean = cfg->maint_mean; c.maint_mu = cfg->maint_mu; c.maint_sigma = cfg->maint_sigma;
  c.jam_sd = cfg->jam_sd; c.jam_scale = cfg->jam_scale;
  c.sim_time = cfg->sim_time;
#define UP(field, src, count) if (rc == OM_OK) rc = upload(h, cfg->src, (size_t)(count), &c.field)
  UP(cap, truck_capacity, T); UP(speed, truck_speed, T); UP(ttype, truck_type, T);
  UP(ubeg, site_unit_begin, L + D + 1); UP(dcycle, dumper_cycle, M); UP(rdist, road_distance, c.R);
  UP(load_time, load_time, (size_t)c.NT * S); UP(sptf_cost, sptf_cost, (size_t)c.NT * L);
  UP(jam_chance, jam_chance, T + 1); UP(near_haul, nearest_haul, L); UP(near_back, nearest_back, D);
  UP(fixed_group, fixed_group, T); UP(shovel_rate, shovel_rate, S); UP(site_prod, site_productivity, L);
#undef UP
  if (rc == OM_OK) {
    // service time of truck type x unit: (cap/shovel_tons)*cycle for shovels (truck.py:223), the dumper's cycle time otherwise
    std::vector<double> svc((size_t)c.NT * c.U);
    for (int tt = 0; tt < c.NT; ++tt)
      for (int u = 0; u < c.U; ++u)
        svc[(size_t)tt * c.U + u] = (u < S) ? cfg->load_time[(size_t)tt * S + u] : cfg->dumper_cycle[u - S];
    rc = upload(h, svc.data(), svc.size(), &c.svc_time);
  }
  if (rc == OM_OK) {
    // undisturbed trip time per (speed class, road): (distance / speed) * 60 with the reference's operation order (truck.py:153)
    std::vector<double> speeds;
    std::vector<int32_t> sc(T);
    for (int t = 0; t < T; ++t) {
      size_t k = 0;
      while (k < speeds.size() && speeds[k] != cfg->truck_speed[t]) ++k;
      if (k == speeds.size()) speeds.push_back(cfg->truck_speed[t]);
      sc[t] = (int32_t)k;
    }
    std::vector<double> trip(speeds.size() * (size_t)c.R);
    for (size_t k = 0; k < speeds.size(); ++k)
      for (int r = 0; r < c.R; ++r) {
        volatile double q = cfg->road_distance[r] / speeds[k];      // no contraction / reassociation across the two steps
        trip[k * (size_t)c.R + r] = q * 60.0;
      }
    rc = upload(h, sc.data(), sc.size(), &c.sclass);
    if (rc == OM_OK) rc = upload(h, trip.data(), trip.size(), &c.trip60);
  }
  if (rc == OM_OK) {
    // observation helpers of the RL surface (see DevCfg)
    std::vector<float> tt(c.R), df(c.R);
    for (int r = 0; r < c.R; ++r) {
      volatile double m = cfg->road_distance[r] * 60.0;
      tt[r] = (float)(m / 25.0);
      df[r] = (float)cfg->road_distance[r];
    }
    std::vector<double> iu(c.NS), ip(L), idc(M);
    for (int s = 0; s < c.NS; ++s) iu[s] = 1.0 / (double)(cfg->site_unit_begin[s + 1] - cfg->site_unit_begin[s]);
    for (int s = 0; s < L; ++s) ip[s] = 1.0 / cfg->site_productivity[s];
    for (int u = 0; u < M; ++u) idc[u] = 1.0 / cfg->dumper_cycle[u];
    c.inv_tn = 1.0 / ((double)T + 1e-8);
    rc = upload(h, tt.data(), tt.size(), &c.obs_tt);
    if (rc == OM_OK) rc = upload(h, df.data(), df.size(), &c.obs_dist);
    if (rc == OM_OK) rc = upload(h, iu.data(), iu.size(), &c.inv_units);
    if (rc == OM_OK) rc = upload(h, ip.data(), ip.size(), &c.inv_prod);
    if (rc == OM_OK) rc = upload(h, idc.data(), idc.size(), &c.inv_dcycle);
  }
  if (rc == OM_OK) {
    std::vector<int32_t> rb(c.NS + 1);
    rb[0] = -1;
    for (int cur = 1; cur <= c.NS; ++cur)
      rb[cur] = (cur <= L) ? (cur - 1) * D - 1 : L + L * D + (cur - L - 1) * L - 1;
    rc = upload(h, rb.data(), rb.size(), &c.rbase);
  }
  return rc;
}

static int create_impl(const OmConfig* const* cfgs, int32_t n_configs, const int32_t* config_id, int64_t n_instances,
                       int device, OmHandle* out) {
  OmEngine* h = nullptr;
  if (!cfgs || !out || n_configs <= 0) return set_err(h, OM_E_INVALID, "om_create: null argument");
  if (n_instances <= 0 || n_instances > 0x7FFFFFFFLL) return set_err(h, OM_E_INVALID, "om_create: n_instances must be in 1..2^31-1");
  if (n_configs > 1 && !config_id) return set_err(h, OM_E_INVALID, "om_create_multi: config_id is required with more than one config");
  for (int32_t k = 0; k < n_configs; ++k) {
    int rc0 = check_config(cfgs[k]);
    if (rc0 != OM_OK) return rc0;
    if (cfgs[k]->flags != cfgs[0]->flags) return set_err(h, OM_E_INVALID, "om_create_multi: all configs must carry the same flags");
  }
  if (n_configs > 1 && (cfgs[0]->flags & OM_CFG_RL_SURFACE))
    return set_err(h, OM_E_UNSUPPORTED, "om_create_multi: the RL step surface is per config (the observation width depends on the mine)");
  if (config_id)
    for (int64_t i = 0; i < n_instances; ++i)
      if (config_id[i] < 0 || config_id[i] >= n_configs) return set_err(h, OM_E_INVALID, "om_create_multi: config_id out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return set_err(h, OM_E_CUDA, "om_create: no CUDA device (this engine has no CPU fallback)");
  if (device < 0 || device >= ndev) return set_err(h, OM_E_INVALID, "om_create: bad device index");
  h = new OmEngine();
  h->device = device;
  h->n = n_instances;
  h->multi = n_configs > 1;
  int rc = OM_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { rc = set_err(nullptr, OM_E_CUDA, cudaGetErrorString(e)); delete h; return rc; }
  h->cfgs.resize(n_configs);
  for (int32_t k = 0; k < n_configs && rc == OM_OK; ++k) rc = build_devcfg(h, cfgs[k], h->cfgs[k]);
  DevCfg& c = h->c;
  if (rc == OM_OK) {
    c = h->cfgs[0];
    if (h->multi) {
      // the shared layout is sized for the largest config in every dimension; the table pointers of h->c are unused
      memset(&c, 0, sizeof(c));
      for (const DevCfg& k : h->cfgs) {
        c.T = std::max(c.T, k.T); c.S = std::max(c.S, k.S); c.M = std::max(c.M, k.M); c.U = std::max(c.U, k.U);
        c.R = std::max(c.R, k.R); c.NS = std::max(c.NS, k.NS); c.J = std::max(c.J, k.J); c.L = std::max(c.L, k.L);
        c.D = std::max(c.D, k.D); c.NT = std::max(c.NT, k.NT);
      }
    }
    rc = upload(h, h->cfgs.data(), h->cfgs.size(), &h->d_cfgs);
    if (rc == OM_OK && h->multi) rc = upload(h, config_id, (size_t)n_instances, &h->d_cfg_id);
  }
  if (rc == OM_OK) {
    void* d = nullptr;
    e = cudaMalloc(&d, 2 * sizeof(unsigned long long));          // {work-queue counter, retired warps}: see om_retire_warp
    if (e == cudaSuccess) e = cudaMemset(d, 0, 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) rc = set_err(h, OM_E_CUDA, cudaGetErrorString(e));
    else { h->counter = (unsigned long long*)d; h->dev_allocs.push_back(d); }
  }
  if (rc == OM_OK) {
    void* d = nullptr;
    const size_t nb = sizeof(unsigned long long) * OM_STATS_DISPATCHERS * OM_STATS_PER_DISP + 16;
    e = cudaMalloc(&d, nb);
    if (e == cudaSuccess) e = cudaMemset(d, 0, nb);
    if (e != cudaSuccess) rc = set_err(h, OM_E_CUDA, cudaGetErrorString(e));
    else {
      h->stats_acc = (unsigned long long*)d; h->dev_allocs.push_back(d);
      h->stats_done = (unsigned int*)(h->stats_acc + OM_STATS_DISPATCHERS * OM_STATS_PER_DISP);
    }
  }
  if (rc == OM_OK) {
    void* d = nullptr;
    e = cudaMalloc(&d, sizeof(int32_t) * (size_t)n_instances);
    if (e != cudaSuccess) rc = set_err(h, OM_E_NOMEM, cudaGetErrorString(e));
    else { h->order = (int32_t*)d; h->dev_allocs.push_back(d); }
  }
  if (rc == OM_OK) {
    void* d = nullptr;
    e = cudaMalloc(&d, sizeof(uint32_t) * (size_t)n_instances);
    if (e == cudaSuccess) e = cudaMemset(d, 0, sizeof(uint32_t) * (size_t)n_instances);
    if (e != cudaSuccess) rc = set_err(h, OM_E_NOMEM, cudaGetErrorString(e));
    else { h->seg_done = (uint32_t*)d; h->dev_allocs.push_back(d); }
  }
  if (rc == OM_OK) {
    h->rl = (cfgs[0]->flags & OM_CFG_RL_SURFACE) != 0;
    build_layout(c, h->ly, h->rl, false);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int maxsm = 0;
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    int limit = maxsm < OM_SMEM_LIMIT ? maxsm : OM_SMEM_LIMIT;
    // Whole blob in shared memory when that leaves at least 8 resident warps per SM; otherwise only the hot part
    // (pending entries, unit records, roads, per-truck location / breakdown arrays) is staged and truck records / the
    // jam pool are accessed in place.  The layout depends on the decision (split layouts carry extra hot arrays).
    h->split = (limit / (int)h->ly.total_bytes) < 8;
    if (const char* ev = getenv("OM_SPLIT")) h->split = atoi(ev) != 0;   // tuning knob
    if (h->split) build_layout(c, h->ly, h->rl, true);
    const int staged = (int)(h->split ? h->ly.hot_bytes : h->ly.total_bytes) + (h->multi ? (int)OM_CFG_SLOT : 0);
    int w = limit / staged;
    if (w > OM_MAX_WARPS_PER_CTA) w = OM_MAX_WARPS_PER_CTA;
    if (const char* ev = getenv("OM_WARPS_PER_CTA")) {   // tuning knob (occupancy experiments)
      int wv = atoi(ev);
      if (wv >= 1 && wv < w) w = wv;
    }
    if (w < 1) rc = set_err(h, OM_E_UNSUPPORTED, "om_create: one instance's hot state does not fit in shared memory");
    else {
      // no more resident warps than instances per SM need (balances the last wave)
      long long per_sm = (n_instances + sms - 1) / sms;
      if (per_sm < w) w = (int)(per_sm < 1 ? 1 : per_sm);
      h->warps = w; h->block = w * 32; h->smem = w * staged;
      long long ctas = (n_instances + w - 1) / w;
      h->grid = (int)(ctas < sms ? ctas : sms);
      // more instances than resident warps: the last tenth of every episode is a work item of its own, handed out after all
      // the first parts -- when the queue runs dry the resident warps have a tenth of an episode left, not a whole one
      // (measured on the C3 shard: 2 x {0.9, 0.1} 26.35 ms, 2 x {0.8, 0.2} 26.55, 3 x 26.50, 4 x 26.6-26.8, 1 x 28.2)
      // With only a few waves the long first parts leave stragglers nobody can fill in for (C2: 6144 instances = 2.07 waves;
      // four equal segments 5.44 ms, {0.9, 0.1} 5.75 ms, one 6.06 ms), so short batches are cut into four equal parts.
      const double waves = (double)n_instances / ((double)h->grid * (double)w);
      h->segments = (h->split || waves <= 1.0) ? 1 : (waves >= 4.0 ? 2 : 4);
      if (const char* ev = getenv("OM_SEGMENTS")) { int sv = atoi(ev); if (sv >= 1 && sv <= 16 && !h->split) h->segments = sv; }
      h->seg_tail = (h->segments == 2) ? 0.1 : 1.0 / (double)h->segments;
      if (const char* ev = getenv("OM_SEG_TAIL")) { double tv = atof(ev); if (tv > 0.0 && tv * (h->segments - 1) < 1.0) h->seg_tail = tv; }
      const void* fns[] = {
#define OM_RUN_FNS(MULTI)                                                                                                 \
          (const void*)om_run_kernel<OM_RNG_DET, false, false, MULTI>, (const void*)om_run_kernel<OM_RNG_PHILOX, false, false, MULTI>, \
          (const void*)om_run_kernel<OM_RNG_TAPE, false, false, MULTI>, (const void*)om_run_kernel<OM_RNG_DET, true, false, MULTI>,    \
          (const void*)om_run_kernel<OM_RNG_PHILOX, true, false, MULTI>, (const void*)om_run_kernel<OM_RNG_TAPE, true, false, MULTI>,  \
          (const void*)om_run_kernel<OM_RNG_DET, false, true, MULTI>, (const void*)om_run_kernel<OM_RNG_PHILOX, false, true, MULTI>,   \
          (const void*)om_run_kernel<OM_RNG_TAPE, false, true, MULTI>, (const void*)om_run_kernel<OM_RNG_DET, true, true, MULTI>,      \
          (const void*)om_run_kernel<OM_RNG_PHILOX, true, true, MULTI>, (const void*)om_run_kernel<OM_RNG_TAPE, true, true, MULTI>
          OM_RUN_FNS(false), OM_RUN_FNS(true),
#undef OM_RUN_FNS
          (const void*)om_step_kernel<OM_RNG_DET, false>, (const void*)om_step_kernel<OM_RNG_PHILOX, false>,
          (const void*)om_step_kernel<OM_RNG_TAPE, false>, (const void*)om_step_kernel<OM_RNG_DET, true>,
          (const void*)om_step_kernel<OM_RNG_PHILOX, true>, (const void*)om_step_kernel<OM_RNG_TAPE, true>};
      for (size_t i = 0; i < sizeof(fns) / sizeof(fns[0]) && rc == OM_OK; ++i) {
        // the same value for every handle (a per-handle value would lower the limit under older, larger handles)
        e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
        if (e != cudaSuccess) rc = set_err(h, OM_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
      }
    }
  }
  if (rc != OM_OK) {
    g_create_err = h->err;
    for (void* d : h->dev_allocs) cudaFree(d);
    delete h;
    return rc;
  }
  *out = h;
  return OM_OK;
}

int om_create(const OmConfig* cfg, int64_t n_instances, int device, OmHandle* out) {
  if (!cfg || !out) return set_err(nullptr, OM_E_INVALID, "om_create: null argument");
  return create_impl(&cfg, 1, nullptr, n_instances, device, out);
}

int om_create_multi(const OmConfig* const* cfgs, int32_t n_configs, const int32_t* config_id, int64_t n_instances, int device,
                    OmHandle* out) {
  return create_impl(cfgs, n_configs, config_id, n_instances, device, out);
}

int om_destroy(OmHandle h) {
  if (!h) return OM_E_INVALID;
  cudaSetDevice(h->device);
  for (void* d : h->dev_allocs) cudaFree(d);
  delete h;
  return OM_OK;
}

int om_state_bytes(OmHandle h, size_t* per_instance) {
  if (!h || !per_instance) return OM_E_INVALID;
  *per_instance = h->ly.total_bytes;
  return OM_OK;
}

int om_state_layout(OmHandle h, OmStateLayout* out) {
  if (!h || !out) return OM_E_INVALID;
  *out = h->ly;
  return OM_OK;
}

int om_launch_info(OmHandle h, int32_t* grid, int32_t* block, int32_t* smem_bytes, int32_t* warps_per_cta) {
  if (!h) return OM_E_INVALID;
  if (grid) *grid = h->grid;
  if (block) *block = h->block;
  if (smem_bytes) *smem_bytes = h->smem;
  if (warps_per_cta) *warps_per_cta = h->warps;
  return OM_OK;
}

int om_reset(OmHandle h, const uint64_t* seeds, const int32_t* dispatcher, int rng_mode, const double* tape_values,
             const uint8_t* tape_sites, const int64_t* tape_offsets, void* state, void* stream) {
  if (!h || !seeds || !dispatcher || !state) return set_err(h, OM_E_INVALID, "om_reset: null argument");
  if (rng_mode < OM_RNG_DET || rng_mode > OM_RNG_TAPE) return set_err(h, OM_E_INVALID, "om_reset: bad rng_mode");
  if (rng_mode == OM_RNG_TAPE && (!tape_values || !tape_sites || !tape_offsets))
    return set_err(h, OM_E_INVALID, "om_reset: OM_RNG_TAPE needs tape_values/tape_sites/tape_offsets");
  CK(cudaSetDevice(h->device));
  h->rng_mode = rng_mode;
  h->reset_done = true;
  h->tape_values = tape_values; h->tape_sites = tape_sites; h->tape_offsets = tape_offsets;
  CK(cudaMemsetAsync(h->counter, 0, 2 * sizeof(unsigned long long), (cudaStream_t)stream));   // also after an aborted launch
  CK(cudaMemsetAsync(h->seg_done, 0, sizeof(uint32_t) * (size_t)h->n, (cudaStream_t)stream));
  const int wpb = 8;
  long long blocks = (h->n + wpb - 1) / wpb;
  om_reset_kernel<<<(unsigned)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(h->d_cfgs, h->d_cfg_id, h->ly, (unsigned char*)state,
                                                                          seeds, dispatcher, h->n);
  CK(cudaGetLastError());
  om_order_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(dispatcher, h->n, h->order);
  CK(cudaGetLastError());
  return OM_OK;
}

int om_run(OmHandle h, double until, void* state, const OmOutputs* outs, void* stream) {
  if (!h || !state) return set_err(h, OM_E_INVALID, "om_run: null argument");
  if (!h->reset_done) return set_err(h, OM_E_INVALID, "om_run: om_reset has not been called on this handle (the state blobs are uninitialised)");
  if (((uintptr_t)state & 15u) != 0) return set_err(h, OM_E_INVALID, "om_run: state must be 16-byte aligned (TMA bulk copies)");
  if (h->split && !(until <= 65535.0)) return set_err(h, OM_E_UNSUPPORTED, "om_run: split layouts (big mines) hold breakdown minutes as uint16: until <= 65535");
  CK(cudaSetDevice(h->device));
  RunParams p;
  memset(&p, 0, sizeof(p));
  p.c = h->c; p.ly = h->ly;
  if (outs) p.out = *outs;
  if (p.out.ticks == nullptr) p.out.tick_capacity = 0;
  p.state = (unsigned char*)state;
  p.tape_values = h->tape_values; p.tape_sites = h->tape_sites; p.tape_offsets = h->tape_offsets;
  p.counter = h->counter; p.n_instances = h->n; p.until = until;
  p.order = h->order;
  p.cfgs = h->d_cfgs; p.cfg_id = h->d_cfg_id; p.ns_stride = h->c.NS; p.u_stride_out = h->c.U;
  p.seg_done = h->seg_done; p.n_segments = h->segments; p.seg_epoch = 0; p.seg_tail = h->seg_tail;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.out.draw_values != nullptr && (p.out.draw_sites == nullptr || p.out.draw_capacity <= 0))
    return set_err(h, OM_E_INVALID, "om_run: draw_values needs draw_sites and draw_capacity > 0");
  if (p.out.draw_values != nullptr && h->rng_mode != OM_RNG_PHILOX)
    return set_err(h, OM_E_INVALID, "om_run: the draw-tape export exists in OM_RNG_PHILOX mode only");
  const bool trace = p.out.trace != nullptr || p.out.orders != nullptr || p.out.draw_values != nullptr;   // event-trace / order-log / draw-tape variant
#define OM_LAUNCH2(MODE, MULTI)                                                                     \
  do {                                                                                              \
    if (h->split) {                                                                                 \
      if (trace) om_run_kernel<MODE, true, true, MULTI><<<h->grid, h->block, h->smem, st>>>(p);     \
      else om_run_kernel<MODE, false, true, MULTI><<<h->grid, h->block, h->smem, st>>>(p);          \
    } else {                                                                                        \
      if (trace) om_run_kernel<MODE, true, false, MULTI><<<h->grid, h->block, h->smem, st>>>(p);    \
      else om_run_kernel<MODE, false, false, MULTI><<<h->grid, h->block, h->smem, st>>>(p);         \
    }                                                                                               \
  } while (0)
#define OM_LAUNCH(MODE) do { if (h->multi) OM_LAUNCH2(MODE, true); else OM_LAUNCH2(MODE, false); } while (0)
  switch (h->rng_mode) {
    case OM_RNG_DET: OM_LAUNCH(OM_RNG_DET); break;
    case OM_RNG_PHILOX: OM_LAUNCH(OM_RNG_PHILOX); break;
    default: OM_LAUNCH(OM_RNG_TAPE); break;
  }
#undef OM_LAUNCH
#undef OM_LAUNCH2
  CK(cudaGetLastError());
  return OM_OK;
}

int om_obs_dim(OmHandle h, int32_t* dim) {
  if (!h || !dim) return OM_E_INVALID;
  const DevCfg& c = h->c;
  int C = c.L > c.D ? c.L : c.D;
  *dim = 6 + (1 + c.NS) + 2 + 3 * C + 2 * c.R + 5 * c.NS;
  return OM_OK;
}

int om_step(OmHandle h, double until, void* state, const OmStepIO* io, void* stream) {
  if (!h || !state || !io || !io->obs) return set_err(h, OM_E_INVALID, "om_step: null argument");
  if (!h->reset_done) return set_err(h, OM_E_INVALID, "om_step: om_reset has not been called on this handle (the state blobs are uninitialised)");
  if (!h->rl) return set_err(h, OM_E_INVALID, "om_step: the handle was created without OM_CFG_RL_SURFACE (OmConfig.flags)");
  if (h->multi) return set_err(h, OM_E_UNSUPPORTED, "om_step: not available on a mixed-config handle");
  if (((uintptr_t)state & 15u) != 0) return set_err(h, OM_E_INVALID, "om_step: state must be 16-byte aligned (TMA bulk copies)");
  if (!(until <= 65535.0)) return set_err(h, OM_E_UNSUPPORTED, "om_step: the RL surface holds jam start minutes as uint16: until <= 65535");
  CK(cudaSetDevice(h->device));
  RunParams p;
  memset(&p, 0, sizeof(p));
  p.c = h->c; p.ly = h->ly;
  p.c.sim_time = until;
  p.c.inv_sim_time = 1.0 / until;
  p.io = *io;
  if ((const void*)io->obs != h->last_obs) {      // a buffer this handle has not written yet: its static columns too
    p.io.flags |= OM_STEP_WRITE_STATIC;
    h->last_obs = (const void*)io->obs;
  }
  om_obs_dim(h, &p.obs_dim);
  p.state = (unsigned char*)state;
  p.tape_values = h->tape_values; p.tape_sites = h->tape_sites; p.tape_offsets = h->tape_offsets;
  p.counter = h->counter; p.n_instances = h->n; p.until = until;
  p.cfgs = h->d_cfgs; p.cfg_id = nullptr; p.ns_stride = h->c.NS; p.u_stride_out = h->c.U;
  cudaStream_t st = (cudaStream_t)stream;
  switch (h->rng_mode) {
#define OM_STEP(MODE)                                                                        \
  do {                                                                                       \
    if (h->split) om_step_kernel<MODE, true><<<h->grid, h->block, h->smem, st>>>(p);         \
    else om_step_kernel<MODE, false><<<h->grid, h->block, h->smem, st>>>(p);                 \
  } while (0)
    case OM_RNG_DET: OM_STEP(OM_RNG_DET); break;
    case OM_RNG_PHILOX: OM_STEP(OM_RNG_PHILOX); break;
    default: OM_STEP(OM_RNG_TAPE); break;
#undef OM_STEP
  }
  CK(cudaGetLastError());
  return OM_OK;
}

int om_debug_draws(int kind, uint64_t seed, uint32_t arg, int64_t n, double* out, void* stream) {
  if (!out || n <= 0 || kind < 0 || kind > 6) return OM_E_INVALID;
  om_debug_draws_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, seed, arg, (long long)n, out);
  return cudaGetLastError() == cudaSuccess ? OM_OK : OM_E_CUDA;
}

int om_debug_jam_positions(OmHandle h, void* state, int64_t inst, int32_t road, int32_t n_samples, double* out, int32_t* fallbacks,
                           double* coverage_out, void* stream) {
  if (!h || !state || !out || inst < 0 || inst >= h->n || n_samples <= 0) return set_err(h, OM_E_INVALID, "om_debug_jam_positions: bad argument");
  if (h->multi || h->split || road < 0 || road >= h->c.R) return set_err(h, OM_E_UNSUPPORTED, "om_debug_jam_positions: one-config, unsplit handles only");
  CK(cudaSetDevice(h->device));
  RunParams p;
  memset(&p, 0, sizeof(p));
  p.c = h->c; p.ly = h->ly; p.state = (unsigned char*)state; p.n_instances = h->n;
  p.ns_stride = h->c.NS; p.u_stride_out = h->c.U;
  om_debug_jam_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, (long long)inst, road, n_samples, out, fallbacks, coverage_out);
  CK(cudaGetLastError());
  return OM_OK;
}

int om_debug_bounds_report(int32_t* out8) {
  if (!out8) return OM_E_INVALID;
#if defined(OM_BOUNDS_CHECK) || defined(OM_CONV_PROBES)
  int v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(v, g_om_bounds, sizeof(v)) != cudaSuccess) return OM_E_CUDA;
  for (int k = 0; k < 8; ++k) out8[k] = v[k];
#ifdef OM_BOUNDS_CHECK
  return 1;      // this binary checks indices and convergence
#else
  return 2;      // this binary checks convergence only
#endif
#else
  for (int k = 0; k < 8; ++k) out8[k] = 0;
  return 0;      // this binary does not check
#endif
}

int om_stats_reduce(OmHandle h, const OmSummary* summary, const int32_t* dispatcher, double* stats_out, void* stream) {
  if (!h || !summary || !dispatcher || !stats_out) return set_err(h, OM_E_INVALID, "om_stats_reduce: null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = (h->n + 255) / 256;
  om_stats_kernel<<<(unsigned)blocks, 256, 0, st>>>(summary, dispatcher, h->n, h->stats_acc, h->stats_done, stats_out);
  CK(cudaGetLastError());
  return OM_OK;
}

}  // extern "C"
