// Device-side engine: one warp simulates one mine instance.
//
// The per-instance state is one blob (layout: OmStateLayout; 16-byte pending entries, per-truck and
// per-unit records, a few structure-of-arrays sections -- see the accessor table in Sim) that lives in HBM
// between calls and is staged into shared memory by a TMA bulk copy for the duration of om_run.  All "scalar" DES
// logic (one truck's transition) is executed warp-uniformly -- every lane computes the same values
// and performs the same shared-memory stores (same value, same address), so no intra-warp
// synchronisation is needed around it; lane-parallel phases (next-event reduction, status counts,
// breakdown sampling, wait estimation) are bracketed by __syncwarp().
//
// Event ordering reproduces SimPy 4.0.2: every pending entry carries (time fp64, key u32) with
// key = (priority << 31) | seq, URGENT = 0, NORMAL = 1, seq = per-instance schedule counter
// (simpy.core.Environment.schedule; call sites listed in SURVEY.md section 2 row 10).  Pending
// entries: one per truck (a truck waits on exactly one thing), one pending Release per
// shovel/dumper, one monitor-block entry.
//
// The kernels built from this file are bound by instruction fetch and dependent-instruction latency, not by memory
// (DESIGN.md section 4): loops are deliberately rolled, hot continuations are laid out in execution order, force-inlined
// helpers have one call site, and nothing that a north_pit episode never executes is compiled into its hot loop.
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/openmines_b200.h"
#include "om_rng.cuh"

static_assert(OM_S_LOAD_JITTER == OM_DRAW_LOAD_JITTER && OM_S_BREAK_GAP == OM_DRAW_BREAK_GAP && OM_S_ROAD_GAP == OM_DRAW_ROAD_GAP &&
              OM_S_JAM_COIN == OM_DRAW_JAM_COIN && OM_S_JAM_DUR == OM_DRAW_JAM_DUR && OM_S_DISPATCH == OM_DRAW_DISPATCH,
              "draw-site ids of om_rng.cuh and include/openmines_b200.h");
#define OM_FULL 0xffffffffu
#ifdef OM_NO_RECONV   /* debugging aid: build without the explicit reconvergence points */
#define OM_RECONV() do { } while (0)
#else
#define OM_RECONV() __syncwarp()
#endif
#define OM_NONE16 0xFFFFu
#define OM_NOLOC 0xFFu
#define OM_KEY_NORMAL 0x80000000u
#define OM_KEY_SENTINEL 0xFFFFFFFFu

// truck statuses (truck.py:297-436 string statuses) -- same codes as oracle/trace_format.py
enum : uint8_t {
  ST_IDLE = 0, ST_WAIT_INIT, ST_MOVING, ST_WAIT_SHOVEL, ST_LOADING, ST_WAIT_HAUL, ST_WAIT_DUMPER,
  ST_UNLOADING, ST_WAIT_BACK, ST_REPAIRING, ST_UNREPAIRABLE
};

// truck continuation points (what happens when the truck's pending entry pops); see SURVEY 3.2
enum : uint8_t {
  PH_DEAD = 0,        // Truck.run returned (unrepairable)
  PH_WAITING,         // in a shovel/dumper FIFO, request not yet granted (no pending entry)
  PH_RUN_INIT,        // U  Initialize(run)                truck.py:296-317
  PH_MOVE_BEGIN,      // U  Initialize(move)               truck.py:116-148
  PH_MOVE_REPAIRED,   //    Timeout(repair_time)           truck.py:149-150
  PH_MOVE_ARRIVE,     //    Timeout(duration+pre_jam)      truck.py:211-218
  PH_MOVE_RET,        //    Process(move) -> run resumes   truck.py:319-339 / 404-419
  PH_LOAD_GRANTED,    //    Request granted                truck.py:342-350
  PH_LOAD_BEGIN,      // U  Initialize(load)               truck.py:221-228
  PH_LOAD_DONE,       //    Timeout(load_time)             truck.py:230-235
  PH_LOAD_RET,        //    Process(load) -> run resumes   truck.py:352-370
  PH_MAINT_DONE,      //    Timeout(maintenance)           truck.py:372-379
  PH_DUMP_GRANTED,    //    Request granted                truck.py:422-429
  PH_DUMP_BEGIN,      // U  Initialize(unload)             truck.py:238-242
  PH_DUMP_DONE,       //    Timeout(dump_time)             truck.py:243-251
  PH_DUMP_RET,        //    Process(unload) -> run resumes truck.py:431-453
  PH_AWAIT_ORDER      //    paused for an external decision (OM_DISP_EXTERNAL)
};
// An exception inside Truck.run (IndexError at truck.py:305,390,442, KeyError at sq_dispatcher.py:31) does not stop SimPy
// at once: Process._resume schedules the failed Process event NORMAL at `now`, and env.step() re-raises when that event
// pops -- entries of the same timestamp scheduled earlier still run.  The truck's entry becomes (now, NORMAL | seq) with
// phase PH_CRASH_BASE + OM_IERR_* and the instance stops when it pops.
#define PH_CRASH_BASE 32u

// trace kinds (include/openmines_b200.h OmTraceRecord; same codes as oracle/trace_format.py)
enum : uint8_t {
  K_INIT_ORDER = 0, K_MOVE_INIT, K_MOVE_HAUL, K_MOVE_UNHAUL, K_WAIT_SHOVEL, K_GET_SHOVEL, K_ORDER_HAUL,
  K_ORDER_BACK, K_WAIT_DUMPER, K_GET_DUMPER, K_BREAKDOWN, K_UNREPAIRABLE, K_JAM, K_ROAD_REPAIR,
  K_SHOVEL_MAINT_START, K_SHOVEL_MAINT_DONE
};

// Device view of the packed config (pointers are device pointers owned by the handle).
struct DevCfg {
  int T, L, D, S, M, U, R, NS, NT, J, sum_mode, nearest_init;
  const double* cap;
  const double* speed;
  const int32_t* ttype;
  const int32_t* ubeg;
  const double* dcycle;
  const double* rdist;
  const double* load_time;
  const int32_t* rbase;     // [NS+1] road_id(cur, tgt) = rbase[cur] + tgt
  const double* svc_time;   // [NT*U] load_time for shovels, dumper cycle for dumpers (built in om_create)
  const int32_t* sclass;    // [T] speed class of the truck (built in om_create)
  const double* trip60;     // [n speed classes * R] (road_distance / speed) * 60: the undisturbed trip time, truck.py:153
  const double* sptf_cost;
  const double* jam_chance;
  const int32_t* near_haul;
  const int32_t* near_back;
  const int32_t* fixed_group;
  const double* shovel_rate;
  const double* site_prod;
  // RL observation helpers (built in om_create; reciprocals replace ~10 fp64 divisions per env step -- observations are
  // float32 and pinned to the reference at 2 ulp fp32, the reward and the teacher keep their exact divisions)
  const float* obs_tt;        // [R] (road_distance * 60) / 25 as float: travel minutes at the hard-coded speed (feature_processing.py:43)
  const float* obs_dist;      // [R] road_distance as float (static block of the observation)
  const double* inv_units;    // [NS] 1 / units of the site
  const double* inv_prod;     // [L] 1 / load_site_productivity
  const double* inv_dcycle;   // [M] 1 / dumper cycle time
  double inv_tn;              // 1 / (T + 1e-8)
  double inv_sim_time;        // om_step: 1 / until
  double sim_time;
  double road_lambda, road_mu, road_sigma, detour_mu, detour_sigma;
  double breakdown_mean, repair_mu, repair_sigma, unrepairable_mean, maint_mean, maint_mu, maint_sigma;
  double jam_sd, jam_scale;
};

// Per-instance scalars, first bytes of the state blob.
struct InstHdr {
  double now;
  double tick_time;
  double reported_tons;
  uint64_t seed;
  uint64_t tape_pos;
  uint32_t seq;
  uint32_t tick_key;
  int32_t tick_k;
  int32_t dispatcher;
  uint32_t n_events, n_decisions, n_orders, n_hops;
  uint32_t n_jams, n_road_rep, n_brk, n_unrep;
  uint32_t jam_count;
  uint32_t err;
  uint32_t trace_count;
  uint32_t order_count;
  int32_t pending_truck;
  int32_t pending_kind;
  uint32_t n_ticks;
  uint32_t n_rel;       // pending Release entries in the unit table
  // RL step surface (rl_dispatch.py): window start, last sampled production, closed-interval accumulators
  double rl_last_time;
  double rl_last_prod;
  double acc_wait, acc_shovel, acc_dumper, acc_move, acc_jam;
  uint32_t finished;
  uint32_t rl_steps;
};
static_assert(sizeof(InstHdr) == 184, "InstHdr size");
static_assert(sizeof(OmTickRecord) == 32 && sizeof(OmUnitTick) == 16 && sizeof(OmSummary) == 64, "output record sizes");
#define OM_HDR_BYTES 192

struct RunParams {
  DevCfg c;
  OmStateLayout ly;
  OmOutputs out;
  unsigned char* state;       // [N * ly.total_bytes]
  const double* tape_values;
  const uint8_t* tape_sites;
  const int64_t* tape_offsets;
  unsigned long long* counter; // dynamic instance scheduler
  const int32_t* order;        // instance ids grouped by dispatcher (om_run): position in the work queue -> instance
  long long n_instances;
  double until;
  OmStepIO io;      // om_step only
  int obs_dim;
  // mixed-config batches (om_create_multi, the MULTI kernels): per-instance config index into a device table; p.c then holds
  // only the maxima the shared layout was sized for
  const DevCfg* cfgs;
  const int32_t* cfg_id;
  int ns_stride, u_stride_out;   // row widths of the per-site / per-unit tick outputs (max L+D / max S+M over the configs)
  // episode segments (om_run, unsplit layouts): an episode is cut into n_segments work items so that the last, partially
  // filled wave of the persistent grid is short; seg_done[inst] == s  <=>  segments < s of the current launch are stored
  uint32_t* seg_done;
  uint32_t seg_epoch;
  int n_segments;
  double seg_tail;     // every segment after the first covers this fraction of the episode (the last one sets the tail)
};

// Identity the optimiser cannot see through: a chain of `if (om_opaque(x) == K)` stays a chain of compares and branches
// (the compiler otherwise folds compares of one value back into a switch and lowers that to an indirect branch, BRX).
__device__ __forceinline__ uint32_t om_opaque(uint32_t v) { asm volatile("" : "+r"(v)); return v; }   // volatile: two of them are not merged

__device__ __forceinline__ double om_inf() { return __longlong_as_double(0x7FF0000000000000LL); }

// bounds-check build: {violations, source line of the first, its index, the extent, warp-convergence probes that saw a
// partial active mask, source line of the first, its mask}
__device__ int g_om_bounds[8];
#if defined(OM_BOUNDS_CHECK) || defined(OM_CONV_PROBES)
#define OM_CONV() do { unsigned am_ = __activemask(); if (am_ != 0xffffffffu && atomicAdd(&g_om_bounds[4], 1) == 0) { g_om_bounds[5] = __LINE__; g_om_bounds[6] = (int)am_; } } while (0)
#else
#define OM_CONV() do { } while (0)
#endif
// Branch-free on purpose: a check that branches changes where the compiler reconverges the warp, and the engine's redundant
// all-lane stores are only safe on a converged warp (see emit()).  The violation is latched in a per-lane register (source
// line of the first one) and published when the instance ends (om_publish_violation).
__device__ __forceinline__ long long om_chk(long long i, long long n, int line, uint32_t& latch) {
  const bool bad = (i < 0) | (i >= n);
  latch = (bad && latch == 0u) ? (uint32_t)line : latch;
  return bad ? 0 : i;
}
__device__ __forceinline__ void om_publish_violation(uint32_t latch) {
  if (latch != 0u && atomicAdd(&g_om_bounds[0], 1) == 0) g_om_bounds[1] = (int)latch;
}

// Out-of-line helpers: the hot loop must stay small enough for the instruction caches (profiles/r1_a),
// so the 10-round Philox and the fp64 division (a ~35-instruction Newton sequence) exist once.
template <int UNROLL>
__device__ __noinline__ Philox4 om_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  return philox4x32_10<UNROLL>(c0, c1, c2, c3, k0, k1);
}
__device__ __noinline__ double om_ddiv(double a, double b) { return a / b; }
// 0 / b without the division: a zero (or denormal) numerator sends div.rn.f64 down its ~110-instruction slow path, and the
// counts / tonnages divided on the RL surface are zero most of the time.  Same value (+0.0) for b > 0.
__device__ __forceinline__ double om_ddiv0(double a, double b) { return (a == 0.0) ? 0.0 : om_ddiv(a, b); }
// log(x) in fp64.  Out of line while the RL surface took eight logs per env step; with the two call sites that are left
// (one lane-parallel pass over the observation's log slots, one in the reward) the inlined copies win: 220.7 -> 216.3 ms
// per 500-step rollout of 16384 envs in a same-call A/B
__device__ __forceinline__ double om_dlog(double x) { return log(x); }
// CPython >= 3.12 sum() step over exact floats (Neumaier), see oracle/des_oracle.py py_sum
struct PySum {
  double s, c;
  __device__ __forceinline__ void init() { s = 0.0; c = 0.0; }
  // branch-free step: the compensation is always accumulated and only used when the Neumaier flavour is selected
  __device__ __forceinline__ void add(double x, bool /*neumaier*/) {
    double t = s + x;
    double big = (fabs(s) >= fabs(x)) ? s : x, small = (fabs(s) >= fabs(x)) ? x : s;
    c += (big - t) + small;
    s = t;
  }
  __device__ __forceinline__ double result(bool neumaier) const {
    double r = s;
    if (neumaier && c != 0.0 && isfinite(c)) r += c;
    return r;
  }
};

__device__ __forceinline__ void om_init_state(unsigned char* sm, unsigned char* cold, uint32_t sm_bytes, const DevCfg& c,
                                              const OmStateLayout& ly, int lane, uint64_t seed, int dispatcher);

template <int RNG, bool TRACE, bool RL, bool SPLIT = false>
struct Sim {
  unsigned char* sm;     // shared-memory image of the blob (SPLIT: only its first ly.hot_bytes)
  unsigned char* cold;   // base for the cold fields (truck records, jam pool): sm, or the blob in HBM when SPLIT
  const DevCfg& c;
  const OmStateLayout& ly;
  const OmOutputs& out;
  int lane;
  long long inst;
  // hot scalars (registers); mirrored in InstHdr between calls
  double now, tick_time, reported_tons;
  uint32_t seq, tick_key;
  int tick_k, disp;
  uint32_t k0, k1;
  uint32_t n_events, n_decisions, n_orders, n_hops, n_jams, n_road_rep, n_brk, n_unrep;
  uint32_t jam_count, err, trace_count, order_count, n_ticks, n_rel;
  unsigned long long tape_pos, tape_end;
  const double* tape_v;
  const uint8_t* tape_s;
  uint32_t dbg_viol = 0;   // bounds-check build: source line of this lane's first out-of-range state access
  // Values of the current truck that one continuation stores and a later one of the same chain reads back: the SPLIT
  // kernels (truck records in HBM: every such load is an L2 round trip in the scalar chain) keep them in registers along
  // fused chains; the record fields are still written, and an entry through the event loop loads them.  The unsplit
  // kernels read them back from shared memory (measured: carrying them costs the north_pit kernel 2 %).
  static constexpr bool CARRY = SPLIT;
  uint32_t rngw;  // t_rng: Philox word saved for the next draw of this truck (maintenance gap / unrepairable gap)
  double distw;   // t_dist: length of the move the last order started
  Philox4 ordw;   // words of the current order's Philox call (PHILOX mode)
  bool ordw_ready; // ordw already holds them (computed alongside the load draw earlier in this chain)
  uint32_t dirty; // slots (groups of 32 trucks) whose cached minimum is stale (see run())
  double rl_last, rl_last_prod, acc_wait, acc_shovel, acc_dumper, acc_move, acc_jam;   // RL only

  int ns_stride, u_stride_out;
  static constexpr bool LY_DYNAMIC = false;   // see OM_SPLIT_HOT
  // Which helpers a kernel family inlines and which it calls is a measured choice (same-call A/B on the C3 / C5 shards and
  // the C4 rollout): a call costs a CALL and an indirect RET on the critical path, an inlined copy
  // costs code in a fetch-bound kernel.  fp64 division: inlined in the unsplit episode kernels (-2.0 % on C3), called
  // elsewhere (+1 % when inlined in the step kernel); Box-Muller: inlined in the split episode kernels (-2.3 % on C5),
  // called elsewhere (both inlined in the unsplit episode kernel: +4.4 %).  Philox stays out of line everywhere (+2 %).
#ifndef OM_X_DDIV_INLINE
#define OM_X_DDIV_INLINE (!RL && !SPLIT)
#endif
#ifndef OM_X_NORMAL_INLINE
#define OM_X_NORMAL_INLINE (!RL && SPLIT)
#endif
  static constexpr bool DDIV_INLINE = OM_X_DDIV_INLINE, NORMAL_INLINE = OM_X_NORMAL_INLINE;
  __device__ __forceinline__ double ddiv(double a, double b) const { return DDIV_INLINE ? (a / b) : om_ddiv(a, b); }
  __device__ __forceinline__ double ddiv0(double a, double b) const { return (a == 0.0) ? 0.0 : ddiv(a, b); }
  __device__ __forceinline__ float normal_fast(uint32_t a, uint32_t b) const {
    return NORMAL_INLINE ? om_normal_fast_inl(a, b) : om_normal_fast(a, b);
  }
  __device__ __forceinline__ Sim(unsigned char* sm_, unsigned char* blob, const RunParams& p, const DevCfg& cfg, int lane_, long long inst_)
      : sm(sm_), cold(SPLIT ? blob : sm_), c(cfg), ly(p.ly), out(p.out), lane(lane_), inst(inst_), rngw(0u), distw(0.0), ordw_ready(false),
        ns_stride(p.ns_stride), u_stride_out(p.u_stride_out) {}

  // State accessors.  Per-truck and per-unit scalars live in fixed-size records (array of structures): the scalar
  // DES code touches many fields of ONE truck / unit, so a record base + compile-time field offsets costs one
  // address computation per handler instead of one constant-bank load + add per access (the kernel is
  // instruction-fetch bound).  The pending entry of a truck {time, key, phase} is one 16-byte element (one LDS.128 per truck in the
  // next-event scan, one STS.128 per schedule); u_rel_key (a pending Release always has time == now), roads, jams and per-site samples stay
  // structure-of-arrays.  F64(t_load, i) etc. dispatch on the field name.
#define OM_TSTRIDE 80u
#define OM_USTRIDE 40u
#define OM_ESTRIDE 16u   // pending entry of a truck: {double time; uint32 key; uint32 phase}
#define OM_SOA(type, field, i) (((type*)(sm + ly.field))[i])
#define OM_SOAC(type, field, i) (((type*)(cold + ly.field))[i])
#define OM_TREC(type, off, i) (*(type*)(cold + ly.t_rec + (uint32_t)(i) * OM_TSTRIDE + (off)))
#define OM_UREC(type, off, i) (*(type*)(sm + ly.u_rec + (uint32_t)(i) * OM_USTRIDE + (off)))
#define OM_TEV(type, off, i) (*(type*)(sm + ly.t_time + (uint32_t)(i) * OM_ESTRIDE + (off)))
#define OM_A_t_time(type, i) OM_TEV(type, 0, i)
#define OM_A_t_key(type, i) OM_TEV(type, 8, i)
#define OM_A_t_phase(type, i) OM_TEV(type, 12, i)
#define OM_A_t_load(type, i) OM_TREC(type, 0, i)
#define OM_A_t_jstart(type, i) OM_TREC(type, 8, i)
#define OM_A_t_prejam(type, i) OM_TREC(type, 16, i)
  // Split layouts (big mines) keep what the scalar chains LOAD of a truck in small hot arrays, so that a pop costs no
  // L2 round trip for them: t_loc {status, current, target, flags} (4 B), t_aux {unit, next-in-queue, order epoch, service
  // count} (8 B) and the minute of the last breakdown as a uint16 (2 B; it is always an integer minute, see lastbrk_*()).
  // OM_SPLIT_HOT(f): this kernel / this layout has them (the SPLIT kernels only ever see split layouts; om_init_state()
  // serves both and looks at the layout).
#define OM_THOT(type, field, esz, off, i) (*(type*)(sm + ly.field + (uint32_t)(i) * (esz) + (off)))
#define OM_SPLIT_HOT(f) (SPLIT && (!LY_DYNAMIC || ly.f != 0))
#define OM_A_t_lastbrk(type, i) OM_TREC(type, 24, i)   /* unsplit layouts only: use lastbrk_get() / lastbrk_set_now() */
#define OM_A_t_repair(type, i) OM_TREC(type, 32, i)
#define OM_A_t_total(type, i) OM_TREC(type, 40, i)
#define OM_A_t_cycle(type, i) OM_TREC(type, 48, i)
#define OM_A_t_dist(type, i) OM_TREC(type, 56, i)
#define OM_A_t_since(type, i) OM_SOAC(type, t_since, i)   /* RL surface only */
#define OM_A_t_rng(type, i) OM_TREC(type, 64, i)
#define OM_A_t_svc(type, i) (*(OM_SPLIT_HOT(t_aux) ? &OM_THOT(type, t_aux, 8u, 6, i) : &OM_TREC(type, 68, i)))
#define OM_A_t_epoch(type, i) (*(OM_SPLIT_HOT(t_aux) ? &OM_THOT(type, t_aux, 8u, 4, i) : &OM_TREC(type, 70, i)))
#define OM_A_t_next(type, i) (*(OM_SPLIT_HOT(t_aux) ? &OM_THOT(type, t_aux, 8u, 2, i) : &OM_TREC(type, 72, i)))
#define OM_A_t_unit(type, i) (*(OM_SPLIT_HOT(t_aux) ? &OM_THOT(type, t_aux, 8u, 0, i) : &OM_TREC(type, 74, i)))
#define OM_A_t_status(type, i) (*(OM_SPLIT_HOT(t_loc) ? &OM_THOT(type, t_loc, 4u, 0, i) : &OM_TREC(type, 77, i)))
#define OM_A_t_cur(type, i) (*(OM_SPLIT_HOT(t_loc) ? &OM_THOT(type, t_loc, 4u, 1, i) : &OM_TREC(type, 78, i)))
#define OM_A_t_tgt(type, i) (*(OM_SPLIT_HOT(t_loc) ? &OM_THOT(type, t_loc, 4u, 2, i) : &OM_TREC(type, 79, i)))
#define OM_A_t_flags(type, i) OM_THOT(type, t_loc, 4u, 3, i)   /* split layouts only: bit 0 = a repair time is pending (t_repair != 0) */
#define OM_A_u_rel_key(type, i) OM_SOA(type, u_rel_key, i)
#define OM_A_u_last_start(type, i) OM_UREC(type, 0, i)
#define OM_A_u_last_done(type, i) OM_UREC(type, 8, i)
#define OM_A_u_tons(type, i) OM_UREC(type, 16, i)
#define OM_A_u_lastbrk(type, i) OM_SOA(type, u_lastbrk, i)
#define OM_A_u_svc(type, i) OM_UREC(type, 24, i)
#define OM_A_u_user(type, i) OM_UREC(type, 28, i)
#define OM_A_u_head(type, i) OM_UREC(type, 30, i)
#define OM_A_u_tail(type, i) OM_UREC(type, 32, i)
#define OM_A_u_qlen(type, i) OM_UREC(type, 34, i)
#define OM_A_u_repair(type, i) OM_UREC(type, 36, i)
#define OM_A_r_repair_end(type, i) OM_SOA(type, r_repair_end, i)
#define OM_A_j_pos(type, i) OM_SOAC(type, j_pos, i)
#define OM_A_j_start(type, i) OM_SOAC(type, j_start, i)
#define OM_A_j_end(type, i) OM_SOAC(type, j_end, i)
#define OM_A_j_road(type, i) OM_SOAC(type, j_road, i)
#define OM_A_s_qtotal(type, i) OM_SOA(type, s_qtotal, i)
#define OM_A_s_coming(type, i) OM_SOA(type, s_coming, i)
#define OM_A_s_tons(type, i) OM_SOA(type, s_tons, i)
#define OM_A_s_cap(type, i) OM_SOA(type, s_cap, i)
#define OM_A_s_ability(type, i) OM_SOA(type, s_ability, i)
#define OM_A_scratch(type, i) OM_SOA(type, scratch, i)
  // Debug build (-DOM_BOUNDS_CHECK, libopenmines_b200_dbg.so): every state access through the accessors checks its index
  // against the array's extent for THIS instance's config; the first violation is recorded (source line, index, extent;
  // om_debug_bounds_report) and the access is redirected to element 0.  compute-sanitizer is not available on the GPU pool
  // this was developed on, so this build + tools/sanitize_target.py is the memcheck of the state arrays.
#define OM_B_t_time c.T
#define OM_B_t_key c.T
#define OM_B_t_phase c.T
#define OM_B_t_load c.T
#define OM_B_t_jstart c.T
#define OM_B_t_prejam c.T
#define OM_B_t_lastbrk c.T
#define OM_B_t_repair c.T
#define OM_B_t_total c.T
#define OM_B_t_cycle c.T
#define OM_B_t_dist c.T
#define OM_B_t_since c.T
#define OM_B_t_rng c.T
#define OM_B_t_svc c.T
#define OM_B_t_epoch c.T
#define OM_B_t_next c.T
#define OM_B_t_unit c.T
#define OM_B_t_status c.T
#define OM_B_t_cur c.T
#define OM_B_t_tgt c.T
#define OM_B_t_flags c.T
#define OM_B_u_rel_key c.U
#define OM_B_u_last_start c.U
#define OM_B_u_last_done c.U
#define OM_B_u_tons c.U
#define OM_B_u_lastbrk c.S
#define OM_B_u_svc c.U
#define OM_B_u_user c.U
#define OM_B_u_head c.U
#define OM_B_u_tail c.U
#define OM_B_u_qlen c.U
#define OM_B_u_repair c.S
#define OM_B_r_repair_end c.R
#define OM_B_j_pos c.J
#define OM_B_j_start c.J
#define OM_B_j_end c.J
#define OM_B_j_road c.J
#define OM_B_s_qtotal c.NS
#define OM_B_s_coming c.NS
#define OM_B_s_tons c.NS
#define OM_B_s_cap c.NS
#define OM_B_s_ability c.NS
#define OM_B_scratch (((ly.slot_cache ? ly.slot_cache : ly.u_rel_key) - ly.scratch) / 8u)
#ifdef OM_BOUNDS_CHECK
#define OM_IDX(field, i) om_chk((long long)(i), (long long)(OM_B_##field), __LINE__, dbg_viol)
#else
#define OM_IDX(field, i) (i)
#endif
#define F64(field, i) OM_A_##field(double, OM_IDX(field, i))
#define F32A(field, i) OM_A_##field(float, OM_IDX(field, i))
#define U32A(field, i) OM_A_##field(uint32_t, OM_IDX(field, i))
#define U16A(field, i) OM_A_##field(uint16_t, OM_IDX(field, i))
#define U8A(field, i) OM_A_##field(uint8_t, OM_IDX(field, i))

  // ---------------------------------------------------------------------------------------------
  // header <-> registers
  // ---------------------------------------------------------------------------------------------
  __device__ __forceinline__ void load_hdr(const RunParams& p) {
    const InstHdr* h = (const InstHdr*)(sm + ly.hdr);
    now = h->now; tick_time = h->tick_time; reported_tons = h->reported_tons;
    seq = h->seq; tick_key = h->tick_key; tick_k = h->tick_k; disp = h->dispatcher;
    k0 = (uint32_t)h->seed; k1 = (uint32_t)(h->seed >> 32);
    n_events = h->n_events; n_decisions = h->n_decisions; n_orders = h->n_orders; n_hops = h->n_hops;
    n_jams = h->n_jams; n_road_rep = h->n_road_rep; n_brk = h->n_brk; n_unrep = h->n_unrep;
    jam_count = h->jam_count; err = h->err; trace_count = h->trace_count; order_count = h->order_count;
    n_ticks = h->n_ticks; n_rel = h->n_rel;
    dirty = 0xFFFFFFFFu;            // the slot cache is not part of the persistent state
    if (RL) {
      rl_last = h->rl_last_time; rl_last_prod = h->rl_last_prod; acc_wait = h->acc_wait; acc_shovel = h->acc_shovel;
      acc_dumper = h->acc_dumper; acc_move = h->acc_move; acc_jam = h->acc_jam;
    }
    tape_pos = h->tape_pos; tape_end = 0; tape_v = p.tape_values; tape_s = p.tape_sites;
    if (RNG == OM_RNG_TAPE) {
      tape_pos += (unsigned long long)p.tape_offsets[inst];   // hdr stores the position relative to the instance's tape
      tape_end = (unsigned long long)p.tape_offsets[inst + 1];
    }
  }
  __device__ __forceinline__ void store_hdr(const RunParams& p) {
    InstHdr* h = (InstHdr*)(sm + ly.hdr);
    h->now = now; h->tick_time = tick_time; h->reported_tons = reported_tons;
    h->seq = seq; h->tick_key = tick_key; h->tick_k = tick_k;
    h->n_events = n_events; h->n_decisions = n_decisions; h->n_orders = n_orders; h->n_hops = n_hops;
    h->n_jams = n_jams; h->n_road_rep = n_road_rep; h->n_brk = n_brk; h->n_unrep = n_unrep;
    h->jam_count = jam_count; h->err = err; h->trace_count = trace_count; h->order_count = order_count;
    h->n_ticks = n_ticks; h->n_rel = n_rel;
    if (RL) {
      h->rl_last_time = rl_last; h->rl_last_prod = rl_last_prod; h->acc_wait = acc_wait; h->acc_shovel = acc_shovel;
      h->acc_dumper = acc_dumper; h->acc_move = acc_move; h->acc_jam = acc_jam;
    }
    h->tape_pos = (RNG == OM_RNG_TAPE) ? tape_pos - (unsigned long long)p.tape_offsets[inst] : (TAPE_OUT ? tape_pos : 0ull);
  }

  // ---------------------------------------------------------------------------------------------
  // trace / order log (lane 0 writes to HBM)
  // ---------------------------------------------------------------------------------------------
  __device__ __forceinline__ void emit(uint8_t kind, uint32_t truck, uint32_t arg0, uint32_t arg1, double a, double b) {
    if (kind < 12) n_events++;
    if (TRACE && out.trace != nullptr) {
      if (trace_count < (uint32_t)out.trace_capacity) {
        if (lane == 0) {
          OmTraceRecord r;
          r.time = now; r.a = a; r.b = b; r.truck = (uint16_t)truck; r.kind = kind; r.flags = 0;
          r.arg0 = (uint16_t)arg0; r.arg1 = (uint16_t)arg1;
          out.trace[inst * (long long)out.trace_capacity + trace_count] = r;
        }
        // Explicit reconvergence after a lane-0-only region: the warp-uniform code that follows reads state that the same
        // code, run by another lane, changes (e.g. "is the unit still free?" before a grant).  A lane that fell behind here
        // would find the grant already made by the others and skip its own copy of it -- harmless for the state, but lane 0
        // owns the trace.  (Found with the bounds-checking build, whose extra branches made the compiler leave lane 0 behind.)
        OM_RECONV();
      } else if (err == OM_IERR_NONE) {
        err = OM_IERR_TRACE_OVERFLOW;
      }
      trace_count++;
    }
  }
  __device__ __forceinline__ void log_order(int kind, int site) {
    if (TRACE && out.orders != nullptr && order_count < (uint32_t)out.order_capacity) {
      if (lane == 0) out.orders[inst * (long long)out.order_capacity + order_count] = (int16_t)((kind << 12) | site);
      OM_RECONV();      // reconverge after the lane-0-only store (see emit())
    }
    order_count++;
  }

  // ---------------------------------------------------------------------------------------------
  // RNG
  // ---------------------------------------------------------------------------------------------
  __device__ __forceinline__ double tape_next(int site) {
    if (tape_pos >= tape_end || tape_s[tape_pos] != (uint8_t)site) {
      if (err == OM_IERR_NONE || err == OM_IERR_TRACE_OVERFLOW) err = OM_IERR_TAPE;
      return 0.0;
    }
    double v = tape_v[tape_pos];
    tape_pos++;
    return v;
  }
  // Draw-tape export (TRACE variants in PHILOX mode, OmOutputs.draw_values != NULL): every value the reference would have
  // drawn at this point -- post-transform, in the reference's call order, with its site id -- is appended to the instance's
  // tape, so that the CPU restatement can replay the very trajectory the Philox kernel produced (tests/test_gpu_draw_tape.py).
  // Values are written where they are CONSUMED (the merged Philox calls produce several draws ahead of their use).
  static constexpr bool TAPE_OUT = TRACE && (RNG == OM_RNG_PHILOX);
  __device__ __forceinline__ bool tape_on() const { return TAPE_OUT && out.draw_values != nullptr; }
  __device__ __forceinline__ void tape_put_at(unsigned long long pos, int site, double v) {
    if (pos < (unsigned long long)out.draw_capacity) {
      const long long o = inst * (long long)out.draw_capacity + (long long)pos;
      out.draw_values[o] = v;
      out.draw_sites[o] = (uint8_t)site;
    }
  }
  __device__ __forceinline__ void tape_put(int site, double v) {
    if (TAPE_OUT && out.draw_values != nullptr) {
      if (tape_pos >= (unsigned long long)out.draw_capacity) { if (err == OM_IERR_NONE) err = OM_IERR_TRACE_OVERFLOW; }
      else {
        if (lane == 0) tape_put_at(tape_pos, site, v);
        OM_RECONV();    // reconverge after the lane-0-only store (see emit())
      }
      tape_pos++;
    }
  }
  __device__ __forceinline__ Philox4 px(uint32_t epoch, uint32_t entity, uint32_t site, uint32_t stream = 0) const {
    return om_philox<(SPLIT && !RL) ? 10 : 2>(epoch, entity, site, stream, k0, k1);
  }
  // exponential with mean `mean` (np.random.exponential(scale) / random.expovariate(1/mean))
  __device__ __forceinline__ double draw_exp(int site, uint32_t epoch, uint32_t entity, double mean) {
    if (RNG == OM_RNG_DET) return om_inf();
    if (RNG == OM_RNG_TAPE) return tape_next(site);
    double v = (double)om_exp1(px(epoch, entity, site).x) * mean;
    tape_put(site, v);
    return v;
  }
  __device__ __forceinline__ double draw_normal(int site, uint32_t epoch, uint32_t entity, double mu, double sigma) {
    if (RNG == OM_RNG_TAPE) return tape_next(site);
    Philox4 r = px(epoch, entity, site);
    double v = mu + sigma * (double)normal_fast(r.x, r.y);
    tape_put(site, v);
    return v;
  }
  __device__ __forceinline__ double draw_jitter(uint32_t epoch, uint32_t entity) {
    if (RNG == OM_RNG_DET) return 0.0;
    if (RNG == OM_RNG_TAPE) return tape_next(OM_S_LOAD_JITTER);
    double v = -0.1 + 0.2 * (double)om_u01(px(epoch, entity, OM_S_LOAD_JITTER).x);
    tape_put(OM_S_LOAD_JITTER, v);
    return v;
  }
  __device__ __forceinline__ int draw_randint(uint32_t epoch, uint32_t entity, int n) {
    if (RNG == OM_RNG_TAPE) return (int)tape_next(OM_S_DISPATCH);
    int v = (int)om_randint(px(epoch, entity, OM_S_DISPATCH).x, (uint32_t)n);   // also in DET mode (keyed, exact)
    tape_put(OM_S_DISPATCH, (double)v);
    return v;
  }

  // ---------------------------------------------------------------------------------------------
  // scheduling primitives
  // ---------------------------------------------------------------------------------------------
  // one 16-byte store: {time, key, phase}
  __device__ __forceinline__ void set_entry(int i, double t, uint32_t key, uint32_t phase) {
    uint4 e;
    e.x = (uint32_t)__double2loint(t); e.y = (uint32_t)__double2hiint(t); e.z = key; e.w = phase;
    *(uint4*)(sm + ly.t_time + (uint32_t)i * OM_ESTRIDE) = e;
    if (SPLIT) dirty |= 1u << (((uint32_t)i >> 5) & 31u);
  }
  __device__ __forceinline__ void sched_truck(int i, double t, uint32_t prio, uint8_t phase) {
    set_entry(i, t, prio | seq, phase);
    seq++;
  }
  // RL reward bookkeeping: a truck leaves a timed state (wait shovel / get shovel / get dumper / move record);
  // add the part of [t_since, now] that lies in the current reward window (rl_dispatch.py:221-273)
  __device__ __forceinline__ void rl_close(double& acc, int i) {
    if (RL) {
      double st = F64(t_since, i);
      double lo = (st > rl_last) ? st : rl_last;
      if (lo < now) acc += now - lo;
    }
  }
  __device__ __forceinline__ void fail(uint32_t code) {
    if (err == OM_IERR_NONE || err == OM_IERR_TRACE_OVERFLOW) err = code;
  }
  // Truck.last_breakdown_time (truck.py:522-536): only ever set to the clock of a monitor block, i.e. to an integer minute
  // (or its initial 0), so split layouts hold it as a uint16 in the hot part (om_run refuses until > 65535 there).
  __device__ __forceinline__ double lastbrk_get(int i) {
    if (SPLIT) return (double)((const uint16_t*)(sm + ly.t_lastbrk_hot))[OM_IDX(t_lastbrk, i)];
    return F64(t_lastbrk, i);
  }
  __device__ __forceinline__ void lastbrk_set_now(int i) {
    if (SPLIT) ((uint16_t*)(sm + ly.t_lastbrk_hot))[OM_IDX(t_lastbrk, i)] = (uint16_t)(int)now;
    else F64(t_lastbrk, i) = now;
  }
  // Truck.repair_time: written by the minute's breakdown sampling, consumed and cleared by the truck's next move.  Split
  // layouts flag a pending one in the hot part so that the move (and the jam sampler's coverage()) only goes to the
  // truck's record in HBM when there is something to read.
  __device__ __forceinline__ void repair_set(int i, double d) {
    F64(t_repair, i) = d;
    if (SPLIT) U8A(t_flags, i) = (uint8_t)(U8A(t_flags, i) | 1u);
  }
  __device__ __forceinline__ double repair_get(int i) {
    if (SPLIT && (U8A(t_flags, i) & 1u) == 0) return 0.0;
    return F64(t_repair, i);
  }
  __device__ __forceinline__ void repair_clear(int i) {
    F64(t_repair, i) = 0.0;
    if (SPLIT) U8A(t_flags, i) = (uint8_t)(U8A(t_flags, i) & 0xFEu);
  }
  // RL surface: start time of a stored jam (the reward integrates jam minutes over its window, rl_dispatch.py:258-273).
  // Jams are only ever created by the monitor block, i.e. at an integer minute: a uint16 per jam (om_step refuses until > 65535).
  __device__ __forceinline__ double jstart_get(uint32_t j) { return (double)U16A(j_start, j); }
  __device__ __forceinline__ void jstart_set(uint32_t j, double t) { U16A(j_start, j) = (uint16_t)(int)t; }

  // FIFO of waiting trucks per unit (simpy Resource.put_queue), linked through t_next
  __device__ __forceinline__ void q_push(int u, int i) {
    uint32_t n = U16A(u_qlen, u);
    U16A(t_next, i) = OM_NONE16;
    if (n == 0) U16A(u_head, u) = (uint16_t)i;
    else U16A(t_next, U16A(u_tail, u)) = (uint16_t)i;
    U16A(u_tail, u) = (uint16_t)i;
    U16A(u_qlen, u) = (uint16_t)(n + 1);
  }
  __device__ __forceinline__ int q_pop(int u) {
    int h = U16A(u_head, u);
    uint32_t n = U16A(u_qlen, u) - 1;
    U16A(u_head, u) = U16A(t_next, h);
    U16A(u_qlen, u) = (uint16_t)n;
    if (n == 0) U16A(u_tail, u) = OM_NONE16;
    return h;
  }
  // Resource.release: users cleared now, Release event scheduled NORMAL (hand-over happens when it pops).
  // Returns the Release's key; the caller either keeps it in registers (fused chain) or files it.
  __device__ __forceinline__ uint32_t release(int u) {
    U16A(u_user, u) = OM_NONE16;
    uint32_t key = OM_KEY_NORMAL | seq;
    seq++;
    return key;
  }
  __device__ __forceinline__ void file_release(int u, uint32_t key) {
    U32A(u_rel_key, u) = key;          // its time is `now`: it pops before the clock can advance
    n_rel++;
  }

  // get_available_shovel / get_available_dumper (load_site.py:209-220, dump_site.py:105-116):
  // first idle unit, else first unit with the shortest queue.  Lane-parallel over the site's units.
  __device__ __forceinline__ int pick_unit(int site) {
    const int ub = __ldg(&c.ubeg[site]), ue = __ldg(&c.ubeg[site + 1]);
    int u0 = ub;
    uint32_t best = 0xFFFFFFFFu;
    #pragma unroll 1
    do {                                           // one trip per 32 units of the site (north_pit: always one)
      const int u = u0 + lane;
      const bool valid = u < ue;
      const bool idle = valid && U16A(u_user, u) == OM_NONE16;
      const unsigned m = __ballot_sync(OM_FULL, idle);
      if (m) return u0 + __ffs(m) - 1;
      const uint32_t v = valid ? (((uint32_t)U16A(u_qlen, u) << 16) | (uint32_t)(u - ub)) : 0xFFFFFFFFu;
      const uint32_t b = __reduce_min_sync(OM_FULL, v);
      best = b < best ? b : best;
      u0 += 32;
    } while (__builtin_expect(u0 < ue, 0));
    return ub + (int)(best & 0xFFFFu);
  }

  // ---------------------------------------------------------------------------------------------
  // road: distance + repair state (road.py:235-331)
  // ---------------------------------------------------------------------------------------------
  // directed road (cur -> tgt): charging->load i: i; load i->dump j: L + i*D + j; dump j->load i: L + L*D + j*L + i.
  // rbase[cur] folds the per-origin part (built in om_create), so rid = rbase[cur] + tgt.
  __device__ __forceinline__ int road_id(int cur, int tgt) const { return __ldg(&c.rbase[cur]) + tgt; }
  __device__ __forceinline__ double get_distance(int i, int cur, int tgt, uint32_t epoch) {
    int rid = road_id(cur, tgt);
    double end = F64(r_repair_end, rid);
    bool in_repair;
    if (now < end) {
      in_repair = true;
    } else {
      in_repair = false;
      if (RNG != OM_RNG_DET) {
        double gap = (RNG == OM_RNG_PHILOX) ? (double)om_exp1(ordw.x) / c.road_lambda
                                            : draw_exp(OM_S_ROAD_GAP, epoch, (uint32_t)i, 1.0 / c.road_lambda);   // expovariate(lambda)
        if (RNG == OM_RNG_PHILOX) tape_put(OM_S_ROAD_GAP, gap);
        if (__builtin_expect(now >= end + gap, 0)) {   // a new road repair starts: a handful per episode
          double d = draw_normal(OM_S_ROAD_DUR, epoch, (uint32_t)i, c.road_mu, c.road_sigma);
          double dur = (d > 0.0) ? d : 0.0;   // max(0, x)
          in_repair = true;
          end = now + dur;
          F64(r_repair_end, rid) = end;
          n_road_rep++;
          emit(K_ROAD_REPAIR, 0xFFFFu, (uint32_t)rid, 0, end, dur);
        }
      }
    }
    double dist = __ldg(&c.rdist[rid]);
    if (in_repair) {
      double p = draw_normal(OM_S_DETOUR, epoch, (uint32_t)i, c.detour_mu, c.detour_sigma);
      double pun = (p > 0.0) ? p : 0.0;
      dist = dist * (1.0 + pun);
    }
    return dist;
  }

  // ---------------------------------------------------------------------------------------------
  // queue-wait estimation (ParkingLot.update_queue_wait_status, load_site.py:21-68), lane = unit
  // ---------------------------------------------------------------------------------------------
  // Writes est per unit of [ub, ue) into scratch[], then returns (on every lane) nothing; callers read
  // site minima with site_est().
  __device__ __forceinline__ void unit_est_all(bool load_side) {
    int ub = load_side ? 0 : c.S, ue = load_side ? c.S : c.U;
    // only one side's estimates are live at a time (the RL observation reduces each side to per-site minima before it
    // computes the other): dumpers share the shovels' slots
    const int sbase = (!load_side) ? c.S : 0;
    #pragma unroll 1
    for (int u = ub + lane; u < ue; u += 32) {
      double est;
      int n = U16A(u_qlen, u);
      if (load_side) {
        double ls = F64(u_last_start, u), ld = F64(u_last_done, u);
        double last = (ld > ls) ? ld : ls;               // max(last_service_time, last_service_done_time)
        double total = 0.0;
        int t = U16A(u_head, u);
        #pragma unroll 1
        for (int q = 0; q < n; ++q) {
          total += __ldg(&c.load_time[__ldg(&c.ttype[t]) * c.S + u]);
          t = U16A(t_next, t);
        }
        int usr = U16A(u_user, u);
        if (usr != OM_NONE16) total += __ldg(&c.load_time[__ldg(&c.ttype[usr]) * c.S + u]);
        double v = (last + total) - now;
        est = (0.0 > v) ? 0.0 : v;                       // max(v, 0)
      } else {
        double cyc = __ldg(&c.dcycle[u - c.S]);
        double total = 0.0;
        #pragma unroll 1
        for (int q = 0; q < n; ++q) total += cyc;
        double used = now - F64(u_last_start, u);
        est = (total > used) ? (total - used) : 0.0;
      }
      F64(scratch, u - sbase) = est;
    }
    __syncwarp();
  }
  // min over the site's units (call after unit_est_all); lane-local loop
  __device__ __forceinline__ double site_est(int site) {
    int ub = __ldg(&c.ubeg[site]), ue = __ldg(&c.ubeg[site + 1]);
    const int sbase = (site >= c.L) ? c.S : 0;
    double m = F64(scratch, ub - sbase);
    #pragma unroll 1
    for (int u = ub + 1; u < ue; ++u) {
      double v = F64(scratch, u - sbase);
      if (v < m) m = v;
    }
    return m;
  }

  // first-minimum argmin over lanes [0, n) of a double (np.argmin on a list)
  __device__ __forceinline__ int argmin_lanes(double v, int n) {
    double best = __shfl_sync(OM_FULL, v, 0);
    int idx = 0;
    #pragma unroll 1
    for (int s = 1; s < n; ++s) {
      double x = __shfl_sync(OM_FULL, v, s);
      if (x < best) { best = x; idx = s; }
    }
    return idx;
  }

  // Trucks on the roads from location `from` to every candidate site of an order (candidate q <-> location tgt0 + q):
  // lane q gets the number of trucks whose (current, target) pair is that road (mine.update_road_status, mine.py:182-187),
  // the number of MOVING ones (Road.truck_on_road) and the sum of their capacities in truck order.  Lanes run over
  // trucks, one ballot per candidate; the candidate's lane walks its own mask.
  __device__ __forceinline__ void road_load(int from, int tgt0, int ncand, int& pair_cnt, int& on_cnt, double& capsum) {
    const int T = c.T;
    pair_cnt = 0; on_cnt = 0; capsum = 0.0;
#pragma unroll 1
    for (int base = 0; base < T; base += 32) {
      int o = base + lane;
      int tg = -1;
      bool mv = false;
      if (o < T && U8A(t_cur, o) == from) { tg = (int)U8A(t_tgt, o) - tgt0; mv = U8A(t_status, o) == ST_MOVING; }
      if (!__any_sync(OM_FULL, tg >= 0 && tg < ncand)) continue;
      unsigned mvmask = __ballot_sync(OM_FULL, mv);
      unsigned my_pair = 0;
#pragma unroll 1
      for (int q = 0; q < ncand; ++q) {
        unsigned mp = __ballot_sync(OM_FULL, tg == q);
        if (lane == q) my_pair = mp;
      }
      unsigned my_on = my_pair & mvmask;
      pair_cnt += __popc(my_pair);
      on_cnt += __popc(my_on);
      const double mycap = (o < T) ? __ldg(&c.cap[o]) : 0.0;   // values reach the candidate's lane by shuffle, in truck order
#pragma unroll 1
      for (int left = __reduce_max_sync(OM_FULL, __popc(my_on)); left > 0; --left) {
        double x = __shfl_sync(OM_FULL, mycap, __ffs(my_on) - 1);
        if (my_on) capsum += x;
        my_on &= my_on - 1;
      }
    }
  }

  // ---------------------------------------------------------------------------------------------
  // the six built-in dispatchers as device functions (dispatch_algorithms/*.py) behind
  // give_init_order / give_haul_order / give_back_order.  Returns the site index, or -1 on error.
  // ---------------------------------------------------------------------------------------------
  // Every helper below (unit_est_all, site_est, road_load, argmin_lanes) has ONE call site: they are force-inlined,
  // and duplicated copies of them used to make this function 14 KB of SASS in the middle of the hot order chain.
  // `est_ready`: lane s of `est_all` holds site s's estimated queue wait already; `shared_on` / `shared_cap` (lane q = candidate q): the
  // moving trucks / their capacity sum on the roads from the truck's CURRENT location, as the RL observation just counted
  // them -- the ShortestTrip teacher looks at the same roads for haul and init orders (have_shared).
  __device__ __forceinline__ int decide(int kind, int i, bool est_ready = false, bool have_shared = false, int shared_on = 0,
                                        double shared_cap = 0.0, double est_all = 0.0) {
    int r = 0;
    const int L = c.L, D = c.D;
    const bool haul = kind == OM_ORDER_HAUL;
    const int n = haul ? D : L;                 // candidate sites of this order
    const int sbase_site = haul ? L : 0;        // site index of candidate 0
    const bool est = (disp == OM_DISP_SPTF) | (disp == OM_DISP_SHORTESTTRIP);
    double val = om_inf();                      // lane q < n: cost of candidate q (argmin dispatchers)
    double wait = 0.0;                          // lane q < n: estimated_queue_wait_time of candidate q
    if (est) {
      // dispatcher.update_mine (dispatcher.py:24-33) for the side that matters; the RL step has both sides in scratch already
      if (est_ready) {                            // lane s of est_all = estimated_queue_wait_time of site s (rl_observation)
        const double w = __shfl_sync(OM_FULL, est_all, (lane < n) ? sbase_site + lane : 0);
        if (lane < n) wait = w;
      } else {
        unit_est_all(!haul);
        if (lane < n) wait = site_est(sbase_site + lane);
      }
    }
    // Dispatch on the dispatcher id as a chain of compares, the two dispatchers of the headline workload first: a
    // `switch` over seven ids compiles to an indirect branch (BRX), which stalls instruction fetch until it resolves
    // (same-call A/B on the C3 shard).  OM_X_SWITCH_DISP restores the switch for measurements.
#ifdef OM_X_SWITCH_DISP
#define OM_DISPATCH_BEGIN switch (disp) {
#define OM_DISPATCH_CASE(x) case x:
#define OM_DISPATCH_END(e) default: return (e); }
#else
#define OM_DISPATCH_BEGIN do {
#define OM_DISPATCH_CASE(x) if (om_opaque((uint32_t)disp) == (uint32_t)(x))
#define OM_DISPATCH_END(e) return (e); } while (0);
#endif
    OM_DISPATCH_BEGIN
      OM_DISPATCH_CASE(OM_DISP_SQ) {
        // coming trucks + minute-sampled total queue length (sq_dispatcher.py:18-100)
        int k = (int)now;   // int(env.now)
        if (k != tick_k) return -OM_IERR_SQ_KEYERROR;   // KeyError: queue_status['total'][int(now)] not sampled yet
        uint32_t v = 0xFFFFFFFFu;
        if (lane < n) {
          int site = sbase_site + lane;
          uint32_t coming = (uint32_t)U16A(s_coming, site) - ((U8A(t_tgt, i) == 1 + site) ? 1u : 0u);
          v = ((coming + (uint32_t)U16A(s_qtotal, site)) << 8) | (uint32_t)lane;
        }
        r = (int)(__reduce_min_sync(OM_FULL, v) & 0xFFu);
        break;
      }
      OM_DISPATCH_CASE(OM_DISP_SPTF) {
        if (haul) {
          // len(coming) + estimated_queue_wait_time   (sptf_dispatcher.py:40-57)
          if (lane < D) {
            int site = L + lane;
            int coming = (int)U16A(s_coming, site) - ((U8A(t_tgt, i) == 1 + site) ? 1 : 0);
            val = (double)coming + wait;
          }
        } else {
          // sum(cap/productivity for coming trucks, truck order) + estimated_queue_wait_time (:21-38)
          // Lanes run over trucks: every lane fetches its own truck's cost at the site it is heading to (parallel loads, off
          // the summation chain); lane s then adds the costs of the trucks coming to site s in truck order, taking each value
          // from its owner by shuffle.
          PySum ps; ps.init();
          const bool neu = c.sum_mode != 0;
          #pragma unroll 1
          for (int base = 0; base < c.T; base += 32) {
            int o = base + lane;
            int tg = (o < c.T && o != i) ? (int)U8A(t_tgt, o) : 0;
            double mycost = 0.0;
            if (tg >= 1 && tg <= L) mycost = __ldg(&c.sptf_cost[__ldg(&c.ttype[o]) * L + (tg - 1)]);
            unsigned mine_mask = 0;
            #pragma unroll 1
            for (int s = 0; s < L; ++s) {
              unsigned m = __ballot_sync(OM_FULL, tg == 1 + s);
              if (lane == s) mine_mask = m;
            }
            #pragma unroll 1
            for (int left = __reduce_max_sync(OM_FULL, __popc(mine_mask)); left > 0; --left) {
              int b = __ffs(mine_mask) - 1;          // -1 (-> lane 31, value unused) once this lane's list is done
              double x = __shfl_sync(OM_FULL, mycost, b);
              ps.add(mine_mask ? x : 0.0, neu);      // + 0.0 leaves both the sum and the compensation unchanged
              mine_mask &= mine_mask - 1;
            }
          }
          if (lane < L) val = ps.result(neu) + wait;
        }
        break;
      }
      OM_DISPATCH_CASE(OM_DISP_NAIVE) {
        r = 0;
        break;
      }
      OM_DISPATCH_CASE(OM_DISP_NEAREST) {
        if (kind == OM_ORDER_INIT) r = c.nearest_init;
        else if (haul) r = __ldg(&c.near_haul[U8A(t_cur, i) - 1]);
        else r = __ldg(&c.near_back[U8A(t_cur, i) - 1 - L]);
        break;
      }
      OM_DISPATCH_CASE(OM_DISP_FIXEDGROUP) {
        if (haul) r = __ldg(&c.near_haul[U8A(t_cur, i) - 1]);
        else r = __ldg(&c.fixed_group[i]);
        break;
      }
      OM_DISPATCH_CASE(OM_DISP_RANDOM) {
        // haul draws over len(load_sites), back over len(dump_sites)  (random_dispatcher.py:21-29)
        r = draw_randint(U16A(t_epoch, i), (uint32_t)i, (kind == OM_ORDER_BACK) ? D : L);
        break;
      }
      OM_DISPATCH_CASE(OM_DISP_SHORTESTTRIP) {
        // shortest_trip_dispatcher.py:12-84: earliest expected finish of (travel | queue + on-road trucks) + service.
        // Trucks already on the road: for haul orders those from the current load site, for init AND back orders those
        // on the charging-site roads (:71).
        const double v = __ldg(&c.speed[i]);
        const int cur = U8A(t_cur, i);
        int pc_ = 0, cnt = shared_on;
        double capsum = shared_cap;
        if (!(have_shared && (haul || cur == 0))) road_load(haul ? cur : 0, 1 + sbase_site, n, pc_, cnt, capsum);
        if (lane < n) {
          int site = sbase_site + lane, tgt = 1 + site;
          int ub = __ldg(&c.ubeg[site]), ue = __ldg(&c.ubeg[site + 1]);
          double reach = now + ddiv(60.0 * __ldg(&c.rdist[road_id(cur, tgt)]), v);
          double busy, service_time;
          if (haul) {
            double dt0 = __ldg(&c.dcycle[ub - c.S]);
            busy = ddiv0((double)cnt, (double)(ue - ub)) * dt0;
            service_time = dt0;
          } else {
            double prod = __ldg(&c.site_prod[lane]);
            busy = ddiv0(capsum, prod);
            service_time = ddiv(__ldg(&c.cap[i]), ddiv(prod, (double)(ue - ub)));
          }
          double avail = (now + wait) + busy;
          double service = (avail > reach) ? avail : reach;        // np.maximum
          val = (service + service_time) - now;
        }
        break;
      }
    OM_DISPATCH_END(-OM_IERR_BAD_ORDER)
    if (est) {
      r = argmin_lanes(val, n);
      __syncwarp();
    }
    return r;
  }

  __device__ __forceinline__ void set_target(int i, int tgt) {
    int old = U8A(t_tgt, i);
    if (old != OM_NOLOC && old != 0) U16A(s_coming, old - 1) = (uint16_t)(U16A(s_coming, old - 1) - 1);
    U16A(s_coming, tgt - 1) = (uint16_t)(U16A(s_coming, tgt - 1) + 1);
    U8A(t_tgt, i) = (uint8_t)tgt;
  }

  // wrapper bookkeeping (dispatcher.py:35-89): order counts; SPTF back orders are logged twice
  __device__ __forceinline__ void count_order(int kind, int site) {
    n_decisions++;
    n_orders += (disp == OM_DISP_SPTF && kind == OM_ORDER_BACK) ? 2u : 1u;
    log_order(kind, site);
  }

  // ---------------------------------------------------------------------------------------------
  // truck transitions
  // ---------------------------------------------------------------------------------------------
  // after an order has been decided: truck.py:303-317 (init), 390-402 (haul), 442-453 (back)
  __device__ __forceinline__ void apply_order(int i, int kind, int site) {
    const int L = c.L;
    uint32_t epoch = U16A(t_epoch, i);
    if (RNG == OM_RNG_PHILOX) {
      // one call per order: x -> road-repair gap (haul/back), y -> unrepairable gap of the move that follows.  A haul
      // order issued in the same chain as its LOAD_DONE got these words from the load's call (ordw_ready).
      if (!ordw_ready) ordw = px(epoch, (uint32_t)i, OM_S_ROAD_GAP);
      ordw_ready = false;
      U32A(t_rng, i) = ordw.y;
      rngw = ordw.y;
    }
    count_order(kind, site);
    U16A(t_epoch, i) = (uint16_t)(epoch + 1);
    double dist;
    const int tgt = (kind == OM_ORDER_HAUL) ? 1 + L + site : 1 + site;
    if (kind == OM_ORDER_INIT) {
      U8A(t_status, i) = ST_MOVING;
      dist = __ldg(&c.rdist[site]);      // charging_to_load[site]; no road events on this leg (truck.py:308)
    } else {
      dist = get_distance(i, U8A(t_cur, i), tgt, epoch);
    }
    set_target(i, tgt);
    emit((kind == OM_ORDER_INIT) ? K_INIT_ORDER : ((kind == OM_ORDER_HAUL) ? K_ORDER_HAUL : K_ORDER_BACK), (uint32_t)i,
         (uint32_t)site, 0, dist, 0.0);
    F64(t_dist, i) = dist;
    if (CARRY) distw = dist;
  }

  // Big mines (SPLIT layouts): the trucks under way on every road are kept as a doubly linked list per road in shared memory
  // (r_head / t_rnext / t_rprev, with the count in r_on), maintained at departure and arrival -- so the per-minute jam
  // sampling neither recounts 500 trucks nor scans them once per sampled position (both were 16-trip passes).
  __device__ __forceinline__ void road_link(int i, int rid) {
    if (SPLIT) {
      uint16_t* head = (uint16_t*)(sm + ly.r_head);
      uint16_t* nxt = (uint16_t*)(sm + ly.t_rnext);
      uint16_t* prv = (uint16_t*)(sm + ly.t_rprev);
      uint16_t* cnt = (uint16_t*)(sm + ly.r_on);
      const uint16_t h = head[rid];
      nxt[i] = h; prv[i] = OM_NONE16;
      if (h != OM_NONE16) prv[h] = (uint16_t)i;
      head[rid] = (uint16_t)i;
      cnt[rid] = (uint16_t)(cnt[rid] + 1);
    }
  }
  __device__ __forceinline__ void road_unlink(int i, int rid) {
    if (SPLIT) {
      uint16_t* head = (uint16_t*)(sm + ly.r_head);
      uint16_t* nxt = (uint16_t*)(sm + ly.t_rnext);
      uint16_t* prv = (uint16_t*)(sm + ly.t_rprev);
      uint16_t* cnt = (uint16_t*)(sm + ly.r_on);
      const uint16_t p = prv[i], n = nxt[i];
      if (p != OM_NONE16) nxt[p] = n; else head[rid] = n;
      if (n != OM_NONE16) prv[n] = p;
      cnt[rid] = (uint16_t)(cnt[rid] - 1);
    }
  }

  // move(): jam scan + departure (truck.py:151-210)
  __device__ __forceinline__ void move_depart(int i, bool dead) {
    int cur = U8A(t_cur, i), tgt = U8A(t_tgt, i);
    double dist = CARRY ? distw : F64(t_dist, i);   // CARRY: every path into the departure has t_dist in the register
    const int rid = road_id(cur, tgt);
    // (distance / speed) * 60: tabulated per (speed class, road) for the road's own length; only a detour divides
    double duration = __ldg(&c.trip60[__ldg(&c.sclass[i]) * c.R + rid]);
    if (__builtin_expect(dist != __ldg(&c.rdist[rid]), 0)) duration = ddiv(dist, __ldg(&c.speed[i])) * 60.0;
    double pre = 0.0;
    if (RNG != OM_RNG_DET && jam_count > 0) {
      const bool maybe = !SPLIT || ((const uint16_t*)(sm + ly.r_jamcnt))[rid] != 0;
      #pragma unroll 1
      for (uint32_t base = 0; maybe && base < jam_count; base += 32) {
        uint32_t j = base + lane;
        bool hit = false;
        if (j < jam_count)
          hit = (U16A(j_road, j) == rid) && (F64(j_end, j) >= now);   // start_time <= now holds for every stored jam
        unsigned m = __ballot_sync(OM_FULL, hit);
        #pragma unroll 1
        while (m) {
          int b = __ffs(m) - 1;
          m &= m - 1;
          double y = (F64(j_end, base + b) - now) - F64(j_pos, base + b) * duration;
          pre += (y > 0.0) ? y : 0.0;        // max(0, jam_time - time_to_jam)
        }
      }
      F64(t_prejam, i) = pre;
    }
    if (__builtin_expect(dead, 0)) {
      n_unrep++;
      emit(K_UNREPAIRABLE, (uint32_t)i, 0, 0, 0.0, 0.0);
      U8A(t_status, i) = ST_UNREPAIRABLE;
      set_entry(i, om_inf(), OM_KEY_SENTINEL, PH_DEAD);
      return;
    }
    uint8_t kind = (cur > c.L) ? K_MOVE_UNHAUL : (cur == 0 ? K_MOVE_INIT : K_MOVE_HAUL);
    int idx = (tgt <= c.L) ? tgt - 1 : tgt - 1 - c.L;
    emit(kind, (uint32_t)i, (uint32_t)idx, 0, now + duration + pre, dist);
    if (RL) F64(t_since, i) = now;
    U8A(t_status, i) = ST_MOVING;
    road_link(i, rid);
    sched_truck(i, now + (duration + pre), OM_KEY_NORMAL, PH_MOVE_ARRIVE);
  }

  // ---------------------------------------------------------------------------------------------
  // jam pool
  // ---------------------------------------------------------------------------------------------
  __device__ __forceinline__ void jam_append(int rid, double pos, double end) {
    // Big mines (SPLIT): per-road count of stored jams in shared memory, so a departure on a road without jams (most
    // of them) does not scan the jam pool in HBM.  Exact after a compaction, an over-estimate (expired jams) between.
    uint16_t* rjc = SPLIT ? (uint16_t*)(sm + ly.r_jamcnt) : nullptr;   // every split layout has it
    if (__builtin_expect(jam_count == (uint32_t)c.J, 0)) {   // drop expired entries, keep creation order
      if (SPLIT) {
        #pragma unroll 1
        for (int r = lane; r < c.R; r += 32) rjc[r] = 0;
        __syncwarp();
      }
      uint32_t w = 0;
      #pragma unroll 1
      for (uint32_t j = 0; j < jam_count; ++j) {
        if (RL && !(F64(j_end, j) >= now)) {
          double st = jstart_get(j), lo = (st > rl_last) ? st : rl_last;
          if (lo < F64(j_end, j)) acc_jam += F64(j_end, j) - lo;
        }
        if (F64(j_end, j) >= now) {
          if (w != j) {
            F64(j_pos, w) = F64(j_pos, j); if (RL) U16A(j_start, w) = U16A(j_start, j);
            F64(j_end, w) = F64(j_end, j); U16A(j_road, w) = U16A(j_road, j);
          }
          if (SPLIT) rjc[U16A(j_road, w)] = (uint16_t)(rjc[U16A(j_road, w)] + 1);
          w++;
        }
      }
      jam_count = w;
    }
    if (jam_count == (uint32_t)c.J) { fail(OM_IERR_JAM_OVERFLOW); return; }
    F64(j_pos, jam_count) = pos; if (RL) jstart_set(jam_count, now); F64(j_end, jam_count) = end;
    U16A(j_road, jam_count) = (uint16_t)rid;
    if (SPLIT) rjc[rid] = (uint16_t)(rjc[rid] + 1);
    jam_count++;
  }

  // Big mines (SPLIT): drop the expired jams once per minute, lanes over the pool, creation order kept (a move adds the
  // delays of the jams on its road in that order).  The pool of the 500-truck mine otherwise fills up with expired entries
  // (it is only compacted when full) and every departure on a road with a stored jam scans all 512 of them.
  __device__ __forceinline__ void jam_compact_parallel() {
    uint16_t* rjc = (uint16_t*)(sm + ly.r_jamcnt);     // only called by the SPLIT kernels, whose layouts have it
    uint32_t w = 0;
    double expired = 0.0;                              // RL: jam minutes of the dropped entries inside the reward window
    #pragma unroll 1
    for (uint32_t base = 0; base < jam_count; base += 32) {
      const uint32_t j = base + lane;
      const bool valid = j < jam_count;
      double en = 0.0, pos = 0.0, st = 0.0;
      uint32_t rd = 0;
      if (valid) { en = F64(j_end, j); pos = F64(j_pos, j); rd = U16A(j_road, j); if (RL) st = jstart_get(j); }
      const bool keep = valid && en >= now;
      if (RL && valid && !keep) {
        const double lo = (st > rl_last) ? st : rl_last;
        if (lo < en) expired += en - lo;
      }
      const unsigned m = __ballot_sync(OM_FULL, keep);
      unsigned md = __ballot_sync(OM_FULL, valid && !keep);
      const uint32_t dst = w + (uint32_t)__popc(m & ((1u << lane) - 1u));
      __syncwarp();                                    // the chunk is in registers before any lane overwrites part of it
      if (keep) {
        F64(j_end, dst) = en; F64(j_pos, dst) = pos; U16A(j_road, dst) = (uint16_t)rd; if (RL) jstart_set(dst, st);
      }
      w += (uint32_t)__popc(m);
      {
        #pragma unroll 1
        while (md) {                                   // stored jams per road: the dropped ones leave (warp-uniform)
          const int b = __ffs(md) - 1;
          md &= md - 1;
          const uint32_t r = __shfl_sync(OM_FULL, rd, b);
          rjc[r] = (uint16_t)(rjc[r] - 1);
        }
      }
    }
    if (RL) {
      #pragma unroll 1
      for (int o = 16; o > 0; o >>= 1) expired += __shfl_xor_sync(OM_FULL, expired, o);
      acc_jam += expired;
    }
    jam_count = w;
    __syncwarp();
  }

  // ---------------------------------------------------------------------------------------------
  // the per-minute monitor block (mine.py:50-155 + site monitors), atomic
  // ---------------------------------------------------------------------------------------------
  __device__ __forceinline__ void tick_block() {
    const int T = c.T, L = c.L, D = c.D, NS = c.NS;
    const int k = (int)now;
    const bool neu = c.sum_mode != 0;
    // (1) queue totals per site (load_site.py:81-83) and site tons (load_site.py:175, dump_site.py:69)
    double site_tons = 0.0, rl_cap = 0.0;
    uint32_t qtot = 0, site_svc = 0;
    if (lane < NS) {
      int ub = __ldg(&c.ubeg[lane]), ue = __ldg(&c.ubeg[lane + 1]);
      PySum ps; ps.init();
      #pragma unroll 1
      for (int u = ub; u < ue; ++u) {
        qtot += U16A(u_qlen, u);
        ps.add(F64(u_tons, u), neu);
        site_svc += U32A(u_svc, u);
      }
      site_tons = ps.result(neu);
      U16A(s_qtotal, lane) = (uint16_t)qtot;
      if (RL) {
        // minute-sampled values the observation reads (load_site.py:175-185, dump_site.py:69-78), held as float32 (what
        // the observation is).  They change once per
        // minute but are observed at every decision, so the observation's log(x + 1) of them is taken here, once:
        // s_tons <- log(site tons + 1), s_cap <- log(capacity + 1), s_ability <- service ability ratio.
        if (lane < L) {
          PySum pa; pa.init();
#pragma unroll 1
          for (int u = ub; u < ue; ++u) pa.add(__ldg(&c.shovel_rate[u]) * (U8A(u_repair, u) ? 0.0 : 1.0), neu);
          rl_cap = __ldg(&c.site_prod[lane]);
          F32A(s_ability, lane) = (float)(pa.result(neu) * __ldg(&c.inv_prod[lane]));
        } else {
          PySum pd; pd.init();
#pragma unroll 1
          for (int u = ub; u < ue; ++u) pd.add(F64(u_tons, u) * __ldg(&c.inv_dcycle[u - c.S]), neu);
          rl_cap = pd.result(neu);
          F32A(s_ability, lane) = 1.0f;
        }
      }
    }
    if (RL) {
      // one log call for both blocks: lanes [0, NS) capacity, the same lanes' tons in a second trip of a rolled loop
#pragma unroll 1
      for (int w = 0; w < 2; ++w) {
        const double v = om_dlog(((w == 0) ? rl_cap : site_tons) + 1.0);
        if (lane < NS) { if (w == 0) F32A(s_cap, lane) = (float)v; else F32A(s_tons, lane) = (float)v; }
      }
    }
    tick_k = k;
    // (2) mine KPI sample (mine.py:61-119)
    PySum pm; pm.init();
    int svc_total = 0;
    #pragma unroll 1
    for (int j = 0; j < D; ++j) {
      pm.add(__shfl_sync(OM_FULL, site_tons, L + j), neu);
      svc_total += (int)__shfl_sync(OM_FULL, site_svc, L + j);
    }
    double produce = pm.result(neu);
    reported_tons = produce;
    // truck status counts (mine.py:69-86): every lane counts its trucks in two packed accumulators (16 bits per class),
    // two warp sums at the end.  A truck turns unrepairable once and stays (truck.py:545-547), so the count of trucks
    // still in service is T - n_unrep.
    const uint32_t WAIT_SET = (1u << ST_WAIT_INIT) | (1u << ST_WAIT_SHOVEL) | (1u << ST_WAIT_HAUL) | (1u << ST_WAIT_DUMPER) |
                              (1u << ST_WAIT_BACK);
    const uint32_t LU_SET = (1u << ST_LOADING) | (1u << ST_UNLOADING);
    uint32_t acc_a = 0, acc_b = 0;                 // {waiting, loading/unloading}, {moving, repairing}
    #pragma unroll 1
    for (int i = lane; i < T; i += 32) {
      const uint32_t st = U8A(t_status, i);
      acc_a += ((WAIT_SET >> st) & 1u) | (((LU_SET >> st) & 1u) << 16);
      acc_b += (st == ST_MOVING ? 1u : 0u) | (st == ST_REPAIRING ? 0x10000u : 0u);
    }
    acc_a = __reduce_add_sync(OM_FULL, acc_a);
    acc_b = __reduce_add_sync(OM_FULL, acc_b);
    const uint32_t n_work = (uint32_t)T - n_unrep, n_wait = acc_a & 0xFFFFu, n_lu = acc_a >> 16, n_mov = acc_b & 0xFFFFu,
                   n_rep = acc_b >> 16;
    if (out.ticks != nullptr && k < out.tick_capacity) {
      if (lane == 0) {
        // OmTickRecord (32 B) packed in registers and written as two 16-byte stores
        uint4 a, b;
        a.x = (uint32_t)__double2loint(produce); a.y = (uint32_t)__double2hiint(produce);
        a.z = (uint32_t)svc_total; a.w = (n_work & 0xFFFFu) | (n_wait << 16);
        // the cumulative event counters saturate at 65535 instead of wrapping
        b.x = (n_lu & 0xFFFFu) | (n_mov << 16); b.y = (n_rep & 0xFFFFu) | (min(n_jams, 0xFFFFu) << 16);
        b.z = min(n_road_rep, 0xFFFFu) | (min(n_brk, 0xFFFFu) << 16); b.w = n_unrep & 0xFFFFu;
        uint4* dst = (uint4*)(out.ticks + inst * (long long)out.tick_capacity + k);
        dst[0] = a; dst[1] = b;
      }
      if (lane < NS) {
        long long o = (inst * (long long)out.tick_capacity + k) * ns_stride + lane;
        if (out.site_queue_ticks != nullptr) out.site_queue_ticks[o] = (uint16_t)qtot;
        if (out.site_tons_ticks != nullptr) out.site_tons_ticks[o] = site_tons;
      }
      if (__builtin_expect(out.unit_ticks != nullptr, 0)) {
#pragma unroll 1
        for (int u = lane; u < c.U; u += 32) {
          OmUnitTick r;
          r.tons = F64(u_tons, u); r.service_count = U32A(u_svc, u); r.queue_length = U16A(u_qlen, u);
          r.repair = (u < c.S) ? U8A(u_repair, u) : (uint8_t)0; r.pad = 0;
          out.unit_ticks[(inst * (long long)out.tick_capacity + k) * u_stride_out + u] = r;
        }
      }
      OM_RECONV();      // reconverge after the lane-specific output stores
    }
    n_ticks++;
    if (RNG != OM_RNG_DET) {
      // (3) truck breakdown sampling (truck.py:522-536), every truck every minute, truck order
      if (RNG == OM_RNG_TAPE) {
        #pragma unroll 1
        for (int i = 0; i < T; ++i) {
          double gap = tape_next(OM_S_BREAK_GAP);
          if (now >= lastbrk_get(i) + gap) {
            double d = tape_next(OM_S_BREAK_DUR);
            repair_set(i, (0.0 > d) ? 0.0 : d);
            lastbrk_set_now(i);
          }
        }
      } else {
        // one Philox call serves three trucks (words x,y,z -> gaps of trucks q, q+C, q+2C); the repair-time normal of
        // the lane's first breakdown in this minute comes from word w and the gap words' unused low bytes, a second
        // or third one (rare) draws its own call.  Rolled loops on purpose (instruction-fetch bound).
        const int C = (T + 2) / 3;
        // tape export: the reference draws truck by truck (gap, then the repair time if the breakdown fires); lanes work on
        // trucks out of order, so the minute reserves two slots per truck and leaves 0xFF in the second when nothing fired
        const bool tp = tape_on();
        const unsigned long long tbase = tape_pos;
        if (tp && tbase + 2ull * (unsigned long long)T > (unsigned long long)out.draw_capacity && err == OM_IERR_NONE)
          err = OM_IERR_TRACE_OVERFLOW;
        #pragma unroll 1
        for (int q = lane; q < C; q += 32) {
          Philox4 r = px((uint32_t)k, (uint32_t)q, OM_S_BREAK_GAP);
          uint32_t na = r.w, nb2 = ((r.x & 0xFFu) << 24) | ((r.y & 0xFFu) << 16) | ((r.z & 0xFFu) << 8);
          bool used = false;
          #pragma unroll 1
          for (int w = 0; w < 3; ++w) {
            int i = q + w * C;
            uint32_t word = (w == 0) ? r.x : ((w == 1) ? r.y : r.z);
            if (i < T) {
              double gap = (double)om_exp1(word) * c.breakdown_mean;
              if (TAPE_OUT && tp) {
                tape_put_at(tbase + 2ull * (unsigned long long)i, OM_S_BREAK_GAP, gap);
                tape_put_at(tbase + 2ull * (unsigned long long)i + 1ull, OM_DRAW_SKIP, 0.0);
              }
              if (now >= lastbrk_get(i) + gap) {
                if (used) {
                  Philox4 r2 = px((uint32_t)k, (uint32_t)i, OM_S_BREAK_DUR);
                  na = r2.x; nb2 = r2.y;
                }
                used = true;
                double d = c.repair_mu + c.repair_sigma * (double)normal_fast(na, nb2);
                if (TAPE_OUT && tp) tape_put_at(tbase + 2ull * (unsigned long long)i + 1ull, OM_S_BREAK_DUR, d);
                repair_set(i, (0.0 > d) ? 0.0 : d);
                lastbrk_set_now(i);
              }
            }
          }
        }
        if (TAPE_OUT && tp) tape_pos = tbase + 2ull * (unsigned long long)T;
        __syncwarp();
      }
      // (4) jam sampling per road in monitor order (mine.py:138-154, road.py:137-233)
      if (SPLIT && jam_count > 32u) jam_compact_parallel();
      if (now >= 2.0) jam_sampling(k);
    }
    // re-arm: every monitor generator does `yield env.timeout(1)`; the block's eids are contiguous
    tick_time = now + 1.0;
    tick_key = OM_KEY_NORMAL | seq;
    seq++;
  }

  __device__ __forceinline__ void jam_sampling(int k) {
    const int T = c.T, R = c.R;
    uint32_t* rcnt = (uint32_t*)(sm + ly.scratch);   // scratch is sized >= max(U, R/2 + 1, 32) doubles
    const uint16_t* ron = SPLIT ? (const uint16_t*)(sm + ly.r_on) : nullptr;   // big mines: kept incrementally
    if (!SPLIT) {
      #pragma unroll 1
      for (int r = lane; r < R; r += 32) rcnt[r] = 0;
      __syncwarp();
      #pragma unroll 1
      for (int i = lane; i < T; i += 32) {
        int cur = U8A(t_cur, i), tgt = U8A(t_tgt, i);
        int rid = (U8A(t_status, i) == ST_MOVING && cur != tgt) ? road_id(cur, tgt) : -1;
        if (rid >= 0) atomicAdd(&rcnt[rid], 1u);
      }
      __syncwarp();
    }
    // every candidate road flips its own keyed coin (road.py:179): lane l of the 64-road group g serves roads
    // 64g + l (words x: coin, z: duration) and 64g + 32 + l (words y, w) from ONE Philox call
    Philox4 rr;
    rr.x = rr.y = rr.z = rr.w = 0u;
    #pragma unroll 1
    for (int base = 0; base < R; base += 32) {
      int r = base + lane;
      uint32_t n_on = (r < R) ? (SPLIT ? (uint32_t)ron[r] : rcnt[r]) : 0u;
      bool candidate = n_on >= 2u;                          // road.py:148-176: >= 2 moving trucks
      bool hit = candidate;
      double coin = 0.0;
      const bool second = (base & 32) != 0;
      if (RNG == OM_RNG_PHILOX) {
        if (!second) {
          uint32_t n2 = (r + 32 < R) ? (SPLIT ? (uint32_t)ron[r + 32] : rcnt[r + 32]) : 0u;
          if (__any_sync(OM_FULL, candidate || n2 >= 2u)) rr = px((uint32_t)k, (uint32_t)r, OM_S_JAM_COIN);
        }
        uint32_t cw = second ? rr.y : rr.x;
        coin = (double)((float)(cw >> 8) * 5.9604644775390625e-08f);   // [0,1)
        hit = candidate && coin < __ldg(&c.jam_chance[candidate ? n_on : 0u]);
      }
      // tape export: the reference flips the coin of every candidate road, in road order, whether or not a jam follows
      const bool tp = tape_on();
      unsigned m = __ballot_sync(OM_FULL, (TAPE_OUT && tp) ? candidate : hit);
      #pragma unroll 1
      while (m) {
        int b = __ffs(m) - 1;
        m &= m - 1;
        int rid = base + b;
        double pos, dur;
        if (TAPE_OUT && tp) {
          tape_put(OM_S_JAM_COIN, __shfl_sync(OM_FULL, coin, b));
          if (!__shfl_sync(OM_FULL, (int)hit, b)) continue;
        }
        if (RNG == OM_RNG_TAPE) {
          uint32_t n = __shfl_sync(OM_FULL, n_on, b);
          double coin = tape_next(OM_S_JAM_COIN);
          if (coin >= __ldg(&c.jam_chance[n])) continue;
          pos = tape_next(OM_S_JAM_POS);
          dur = tape_next(OM_S_JAM_DUR) * c.jam_scale;
        } else {
          uint32_t wz = __shfl_sync(OM_FULL, second ? rr.w : rr.z, b);
          pos = sample_jam_position(k, rid, __shfl_sync(OM_FULL, n_on, b));
          const double wb = (double)sqrtf(om_exp1(wz));                           // Weibull(shape 2)
          dur = wb * c.jam_scale;
          tape_put(OM_S_JAM_POS, pos);
          tape_put(OM_S_JAM_DUR, wb);
        }
        double end = now + dur;
        jam_append(rid, pos, end);
        n_jams++;
        emit(K_JAM, 0xFFFFu, (uint32_t)rid, 0, pos, end);
      }
    }
    __syncwarp();
  }

  // Position of a new jam on road `rid`: the reference draws an index on linspace(0,1,1000) with
  // p ~ sum_i N(x; coverage_i, sd) over the trucks on the road (road.py:185-206).  Equivalent rejection
  // sampler: pick one of the N trucks uniformly, draw x ~ N(coverage_i, sd), accept iff x falls in the
  // grid window [-h/2, 1+h/2] (else pick again), round to the grid.  Accepted x has density
  // ~ sum_i N(x; c_i, sd) on the window, i.e. the reference's categorical up to the mid-point rule
  // (relative error O((h/sd)^2/24) = 4e-6).  Acceptance >= 1/2 per attempt.
  template <bool DBG = false>
  __device__ __forceinline__ double sample_jam_position(int k, int rid, uint32_t n_on, int* fallbacks = nullptr) {
    const int T = c.T;
    const float sd = (float)c.jam_sd, h = 1.0f / 999.0f;
    float x = 0.5f;
#pragma unroll 1
    for (uint32_t it = 0; it < 64; ++it) {
      Philox4 r = px((uint32_t)k, (uint32_t)rid, OM_S_JAM_POS, it);
      int pick = (int)om_randint(r.x, n_on);               // pick-th truck on this road (truck order; list order for big mines)
      int found = -1;
      if (SPLIT) {
        // big mines: walk the road's list (a uniform pick is uniform whatever the order of the list)
        const uint16_t* nxt = (const uint16_t*)(sm + ly.t_rnext);
        uint32_t tk = ((const uint16_t*)(sm + ly.r_head))[rid];
#pragma unroll 1
        for (int q = 0; q < pick && tk != OM_NONE16; ++q) tk = nxt[tk];
        found = (tk != OM_NONE16) ? (int)tk : -1;
      } else {
#pragma unroll 1
        for (int base = 0; base < T && found < 0; base += 32) {
          int i = base + lane;
          bool on = false;
          if (i < T) {
            int cur = U8A(t_cur, i), tgt = U8A(t_tgt, i);
            on = U8A(t_status, i) == ST_MOVING && cur != tgt && road_id(cur, tgt) == rid;
          }
          unsigned m = __ballot_sync(OM_FULL, on);
          int cnt = __popc(m);
          if (pick < cnt) {
            // position of the pick-th set bit
#pragma unroll 1
            for (int q = 0; q < pick; ++q) m &= m - 1;
            found = base + __ffs(m) - 1;
          } else {
            pick -= cnt;
          }
        }
      }
      if (found < 0) break;                                 // cannot happen: n_on trucks are on the road
      float cv = coverage(found, rid);
      float xx = cv + sd * normal_fast(r.y, r.z);
      if (xx >= -0.5f * h && xx <= 1.0f + 0.5f * h) { x = xx; break; }
      if (DBG && it == 63 && fallbacks != nullptr) (*fallbacks)++;      // 64 rejections in a row (p <= 2^-64): x stays 0.5
    }
    int idx = __float2int_rn(x * 999.0f);
    idx = idx < 0 ? 0 : (idx > 999 ? 999 : idx);
    return (double)idx * (1.0 / 999.0);       // np.linspace(0, 1, 1000)[idx]
  }

  // test hook (om_debug_jam_positions): moving trucks on road `rid`, as jam_sampling() counts them
  __device__ __forceinline__ uint32_t debug_trucks_on_road(int rid) {
    uint32_t n = 0;
    for (int i = 0; i < c.T; ++i) {
      int cur = U8A(t_cur, i), tgt = U8A(t_tgt, i);
      if (U8A(t_status, i) == ST_MOVING && cur != tgt && road_id(cur, tgt) == rid) n++;
    }
    return n;
  }

  __device__ __forceinline__ double debug_coverage(int i, int rid) {
    int cur = U8A(t_cur, i), tgt = U8A(t_tgt, i);
    const bool on = U8A(t_status, i) == ST_MOVING && cur != tgt && road_id(cur, tgt) == rid;
    return on ? (double)coverage(i, rid) : -1.0;
  }

  // Truck.get_route_coverage (truck.py:496-520)
  __device__ __forceinline__ float coverage(int i, int rid) {
    double total = __ldg(&c.trip60[__ldg(&c.sclass[i]) * c.R + rid]) + F64(t_prejam, i) + repair_get(i);
    double cov = ddiv0(now - F64(t_jstart, i), total);
    if (cov > 1.0) cov = 1.0;
    return (float)cov;
  }

  // ---------------------------------------------------------------------------------------------
  // main loop: pop the (time, key)-minimum entry until `until` (env.run(until), mine.py:391) and run the
  // popped continuation.
  //
  // Exact hop fusion.  SimPy needs 16 truck hops + 2 Release hops per truck cycle, but most of them
  // are forced: (1) an URGENT Initialize scheduled at t > 0 is always the very next pop (no other URGENT
  // entry can be pending), so MOVE_BEGIN / LOAD_BEGIN / DUMP_BEGIN run inline; (2) when the
  // popped entry is the ONLY pending entry with time == now (trucks, Releases and the monitor block
  // included), every same-time successor it creates is again the unique minimum, so the whole chain
  //   ARRIVE -> RET -> (grant) -> GRANTED -> BEGIN      and
  //   DONE -> RET -> Release/order -> MOVE_BEGIN -> Release pops -> grant head -> GRANTED -> BEGIN
  // runs back to back in SimPy's order without returning to the reduction.  With ties present every hop
  // goes through the reduction exactly as SimPy's heap would order it.  `seq` advances identically in
  // both modes, so the state is the same whichever path ran.
  // ---------------------------------------------------------------------------------------------
  // Code layout: the kernel is instruction-fetch bound (profiles/r1_d: ~1 IPC per SM whatever the occupancy), so the
  // continuations are laid out in the order the two hot chains execute them and fall through from one to the
  // next (no dispatch loop): shovels and dumpers share the GRANTED / BEGIN / DONE / RET code, cold paths (external
  // pause, errors, maintenance) are marked unlikely.  (No __noinline__ member functions: they would force the
  // register-resident Sim scalars onto the local-memory stack.)
  __device__ __forceinline__ void pause_for_order(int i, int kind, int rel_u, uint32_t rel_key, int& pend_truck, int& pend_kind) {
    set_entry(i, om_inf(), OM_KEY_SENTINEL, PH_AWAIT_ORDER);
    pend_truck = i; pend_kind = kind;
    if (rel_u >= 0) file_release(rel_u, rel_key);
  }
  // `site` < 0: the dispatcher itself raised (-OM_IERR_*); otherwise truck.py indexes the site list out of range
  __device__ __forceinline__ void bad_order(int i, int kind, int site, int rel_u, uint32_t rel_key) {
    if (site >= 0) count_order(kind, site);   // the wrapper logs the order before truck.py indexes the site list
    uint32_t code = (site < 0) ? (uint32_t)(-site) : (uint32_t)OM_IERR_BAD_ORDER;
    sched_truck(i, now, OM_KEY_NORMAL, (uint8_t)(PH_CRASH_BASE + code));   // the failed Process event
    if (rel_u >= 0) file_release(rel_u, rel_key);
  }

  // CARRY kernels, a move continued through the event loop: what apply_order() left in registers comes from the record
  __device__ __forceinline__ void load_move_carry(int i) {
    if (RNG == OM_RNG_PHILOX) rngw = U32A(t_rng, i);
    distw = F64(t_dist, i);
  }

  __device__ __forceinline__ void run(double until) {
    const int T = c.T, U = c.U, L = c.L, D = c.D, S = c.S;
    int pend_truck = -1, pend_kind = -1;
    for (;;) {
      // ---- next pending entry.  Times are >= 0, so their bit patterns order as unsigned integers.
      uint32_t bhi = 0x7FF00000u, blo = 0u, bk = OM_KEY_SENTINEL;
      int bid = -1;
      bool dup = false;
      const int nslot = (T + 31) >> 5;
      if (!SPLIT) {
        // every lane scans its truck entries (one LDS.128 each; 3 for north_pit)
        const uint4* ev = (const uint4*)(sm + ly.t_time) + lane;
        #pragma unroll 1
        for (int i = lane; i < T; i += 32, ev += 32) {
          const uint4 e = *ev;                         // {time lo, time hi, key, phase}
          bool teq = (e.y == bhi) && (e.x == blo);
          bool tlt = (e.y < bhi) || (e.y == bhi && e.x < blo);
          if (teq) dup = true;
          if (tlt) dup = false;
          if (tlt || (teq && e.z < bk)) { bhi = e.y; blo = e.x; bk = e.z; bid = i; }
        }
      } else {
        // big mines (the SPLIT kernels): truck entries are cached per SLOT of 32 trucks -- {min time, its smallest key, that truck, "several
        // entries share the slot's min time"} in shared memory, recomputed (one LDS.128 per lane + 4 redux + 1 ballot)
        // only for the slots an entry of which changed since the last pop (one or two) instead of scanning all T
        // entries: -10 % on the 500-truck mine.  Compiled only into the SPLIT kernels: on north_pit (3 slots) it saves
        // nothing, and merely having the code in the hot loop cost 13 % (instruction-cache layout).
        uint4* slots = (uint4*)(sm + ly.slot_cache);
        uint32_t d = (nslot > 32) ? 0xFFFFFFFFu : (dirty & ((nslot == 32) ? 0xFFFFFFFFu : ((1u << nslot) - 1u)));
        dirty = 0u;
        #pragma unroll 1
        while (d) {
          const int b = __ffs(d) - 1;
          d &= d - 1;
          #pragma unroll 1
          for (int sl = b; sl < nslot; sl += 32) {       // more than one trip only for T > 1024
            const int idx = (sl << 5) + lane;
            uint4 e;
            e.x = 0u; e.y = 0x7FF00000u; e.z = OM_KEY_SENTINEL; e.w = 0u;
            if (idx < T) e = *((const uint4*)(sm + ly.t_time) + idx);
            uint32_t shi = __reduce_min_sync(OM_FULL, e.y);
            uint32_t slo = __reduce_min_sync(OM_FULL, (e.y == shi) ? e.x : 0xFFFFFFFFu);
            bool sc = (e.y == shi) && (e.x == slo);
            uint32_t sk = __reduce_min_sync(OM_FULL, sc ? e.z : OM_KEY_SENTINEL);
            uint32_t sid = __reduce_min_sync(OM_FULL, (sc && e.z == sk) ? (uint32_t)idx : 0x7FFFFFFFu);
            uint32_t multi = (__popc(__ballot_sync(OM_FULL, sc)) > 1) ? 0x80000000u : 0u;
            if (lane == 0) slots[sl] = make_uint4(slo, shi, sk, sid | multi);
          }
        }
        __syncwarp();
        #pragma unroll 1
        for (int sl = lane; sl < nslot; sl += 32) {
          const uint4 e = slots[sl];
          bool teq = (e.y == bhi) && (e.x == blo);
          bool tlt = (e.y < bhi) || (e.y == bhi && e.x < blo);
          if (teq) dup = true;
          if (tlt) dup = (e.w >> 31) != 0u;
          else if (teq && (e.w >> 31)) dup = true;
          if (tlt || (teq && e.z < bk)) { bhi = e.y; blo = e.x; bk = e.z; bid = (int)(e.w & 0x7FFFFFFFu); }
        }
      }
      if (n_rel != 0) {
        #pragma unroll 1
        for (int u = lane; u < U; u += 32) {
          // a pending Release was filed at the current clock and nothing earlier can be pending
          uint32_t kk = U32A(u_rel_key, u);
          if (kk == OM_KEY_SENTINEL) continue;
          uint2 tv;
          tv.x = (uint32_t)__double2loint(now); tv.y = (uint32_t)__double2hiint(now);
          bool teq = (tv.y == bhi) && (tv.x == blo);
          bool tlt = (tv.y < bhi) || (tv.y == bhi && tv.x < blo);
          if (teq) dup = true;
          if (tlt) dup = false;
          if (tlt || (teq && kk < bk)) { bhi = tv.y; blo = tv.x; bk = kk; bid = T + u; }
        }
      }
      uint32_t mhi = __reduce_min_sync(OM_FULL, bhi);
      uint32_t mlo = __reduce_min_sync(OM_FULL, (bhi == mhi) ? blo : 0xFFFFFFFFu);
      bool cand = (bhi == mhi) && (blo == mlo);
      uint32_t mk = __reduce_min_sync(OM_FULL, cand ? bk : OM_KEY_SENTINEL);
      unsigned cands = __ballot_sync(OM_FULL, cand);
      // keys are unique, so exactly one lane holds (tmin, mk); a tie inside the only candidate lane shows in its dup flag
      int id = (int)__reduce_min_sync(OM_FULL, (cand && bk == mk) ? (uint32_t)bid : 0x7FFFFFFFu);
      bool anydup = __any_sync(OM_FULL, cand && dup);
      double tmin = __hiloint2double((int)mhi, (int)mlo);
      bool is_tick = (tick_time < tmin) || (tick_time == tmin && tick_key < mk);
      bool unique = (__popc(cands) == 1) && !anydup && (tick_time != tmin);
      if (is_tick) tmin = tick_time;
      if (!(tmin < until) || (err != OM_IERR_NONE && err != OM_IERR_TRACE_OVERFLOW)) break;
      now = tmin;
      n_hops++;
      OM_CONV();
      if (__builtin_expect(is_tick, 0)) { tick_block(); OM_CONV(); continue; }   // 15 % of the pops: keep its 12 KB out of the main path

      const bool inl = now > 0.0;            // URGENT successors run inline
      const bool chain = unique && inl;      // same-time NORMAL successors run inline
      int i = id;
      int rel_u = -1;                        // Release created in this chain, not yet popped (kept in registers)
      ordw_ready = false;
      uint32_t rel_key = 0;
      int kind = 0, u = 0;
      uint32_t ph0 = 0;
      bool dead = false, dump = false;
      if (id >= T) {
        // Release event popped: Resource._trigger_put hands the unit to the queue head
        u = id - T;
        U32A(u_rel_key, u) = OM_KEY_SENTINEL;
        n_rel--;
        if (U16A(u_user, u) != OM_NONE16 || U16A(u_qlen, u) == 0) continue;
        i = q_pop(u);
        U16A(u_user, u) = (uint16_t)i;
        if (!chain) { sched_truck(i, now, OM_KEY_NORMAL, (u < S) ? PH_LOAD_GRANTED : PH_DUMP_GRANTED); continue; }
        seq++;
        goto L_GRANTED;
      }
      ph0 = U32A(t_phase, i);
      // The continuation is picked by a chain of compares, most frequent first (live RNG: four pops per truck cycle, two
      // arrivals and two service ends): as a `switch` this was two indirect branches (BRX) per pop.
      // OM_X_SWITCH_PHASE restores the switch for measurements.
#ifdef OM_X_SWITCH_PHASE
#define OM_PH_BEGIN switch (ph0) {
#define OM_PH1(a) case a:
#define OM_PH2(a, b) case a: case b:
#define OM_PH_DEFAULT default:
#define OM_PH_END }
#else
#define OM_PH_BEGIN {
#define OM_PH1(a) if (om_opaque(ph0) == (uint32_t)(a))
#define OM_PH2(a, b) if (om_opaque(ph0) == (uint32_t)(a) || om_opaque(ph0) == (uint32_t)(b))
#define OM_PH_DEFAULT
#define OM_PH_END }
#endif
      OM_PH_BEGIN
        OM_PH1(PH_MOVE_ARRIVE) goto L_ARRIVE;
        OM_PH2(PH_LOAD_DONE, PH_DUMP_DONE) goto L_SVC_DONE;
        OM_PH1(PH_MOVE_REPAIRED) {
          if (CARRY) load_move_carry(i);
          repair_clear(i);
          U8A(t_status, i) = ST_MOVING;
          goto L_DEPART;
        }
        OM_PH1(PH_MOVE_RET) goto L_MOVE_RET;
        // CARRY kernels: `u` (the truck's unit), `rngw` and `distw` are in registers on every fused path into the labels
        // below; only the entries through the event loop load them
        OM_PH2(PH_LOAD_GRANTED, PH_DUMP_GRANTED) { if (CARRY) u = U16A(t_unit, i); goto L_GRANTED; }
        OM_PH2(PH_LOAD_BEGIN, PH_DUMP_BEGIN) { if (CARRY) u = U16A(t_unit, i); goto L_BEGIN; }
        OM_PH2(PH_LOAD_RET, PH_DUMP_RET) {
          if (CARRY) { u = U16A(t_unit, i); if (RNG == OM_RNG_PHILOX) rngw = U32A(t_rng, i); }
          goto L_SVC_RET;
        }
        OM_PH1(PH_MOVE_BEGIN) { if (CARRY) load_move_carry(i); goto L_MOVE_BEGIN; }
        OM_PH1(PH_RUN_INIT) {                                          // truck.py:296-302
          U8A(t_cur, i) = 0;
          U8A(t_status, i) = ST_WAIT_INIT;
          kind = OM_ORDER_INIT;
          goto L_ORDER;
        }
        OM_PH1(PH_MAINT_DONE) {                                        // truck.py:372-379
          u = U16A(t_unit, i);
          emit(K_SHOVEL_MAINT_DONE, (uint32_t)i, (uint32_t)u, 0, 0.0, 0.0);
          F64(u_lastbrk, u) = now;
          goto L_RELEASE;
        }
        OM_PH_DEFAULT {
          if (ph0 >= PH_CRASH_BASE) fail(ph0 - PH_CRASH_BASE);           // the failed Process event pops: env.step() raises
          continue;
        }
      OM_PH_END

      // ================= chain 1: ARRIVE -> RET -> (grant) -> GRANTED -> BEGIN =================
    L_ARRIVE:                                                          // truck.py:211-218
      OM_CONV();
      rl_close(acc_move, i);
      if (SPLIT) road_unlink(i, road_id(U8A(t_cur, i), U8A(t_tgt, i)));
      U8A(t_cur, i) = U8A(t_tgt, i);
      F64(t_prejam, i) = 0.0;
      if (!chain) { sched_truck(i, now, OM_KEY_NORMAL, PH_MOVE_RET); continue; }
      seq++;                                                           // Process(move) event
    L_MOVE_RET: {                                                      // truck.py:324-339 / 409-419
      OM_CONV();
      int cur = U8A(t_cur, i);
      u = pick_unit(cur - 1);
      q_push(u, i);
      U16A(t_unit, i) = (uint16_t)u;
      bool me = false;
      if (U16A(u_user, u) == OM_NONE16) {                              // head-only _trigger_put
        int h = q_pop(u);
        U16A(u_user, u) = (uint16_t)h;
        me = (h == i);
        if (!(me && chain)) sched_truck(h, now, OM_KEY_NORMAL, (u < S) ? PH_LOAD_GRANTED : PH_DUMP_GRANTED);
      }
      if (!me) {
        set_entry(i, om_inf(), OM_KEY_SENTINEL, PH_WAITING);
      }
      int qi = me ? 0 : (int)U16A(u_qlen, u) - 1;
      if (cur <= L) {
        emit(K_WAIT_SHOVEL, (uint32_t)i, (uint32_t)u, (uint32_t)qi, 0.0, 0.0);
        if (RL) F64(t_since, i) = now;
        U8A(t_status, i) = ST_WAIT_SHOVEL;
      } else {
        emit(K_WAIT_DUMPER, (uint32_t)i, (uint32_t)(u - S), 0, 0.0, 0.0);
        U8A(t_status, i) = ST_WAIT_DUMPER;
      }
      if (!(me && chain)) continue;
      seq++;                                                           // Request.succeed()
    }
    L_GRANTED:                                                         // truck.py:342-350 / 422-429
      OM_CONV();
      if (!CARRY) u = U16A(t_unit, i);
      dump = u >= S;
      if (!dump) rl_close(acc_wait, i);
      if (RL) F64(t_since, i) = now;
      emit(dump ? K_GET_DUMPER : K_GET_SHOVEL, (uint32_t)i, (uint32_t)(dump ? u - S : u), 0, 0.0, 0.0);
      seq++;                                                           // env.process(load/unload): Initialize URGENT
      if (!inl) {
        set_entry(i, now, seq - 1, dump ? PH_DUMP_BEGIN : PH_LOAD_BEGIN);
        continue;
      }
    L_BEGIN: {                                                         // truck.py:221-228 / 238-242
      if (!CARRY) u = U16A(t_unit, i);
      dump = u >= S;
      U8A(t_status, i) = dump ? ST_UNLOADING : ST_LOADING;
      F64(u_last_start, u) = now;
      double dt = __ldg(&c.svc_time[__ldg(&c.ttype[i]) * U + u]);      // (cap/shovel_tons)*cycle | dumper cycle time
      sched_truck(i, now + dt, OM_KEY_NORMAL, dump ? PH_DUMP_DONE : PH_LOAD_DONE);
      continue;
    }

      // ====== chain 2: DONE -> RET -> release + order -> MOVE_BEGIN -> depart -> Release pops -> (GRANTED) ======
    L_SVC_DONE: {                                                      // truck.py:230-235 / 243-251
      OM_CONV();
      u = U16A(t_unit, i);
      dump = u >= S;
      double load;
      if (!dump) {
        double jit;
        if (RNG == OM_RNG_PHILOX) {
          // one call per load: x -> load jitter, y -> shovel-maintenance gap (read at LOAD_RET).  Odd lanes compute the
          // call of the haul order that follows with the same epoch instead (same counters as apply_order would use), so
          // a chain that runs on to that order in this pop does not pay a second Philox call.
          Philox4 r = px(U16A(t_epoch, i), (uint32_t)i, (lane & 1) ? OM_S_ROAD_GAP : OM_S_LOAD_JITTER);
          ordw.x = __shfl_sync(OM_FULL, r.x, 1);
          ordw.y = __shfl_sync(OM_FULL, r.y, 1);
          ordw_ready = chain;
          r.x = __shfl_sync(OM_FULL, r.x, 0);
          r.y = __shfl_sync(OM_FULL, r.y, 0);
          jit = -0.1 + 0.2 * (double)om_u01(r.x);
          tape_put(OM_S_LOAD_JITTER, jit);
          U32A(t_rng, i) = r.y;
          rngw = r.y;
        } else {
          jit = draw_jitter(U16A(t_epoch, i), (uint32_t)i);
        }
        load = __ldg(&c.cap[i]) * (1.0 + jit);
        F64(t_load, i) = load;
      } else {
        load = F64(t_load, i);
        F64(t_total, i) = F64(t_total, i) + load;
        uint32_t sc = (uint32_t)U16A(t_svc, i) + 1;
        U16A(t_svc, i) = (uint16_t)sc;
        F64(t_cycle, i) = ddiv(now - 0.0, (double)sc);              // first_order_time is always 0 (truck.py:304)
        F64(t_load, i) = 0.0;
      }
      F64(u_tons, u) = F64(u_tons, u) + load;
      U32A(u_svc, u) = U32A(u_svc, u) + 1;
      F64(u_last_done, u) = now;
      if (!chain) { sched_truck(i, now, OM_KEY_NORMAL, dump ? PH_DUMP_RET : PH_LOAD_RET); continue; }
      seq++;                                                           // Process(load/unload) event
    }
    L_SVC_RET:                                                         // truck.py:352-370 / 431-441
      if (!CARRY) u = U16A(t_unit, i);
      dump = u >= S;
      if (dump) {
        rl_close(acc_dumper, i);
      } else {
        rl_close(acc_shovel, i);
        if (RNG != OM_RNG_DET) {
          double gap = (RNG == OM_RNG_PHILOX) ? (double)om_exp1(CARRY ? rngw : U32A(t_rng, i)) * c.maint_mean
                                              : draw_exp(OM_S_MAINT_GAP, U16A(t_epoch, i), (uint32_t)i, c.maint_mean);
          if (RNG == OM_RNG_PHILOX) tape_put(OM_S_MAINT_GAP, gap);
          if (__builtin_expect(now >= F64(u_lastbrk, u) + gap, 0)) { shovel_maintenance(i, u); continue; }
        }
      }
    L_RELEASE:                                                         // leave `with req:` -> release; truck.py:383-389 / 436-441
      dump = u >= S;
      rel_key = release(u);
      if (chain) rel_u = u; else file_release(u, rel_key);
      U8A(t_status, i) = dump ? ST_WAIT_BACK : ST_WAIT_HAUL;
      kind = dump ? OM_ORDER_BACK : OM_ORDER_HAUL;
    L_ORDER: {                                                         // dispatcher.py:35-89 + truck.py:303-317,389-402,441-453
      // om_step (the RL kernels) drives EVERY instance by actions: the dispatcher id an instance was reset with plays no
      // part there, and the device dispatchers below are not compiled into those kernels' event loop -- the teacher's
      // decide() in rl_step() is their only copy (5936 -> 4952 SASS instructions; +2 % env steps per second)
      if (RL || __builtin_expect(disp == OM_DISP_EXTERNAL, 0)) { pause_for_order(i, kind, rel_u, rel_key, pend_truck, pend_kind); break; }
      OM_CONV();
      int site = decide(kind, i);
      OM_CONV();
      int nmax = (kind == OM_ORDER_HAUL) ? D : L;
      if (__builtin_expect(site < 0 || site >= nmax, 0)) { bad_order(i, kind, site, rel_u, rel_key); continue; }
      apply_order(i, kind, site);
      OM_CONV();
      seq++;                                                           // env.process(move): Initialize URGENT
      if (!inl) {
        set_entry(i, now, seq - 1, PH_MOVE_BEGIN);
        goto L_AFTER_DEPART;
      }
    }
    L_MOVE_BEGIN: {                                                    // truck.py:116-150
      F64(t_jstart, i) = now;
      dead = false;
      if (RNG != OM_RNG_DET) {
        double g = (RNG == OM_RNG_PHILOX) ? (double)om_exp1(CARRY ? rngw : U32A(t_rng, i)) * c.unrepairable_mean
                                          : draw_exp(OM_S_UNREPAIRABLE, U16A(t_epoch, i), (uint32_t)i, c.unrepairable_mean);
        if (RNG == OM_RNG_PHILOX) tape_put(OM_S_UNREPAIRABLE, g);
        dead = now >= g;                                               // check_vehicle_availability, truck.py:538-547
      }
      double rep = repair_get(i);
      if (!dead && rep != 0.0) { begin_repair(i, rep); goto L_AFTER_DEPART; }
    }
    L_DEPART:                                                          // truck.py:151-210
      OM_CONV();
      move_depart(i, dead);
    L_AFTER_DEPART:
      OM_CONV();
      // the Release issued earlier in this chain is now the unique minimum: pop it
      if (rel_u < 0) continue;
      u = rel_u;
      rel_u = -1;
      if (U16A(u_user, u) != OM_NONE16 || U16A(u_qlen, u) == 0) continue;
      i = q_pop(u);
      U16A(u_user, u) = (uint16_t)i;
      seq++;                                                           // Request.succeed()
      goto L_GRANTED;
    }
    InstHdr* h = (InstHdr*)(sm + ly.hdr);
    h->pending_truck = pend_truck;
    h->pending_kind = pend_kind;
  }

  // rare continuations
  __device__ __forceinline__ void shovel_maintenance(int i, int u) {       // truck.py:358-372
    U8A(u_repair, u) = 1;
    double dn = draw_normal(OM_S_MAINT_DUR, U16A(t_epoch, i), (uint32_t)i, c.maint_mu, c.maint_sigma);
    double dur = (0.0 > dn) ? 0.0 : dn;                                // max(x, 0)
    emit(K_SHOVEL_MAINT_START, (uint32_t)i, (uint32_t)u, 0, now + dur, 0.0);
    sched_truck(i, now + dur, OM_KEY_NORMAL, PH_MAINT_DONE);
  }
  __device__ __forceinline__ void begin_repair(int i, double rep) {    // truck.py:143-150
    n_brk++;
    emit(K_BREAKDOWN, (uint32_t)i, 0, 0, lastbrk_get(i), rep);
    U8A(t_status, i) = ST_REPAIRING;
    sched_truck(i, now + rep, OM_KEY_NORMAL, PH_MOVE_REPAIRED);
  }

  // ---------------------------------------------------------------------------------------------
  // RL step surface: dense reward (rl_dispatch.py:194-308), observation float32[obs_dim]
  // (rl_dispatch.py:309-509 + utils/feature_processing.py:3-75), teacher suggestion (:169-184)
  // ---------------------------------------------------------------------------------------------
  __device__ __forceinline__ double rl_reward(bool final_step, bool sparse = false) {
    const int T = c.T;
    if (sparse && !final_step) {          // the reward function is not evaluated: only the window restarts
      acc_wait = acc_shovel = acc_dumper = acc_move = acc_jam = 0.0;
      return 0.0;
    }
    // The four duration penalties are one weighted sum: -0.1 wait - 0.05 (shovel + dumper) - 0.05 jam - 0.01 move
    // (rl_dispatch.py:282-291).  Every lane accumulates its trucks' / jams' share already weighted, so ONE warp sum is
    // needed instead of five (the reward is pinned to the reference at 1e-9 relative; the order of these additions moves
    // it by ~1e-16).
    double pen = 0.0;
#pragma unroll 1
    for (int i = lane; i < T; i += 32) {           // intervals still open at this decision
      int ph = (int)U32A(t_phase, i), cur = U8A(t_cur, i);
      double st = F64(t_since, i);
      double lo = (st > rl_last) ? st : rl_last;
      double part = (lo < now) ? (now - lo) : 0.0;
      double wt = 0.0;
      if ((ph == PH_WAITING || ph == PH_LOAD_GRANTED) && cur >= 1 && cur <= c.L) wt = 0.1;
      else if (ph == PH_LOAD_BEGIN || ph == PH_LOAD_DONE || ph == PH_LOAD_RET || ph == PH_DUMP_BEGIN || ph == PH_DUMP_DONE ||
               ph == PH_DUMP_RET) wt = 0.05;
      else if (ph == PH_MOVE_ARRIVE) wt = 0.01;
      pen += wt * part;
    }
#pragma unroll 1
    for (uint32_t j = lane; j < jam_count; j += 32) {
      double st = jstart_get(j), en = F64(j_end, j);
      double lo = (st > rl_last) ? st : rl_last, hi = (en < now) ? en : now;
      if (lo < hi) pen += 0.05 * (hi - lo);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) pen += __shfl_xor_sync(OM_FULL, pen, o);
    pen += 0.1 * acc_wait + 0.05 * (acc_shovel + acc_dumper) + 0.05 * acc_jam + 0.01 * acc_move;
    double final_reward = 0.0;
    if (final_step && now != 0.0) {
      double ft = rl_last_prod - 10000.0;
      if (ft > 0.0) final_reward = 0.5 * ft;
    }
    double delta = reported_tons - rl_last_prod;   // mine.produce_tons is the minute-sampled value
    rl_last_prod = reported_tons;
    double reward = -pen + 5.0 * om_dlog(delta + 1.0) + final_reward;
    double dt = now - rl_last;
    if (dt > 0.0) reward += 0.1 * ddiv0(delta, dt + 1e-6);
    acc_wait = acc_shovel = acc_dumper = acc_move = acc_jam = 0.0;
    return reward;
  }

  // Observation of the pending decision (truck i, order kind).  Layout generalises the reference's
  // L = D = 5 vector: C = max(L, D) candidate slots (zero padded), NS = L + D, R roads.
  // Returns (through the reference parameters) the trucks on the candidate roads from the current location as
  // road_load() found them (lane q = candidate q), for the teacher to reuse when it looks at the same roads.
  __device__ __forceinline__ void rl_observation(int i, int kind, float* ob, bool write_static, int& pair_cnt, int& on_cnt,
                                                 double& capsum, double& est_site_out) {
    const int L = c.L, D = c.D, NS = c.NS, R = c.R;
    const int C = L > D ? L : D;
    const int o1 = 6, o2 = o1 + 1 + NS, o3 = o2 + 2, o4 = o3 + C, o5 = o4 + C, o6 = o5 + R, o7 = o6 + R,
              o8 = o7 + C, o9 = o8 + NS, o10 = o9 + NS, o11 = o10 + NS, o12 = o11 + NS;
    const int cur = U8A(t_cur, i);
    // all roads: active-jam count (lanes over the jam pool, shared-memory counters per road); the road distances never
    // change, so they are written only when the caller hands in a new observation buffer (OM_STEP_WRITE_STATIC)
    {
      uint32_t* rj = (uint32_t*)(sm + ly.scratch);   // scratch is reused by unit_est_all() right after
#pragma unroll 1
      for (int r = lane; r < R; r += 32) rj[r] = 0u;
      __syncwarp();
#pragma unroll 1
      for (uint32_t j = lane; j < jam_count; j += 32)
        if (F64(j_end, j) >= now) atomicAdd(&rj[U16A(j_road, j)], 1u);
      __syncwarp();
#pragma unroll 1
      for (int r = lane; r < R; r += 32) {
        if (write_static) ob[o5 + r] = __ldg(&c.obs_dist[r]);
        ob[o6 + r] = (float)rj[r];
      }
      __syncwarp();
    }
    // dispatcher.update_mine (dispatcher.py:24-33): both sides' wait estimates (one inlined copy, two trips), each side
    // reduced to estimated_queue_wait_time per site (lane = site, NS <= 32) before the other side reuses the scratch slots
    double est_site = 0.0;
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      unit_est_all(side == 0);
      if (lane < NS && ((lane < L) == (side == 0))) est_site = site_est(lane);
      __syncwarp();
    }
    est_site_out = est_site;
    if (lane == 0) {
      ob[0] = kind == OM_ORDER_INIT ? 1.f : 0.f; ob[1] = kind == OM_ORDER_HAUL ? 1.f : 0.f;
      ob[2] = kind == OM_ORDER_BACK ? 1.f : 0.f;
      ob[3] = (float)(now - rl_last);
      double time_now = now * c.inv_sim_time;
      ob[4] = (float)time_now; ob[5] = (float)(1.0 - time_now);
    }
#pragma unroll 1
    for (int q = lane; q < 1 + NS; q += 32) ob[o1 + q] = (q == cur) ? 1.f : 0.f;
    // candidate roads from the current location
    const int ncand = (kind == OM_ORDER_HAUL) ? D : L;
    const int tgt0 = (kind == OM_ORDER_HAUL) ? 1 + L : 1;
    road_load(cur, tgt0, ncand, pair_cnt, on_cnt, capsum);
    // Every log(x + 1) of the observation in ONE pass over "log slots" (a single copy of the fp64 log, one trip for
    // north_pit's 17 slots instead of six calls): slot 0 truck load, 1 truck cycle time, 2 .. 2+C candidate wait,
    // 2+C .. 2+C+NS site wait.  Lanes take their slot's argument from the lane that owns it by shuffle.
    __syncwarp();
    const int nlog = 2 + C + NS;
#pragma unroll 1
    for (int base = 0; base < nlog; base += 32) {
      const int q = base + lane;
      const int cq = q - 2, sq = q - 2 - C;                      // candidate / site index of this slot
      // operands owned by other lanes (all lanes take part in the shuffles)
      const int src_c = (cq >= 0 && cq < C) ? cq : 0, src_s = (sq >= 0 && sq < NS) ? sq : 0;
      const int site_of_c = tgt0 - 1 + src_c;                     // site of candidate src_c
      const double est_c = __shfl_sync(OM_FULL, est_site, site_of_c < NS ? site_of_c : 0);
      const double est_s = __shfl_sync(OM_FULL, est_site, src_s);
      const int pc_c = __shfl_sync(OM_FULL, pair_cnt, src_c), on_c = __shfl_sync(OM_FULL, on_cnt, src_c);
      const double cap_c = __shfl_sync(OM_FULL, capsum, src_c);
      double arg = 0.0;
      int dst = -1;
      if (q == 0) { arg = F64(t_load, i); dst = o2; }
      else if (q == 1) { arg = F64(t_cycle, i); dst = o2 + 1; }
      else if (cq < C) {
        dst = o7 + cq;
        float tt = 0.f, tc = 0.f;
        if (cq < ncand) {
          const int tgt = tgt0 + cq, site = tgt - 1;
          tt = __ldg(&c.obs_tt[road_id(cur, tgt)]);                      // (distance * 60) / 25
          tc = (float)((double)pc_c * c.inv_tn);
          double extra;
          if (kind == OM_ORDER_HAUL) extra = ((double)on_c * __ldg(&c.inv_units[site])) * __ldg(&c.dcycle[__ldg(&c.ubeg[site]) - c.S]);
          else extra = cap_c * __ldg(&c.inv_prod[cq]);
          arg = est_c + extra;
        } else {
          dst = -(o7 + cq) - 2;                                          // padded slot: 0, no log
        }
        ob[o3 + cq] = tt; ob[o4 + cq] = tc;
      } else if (sq < NS) { arg = est_s; dst = o8 + sq; }
      const double lg = om_dlog(arg + 1.0);
      if (dst >= 0) ob[dst] = (float)lg;
      else if (dst < -1) ob[-dst - 2] = 0.f;
    }
    // per-site blocks: minute-sampled values (their logs were taken by the monitor block)
#pragma unroll 1
    for (int s2 = lane; s2 < NS; s2 += 32) {
      ob[o9 + s2] = (float)((double)U16A(s_qtotal, s2) * c.inv_tn);
      ob[o10 + s2] = F32A(s_cap, s2);            // float32 already (the observation is all they feed)
      ob[o11 + s2] = F32A(s_ability, s2);
      ob[o12 + s2] = F32A(s_tons, s2);
    }
    __syncwarp();
  }

  // One env step of this instance (called by om_step_kernel after the state blob is staged).
  __device__ __forceinline__ void rl_step(const RunParams& p) {
    InstHdr* h = (InstHdr*)(sm + ly.hdr);
    const OmStepIO& io = p.io;
    if (h->finished) {
      if (lane == 0) {
        if (io.reward) io.reward[inst] = 0.0;
        if (io.done) io.done[inst] = 1;
        if (io.order_kind) io.order_kind[inst] = -1;
        if (io.truck) io.truck[inst] = -1;
        if (io.sug_action) io.sug_action[inst] = -1;
      }
      return;
    }
    int pt = h->pending_truck, pk = h->pending_kind;
    if (pt >= 0) {
      // the agent's action arrives: RLDispatcher._step returns, truck.py:303-317 / 390-402 / 442-453 continue
      int a = io.actions ? io.actions[inst] : 0;
      int nmax = (pk == OM_ORDER_HAUL) ? c.D : c.L;
      if (a < 0 || a >= nmax) {
        if (a >= 0) count_order(pk, a);
        sched_truck(pt, now, OM_KEY_NORMAL, (uint8_t)(PH_CRASH_BASE + OM_IERR_BAD_ORDER));
      } else {
        apply_order(pt, pk, a);
        set_entry(pt, now, seq, PH_MOVE_BEGIN);   // Initialize(move), URGENT
        seq++;
      }
      h->pending_truck = -1; h->pending_kind = -1;
    }
    bool fresh = false;       // true once this call has auto-reset the env: the finished episode's outputs stay
  L_RUN:
    run(p.until);
    pt = h->pending_truck; pk = h->pending_kind;
    const bool failed = (err != OM_IERR_NONE && err != OM_IERR_TRACE_OVERFLOW);
    float* ob = io.obs + (size_t)inst * p.obs_dim;
    if (pt >= 0 && !failed) {
      double delta_time = now - rl_last;
      int pair_cnt, on_cnt;
      double capsum;
#ifdef OM_STEP_NO_OBS   /* measurement aid: the DES part of an env step alone (outputs are meaningless) */
      pair_cnt = on_cnt = 0; capsum = 0.0;
      double est_all = 0.0;
      double reward = 0.0;
      int sug = -1;
      if (false) {
#else
      double est_all;
      rl_observation(pt, pk, ob, (io.flags & OM_STEP_WRITE_STATIC) != 0, pair_cnt, on_cnt, capsum, est_all);
      double reward = rl_reward(false, (io.flags & OM_STEP_SPARSE_REWARD) != 0);
      int sug = -1;
      if (io.teacher >= 0 && io.teacher != OM_DISP_RANDOM) {
#endif
        int saved_disp = disp;
        uint32_t saved_err = err;
        disp = io.teacher;
        sug = decide(pk, pt, true, true, on_cnt, capsum, est_all);   // rl_observation() computed every site's wait estimate
        if (sug < 0) sug = -1;                       // the teacher raised (e.g. SQ before its minute's sample)
        disp = saved_disp; err = saved_err;
      }
      rl_last = now;
      if (lane == 0) {
        if (!fresh) {
          if (io.reward) io.reward[inst] = reward;
          if (io.done) io.done[inst] = 0;
          if (io.info) { io.info[inst * 4 + 0] = reported_tons; io.info[inst * 4 + 1] = now; io.info[inst * 4 + 2] = delta_time; io.info[inst * 4 + 3] = 0.0; }
        }
        if (io.order_kind) io.order_kind[inst] = pk;
        if (io.truck) io.truck[inst] = pt;
        if (io.sug_action) io.sug_action[inst] = sug;
        h->rl_steps++;
      }
    } else {
      // Mine.start_rl after env.run: final reward over (last decision, until], done (mine.py:453-476)
      double delta_time = p.until - rl_last;
      now = p.until;
      double reward = rl_reward(true);
      rl_last = now;
      const bool again = (io.flags & OM_STEP_AUTO_RESET) && !fresh && !failed;
      if (lane == 0) {
        if (!fresh) {
          if (io.reward) io.reward[inst] = reward;
          if (io.done) io.done[inst] = 1;
          if (io.info) { io.info[inst * 4 + 0] = reported_tons; io.info[inst * 4 + 1] = now; io.info[inst * 4 + 2] = delta_time; io.info[inst * 4 + 3] = 0.0; }
        }
        if (io.order_kind) io.order_kind[inst] = -1;
        if (io.truck) io.truck[inst] = -1;
        if (io.sug_action) io.sug_action[inst] = -1;
        if (io.episode_totals) {
          io.episode_totals[inst * 3 + 0] += 1;
          io.episode_totals[inst * 3 + 1] += (long long)n_events;
          io.episode_totals[inst * 3 + 2] += (long long)n_decisions;
        }
        h->finished = again ? 0u : 1u;
        h->rl_steps++;
      }
      if (again) {
        // same-step automatic reset: keep the terminal observation, start the next episode of this env and hand out its
        // first decision (the outputs of the finished episode -- reward, done, info -- are already written)
        if (io.final_obs) {
          float* fo = io.final_obs + (size_t)inst * p.obs_dim;
#pragma unroll 1
          for (int q = lane; q < p.obs_dim; q += 32) fo[q] = ob[q];
        }
        const uint64_t seed2 = (((uint64_t)k1 << 32) | (uint64_t)k0) + io.seed_stride;
        __syncwarp();
        om_init_state(sm, cold, SPLIT ? ly.hot_bytes : ly.total_bytes, c, ly, lane, seed2, OM_DISP_EXTERNAL);
        load_hdr(p);
        fresh = true;
        goto L_RUN;
      }
    }
    __syncwarp();
  }

  // dumped tons at the stop time: sum() of the dumper tons in dumper order (same flavour as the monitors)
  __device__ __forceinline__ double dumped_tons() {
    PySum ps; ps.init();
    const bool neu = c.sum_mode != 0;
    #pragma unroll 1
    for (int u = c.S; u < c.U; ++u) ps.add(F64(u_tons, u), neu);
    return ps.result(neu);
  }
};

// Initial state of one instance at t=0 (Mine.__init__/Truck.__init__/registration order of mine.py:357-389):
// the monitor block's Initialize events are URGENT at 0 with the smallest eids, then one URGENT
// Initialize(run) per truck in index order.
// `sm` addresses the first `sm_bytes` of the blob (the whole blob for the reset kernel, the staged part inside the step
// kernel), `cold` the blob in HBM for everything behind it (== sm when nothing is behind).
__device__ __forceinline__ void om_init_state(unsigned char* sm, unsigned char* cold, uint32_t sm_bytes, const DevCfg& c,
                                              const OmStateLayout& ly, int lane, uint64_t seed, int dispatcher) {
  const double inf = om_inf();
  (void)inf;
#ifdef OM_BOUNDS_CHECK
  uint32_t dbg_viol = 0;
#endif
  const bool SPLIT = true, LY_DYNAMIC = true;   // accessors: follow the layout (ly.t_loc != 0 <=> the handle runs the SPLIT kernels)
  // zero everything first
  #pragma unroll 1
  for (uint32_t o = lane * 8; o < sm_bytes; o += 256) *(unsigned long long*)(sm + o) = 0ull;
  #pragma unroll 1
  for (uint32_t o = sm_bytes + lane * 8; o < ly.total_bytes; o += 256) *(unsigned long long*)(cold + o) = 0ull;
  __syncwarp();
  #pragma unroll 1
  for (int i = lane; i < c.T; i += 32) {
    F64(t_time, i) = 0.0;
    U32A(t_key, i) = (uint32_t)(1 + i);          // URGENT, after the monitors
    U32A(t_phase, i) = PH_RUN_INIT;
    U8A(t_status, i) = ST_IDLE;
    U8A(t_cur, i) = OM_NOLOC;
    U8A(t_tgt, i) = OM_NOLOC;
    U16A(t_next, i) = OM_NONE16;
    U16A(t_unit, i) = OM_NONE16;
  }
  if (ly.r_head != 0) {                          // split layouts: empty per-road lists of the trucks under way
    #pragma unroll 1
    for (int r = lane; r < c.R; r += 32) ((uint16_t*)(sm + ly.r_head))[r] = OM_NONE16;
  }
  #pragma unroll 1
  for (int u = lane; u < c.U; u += 32) {
    U32A(u_rel_key, u) = OM_KEY_SENTINEL;
    U16A(u_user, u) = OM_NONE16;
    U16A(u_head, u) = OM_NONE16;
    U16A(u_tail, u) = OM_NONE16;
  }
  if (lane == 0) {
    InstHdr* h = (InstHdr*)(sm + ly.hdr);
    h->now = 0.0;
    h->tick_time = 0.0;
    h->tick_key = 0u;                            // URGENT, eid 0
    h->tick_k = -1;
    h->seq = (uint32_t)(1 + c.T);
    h->seed = seed;
    h->dispatcher = dispatcher;
    h->pending_truck = -1;
    h->pending_kind = -1;
  }
#ifdef OM_BOUNDS_CHECK
  om_publish_violation(dbg_viol);
#endif
  __syncwarp();
}
