/*
 * openmines_b200 -- C ABI of the batched, B200-native discrete-event engine for OpenMines.
 *
 * This is the drop-in boundary for ONE hot path of 370025263/openmines: stepping many
 * independent mine instances (seed x dispatcher x config).  Each entry point below names the
 * reference interface (file:line, relative to the reference tree) that it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative OM_E_* code on failure; the message is
 *     available from om_last_error(handle) (or om_last_error(NULL) for om_create failures);
 *   - no exception crosses the boundary; the caller (PyTorch on the host side) owns every device
 *     buffer, the library owns only the opaque handle and its packed copy of the mine config;
 *   - all pointers marked "device" are device pointers on the handle's GPU; all work is enqueued
 *     on the cudaStream_t passed in (as void*, 0 = legacy default stream); calls are not
 *     thread-safe per handle;
 *   - there is NO CPU fallback: without a CUDA device om_create fails with OM_E_CUDA.
 */
#ifndef OPENMINES_B200_H
#define OPENMINES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OM_ABI_VERSION 9

/* OmConfig.flags */
#define OM_CFG_RL_SURFACE 1 /* allocate the fields om_step needs (reward windows, minute-sampled site values): bigger
                               state blobs, fewer resident instances per SM; om_step fails on a handle created without it */

/* error codes */
#define OM_OK 0
#define OM_E_INVALID (-1)
#define OM_E_CUDA (-2)
#define OM_E_NOMEM (-3)
#define OM_E_UNSUPPORTED (-4)

/* dispatcher ids: the six built-ins of openmines/src/dispatch_algorithms/ as device functions
 * behind BaseDispatcher.give_init_order / give_haul_order / give_back_order (dispatcher.py:91-98) */
#define OM_DISP_NAIVE 0       /* naive_dispatcher.py:12-22 */
#define OM_DISP_RANDOM 1      /* random_dispatcher.py:12-30 */
#define OM_DISP_NEAREST 2     /* nearest_dispatcher.py:17-67 */
#define OM_DISP_FIXEDGROUP 3  /* fixed_group_dispatcher.py:21-110 */
#define OM_DISP_SPTF 4        /* sptf_dispatcher.py:21-61 */
#define OM_DISP_SQ 5          /* sq_dispatcher.py:18-100 */
#define OM_DISP_SHORTESTTRIP 6 /* shortest_trip_dispatcher.py:12-84 (default teacher of the gym env) */
#define OM_DISP_EXTERNAL 255  /* decisions supplied by the host through om_step (rl_dispatch.py:58-123) */

/* RNG modes */
#define OM_RNG_DET 0     /* no stochastic event can fire; RandomDispatcher uses keyed Philox */
#define OM_RNG_PHILOX 1  /* counter-based Philox4x32-10, key=(seed), counter=(epoch, entity, site, 0) */
#define OM_RNG_TAPE 2    /* replay recorded draws (post-transform values) in call order */

/* order kinds */
#define OM_ORDER_INIT 0
#define OM_ORDER_HAUL 1
#define OM_ORDER_BACK 2

/* per-instance error codes (OmSummary.err) -- the reference would raise at these points */
#define OM_IERR_NONE 0
#define OM_IERR_SQ_KEYERROR 1    /* sq_dispatcher.py:31,60,88: order at an integer minute before its monitor block */
#define OM_IERR_BAD_ORDER 2      /* IndexError at truck.py:305,390,442 */
#define OM_IERR_JAM_OVERFLOW 3   /* engine limit: more simultaneously active jams than jam_capacity */
#define OM_IERR_TAPE 4           /* tape exhausted or site-id mismatch */
#define OM_IERR_TRACE_OVERFLOW 5 /* trace buffer too small (simulation itself is unaffected) */

/* Packed mine configuration: what run_dispatch_sim (cli/run.py:48-107) builds from the mine JSON
 * (conf/north_pit_mine.json), flattened into structure-of-arrays.  All arrays are HOST pointers and
 * are copied by om_create.  Locations are coded 0 = charging site, 1..L = load sites,
 * L+1..L+D = dump sites.  Units are shovels 0..S-1 (grouped by load site) then dumpers S..S+M-1
 * (grouped by dump site).  Roads are numbered in the monitor's order (mine.py:138-154):
 * [0,L) charging->load i; L+i*D+j load i->dump j; L+L*D+j*L+i dump j->load i. */
typedef struct OmConfig {
  int32_t abi_version;      /* OM_ABI_VERSION */
  int32_t n_trucks;         /* T */
  int32_t n_load_sites;     /* L */
  int32_t n_dump_sites;     /* D */
  int32_t n_shovels;        /* S */
  int32_t n_dumpers;        /* M */
  int32_t n_truck_types;    /* distinct (capacity) classes used by the tables below */
  int32_t jam_capacity;     /* max simultaneously tracked jam events per instance */
  int32_t sum_mode;         /* 1 = CPython>=3.12 Neumaier sum() (default), 0 = naive left-to-right */
  int32_t nearest_init;     /* nearest_dispatcher.py:24-26 */
  int32_t flags;            /* OM_CFG_* */
  int32_t reserved1;
  const double* truck_capacity;   /* [T]  truck.py:56 */
  const double* truck_speed;      /* [T]  truck.py:57 */
  const int32_t* truck_type;      /* [T]  index into load_time / sptf_cost */
  const int32_t* site_unit_begin; /* [L+D+1] first unit of each site (load sites then dump sites) */
  const double* dumper_cycle;     /* [M]  dump_site.py:15 */
  const double* road_distance;    /* [R]  road.py:311-323 (d2l indexed [load][dump], road.py:314-316) */
  const double* load_time;        /* [n_types*S] (cap/shovel_tons)*cycle, truck.py:223 */
  const double* sptf_cost;        /* [n_types*L] cap/load_site_productivity, sptf_dispatcher.py:35 */
  const double* jam_chance;       /* [T+1] 1-exp(-0.05*n), road.py:164-167 */
  const int32_t* nearest_haul;    /* [L]  argmin l2d[i,:], nearest_dispatcher.py:42-45 */
  const int32_t* nearest_back;    /* [D]  argmin d2l[:,j], nearest_dispatcher.py:64-66 */
  const int32_t* fixed_group;     /* [T]  fixed_group_dispatcher.py:21-56 (0 when unassigned) */
  const double* shovel_rate;      /* [S]  shovel_tons / cycle_time, load_site.py:182-185 */
  const double* site_productivity; /* [L] sum of shovel_rate per load site, load_site.py:182-184 */
  /* stochastic-event parameters (reference defaults are hard-coded where cited) */
  double road_lambda;        /* road.py:38 / config road_event_params */
  double road_mu;            /* road.py:39 */
  double road_sigma;         /* road.py:40 */
  double detour_mu;          /* road.py:41 */
  double detour_sigma;       /* road.py:42 */
  double breakdown_mean;     /* truck.py:66  (360) */
  double repair_mu;          /* truck.py:68  (10) */
  double repair_sigma;       /* truck.py:69  (3) */
  double unrepairable_mean;  /* truck.py:67  (70560) */
  double maint_mean;         /* truck.py:358 (1440) */
  double maint_mu;           /* truck.py:362 (45) */
  double maint_sigma;        /* truck.py:362 (5) */
  double jam_sd;             /* road.py:185  (0.1) */
  double jam_scale;          /* road.py:212  (10, Weibull shape 2) */
  double sim_time;           /* the mine JSON's sim_time (cli/run.py:111): instances of a MIXED-config batch stop at
                                min(until, sim_time) of their own config; <= 0: no limit of its own.  One-config handles
                                ignore it (om_run's `until` rules). */
} OmConfig;

/* One record per entry the reference appends to an EventPool during env.run (truck.py:143,188,
 * 203,311,332,346,395,414,425,447; road.py:218,259; truck.py:364,373).  Kinds 0..11 are the
 * "truck events" of the headline metric. */
typedef struct OmTraceRecord {
  double time;    /* env.now when the record was created */
  double a, b;    /* kind-specific payload, see openmines_b200/trace.py */
  uint16_t truck; /* 0xFFFF for road events */
  uint8_t kind;
  uint8_t flags;
  uint16_t arg0, arg1;
} OmTraceRecord;

/* Per-minute KPI sample written by the monitor block (mine.py:102-119); 32 bytes. */
typedef struct OmTickRecord {
  double produced_tons;      /* mine.produce_tons */
  int32_t service_count;
  uint16_t working, waiting, load_unload, moving, repairing;        /* truck status counts */
  uint16_t n_jams, n_road_repairs, n_truck_repairs, n_unrepairable; /* cumulative event counts */
  uint16_t pad;
} OmTickRecord;

/* Per-minute sample of one shovel / dumper (Shovel.monitor_status load_site.py:130-141, Dumper.monitor_status
 * dump_site.py:25-34, ParkingLot.monitor_resource load_site.py:88-104); 16 bytes. */
typedef struct OmUnitTick {
  double tons;
  uint32_t service_count;
  uint16_t queue_length;
  uint8_t repair;
  uint8_t pad;
} OmUnitTick;

/* Per-instance end-of-run summary; 64 bytes. */
typedef struct OmSummary {
  double reported_tons; /* mine.status[last tick]["produced_tons"] (the reported production) */
  double dumped_tons;   /* sum of dumper tons at the stop time */
  double now;           /* simulation clock when the call returned */
  uint32_t n_truck_events;
  uint32_t n_decisions;      /* dispatch decisions taken */
  uint32_t n_wrapper_orders; /* dispatcher.total_order_count (SPTF back orders count twice) */
  uint32_t n_hops;           /* scheduler pops executed (truck/unit/tick entries) */
  uint32_t n_jams, n_road_repairs, n_truck_repairs, n_unrepairable;
  uint32_t n_ticks;
  uint32_t err; /* OM_IERR_* */
} OmSummary;

/* Output buffers of om_run; any pointer may be NULL to skip that output.  All device pointers. */
typedef struct OmOutputs {
  OmSummary* summary;         /* [N] */
  OmTickRecord* ticks;        /* [N * tick_capacity], sample k of instance n at n*tick_capacity+k */
  int32_t tick_capacity;
  int32_t trace_capacity;     /* records per instance */
  OmTraceRecord* trace;       /* [N * trace_capacity] */
  uint16_t* site_queue_ticks; /* [N * tick_capacity * (L+D)] minute-sampled queue totals (load_site.py:81-83) */
  double* site_tons_ticks;    /* [N * tick_capacity * (L+D)] minute-sampled site tons (load_site.py:175, dump_site.py:69) */
  int16_t* orders;            /* [N * order_capacity] (kind<<12 | site) per decision, in order; NULL to skip */
  int32_t order_capacity;
  int32_t reserved;
  OmUnitTick* unit_ticks;     /* [N * tick_capacity * (S+M)] per-minute shovel/dumper samples; NULL to skip */
  /* Draw-tape export (OM_RNG_PHILOX only; NULL to skip): every random value the reference would have drawn, post-transform,
   * in the reference's call order with its site id (OM_DRAW_*) -- replaying this tape through the reference
   * (oracle/ref_harness.py RngControl "replay") or its CPU restatement reproduces the instance's trajectory bit for bit.
   * The per-minute truck-breakdown block reserves two slots per truck (gap, repair time); a slot whose draw the reference
   * would not have made carries site OM_DRAW_SKIP and must be dropped by the reader.  Overflow sets
   * OM_IERR_TRACE_OVERFLOW (the simulation itself is unaffected). */
  double* draw_values;        /* [N * draw_capacity] */
  uint8_t* draw_sites;        /* [N * draw_capacity] */
  uint32_t* draw_count;       /* [N] slots used (including OM_DRAW_SKIP ones) */
  int64_t draw_capacity;      /* slots per instance */
} OmOutputs;

/* draw-site ids of the tape formats (OM_RNG_TAPE input, draw-tape export): one per random call site of the reference */
#define OM_DRAW_LOAD_JITTER 0   /* truck.py:230  np.random.uniform(-0.1, 0.1) */
#define OM_DRAW_MAINT_GAP 1     /* truck.py:358  np.random.exponential(1440) */
#define OM_DRAW_MAINT_DUR 2     /* truck.py:362  np.random.normal(45, 5) */
#define OM_DRAW_BREAK_GAP 3     /* truck.py:528  np.random.exponential(360) */
#define OM_DRAW_BREAK_DUR 4     /* truck.py:532  np.random.normal(10, 3) */
#define OM_DRAW_UNREPAIRABLE 5  /* truck.py:545  random.expovariate(1/70560) */
#define OM_DRAW_ROAD_GAP 6      /* road.py:253   random.expovariate(lambda_repair) */
#define OM_DRAW_ROAD_DUR 7      /* road.py:256   np.random.normal(mu, sigma) */
#define OM_DRAW_DETOUR 8        /* road.py:329   np.random.normal(0.4, 0.1) */
#define OM_DRAW_JAM_COIN 9      /* road.py:179   random.random() */
#define OM_DRAW_JAM_POS 10      /* road.py:203   np.random.choice(1000, p) -> the chosen grid value */
#define OM_DRAW_JAM_DUR 11      /* road.py:212   np.random.weibull(2) */
#define OM_DRAW_DISPATCH 12     /* random_dispatcher.py:16,22,29  RandomState.randint */
#define OM_DRAW_SKIP 255

/* Byte offsets of the per-instance state arrays inside one state blob (om_state_bytes).  Truck
 * arrays have n_trucks elements, unit arrays n_shovels+n_dumpers, road arrays R. */
typedef struct OmStateLayout {
  /* Byte offsets inside one instance's state blob.  Every field is the offset of ITS ELEMENT 0; element i lives at
   * offset + i * stride, where stride = e_stride for t_time / t_key / t_phase (the truck's pending entry), t_stride for
   * the other t_* fields except t_since (per-truck record), u_stride for the u_* fields except u_rel_key / u_lastbrk
   * (per-unit record), and the element size otherwise.  t_phase is a uint32.  t_since, j_start, s_tons, s_cap and
   * s_ability exist only with OM_CFG_RL_SURFACE (offset 0 otherwise); the header is hdr_bytes long (128 without the
   * RL fields, 192 with them).  A pending Release has no stored time: it always equals the instance clock. */
  uint32_t total_bytes;
  uint32_t hdr;
  uint32_t t_time, t_load, t_jstart, t_prejam, t_lastbrk, t_repair, t_total, t_cycle, t_dist, t_since; /* f64 */
  uint32_t t_key, t_rng;                                                                     /* u32 */
  uint32_t t_svc, t_epoch, t_next, t_unit;                                                   /* u16 */
  uint32_t t_phase, t_status, t_cur, t_tgt;                                                  /* u8  */
  uint32_t u_last_start, u_last_done, u_tons, u_lastbrk;                                     /* f64 */
  uint32_t u_rel_key, u_svc;                                                                 /* u32 */
  uint32_t u_user, u_head, u_tail, u_qlen;                                                   /* u16 */
  uint32_t u_repair;                                                                         /* u8  */
  uint32_t r_repair_end;                                                                     /* f64 */
  uint32_t j_pos, j_start, j_end;                                                            /* f64 [jam_capacity]; j_start: uint16 minute */
  uint32_t j_road;                                                                           /* u16 [jam_capacity] */
  uint32_t s_qtotal, s_coming;                                                               /* u16 [L+D] */
  uint32_t s_tons, s_cap, s_ability;                                                         /* f32 [L+D] minute-sampled: log(tons + 1), log(capacity + 1), ability ratio */
  uint32_t scratch;                                                                          /* f64 [max(S, M, R/2 + 1, 32)] kernel scratch */
  uint32_t t_rec, u_rec;       /* start of the truck / unit record arrays */
  uint32_t t_stride, u_stride; /* record sizes in bytes */
  uint32_t e_stride;           /* pending-entry size in bytes */
  uint32_t hot_bytes;          /* prefix of the blob the kernels stage in shared memory (== total_bytes unless split) */
  uint32_t hdr_bytes;          /* bytes of the per-instance header present in the blob */
  uint32_t slot_cache;         /* split layouts only (0 otherwise): kernel-internal cache (16 B per 32 trucks); contents are meaningless between calls */
  uint32_t t_aux, r_jamcnt;    /* split layouts only (0 otherwise): t_aux {unit, next, epoch, service count} u16[4] per truck in the hot
                                  part -- t_unit / t_next / t_epoch / t_svc then point here (stride 8); r_jamcnt u16 [R] stored jams per road */
  uint32_t t_loc, t_lastbrk_hot; /* split layouts only (0 otherwise): {status, cur, tgt, flags} u8[4] per truck (flags bit 0: a repair
                                  time is pending) and the MINUTE of the last breakdown as uint16 per truck, both in the hot part;
                                  t_status / t_cur / t_tgt / t_lastbrk then point here (strides 4 / 2; t_lastbrk is NOT an f64 then) */
  uint32_t r_on, r_head, t_rnext, t_rprev; /* split layouts only (0 otherwise), all u16: moving trucks per road [R], head of the
                                  road's list of them [R] (0xFFFF: empty), next / previous truck on the same road [T] */
} OmStateLayout;

typedef struct OmEngine* OmHandle;

/* Replaces building Mine + sites + road (cli/run.py:48-107).  Copies the config to the device. */
int om_create(const OmConfig* cfg, int64_t n_instances, int device, OmHandle* out);
/* Mixed-config batch ("seed x dispatcher x CONFIG"): n_configs packed mines and, per instance, the index of the one it
 * simulates (config_id: HOST array [n_instances]).  This is what the reference's fleet-size and scene sweeps
 * (cli/run.py:283-472: ten fleet sizes x dispatchers [x scenes], one run_dispatch_sim each) become: ONE batch.  All
 * instances share one state layout, sized for the largest config in every dimension (om_state_layout / om_state_bytes);
 * instance n uses the first n_trucks / n_shovels ... entries of each array according to ITS config.  Per-site and per-unit
 * tick outputs have rows of max(L+D) / max(S+M) entries.  om_step is not available on such a handle. */
int om_create_multi(const OmConfig* const* cfgs, int32_t n_configs, const int32_t* config_id, int64_t n_instances,
                    int device, OmHandle* out);
int om_destroy(OmHandle h);
const char* om_last_error(OmHandle h);

/* ABI self-description for FFI bindings: writes up to n of {OM_ABI_VERSION, sizeof OmConfig, OmOutputs, OmStepIO,
 * OmStateLayout, OmSummary, OmTickRecord, OmTraceRecord, OmUnitTick, per-instance header, stats rows, stats columns} and
 * returns how many values exist.  A binding compares them with its own mirrors before the first call (_lib.py). */
int om_abi_info(int32_t* out, int32_t n);

/* Size / layout of one instance's state blob; the caller allocates n_instances * bytes, 16-byte aligned (the kernels
 * move blobs with TMA bulk copies). */
int om_state_bytes(OmHandle h, size_t* per_instance);
int om_state_layout(OmHandle h, OmStateLayout* out);

/* Replaces Mine.__init__ + Truck.__init__ + registration of the generator processes
 * (mine.py:30-48, 357-389): writes the t=0 state of every instance.
 *   seeds       device [N]  Philox key per instance
 *   dispatcher  device [N]  OM_DISP_* per instance
 *   tape_values/tape_sites/tape_offsets: device, only for OM_RNG_TAPE (offsets has N+1 entries); the handle keeps the
 *     pointers, so the buffers must stay valid until the last om_run / om_step of this episode batch has completed.
 * Also computes the work-queue order of om_run (instances grouped by dispatcher, most expensive group first). */
int om_reset(OmHandle h, const uint64_t* seeds, const int32_t* dispatcher, int rng_mode,
             const double* tape_values, const uint8_t* tape_sites, const int64_t* tape_offsets,
             void* state, void* stream);

/* Replaces env.run(until) (mine.py:391): advances every instance until no pending entry has
 * time < until (the reference's stop event is URGENT at `until`).  May be called repeatedly
 * with increasing `until`.  Handles whose layout is split (big mines: om_state_layout().t_aux != 0) hold breakdown
 * minutes as uint16 and return OM_E_UNSUPPORTED for until > 65535 (45 simulated days); om_step has the same limit on
 * every handle (jam start minutes of the RL surface). */
int om_run(OmHandle h, double until, void* state, const OmOutputs* outs, void* stream);

/* Local partial statistics for the cross-GPU all-reduce (issued by the caller with NCCL):
 * stats_out device [7*16] doubles, one row per device dispatcher id (0..6), every entry an INTEGER so that sums are exact
 * and order-independent (bit-identical totals on 1 or 8 GPUs over the same global instance ids):
 *   0 count, 1/2 reported tons in units of 2^-24 t (value >> 25, value & (2^25-1)), 3 truck events, 4 decisions, 5 jams,
 *   6 road repairs, 7 truck repairs, 8 unrepairable, 9 instances with an engine error, 10 scheduler hops,
 *   11/12 dumped tons (as 1/2), 13/14 sum of (tons rounded to 2^-8 t)^2 in units of 2^-16 t^2 (value >> 32, value & (2^32-1)).
 * Decoder: openmines_b200.parallel.describe. */
int om_stats_reduce(OmHandle h, const OmSummary* summary, const int32_t* dispatcher,
                    double* stats_out, void* stream);

#define OM_STEP_WRITE_STATIC 4 /* (re)write the observation columns that never change (the road distances); om_step sets it
                                  by itself whenever it is handed an observation buffer it has not written before */
#define OM_STEP_AUTO_RESET 1
#define OM_STEP_SPARSE_REWARD 2 /* reward_mode="sparse" (rl_dispatch.py:186-192): 0 at every decision; the final step carries the
                                   dense reward of Mine.start_rl (mine.py:456-459) with last_production never updated */

/* Batched gym-style step (replaces RLDispatcher._step + Mine.start_rl + preprocess_observation:
 * rl_dispatch.py:58-123,194-509, mine.py:397-476, utils/feature_processing.py:3-75).  Under om_step EVERY instance
 * pauses at every dispatch decision and is driven by `actions` -- reset them with OM_DISP_EXTERNAL; an instance reset
 * with a device dispatcher id is treated the same (the device dispatchers run in om_run and, as the teacher, in
 * om_step's sug_action).  One call = apply `actions[n]` to instance n's pending
 * order (ignored when none is pending: first call after om_reset), advance to the next decision point (or to
 * `until`), and emit that decision's observation / the reward of the previous decision.  All device pointers;
 * obs rows are written only when a new decision is pending (on the final step the reference re-sends the
 * previous observation, so the row is left untouched). */
typedef struct OmStepIO {
  const int32_t* actions; /* [N] in */
  float* obs;             /* [N * obs_dim] */
  double* reward;         /* [N] dense reward (rl_dispatch.py:194-308), final bonus on the last step */
  uint8_t* done;          /* [N] 1 on the final step and afterwards */
  int32_t* order_kind;    /* [N] OM_ORDER_* of the pending decision, -1 when done */
  int32_t* truck;         /* [N] truck asking, -1 when done */
  int32_t* sug_action;    /* [N] teacher's decision (info["sug_action"]), -1 when done / no teacher */
  double* info;           /* [N * 4] produce_tons, time, delta_time, reserved */
  int32_t teacher;        /* OM_DISP_* of the teacher (0..6), or -1 */
  int32_t flags;          /* OM_STEP_* */
  /* per-env automatic reset (OM_STEP_AUTO_RESET), the behaviour of gym.vector.AsyncVectorEnv that the reference's
   * trainers run under (test/cleanrl/ppo_single_net.py:593): on the step an episode ends the env reports the final
   * reward and done = 1, then starts a fresh episode with seed += seed_stride in the same call and returns THAT
   * episode's first observation / pending decision; the terminal observation goes to final_obs. */
  float* final_obs;            /* [N * obs_dim] or NULL: rows are written on the step an episode ends */
  uint64_t seed_stride;        /* added to an instance's seed at every automatic reset */
  int64_t* episode_totals;     /* [N * 3] or NULL: finished episodes, their truck events, their decisions (accumulated) */
} OmStepIO;
int om_obs_dim(OmHandle h, int32_t* dim);
int om_step(OmHandle h, double until, void* state, const OmStepIO* io, void* stream);

/* Test hooks: the engine's own draw transforms and its jam-position sampler, for statistical tests of the marginal laws.
 * om_debug_draws: out[i] = transform(Philox(counter i, key seed)) on the device; kind 0 uniform (0,1], 1 exponential(1),
 *   2 standard normal, 3 Weibull(shape 2), 4 load jitter, 5 randint [0, arg), 6 jam coin.  out: device [n].
 * om_debug_jam_positions: n_samples draws of the position of a new jam on `road` for instance `inst` as its state blob
 *   stands (Road.road_jam_sampling's np.random.choice over the 1000-point grid, road.py:185-206); *fallbacks = how often the
 *   rejection sampler gave up.  out: device [n_samples], fallbacks: device int. */
/* om_debug_bounds_report: returns 1 when the library was built with -DOM_BOUNDS_CHECK (every state access checks its index,
 * and the warp-uniform continuations check that the warp arrives converged), 0 otherwise; out8 = {index violations so far,
 * source line of the first, its index, the extent, convergence violations, source line of the first, its active mask, 0}. */
int om_debug_bounds_report(int32_t* out8);
int om_debug_draws(int kind, uint64_t seed, uint32_t arg, int64_t n, double* out, void* stream);
int om_debug_jam_positions(OmHandle h, void* state, int64_t inst, int32_t road, int32_t n_samples, double* out,
                           int32_t* fallbacks, double* coverage_out /* device [T] or NULL: Truck.get_route_coverage
                           (truck.py:496-520) of every truck under way on the road, -1 for the others */, void* stream);

/* Kernel launch geometry actually used by om_run (for bench.py's launch accounting). */
int om_launch_info(OmHandle h, int32_t* grid, int32_t* block, int32_t* smem_bytes, int32_t* warps_per_cta);

#ifdef __cplusplus
}
#endif
#endif /* OPENMINES_B200_H */
