/* The C ABI of libopenmines_b200.so used from plain C: no Python, no PyTorch -- the CUDA runtime for device memory only.
 *
 * Packs the smallest mine of the parity suite by hand (tests/golden/configs/one_truck_mine.json: one 40 t truck, one
 * shovel, one dumper, 120 simulated minutes; the arithmetic of the tables is the reference's, cited per field in
 * include/openmines_b200.h) and runs N independent instances of it under NaiveDispatcher with stochastic events off.
 * Every instance must report what the unmodified reference does for this mine (tests/test_gpu_parity.py compares the
 * printed numbers with the golden fixture).
 *
 *   gcc -O2 -Iinclude -I/usr/local/cuda/include examples/c_api_demo.c -o /tmp/c_api_demo \
 *       -Lopenmines_b200 -lopenmines_b200 -L/usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/openmines_b200 -Wl,-rpath,/usr/local/cuda/lib64
 *   /tmp/c_api_demo 4
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "openmines_b200.h"

#define CHECK_CUDA(call)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                         \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

#define CHECK_OM(h, call)                                                                 \
  do {                                                                                    \
    int rc_ = (call);                                                                     \
    if (rc_ != OM_OK) {                                                                   \
      fprintf(stderr, "%s: error %d: %s\n", #call, rc_, om_last_error(h));                \
      return 3;                                                                           \
    }                                                                                     \
  } while (0)

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4;
  const double sim_time = 120.0;

  /* ---- the mine, packed the way run_dispatch_sim builds its objects (cli/run.py:48-107) ---- */
  const double cap[1] = {40.0}, speed[1] = {20.0};
  const int32_t ttype[1] = {0};
  const int32_t unit_begin[3] = {0, 1, 2};             /* load site 0: shovel 0; dump site 0: dumper 1 */
  const double shovel_tons = 8.0, shovel_cycle = 1.25;
  const double dumper_cycle[1] = {2.0};
  const double roads[3] = {1.5, 4.0, 4.5};             /* charging->load, load->dump, dump->load */
  const double load_time[1] = {(cap[0] / shovel_tons) * shovel_cycle};           /* truck.py:223 */
  const double shovel_rate[1] = {shovel_tons / shovel_cycle};                    /* load_site.py:182-185 */
  const double site_prod[1] = {shovel_rate[0]};
  const double sptf_cost[1] = {cap[0] / site_prod[0]};                           /* sptf_dispatcher.py:35 */
  const double jam_chance[2] = {0.0, 1.0 - exp(-0.05 * 1.0)};                    /* road.py:164-167 */
  const int32_t zero_i[1] = {0};

  OmConfig cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = OM_ABI_VERSION;
  cfg.n_trucks = 1; cfg.n_load_sites = 1; cfg.n_dump_sites = 1; cfg.n_shovels = 1; cfg.n_dumpers = 1;
  cfg.n_truck_types = 1;
  cfg.jam_capacity = 64;
  cfg.sum_mode = 1;
  cfg.nearest_init = 0;
  cfg.truck_capacity = cap; cfg.truck_speed = speed; cfg.truck_type = ttype;
  cfg.site_unit_begin = unit_begin; cfg.dumper_cycle = dumper_cycle; cfg.road_distance = roads;
  cfg.load_time = load_time; cfg.sptf_cost = sptf_cost; cfg.jam_chance = jam_chance;
  cfg.nearest_haul = zero_i; cfg.nearest_back = zero_i; cfg.fixed_group = zero_i;
  cfg.shovel_rate = shovel_rate; cfg.site_productivity = site_prod;
  cfg.road_lambda = 1.0 / 120.0; cfg.road_mu = 60.0; cfg.road_sigma = 10.0;      /* road.py:37-43 */
  cfg.detour_mu = 0.4; cfg.detour_sigma = 0.1;
  cfg.breakdown_mean = 360.0; cfg.repair_mu = 10.0; cfg.repair_sigma = 3.0;      /* truck.py:66-69 */
  cfg.unrepairable_mean = 70560.0;
  cfg.maint_mean = 1440.0; cfg.maint_mu = 45.0; cfg.maint_sigma = 5.0;           /* truck.py:358-362 */
  cfg.jam_sd = 0.1; cfg.jam_scale = 10.0;                                        /* road.py:185,212 */

  /* ---- engine ---- */
  OmHandle h = NULL;
  CHECK_OM(NULL, om_create(&cfg, n, 0, &h));
  size_t state_bytes = 0;
  CHECK_OM(h, om_state_bytes(h, &state_bytes));
  int32_t grid, block, smem, warps;
  CHECK_OM(h, om_launch_info(h, &grid, &block, &smem, &warps));
  printf("instances %d  state %zu B/instance  launch grid %d x block %d, %d B shared memory, %d warps per CTA\n", n,
         state_bytes, grid, block, smem, warps);

  uint64_t* h_seeds = (uint64_t*)malloc(sizeof(uint64_t) * n);
  int32_t* h_disp = (int32_t*)malloc(sizeof(int32_t) * n);
  for (int i = 0; i < n; ++i) { h_seeds[i] = (uint64_t)i; h_disp[i] = OM_DISP_NAIVE; }
  void *d_state, *d_seeds, *d_disp, *d_summary, *d_ticks;
  const int T = (int)sim_time;
  CHECK_CUDA(cudaMalloc(&d_state, state_bytes * n));
  CHECK_CUDA(cudaMalloc(&d_seeds, sizeof(uint64_t) * n));
  CHECK_CUDA(cudaMalloc(&d_disp, sizeof(int32_t) * n));
  CHECK_CUDA(cudaMalloc(&d_summary, sizeof(OmSummary) * n));
  CHECK_CUDA(cudaMalloc(&d_ticks, sizeof(OmTickRecord) * n * T));
  CHECK_CUDA(cudaMemcpy(d_seeds, h_seeds, sizeof(uint64_t) * n, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(d_disp, h_disp, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemset(d_ticks, 0, sizeof(OmTickRecord) * n * T));

  OmOutputs outs;
  memset(&outs, 0, sizeof outs);
  outs.summary = (OmSummary*)d_summary;
  outs.ticks = (OmTickRecord*)d_ticks;
  outs.tick_capacity = T;

  CHECK_OM(h, om_reset(h, (const uint64_t*)d_seeds, (const int32_t*)d_disp, OM_RNG_DET, NULL, NULL, NULL, d_state, NULL));
  CHECK_OM(h, om_run(h, sim_time, d_state, &outs, NULL));
  CHECK_CUDA(cudaDeviceSynchronize());

  OmSummary* s = (OmSummary*)malloc(sizeof(OmSummary) * n);
  OmTickRecord* ticks = (OmTickRecord*)malloc(sizeof(OmTickRecord) * n * T);
  CHECK_CUDA(cudaMemcpy(s, d_summary, sizeof(OmSummary) * n, cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(ticks, d_ticks, sizeof(OmTickRecord) * n * T, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    printf("instance %d: reported_tons %.17g dumped_tons %.17g truck_events %u decisions %u ticks %u err %u "
           "tons@60min %.17g\n", i, s[i].reported_tons, s[i].dumped_tons, s[i].n_truck_events, s[i].n_decisions,
           s[i].n_ticks, s[i].err, ticks[(size_t)i * T + 60].produced_tons);

  CHECK_OM(h, om_destroy(h));
  cudaFree(d_state); cudaFree(d_seeds); cudaFree(d_disp); cudaFree(d_summary); cudaFree(d_ticks);
  free(h_seeds); free(h_disp); free(s); free(ticks);
  return 0;
}
