// Measurement helper (NOT part of the product path): hardware counters of the kernels launched inside one user range,
// collected in-process with the CUPTI range profiler (user range, user replay: the caller repeats the work once per
// hardware pass).  bench.py uses it AFTER its timed region to read smsp__inst_executed / dram__bytes of the
// episode kernel for the roofline line on the very box, binary and workload it has just timed -- a number taken under
// the profiler is never a timing.  C ABI so that it loads with ctypes; every failure is an error code + message (the
// caller falls back to the committed ncu capture and says so).
//
// Build: g++ -shared -fPIC -I$CUDA/include om_counters.cpp -L$CUDA/lib64 -lcupti -lcuda   (openmines_b200/_lib.py)
#include <cuda.h>
#include <cupti_profiler_host.h>
#include <cupti_profiler_target.h>
#include <cupti_range_profiler.h>
#include <cupti_target.h>

#include <string>
#include <vector>

namespace {
std::string g_err;
std::vector<std::string> g_metric_names;
std::vector<const char*> g_metrics;
std::vector<uint8_t> g_config, g_counter_data, g_avail;
CUpti_Profiler_Host_Object* g_host = nullptr;
CUpti_RangeProfiler_Object* g_range = nullptr;
bool g_started = false;
int g_passes = 0;

int fail(const char* what, CUptiResult r) {
  const char* s = nullptr;
  cuptiGetResultString(r, &s);
  g_err = std::string(what) + ": " + (s ? s : "unknown CUPTI error");
  return -1;
}
#define CU(call)                                             \
  do {                                                       \
    CUptiResult r_ = (call);                                 \
    if (r_ != CUPTI_SUCCESS) return fail(#call, r_);         \
  } while (0)

void teardown() {
  if (g_range) {
    CUpti_RangeProfiler_Disable_Params p{CUpti_RangeProfiler_Disable_Params_STRUCT_SIZE};
    p.pRangeProfilerObject = g_range;
    cuptiRangeProfilerDisable(&p);
    g_range = nullptr;
  }
  if (g_host) {
    CUpti_Profiler_Host_Deinitialize_Params p{CUpti_Profiler_Host_Deinitialize_Params_STRUCT_SIZE};
    p.pHostObject = g_host;
    cuptiProfilerHostDeinitialize(&p);
    g_host = nullptr;
  }
  g_started = false;
}
}  // namespace

extern "C" {

const char* omc_last_error() { return g_err.c_str(); }

// Set up the collection of `metrics` over ONE user range on `device`'s current context.  The caller then repeats
//   omc_pass_begin();  <launch the work>;  <synchronise>;  omc_pass_end(&done);
// until done != 0 (user replay: one repetition per hardware pass the metric set needs; the caller restores whatever the
// work consumes -- bench.py re-runs om_reset outside the range), and reads the values with omc_finish().
int omc_setup(int device, const char* const* metrics, int n_metrics) {
  if (g_range) { g_err = "omc_setup: already set up"; return -1; }
  if (!metrics || n_metrics <= 0) { g_err = "omc_setup: bad arguments"; return -1; }
  CUcontext ctx = nullptr;
  if (cuCtxGetCurrent(&ctx) != CUDA_SUCCESS || ctx == nullptr) { g_err = "omc_setup: no current CUDA context"; return -1; }
  g_metric_names.assign(metrics, metrics + n_metrics);
  g_metrics.clear();
  for (auto& m : g_metric_names) g_metrics.push_back(m.c_str());

  CUpti_Profiler_Initialize_Params init{CUpti_Profiler_Initialize_Params_STRUCT_SIZE};
  CU(cuptiProfilerInitialize(&init));
  CUpti_Profiler_DeviceSupported_Params sup{CUpti_Profiler_DeviceSupported_Params_STRUCT_SIZE};
  sup.cuDevice = device;
  sup.api = CUPTI_PROFILER_RANGE_PROFILING;
  CU(cuptiProfilerDeviceSupported(&sup));
  if (sup.isSupported != CUPTI_PROFILER_CONFIGURATION_SUPPORTED) { g_err = "omc_setup: range profiling is not supported on this device / configuration"; return -1; }
  CUpti_Device_GetChipName_Params chip{CUpti_Device_GetChipName_Params_STRUCT_SIZE};
  chip.deviceIndex = (size_t)device;
  CU(cuptiDeviceGetChipName(&chip));
  CUpti_Profiler_GetCounterAvailability_Params av{CUpti_Profiler_GetCounterAvailability_Params_STRUCT_SIZE};
  av.ctx = ctx;
  CU(cuptiProfilerGetCounterAvailability(&av));
  g_avail.assign(av.counterAvailabilityImageSize, 0);
  av.pCounterAvailabilityImage = g_avail.data();
  CU(cuptiProfilerGetCounterAvailability(&av));
  CUpti_Profiler_Host_Initialize_Params hi{CUpti_Profiler_Host_Initialize_Params_STRUCT_SIZE};
  hi.profilerType = CUPTI_PROFILER_TYPE_RANGE_PROFILER;
  hi.pChipName = chip.pChipName;
  hi.pCounterAvailabilityImage = g_avail.data();
  CU(cuptiProfilerHostInitialize(&hi));
  g_host = hi.pHostObject;
#define CUT(call)                                                          \
  do {                                                                     \
    CUptiResult r_ = (call);                                               \
    if (r_ != CUPTI_SUCCESS) { teardown(); return fail(#call, r_); }       \
  } while (0)
  CUpti_Profiler_Host_ConfigAddMetrics_Params add{CUpti_Profiler_Host_ConfigAddMetrics_Params_STRUCT_SIZE};
  add.pHostObject = g_host;
  add.ppMetricNames = g_metrics.data();
  add.numMetrics = g_metrics.size();
  CUT(cuptiProfilerHostConfigAddMetrics(&add));
  CUpti_Profiler_Host_GetConfigImageSize_Params cs{CUpti_Profiler_Host_GetConfigImageSize_Params_STRUCT_SIZE};
  cs.pHostObject = g_host;
  CUT(cuptiProfilerHostGetConfigImageSize(&cs));
  g_config.assign(cs.configImageSize, 0);
  CUpti_Profiler_Host_GetConfigImage_Params ci{CUpti_Profiler_Host_GetConfigImage_Params_STRUCT_SIZE};
  ci.pHostObject = g_host;
  ci.pConfigImage = g_config.data();
  ci.configImageSize = g_config.size();
  CUT(cuptiProfilerHostGetConfigImage(&ci));
  CUpti_Profiler_Host_GetNumOfPasses_Params np{CUpti_Profiler_Host_GetNumOfPasses_Params_STRUCT_SIZE};
  np.pConfigImage = g_config.data();
  np.configImageSize = g_config.size();
  CUT(cuptiProfilerHostGetNumOfPasses(&np));
  g_passes = (int)np.numOfPasses;
  CUpti_RangeProfiler_Enable_Params en{CUpti_RangeProfiler_Enable_Params_STRUCT_SIZE};
  en.ctx = ctx;
  CUT(cuptiRangeProfilerEnable(&en));
  g_range = en.pRangeProfilerObject;
  CUpti_RangeProfiler_GetCounterDataSize_Params ds{CUpti_RangeProfiler_GetCounterDataSize_Params_STRUCT_SIZE};
  ds.pRangeProfilerObject = g_range;
  ds.pMetricNames = g_metrics.data();
  ds.numMetrics = g_metrics.size();
  ds.maxNumOfRanges = 4;
  ds.maxNumRangeTreeNodes = 4;
  CUT(cuptiRangeProfilerGetCounterDataSize(&ds));
  g_counter_data.assign(ds.counterDataSize, 0);
  CUpti_RangeProfiler_CounterDataImage_Initialize_Params di{CUpti_RangeProfiler_CounterDataImage_Initialize_Params_STRUCT_SIZE};
  di.pRangeProfilerObject = g_range;
  di.pCounterData = g_counter_data.data();
  di.counterDataSize = g_counter_data.size();
  CUT(cuptiRangeProfilerCounterDataImageInitialize(&di));
  CUpti_RangeProfiler_SetConfig_Params sc{CUpti_RangeProfiler_SetConfig_Params_STRUCT_SIZE};
  sc.pRangeProfilerObject = g_range;
  sc.pConfig = g_config.data();
  sc.configSize = g_config.size();
  sc.pCounterDataImage = g_counter_data.data();
  sc.counterDataImageSize = g_counter_data.size();
  sc.maxRangesPerPass = 4;
  sc.numNestingLevels = 1;
  sc.minNestingLevel = 1;
  sc.passIndex = 0;
  sc.targetNestingLevel = 1;
  sc.range = CUPTI_UserRange;
  sc.replayMode = CUPTI_UserReplay;
  CUT(cuptiRangeProfilerSetConfig(&sc));
  return 0;
}

int omc_num_passes() { return g_passes; }

int omc_pass_begin() {
  if (!g_range) { g_err = "omc_pass_begin: not set up"; return -1; }
  CUpti_RangeProfiler_Start_Params st{CUpti_RangeProfiler_Start_Params_STRUCT_SIZE};
  st.pRangeProfilerObject = g_range;
  CUT(cuptiRangeProfilerStart(&st));
  CUpti_RangeProfiler_PushRange_Params pu{CUpti_RangeProfiler_PushRange_Params_STRUCT_SIZE};
  pu.pRangeProfilerObject = g_range;
  pu.pRangeName = "om";
  CUT(cuptiRangeProfilerPushRange(&pu));
  return 0;
}

int omc_pass_end(int* all_passes_done) {
  if (!g_range) { g_err = "omc_pass_end: not set up"; return -1; }
  CUpti_RangeProfiler_PopRange_Params po{CUpti_RangeProfiler_PopRange_Params_STRUCT_SIZE};
  po.pRangeProfilerObject = g_range;
  CUT(cuptiRangeProfilerPopRange(&po));
  CUpti_RangeProfiler_Stop_Params sp{CUpti_RangeProfiler_Stop_Params_STRUCT_SIZE};
  sp.pRangeProfilerObject = g_range;
  CUT(cuptiRangeProfilerStop(&sp));
  if (all_passes_done) *all_passes_done = sp.isAllPassSubmitted ? 1 : 0;
  return 0;
}

// values[m] = metric m over the range; tears the profiler down.
int omc_finish(double* values, int* n_ranges) {
  if (!g_range) { g_err = "omc_finish: not set up"; return -1; }
  CUpti_RangeProfiler_DecodeData_Params dd{CUpti_RangeProfiler_DecodeData_Params_STRUCT_SIZE};
  dd.pRangeProfilerObject = g_range;
  CUT(cuptiRangeProfilerDecodeData(&dd));
  CUpti_RangeProfiler_GetCounterDataInfo_Params info{CUpti_RangeProfiler_GetCounterDataInfo_Params_STRUCT_SIZE};
  info.pCounterDataImage = g_counter_data.data();
  info.counterDataImageSize = g_counter_data.size();
  CUT(cuptiRangeProfilerGetCounterDataInfo(&info));
  if (n_ranges) *n_ranges = (int)info.numTotalRanges;
  if (info.numTotalRanges < 1) { teardown(); g_err = "omc_finish: no range was collected"; return -1; }
  CUpti_Profiler_Host_EvaluateToGpuValues_Params ev{CUpti_Profiler_Host_EvaluateToGpuValues_Params_STRUCT_SIZE};
  ev.pHostObject = g_host;
  ev.pCounterDataImage = g_counter_data.data();
  ev.counterDataImageSize = g_counter_data.size();
  ev.ppMetricNames = g_metrics.data();
  ev.numMetrics = g_metrics.size();
  ev.rangeIndex = 0;
  ev.pMetricValues = values;
  CUT(cuptiProfilerHostEvaluateToGpuValues(&ev));
  teardown();
  return 0;
}

void omc_abort() { teardown(); }

}  // extern "C"
