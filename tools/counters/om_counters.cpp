// Measurement helper (NOT part of the product path): hardware counters of the kernels launched between omc_begin() and
// omc_end(), collected in-process with the CUPTI range profiler (auto ranges = one per kernel launch, kernel replay for
// multi-pass metric sets).  bench.py uses it AFTER its timed region to read smsp__inst_executed / dram__bytes of the
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

// Start collecting `metrics` for every kernel launched on `device`'s current context until omc_end().
int omc_begin(int device, const char* const* metrics, int n_metrics, int max_ranges) {
  if (g_started) { g_err = "omc_begin: already collecting"; return -1; }
  if (!metrics || n_metrics <= 0 || max_ranges <= 0) { g_err = "omc_begin: bad arguments"; return -1; }
  CUcontext ctx = nullptr;
  if (cuCtxGetCurrent(&ctx) != CUDA_SUCCESS || ctx == nullptr) { g_err = "omc_begin: no current CUDA context"; return -1; }
  g_metric_names.assign(metrics, metrics + n_metrics);
  g_metrics.clear();
  for (auto& m : g_metric_names) g_metrics.push_back(m.c_str());

  CUpti_Profiler_Initialize_Params init{CUpti_Profiler_Initialize_Params_STRUCT_SIZE};
  CU(cuptiProfilerInitialize(&init));
  CUpti_Profiler_DeviceSupported_Params sup{CUpti_Profiler_DeviceSupported_Params_STRUCT_SIZE};
  sup.cuDevice = device;
  sup.api = CUPTI_PROFILER_RANGE_PROFILING;
  CU(cuptiProfilerDeviceSupported(&sup));
  if (sup.isSupported != CUPTI_PROFILER_CONFIGURATION_SUPPORTED) { g_err = "omc_begin: range profiling is not supported on this device / configuration"; return -1; }
  CUpti_Device_GetChipName_Params chip{CUpti_Device_GetChipName_Params_STRUCT_SIZE};
  chip.deviceIndex = (size_t)device;
  CU(cuptiDeviceGetChipName(&chip));
  // counter availability of this context
  CUpti_Profiler_GetCounterAvailability_Params av{CUpti_Profiler_GetCounterAvailability_Params_STRUCT_SIZE};
  av.ctx = ctx;
  CU(cuptiProfilerGetCounterAvailability(&av));
  g_avail.assign(av.counterAvailabilityImageSize, 0);
  av.pCounterAvailabilityImage = g_avail.data();
  CU(cuptiProfilerGetCounterAvailability(&av));
  // host object + config image
  CUpti_Profiler_Host_Initialize_Params hi{CUpti_Profiler_Host_Initialize_Params_STRUCT_SIZE};
  hi.profilerType = CUPTI_PROFILER_TYPE_RANGE_PROFILER;
  hi.pChipName = chip.pChipName;
  hi.pCounterAvailabilityImage = g_avail.data();
  CU(cuptiProfilerHostInitialize(&hi));
  g_host = hi.pHostObject;
  CUpti_Profiler_Host_ConfigAddMetrics_Params add{CUpti_Profiler_Host_ConfigAddMetrics_Params_STRUCT_SIZE};
  add.pHostObject = g_host;
  add.ppMetricNames = g_metrics.data();
  add.numMetrics = g_metrics.size();
  if (CUptiResult r = cuptiProfilerHostConfigAddMetrics(&add); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiProfilerHostConfigAddMetrics", r); }
  CUpti_Profiler_Host_GetConfigImageSize_Params cs{CUpti_Profiler_Host_GetConfigImageSize_Params_STRUCT_SIZE};
  cs.pHostObject = g_host;
  if (CUptiResult r = cuptiProfilerHostGetConfigImageSize(&cs); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiProfilerHostGetConfigImageSize", r); }
  g_config.assign(cs.configImageSize, 0);
  CUpti_Profiler_Host_GetConfigImage_Params ci{CUpti_Profiler_Host_GetConfigImage_Params_STRUCT_SIZE};
  ci.pHostObject = g_host;
  ci.pConfigImage = g_config.data();
  ci.configImageSize = g_config.size();
  if (CUptiResult r = cuptiProfilerHostGetConfigImage(&ci); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiProfilerHostGetConfigImage", r); }
  // range profiler on the context
  CUpti_RangeProfiler_Enable_Params en{CUpti_RangeProfiler_Enable_Params_STRUCT_SIZE};
  en.ctx = ctx;
  if (CUptiResult r = cuptiRangeProfilerEnable(&en); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerEnable", r); }
  g_range = en.pRangeProfilerObject;
  CUpti_RangeProfiler_GetCounterDataSize_Params ds{CUpti_RangeProfiler_GetCounterDataSize_Params_STRUCT_SIZE};
  ds.pRangeProfilerObject = g_range;
  ds.pMetricNames = g_metrics.data();
  ds.numMetrics = g_metrics.size();
  ds.maxNumOfRanges = (size_t)max_ranges;
  ds.maxNumRangeTreeNodes = (size_t)max_ranges;
  if (CUptiResult r = cuptiRangeProfilerGetCounterDataSize(&ds); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerGetCounterDataSize", r); }
  g_counter_data.assign(ds.counterDataSize, 0);
  CUpti_RangeProfiler_CounterDataImage_Initialize_Params di{CUpti_RangeProfiler_CounterDataImage_Initialize_Params_STRUCT_SIZE};
  di.pRangeProfilerObject = g_range;
  di.pCounterData = g_counter_data.data();
  di.counterDataSize = g_counter_data.size();
  if (CUptiResult r = cuptiRangeProfilerCounterDataImageInitialize(&di); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerCounterDataImageInitialize", r); }
  CUpti_RangeProfiler_SetConfig_Params sc{CUpti_RangeProfiler_SetConfig_Params_STRUCT_SIZE};
  sc.pRangeProfilerObject = g_range;
  sc.pConfig = g_config.data();
  sc.configSize = g_config.size();
  sc.pCounterDataImage = g_counter_data.data();
  sc.counterDataImageSize = g_counter_data.size();
  sc.maxRangesPerPass = (size_t)max_ranges;
  sc.numNestingLevels = 1;
  sc.minNestingLevel = 1;
  sc.passIndex = 0;
  sc.targetNestingLevel = 1;
  sc.range = CUPTI_AutoRange;
  sc.replayMode = CUPTI_KernelReplay;
  if (CUptiResult r = cuptiRangeProfilerSetConfig(&sc); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerSetConfig", r); }
  CUpti_RangeProfiler_Start_Params st{CUpti_RangeProfiler_Start_Params_STRUCT_SIZE};
  st.pRangeProfilerObject = g_range;
  if (CUptiResult r = cuptiRangeProfilerStart(&st); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerStart", r); }
  g_started = true;
  return 0;
}

// Stop collecting; values[r * n_metrics + m] = metric m of the r-th kernel launched since omc_begin (launch order).
int omc_end(double* values, int max_ranges, int* n_ranges) {
  if (!g_started) { g_err = "omc_end: not collecting"; return -1; }
  CUpti_RangeProfiler_Stop_Params sp{CUpti_RangeProfiler_Stop_Params_STRUCT_SIZE};
  sp.pRangeProfilerObject = g_range;
  if (CUptiResult r = cuptiRangeProfilerStop(&sp); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerStop", r); }
  CUpti_RangeProfiler_DecodeData_Params dd{CUpti_RangeProfiler_DecodeData_Params_STRUCT_SIZE};
  dd.pRangeProfilerObject = g_range;
  if (CUptiResult r = cuptiRangeProfilerDecodeData(&dd); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerDecodeData", r); }
  CUpti_RangeProfiler_GetCounterDataInfo_Params info{CUpti_RangeProfiler_GetCounterDataInfo_Params_STRUCT_SIZE};
  info.pCounterDataImage = g_counter_data.data();
  info.counterDataImageSize = g_counter_data.size();
  if (CUptiResult r = cuptiRangeProfilerGetCounterDataInfo(&info); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiRangeProfilerGetCounterDataInfo", r); }
  int nr = (int)info.numTotalRanges;
  if (n_ranges) *n_ranges = nr;
  const size_t nm = g_metrics.size();
  for (int ri = 0; ri < nr && ri < max_ranges; ++ri) {
    CUpti_Profiler_Host_EvaluateToGpuValues_Params ev{CUpti_Profiler_Host_EvaluateToGpuValues_Params_STRUCT_SIZE};
    ev.pHostObject = g_host;
    ev.pCounterDataImage = g_counter_data.data();
    ev.counterDataImageSize = g_counter_data.size();
    ev.ppMetricNames = g_metrics.data();
    ev.numMetrics = nm;
    ev.rangeIndex = (size_t)ri;
    ev.pMetricValues = values + (size_t)ri * nm;
    if (CUptiResult r = cuptiProfilerHostEvaluateToGpuValues(&ev); r != CUPTI_SUCCESS) { teardown(); return fail("cuptiProfilerHostEvaluateToGpuValues", r); }
  }
  teardown();
  return 0;
}

}  // extern "C"
