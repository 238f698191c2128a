"""Loader of the C-ABI shared library ``libopenmines_b200.so`` (include/openmines_b200.h).

There is no CPU fallback: if the library has not been built (``python -c "import
__graft_entry__ as g; g.build()"``, which calls ``build()`` below) loading the engine raises.  The
library is built in-tree next to this file so it travels with the repo snapshot.  A library older
than its sources is rebuilt when nvcc is available and refused otherwise, and the sizes of the ABI
structs it was compiled with are compared with the ctypes mirrors before the first call.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# OM_DEBUG_LIB=1: the bounds-checking build (-DOM_BOUNDS_CHECK, see csrc/om_engine.cuh), used by tools/sanitize.sh
DEBUG = bool(os.environ.get("OM_DEBUG_LIB"))
LIB_PATH = os.environ.get("OM_LIB_PATH") or os.path.join(_HERE, "libopenmines_b200_dbg.so" if DEBUG else "libopenmines_b200.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC",
              "-diag-suppress", "128",   # "loop is not reachable": the OM_CONV() probe stub behind a compile-time-certain break
              # immediates stay in the instructions instead of a compiler-made constant bank (measured, same-call A/B:
              # -0.6 % on the north_pit episode kernel, -1.9 % on the split kernel)
              "-Xptxas", "--disable-optimizer-constants"]

_lib = None


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
        [os.path.join(INCLUDE, "openmines_b200.h")]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """Compile every CUDA source of the engine for sm_100a with nvcc into the in-tree .so."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cus = [s for s in sources() if s.endswith(".cu")]
    # nvcc writes a private file that then replaces the library in one step: a process that loads the library while
    # another one rebuilds it (eight ranks starting together) sees the old file or the new one, never a partial one
    tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
    cmd = [nvcc] + NVCC_FLAGS + (["-DOM_BOUNDS_CHECK=" + (os.environ.get("OM_DEBUG_LIB") or "1")] if DEBUG and os.environ.get("OM_DEBUG_LIB") != "probes" else []) + \
        (os.environ.get("OM_EXTRA_NVCC_FLAGS", "").split()) + (["-Xptxas", "-v"] if verbose else []) + \
        ["-I", INCLUDE, "-o", tmp] + cus
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(r.stderr)
    return LIB_PATH


COUNTERS_SRC = os.path.join(os.path.dirname(_HERE), "tools", "counters", "om_counters.cpp")
COUNTERS_LIB = os.path.join(os.path.dirname(_HERE), "tools", "counters", "libom_counters.so")


def build_counters(force=False):
    """Measurement helper of bench.py (CUPTI range profiler behind a C ABI, tools/counters/om_counters.cpp): g++ against the
    toolkit's libcupti.  Best effort -- returns None when the toolkit has no CUPTI; bench.py then falls back to the committed
    ncu capture."""
    if not os.path.exists(COUNTERS_SRC):
        return None
    if not force and os.path.exists(COUNTERS_LIB) and os.path.getmtime(COUNTERS_LIB) >= os.path.getmtime(COUNTERS_SRC):
        return COUNTERS_LIB
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I" + os.path.join(cuda, "include"),
           "-I" + os.path.join(cuda, "extras", "CUPTI", "include"), COUNTERS_SRC, "-o", COUNTERS_LIB,
           "-L" + os.path.join(cuda, "lib64"), "-L" + os.path.join(cuda, "extras", "CUPTI", "lib64"),
           "-L" + os.path.join(cuda, "lib64", "stubs"), "-lcupti", "-lcuda",
           "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-Wl,-rpath," + os.path.join(cuda, "extras", "CUPTI", "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return COUNTERS_LIB if r.returncode == 0 else None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("openmines_b200: %s is missing -- build it first (__graft_entry__.build()); "
                           "the engine has no CPU fallback" % LIB_PATH)
    if needs_build() and not os.environ.get("OM_LIB_PATH"):
        # a stale binary with changed struct layouts would corrupt memory silently.  (An explicit OM_LIB_PATH names a
        # variant build for an A/B measurement -- tools/build_variant.sh -- and is loaded as it is: rebuilding it from the
        # tree's sources would silently turn it into a copy of the default build; the ABI check below still applies.)
        import shutil
        if shutil.which(os.environ.get("NVCC", "nvcc")) is None:
            raise RuntimeError("openmines_b200: %s is older than its sources and nvcc is not available to rebuild it"
                               % LIB_PATH)
        # one process rebuilds, the others of the node (ranks of one launch) wait for it and find the library current
        import fcntl
        with open(LIB_PATH + ".lock", "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if needs_build():
                    build(force=True)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    lib.om_last_error.restype = ctypes.c_char_p
    lib.om_last_error.argtypes = [vp]
    lib.om_create.argtypes = [vp, i64, ctypes.c_int, ctypes.POINTER(vp)]
    lib.om_create_multi.argtypes = [vp, i32, vp, i64, ctypes.c_int, ctypes.POINTER(vp)]
    lib.om_destroy.argtypes = [vp]
    lib.om_state_bytes.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t)]
    lib.om_state_layout.argtypes = [vp, vp]
    lib.om_launch_info.argtypes = [vp] + [ctypes.POINTER(i32)] * 4
    lib.om_reset.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp]
    lib.om_run.argtypes = [vp, dbl, vp, vp, vp]
    lib.om_stats_reduce.argtypes = [vp, vp, vp, vp, vp]
    lib.om_obs_dim.argtypes = [vp, ctypes.POINTER(i32)]
    lib.om_step.argtypes = [vp, dbl, vp, vp, vp]
    for f in ("om_create", "om_create_multi", "om_destroy", "om_state_bytes", "om_state_layout", "om_launch_info", "om_reset",
              "om_run", "om_stats_reduce", "om_obs_dim", "om_step"):
        getattr(lib, f).restype = ctypes.c_int
    lib.om_debug_draws.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32, i64, vp, vp]
    lib.om_debug_jam_positions.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp]
    lib.om_debug_draws.restype = lib.om_debug_jam_positions.restype = ctypes.c_int
    lib.om_debug_bounds_report.argtypes = [ctypes.POINTER(i32)]
    lib.om_debug_bounds_report.restype = ctypes.c_int
    lib.om_abi_info.argtypes = [ctypes.POINTER(i32), i32]
    lib.om_abi_info.restype = i32
    _lib = lib
    _check_abi()
    return lib


ABI_INFO_FIELDS = ["abi_version", "sizeof_OmConfig", "sizeof_OmOutputs", "sizeof_OmStepIO", "sizeof_OmStateLayout",
                   "sizeof_OmSummary", "sizeof_OmTickRecord", "sizeof_OmTraceRecord", "sizeof_OmUnitTick", "sizeof_InstHdr",
                   "stats_rows", "stats_cols"]


def abi_info():
    """What the loaded binary was compiled with (``om_abi_info``)."""
    lib = load()
    buf = (ctypes.c_int32 * len(ABI_INFO_FIELDS))()
    n = lib.om_abi_info(buf, len(ABI_INFO_FIELDS))
    return {k: int(buf[i]) for i, k in enumerate(ABI_INFO_FIELDS[:n])}


def _check_abi():
    global _lib
    from . import config, engine, trace, parallel
    info = abi_info()
    want = {"abi_version": config.OM_ABI_VERSION, "sizeof_OmConfig": ctypes.sizeof(config.OmConfigStruct),
            "sizeof_OmOutputs": ctypes.sizeof(engine.OmOutputsStruct), "sizeof_OmStepIO": ctypes.sizeof(engine.OmStepIOStruct),
            "sizeof_OmStateLayout": ctypes.sizeof(engine.OmStateLayoutStruct), "sizeof_OmSummary": trace.SUMMARY_DTYPE.itemsize,
            "sizeof_OmTickRecord": trace.TICK_DTYPE.itemsize, "sizeof_OmTraceRecord": trace.TRACE_DTYPE.itemsize,
            "sizeof_OmUnitTick": trace.UNIT_TICK_DTYPE.itemsize, "stats_rows": parallel.N_DISPATCHERS,
            "stats_cols": parallel.STATS_WIDTH}
    bad = {k: (info.get(k), v) for k, v in want.items() if info.get(k) != v}
    if info.get("sizeof_InstHdr", 0) > engine.HDR_DTYPE.itemsize:
        bad["sizeof_InstHdr"] = (info["sizeof_InstHdr"], engine.HDR_DTYPE.itemsize)
    if bad:
        _lib = None
        raise RuntimeError("openmines_b200: %s does not match the Python bindings (binary, bindings): %r -- rebuild it"
                           % (LIB_PATH, bad))


EXPORTED_SYMBOLS = ["om_create", "om_create_multi", "om_destroy", "om_last_error", "om_abi_info", "om_state_bytes", "om_state_layout", "om_reset",
                    "om_run", "om_stats_reduce", "om_launch_info", "om_obs_dim", "om_step", "om_debug_draws", "om_debug_jam_positions", "om_debug_bounds_report"]
