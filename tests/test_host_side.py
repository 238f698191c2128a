"""CPU tests (no GPU) of the host side: config packer anchors, dispatcher surface, C-ABI exports."""
import ctypes
import os

import numpy as np
import pytest

import openmines_b200 as om
from openmines_b200 import _lib, config as cfgmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packer_static_anchors(north_pit):
    """SURVEY.md Appendix D anchors, recomputed with the reference's arithmetic."""
    pc = om.pack_config(north_pit)
    assert (pc.T, pc.L, pc.D, pc.S, pc.M, pc.R) == (71, 5, 5, 20, 37, 55)
    mins = (pc.road_distance[:5] / 25) * 60
    assert list(mins) == [7.199999999999999, 10.319999999999999, 7.536000000000001, 5.4239999999999995, 5.04]
    lt = pc.load_time.reshape(pc.n_types, pc.S)
    assert lt[0, 0] == 34.22222222222222 and lt[0, 5] == 5.684055118110236
    assert lt[1, 0] == 15.555555555555555 and lt[1, 5] == 2.5836614173228347
    assert lt[2, 0] == 24.444444444444443 and lt[2, 5] == 4.060039370078741
    assert pc.load_site_productivity == [11.25, 22.546666666666667, 18.046666666666667, 22.546666666666667,
                                         27.093333333333334]
    assert pc.nearest_init == 4
    assert list(pc.nearest_haul) == [4, 4, 4, 1, 0] and list(pc.nearest_back) == [2, 2, 2, 3, 3]
    assert list(np.bincount(pc.fixed_group)) == [6, 19, 17, 15, 14]
    assert list(pc.site_unit_begin) == [0, 5, 10, 13, 18, 20, 25, 33, 41, 49, 57]
    # dump j -> load i uses d2l[i][j] (road.py:314-316)
    d2l = np.array(north_pit["road"]["d2l_road_matrix"])
    assert pc.road_distance[5 + 25 + 4 * 5 + 0] == d2l[0][4] == 34.26


def test_packer_rejects_duplicate_truck_names(north_pit):
    cfg = {**north_pit, "charging_site": {**north_pit["charging_site"], "trucks": [
        {"type": "A", "count": 2, "capacity": 10, "speed": 20}, {"type": "A", "count": 1, "capacity": 20, "speed": 20}]}}
    with pytest.raises(NotImplementedError):
        om.pack_config(cfg)


def test_dispatcher_surface():
    d = om.by_name("SQDispatcher")
    assert isinstance(d, om.BaseDispatcher) and d.name == "SQDispatcher" and d.device_id == 5
    assert d.total_order_count == 0 and d.init_orders == []
    with pytest.raises(RuntimeError):
        d.give_init_order(None, None)
    d = om.SPTFDispatcher()
    d.record_orders([(0 << 12) | 3, (1 << 12) | 2, (2 << 12) | 1, -1], sptf_double_log=True)
    assert d.init_orders == [3, 1] and d.haul_orders == [2] and d.back_orders == [1] and d.total_order_count == 4

    class Mine(om.BaseDispatcher):
        pass
    assert Mine().device_id == 255
    with pytest.raises(NotImplementedError):
        Mine().give_haul_order(None, None)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports what include/openmines_b200.h declares."""
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(_lib.INCLUDE, "openmines_b200.h")).read()
    import re
    declared = set(re.findall(r"^(?:int|const char\*)\s+(om_[a-z_]+)\s*\(", header, re.M))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_struct_sizes_match_header():
    from openmines_b200.engine import OmOutputsStruct, OmStateLayoutStruct
    from openmines_b200.trace import SUMMARY_DTYPE, TICK_DTYPE, TRACE_DTYPE
    assert ctypes.sizeof(OmOutputsStruct) == 104
    info = _lib.abi_info()
    assert info["abi_version"] == cfgmod.OM_ABI_VERSION
    assert info["sizeof_OmOutputs"] == ctypes.sizeof(OmOutputsStruct)
    assert info["sizeof_OmStateLayout"] == ctypes.sizeof(OmStateLayoutStruct)
    assert info["sizeof_OmConfig"] == ctypes.sizeof(cfgmod.OmConfigStruct)
    assert ctypes.sizeof(OmStateLayoutStruct) == 4 * 60
    assert ctypes.sizeof(cfgmod.OmConfigStruct) == 12 * 4 + 14 * 8 + 15 * 8
    assert (TRACE_DTYPE.itemsize, TICK_DTYPE.itemsize, SUMMARY_DTYPE.itemsize) == (32, 32, 64)


def test_create_fails_loudly_without_gpu(north_pit):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        om.BatchedMine(north_pit, 4)
    lib = _lib.load()
    pc = om.pack_config(north_pit)
    s = pc.c_struct()
    h = ctypes.c_void_p()
    rc = lib.om_create(ctypes.byref(s), 4, 0, ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.om_last_error(None)


def test_product_never_imports_oracle():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "openmines_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_external_runner_needs_a_gpu(north_pit):
    """run_external goes through the same C ABI: without a CUDA device it fails loudly (no host-side simulation)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import openmines_b200 as om

    class D(om.BaseDispatcher):
        def give_init_order(self, truck, mine):
            return 0
    with pytest.raises(RuntimeError):
        om.run_external(D, north_pit, n_instances=1)


def _build_c_demo(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    libdir = os.path.join(ROOT, "openmines_b200")
    exe = str(tmp_path / "c_api_demo")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "c_api_demo.c"), "-o", exe, "-L" + libdir, "-lopenmines_b200",
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + libdir,
           "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_abi_compiles_as_plain_c_and_fails_loudly_without_gpu(tmp_path):
    """include/openmines_b200.h is a C header (no C++, no torch types): examples/c_api_demo.c builds with gcc -Wall -Werror
    against the shared library; without a GPU om_create reports OM_E_CUDA instead of falling back to anything."""
    import subprocess
    import torch
    from openmines_b200 import _lib
    _lib.load()
    exe = _build_c_demo(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert r.returncode == 3
    assert "no CUDA device" in r.stderr


def _limits_mine(n_load=16, n_dump=16, units_site0=32):
    from openmines_b200.synthetic import scaled_mine
    cfg = scaled_mine(n_trucks=(60, 80, 60), n_load=n_load, n_dump=n_dump, sim_time=60)
    sh = cfg["load_sites"][0]["shovels"]
    cfg["load_sites"][0]["shovels"] = [dict(sh[1], name="LS1-Shovel-%d" % (k + 1)) for k in range(units_site0)]
    cfg["dump_sites"][0]["dumpers"][0]["count"] = units_site0
    for j in range(1, n_dump):
        cfg["dump_sites"][j]["dumpers"][0]["count"] = 1
    return cfg


def test_om_create_rejects_mines_beyond_the_engine_limits():
    """More than 32 sites: OM_E_UNSUPPORTED with a message, checked before any CUDA call (so this runs without a GPU); an
    ABI version mismatch, an empty mine and out-of-range index tables: OM_E_INVALID.  More than 32 shovels / dumpers at one
    site is supported (unit selection takes one lane trip per 32 units)."""
    lib = _lib.load()

    def create(cfg=None, mutate=None):
        pc = om.pack_config(cfg if cfg is not None else _limits_mine())
        s = pc.c_struct()
        if mutate:
            mutate(s)
        h = ctypes.c_void_p()
        return lib.om_create(ctypes.byref(s), 2, 0, ctypes.byref(h)), lib.om_last_error(None)

    rc, msg = create(_limits_mine(n_load=17, n_dump=16))
    assert rc == -4 and b"L + D <= 32" in msg
    rc, msg = create(_limits_mine(units_site0=33))
    assert rc == -2 and b"no CUDA device" in msg          # accepted by the argument checks; only the GPU is missing here
    import numpy as np

    def bad_table(name, value):
        def mutate(s):
            arr = np.array([value] * 400, dtype=np.int32)
            mutate.keep = arr
            setattr(s, name, arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        return mutate
    for name, value in (("truck_type", 99), ("fixed_group", -1), ("nearest_haul", 16), ("nearest_back", 77)):
        rc, msg = create(mutate=bad_table(name, value))
        assert rc == -1 and b"out of range" in msg, (name, rc, msg)
    rc, msg = create(mutate=lambda s: setattr(s, "nearest_init", 16))
    assert rc == -1 and b"nearest_init" in msg
    rc, msg = create(mutate=lambda s: setattr(s, "abi_version", 1))
    assert rc == -1 and b"ABI" in msg
    rc, msg = create(mutate=lambda s: setattr(s, "n_trucks", 0))
    assert rc == -1 and b"empty mine" in msg
    rc, msg = create(mutate=lambda s: setattr(s, "jam_capacity", 0))
    assert rc == -4


def test_stale_library_is_rebuilt_once_and_never_loaded_half_written():
    """Ranks of one launch that start together on a library older than its sources (it happened: eight ranks, eight nvcc
    runs into the same file, `file too short`): one of them rebuilds under a lock into a private file that replaces the
    library in one step, the others wait and load the finished one."""
    import shutil
    import subprocess
    import sys
    import time
    if shutil.which(os.environ.get("NVCC", "nvcc")) is None:
        pytest.skip("no nvcc")
    src = os.path.join(_lib.CSRC, "om_engine.cu")
    st = os.stat(src)
    code = ("import time; from openmines_b200 import _lib; lib = _lib.load(); "
            "assert not _lib.needs_build(); print('abi', _lib.abi_info()['abi_version'])")
    env = dict(os.environ)
    env.pop("OM_LIB_PATH", None)
    env.pop("OM_DEBUG_LIB", None)
    root = os.path.dirname(os.path.dirname(os.path.abspath(_lib.__file__)))
    try:
        os.utime(src, (time.time(), time.time()))            # the sources are now newer than the library
        assert _lib.needs_build()
        procs = [subprocess.Popen([sys.executable, "-c", code], cwd=root, env=env, stdout=subprocess.PIPE,
                                  stderr=subprocess.PIPE, text=True) for _ in range(3)]
        outs = [p.communicate(timeout=600) for p in procs]
        for p, (out, err) in zip(procs, outs):
            assert p.returncode == 0, err[-2000:]
            assert "abi %d" % cfgmod.OM_ABI_VERSION in out
        assert not _lib.needs_build()
        assert not [f for f in os.listdir(os.path.dirname(_lib.LIB_PATH)) if ".so.tmp." in f]     # no private file left behind
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
