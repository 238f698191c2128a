"""Event-trace, tick and summary record formats emitted by the device engine
(``OmTraceRecord`` / ``OmTickRecord`` / ``OmSummary`` in include/openmines_b200.h).

One trace record per entry the reference appends to an ``EventPool`` during ``env.run``
(truck.py:143,188,203,311,332,346,395,414,425,447; road.py:218,259; truck.py:364,373).
"""
import numpy as np

TRACE_DTYPE = np.dtype([("time", "<f8"), ("a", "<f8"), ("b", "<f8"), ("truck", "<u2"), ("kind", "u1"),
                        ("flags", "u1"), ("arg0", "<u2"), ("arg1", "<u2")])
TICK_DTYPE = np.dtype([("produced_tons", "<f8"), ("service_count", "<i4"), ("working", "<u2"), ("waiting", "<u2"),
                       ("load_unload", "<u2"), ("moving", "<u2"), ("repairing", "<u2"), ("n_jams", "<u2"),
                       ("n_road_repairs", "<u2"), ("n_truck_repairs", "<u2"), ("n_unrepairable", "<u2"),
                       ("pad", "<u2")])
SUMMARY_DTYPE = np.dtype([("reported_tons", "<f8"), ("dumped_tons", "<f8"), ("now", "<f8"),
                          ("n_truck_events", "<u4"), ("n_decisions", "<u4"), ("n_wrapper_orders", "<u4"),
                          ("n_hops", "<u4"), ("n_jams", "<u4"), ("n_road_repairs", "<u4"),
                          ("n_truck_repairs", "<u4"), ("n_unrepairable", "<u4"), ("n_ticks", "<u4"), ("err", "<u4")])
UNIT_TICK_DTYPE = np.dtype([("tons", "<f8"), ("service_count", "<u4"), ("queue_length", "<u2"), ("repair", "u1"),
                            ("pad", "u1")])
assert UNIT_TICK_DTYPE.itemsize == 16
assert TRACE_DTYPE.itemsize == 32 and TICK_DTYPE.itemsize == 32 and SUMMARY_DTYPE.itemsize == 64

NO_TRUCK = 0xFFFF
(K_INIT_ORDER, K_MOVE_INIT, K_MOVE_HAUL, K_MOVE_UNHAUL, K_WAIT_SHOVEL, K_GET_SHOVEL, K_ORDER_HAUL, K_ORDER_BACK,
 K_WAIT_DUMPER, K_GET_DUMPER, K_BREAKDOWN, K_UNREPAIRABLE, K_JAM, K_ROAD_REPAIR, K_SHOVEL_MAINT_START,
 K_SHOVEL_MAINT_DONE) = range(16)
KIND_NAMES = ["INIT_ORDER", "MOVE_INIT", "MOVE_HAUL", "MOVE_UNHAUL", "WAIT_SHOVEL", "GET_SHOVEL", "ORDER_HAUL",
              "ORDER_BACK", "WAIT_DUMPER", "GET_DUMPER", "BREAKDOWN", "UNREPAIRABLE", "JAM", "ROAD_REPAIR",
              "SHOVEL_MAINT_START", "SHOVEL_MAINT_DONE"]
TRUCK_EVENT_KINDS = tuple(range(12))

STATUS_NAMES = ["idle", "waiting for init order", "moving", "waiting for shovel", "loading",
                "waiting for haul order", "waiting for dumper", "unloading", "waiting for back order",
                "repairing", "unrepairable"]

IERR_NAMES = {0: None, 1: "KeyError (sq_dispatcher.py:31: order at an integer minute before its monitor block)",
              2: "IndexError (bad order index, truck.py:305,390,442)", 3: "jam pool overflow (engine limit)",
              4: "tape exhausted / site mismatch", 5: "trace buffer overflow"}
