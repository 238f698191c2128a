"""TEST INFRASTRUCTURE ONLY -- canonical event-trace record shared by the oracle legs.

One record per entry appended to an ``EventPool`` of the reference during ``env.run``
(truck pools: truck.py:143,188,203,311,332,346,395,414,425,447; global random-event pool:
road.py:218,259; shovel pools: truck.py:364,373).  The CUDA engine emits the same 32-byte
record (include/openmines_b200.h ``OmTraceRecord``); tests compare them field by field.

``time`` is ``env.now`` at the moment the record is created (NOT the possibly bumped
``time_stamp``); ``a``/``b`` are kind-specific fp64 payloads.
"""
import numpy as np

TRACE_DTYPE = np.dtype([
    ("time", "<f8"), ("a", "<f8"), ("b", "<f8"),
    ("truck", "<u2"), ("kind", "u1"), ("flags", "u1"), ("arg0", "<u2"), ("arg1", "<u2"),
])
assert TRACE_DTYPE.itemsize == 32

NO_TRUCK = 0xFFFF

# kind codes                      arg0            arg1         a                    b
K_INIT_ORDER = 0          # load idx        -            distance             -
K_MOVE_INIT = 1           # load idx        -            est_end_time         distance
K_MOVE_HAUL = 2           # dump idx        -            est_end_time         distance
K_MOVE_UNHAUL = 3         # load idx        -            est_end_time         distance
K_WAIT_SHOVEL = 4         # shovel (global) queue_index  -                    -
K_GET_SHOVEL = 5          # shovel (global) -            -                    -
K_ORDER_HAUL = 6          # dump idx        -            distance             -
K_ORDER_BACK = 7          # load idx        -            distance             -
K_WAIT_DUMPER = 8         # dumper (global) -            -                    -
K_GET_DUMPER = 9          # dumper (global) -            -                    -
K_BREAKDOWN = 10          # -               -            last_breakdown_time  repair_time
K_UNREPAIRABLE = 11       # -               -            -                    -
K_JAM = 12                # road id         -            jam_position         est_end_time   (truck=NO_TRUCK)
K_ROAD_REPAIR = 13        # road id         -            repair_end_time      repair_duration (truck=NO_TRUCK)
K_SHOVEL_MAINT_START = 14 # shovel (global) -            end_time             -              (truck=holder)
K_SHOVEL_MAINT_DONE = 15  # shovel (global) -            -                    -              (truck=holder)

TRUCK_EVENT_KINDS = tuple(range(0, 12))   # the records the headline metric counts

KIND_NAMES = ["INIT_ORDER", "MOVE_INIT", "MOVE_HAUL", "MOVE_UNHAUL", "WAIT_SHOVEL", "GET_SHOVEL",
              "ORDER_HAUL", "ORDER_BACK", "WAIT_DUMPER", "GET_DUMPER", "BREAKDOWN", "UNREPAIRABLE",
              "JAM", "ROAD_REPAIR", "SHOVEL_MAINT_START", "SHOVEL_MAINT_DONE"]

# RNG draw sites
S_LOAD_JITTER = 0      # np uniform(-0.1, 0.1)            truck.py:230
S_MAINT_GAP = 1        # np exponential(1440)              truck.py:358
S_MAINT_DUR = 2        # np normal(45, 5) (pre-max)        truck.py:362
S_BREAK_GAP = 3        # np exponential(360)               truck.py:528
S_BREAK_DUR = 4        # np normal(10, 3) (pre-max)        truck.py:532
S_UNREPAIRABLE = 5     # py expovariate(1/70560)           truck.py:545
S_ROAD_GAP = 6         # py expovariate(lambda_repair)     road.py:253
S_ROAD_DUR = 7         # np normal(mu, sigma) (pre-max)    road.py:256
S_DETOUR = 8           # np normal(0.4, 0.1) (pre-max)     road.py:329
S_JAM_COIN = 9         # py random()                       road.py:179
S_JAM_POS = 10         # np choice -> x[idx]               road.py:203-206
S_JAM_DUR = 11         # np weibull(2) (pre *10)           road.py:212
S_DISPATCH = 12        # RandomState.randint               random_dispatcher.py:16,22,29

# TruckStatus codes used in per-tick arrays (string statuses of truck.py)
STATUS_NAMES = ["idle", "waiting for init order", "moving", "waiting for shovel", "loading",
                "waiting for haul order", "waiting for dumper", "unloading", "waiting for back order",
                "repairing", "unrepairable"]


def road_id(L, D, cur_loc, tgt_loc):
    """Road numbering follows the monitor's iteration order (mine.py:138-154).

    Locations: 0 = charging site, 1..L = load sites, L+1..L+D = dump sites.
    """
    if cur_loc == 0:
        return tgt_loc - 1
    if cur_loc <= L:                      # load i -> dump j
        return L + (cur_loc - 1) * D + (tgt_loc - L - 1)
    return L + L * D + (cur_loc - L - 1) * L + (tgt_loc - 1)   # dump j -> load i
