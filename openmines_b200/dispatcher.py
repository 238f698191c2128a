"""The reference's dispatcher plug-in surface, mapped onto device functions.

``BaseDispatcher`` keeps the reference's interface (openmines/src/dispatcher.py:4-98):
``give_init_order / give_haul_order / give_back_order(truck, mine) -> int`` and the order
bookkeeping attributes (``init_orders``, ``haul_orders``, ``back_orders``, ``*_order_count``,
``total_order_count``).  The six built-in dispatchers of ``openmines/src/dispatch_algorithms``
are *names* here: each class carries the ``device_id`` of its CUDA device function
(``Sim::decide`` in csrc/om_engine.cuh); their ``give_*_order`` methods are not Python code
paths -- asking one directly raises, because the decision is taken on the GPU inside the
batched engine.  After a run ``BatchedMine`` fills the bookkeeping lists from the device order
log so downstream consumers (``summary.init_orders`` ...; ticker.py:440-453) keep working.

Any other ``BaseDispatcher`` subclass (a user's own Python policy) runs in external-decision
mode: ``device_id = 255`` and the engine pauses each instance at its next order.
"""


class BaseDispatcher:
    device_id = 255   # OM_DISP_EXTERNAL

    def __init__(self):
        self.init_orders = []
        self.haul_orders = []
        self.back_orders = []
        self.init_order_count = 0
        self.haul_order_count = 0
        self.back_order_count = 0
        self.init_order_time = 0
        self.haul_order_time = 0
        self.back_order_time = 0
        self.total_order_count = 0
        self.total_order_time = 0

    def give_init_order(self, truck, mine) -> int:
        raise NotImplementedError("Subclass must implement this method.")

    def give_haul_order(self, truck, mine) -> int:
        raise NotImplementedError("Subclass must implement this method.")

    def give_back_order(self, truck, mine) -> int:
        raise NotImplementedError("Subclass must implement this method.")

    def record_orders(self, orders_row, sptf_double_log=False):
        """Fill the wrapper's bookkeeping (dispatcher.py:70-86) from one instance's device order log."""
        for code in orders_row:
            code = int(code)
            if code < 0:
                break
            kind, site = code >> 12, code & 0xFFF
            if kind == 0:
                self.init_orders.append(site)
                self.init_order_count += 1
            elif kind == 1:
                self.haul_orders.append(site)
                self.haul_order_count += 1
            else:
                if sptf_double_log:        # sptf_dispatcher.py:59-61 re-enters give_init_order
                    self.init_orders.append(site)
                    self.init_order_count += 1
                self.back_orders.append(site)
                self.back_order_count += 1
        self.total_order_count = self.init_order_count + self.haul_order_count + self.back_order_count


class _DeviceDispatcher(BaseDispatcher):
    def _on_device(self, *_a, **_k):
        raise RuntimeError("%s decides on the GPU inside BatchedMine; it has no host-side code path" % self.name)
    give_init_order = give_haul_order = give_back_order = _on_device


def _make(name, dev_id, ref):
    cls = type(name, (_DeviceDispatcher,), {"device_id": dev_id, "__doc__": "Device function for %s." % ref})

    def __init__(self):
        BaseDispatcher.__init__(self)
        self.name = name
    cls.__init__ = __init__
    return cls


NaiveDispatcher = _make("NaiveDispatcher", 0, "dispatch_algorithms/naive_dispatcher.py:7-22")
RandomDispatcher = _make("RandomDispatcher", 1, "dispatch_algorithms/random_dispatcher.py:6-30")
NearestDispatcher = _make("NearestDispatcher", 2, "dispatch_algorithms/nearest_dispatcher.py:12-67")
FixedGroupDispatcher = _make("FixedGroupDispatcher", 3, "dispatch_algorithms/fixed_group_dispatcher.py:15-110")
SPTFDispatcher = _make("SPTFDispatcher", 4, "dispatch_algorithms/sptf_dispatcher.py:15-61")
SQDispatcher = _make("SQDispatcher", 5, "dispatch_algorithms/sq_dispatcher.py:12-100")
ShortestTripDispatcher = _make("ShortestTripDispatcher", 6, "dispatch_algorithms/shortest_trip_dispatcher.py:7-84")

BUILTIN = {c.__name__: c for c in (NaiveDispatcher, RandomDispatcher, NearestDispatcher, FixedGroupDispatcher,
                                   SPTFDispatcher, SQDispatcher, ShortestTripDispatcher)}


def by_name(name):
    """Dispatcher discovery by class name, as cli/run.py:142-219 does from config['dispatcher']['type']."""
    return BUILTIN[name]()
