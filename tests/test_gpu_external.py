"""User-defined BaseDispatcher subclasses on the batched engine (external-decision mode, openmines_b200/external.py).

Python dispatchers written against the object graph README section 6 documents (mine.load_sites, truck.target_location,
load_site.estimated_queue_wait_time, mine.road matrices ...) must drive the GPU simulation to the same trajectory as the
device function implementing the same rule: that pins the proxy views (incl. the host restatement of
update_queue_wait_status) and the pause / resume path bit-exactly."""
import numpy as np
import pytest

from helpers import DISPATCHERS  # noqa: F401

pytestmark = pytest.mark.gpu


def _user_dispatchers():
    import openmines_b200 as om

    class UserNearest(om.BaseDispatcher):
        """Shortest road from where the truck stands (same rule as nearest_dispatcher.py:17-67)."""

        def __init__(self):
            super().__init__()
            self.name = "UserNearest"

        def give_init_order(self, truck, mine):
            d = mine.road.charging_to_load
            return d.index(min(d))

        def give_haul_order(self, truck, mine):
            i = mine.load_sites.index(truck.current_location)
            return mine.road.l2d_road_matrix[i, :].argmin()

        def give_back_order(self, truck, mine):
            j = mine.dump_sites.index(truck.current_location)
            return mine.road.d2l_road_matrix[:, j].argmin()

    class UserSPTF(om.BaseDispatcher):
        """Least expected processing time: service time of the trucks already heading to a site plus the site's estimated
        queue wait (same rule as sptf_dispatcher.py:21-61, including back orders reusing the init rule)."""

        def __init__(self):
            super().__init__()
            self.name = "UserSPTF"

        def give_init_order(self, truck, mine):
            others = [t for t in mine.trucks if t.name != truck.name and t.target_location is not None]
            est = []
            for site in mine.load_sites:
                coming = [t for t in others if t.target_location is site]
                est.append(sum([t.truck_capacity / site.load_site_productivity for t in coming]) + site.estimated_queue_wait_time)
            return int(np.argmin(est))

        def give_haul_order(self, truck, mine):
            others = [t for t in mine.trucks if t.name != truck.name and t.target_location is not None]
            est = [len([t for t in others if t.target_location is site]) + site.estimated_queue_wait_time
                   for site in mine.dump_sites]
            return int(np.argmin(est))

        def give_back_order(self, truck, mine):
            return self.give_init_order(truck, mine)

    class UserBroken(om.BaseDispatcher):
        def __init__(self):
            super().__init__()
            self.name = "UserBroken"

        def give_init_order(self, truck, mine):
            return len(mine.load_sites)        # one past the end

    return UserNearest, UserSPTF, UserBroken


@pytest.mark.parametrize("which", ["nearest", "sptf"])
def test_user_dispatcher_matches_device_function(which, short_mine):
    import openmines_b200 as om
    UserNearest, UserSPTF, _ = _user_dispatchers()
    cls, builtin = (UserNearest, "NearestDispatcher") if which == "nearest" else (UserSPTF, "SPTFDispatcher")
    out = om.run_external(cls, short_mine, n_instances=2, rng_mode="det")
    ref = om.BatchedMine(short_mine, 1)
    ref.reset(builtin, rng_mode="det")
    res = ref.run(order_capacity=4096)
    s = res.summary_np()[0]
    orders = res.orders_np()[0]
    orders = orders[orders >= 0]
    want = {k: [int(c) & 0xFFF for c in orders if (int(c) >> 12) == k] for k in range(3)}
    for e in range(2):
        h, d = out["headers"][e], out["dispatchers"][e]
        assert h["err"] == 0
        assert h["reported_tons"] == s["reported_tons"] and h["n_events"] == s["n_truck_events"]
        assert h["n_decisions"] == s["n_decisions"] and h["now"] <= short_mine["sim_time"]
        assert d.init_orders == want[0] and d.haul_orders == want[1] and d.back_orders == want[2]
        assert d.total_order_count == len(orders)
        assert out["produced_tons"][e] == s["reported_tons"]
    assert out["dispatchers"][0] is not out["dispatchers"][1]


def test_user_dispatcher_bad_index_raises_indexerror(short_mine):
    import openmines_b200 as om
    _, _, UserBroken = _user_dispatchers()
    with pytest.raises(IndexError):
        om.run_external(UserBroken, short_mine, n_instances=1, rng_mode="det")


def test_mine_view_surface(north_pit):
    """The attributes README section 6 lists exist and are consistent with the device state after 30 minutes."""
    import openmines_b200 as om
    mine = om.BatchedMine(north_pit, 1)
    mine.reset("SQDispatcher", seeds=[3], rng_mode="philox")
    mine.run(until=30.0)
    v = om.build_mine_view(mine, 0)
    assert len(v.load_sites) == 5 and len(v.dump_sites) == 5 and len(v.trucks) == 71
    assert all(hasattr(v, a) for a in ("charging_site", "road", "dispatcher", "produced_tons", "service_count"))
    ls, ds = v.load_sites[0], v.dump_sites[0]
    for a in ("shovel_list", "parkinglot", "service_ability_ratio", "estimated_queue_wait_time", "avg_queue_wait_time",
              "load_site_productivity"):
        assert hasattr(ls, a), a
    for a in ("shovel_tons", "shovel_cycle_time", "service_count", "last_service_time", "last_service_done_time",
              "est_waiting_time", "last_breakdown_time", "load_site"):
        assert hasattr(ls.shovel_list[0], a), a
    for a in ("dumper_list", "truck_visits", "produce_tons", "service_count"):
        assert hasattr(ds, a), a
    for a in ("dump_site", "dump_time", "dumper_tons", "service_count"):
        assert hasattr(ds.dumper_list[0], a), a
    assert v.road.l2d_road_matrix.shape == (5, 5) and len(v.road.road_repairs) == 50
    moving = [t for t in v.trucks if t.status == "moving"]
    on_roads = sum(len(v.road.truck_on_road(a, b)) for (a, b) in v.road.road_repairs) + \
        sum(len(v.road.truck_on_road(v.charging_site, s)) for s in v.load_sites)
    assert on_roads == len([t for t in moving if t.current_location is not t.target_location])
    assert abs(sum(d.produce_tons for d in v.dump_sites) - float(mine.headers()[0]["reported_tons"])) < 200.0
    assert v.service_count == sum(t.service_count for t in v.trucks)
