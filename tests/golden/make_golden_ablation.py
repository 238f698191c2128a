"""Golden numbers of the reference's OWN fleet-ablation drivers (authoring container only).

Runs the unmodified ``run_scene_based_fleet_ablation_experiment`` (cli/run.py:283-362) and
``run_algo_based_fleet_ablation_experiment`` (cli/run.py:363-472) through ``oracle/ref_harness.py`` in deterministic
RNG mode (no stochastic event can fire -- SURVEY.md 8a-R) and stores what they hand to ``Charter``: per dispatcher the
fleet sizes and productions, per scene the fleet sizes and target / baseline ratios.  ``matplotlib`` is a stub in this
image, so the Charter methods are replaced by recorders; nothing else of the reference is touched.

    python tests/golden/make_golden_ablation.py      ->  tests/golden/ablation_det.json
"""
import contextlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

DET_DISPATCHERS = ["NaiveDispatcher", "NearestDispatcher", "FixedGroupDispatcher", "SPTFDispatcher", "SQDispatcher"]


def _cfg(name):
    with open(os.path.join(HERE, "configs", name + ".json"), encoding="utf-8") as fh:
        return json.load(fh)


def main():
    ns = rh.load_reference()
    # the drivers discover dispatcher classes by importing EVERY module of dispatch_algorithms (cli/run.py:320-327); some of
    # those import packages that are absent here (asyncore: removed in Python 3.12; docutils, openai, torch-free stubs)
    from unittest import mock
    for name in ("asyncore", "docutils", "docutils.nodes", "openai", "tqdm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = mock.MagicMock(name=name)
    from openmines.src.utils.visualization.charter import Charter
    captured = {}

    def scene_draw(self, results, original_fleet_size=None):
        captured["scene"] = (results, original_fleet_size)

    def algo_draw(self, scenes_data, baseline, target):
        captured["algo"] = scenes_data

    Charter.draw_scene_based_fleet_ablation_experiment = scene_draw
    Charter.draw_algo_based_fleet_ablation_experiment = algo_draw
    Charter.save_ablation = lambda self, tag="": None
    Charter.__init__ = lambda self, *a, **k: None
    out = {"python": list(sys.version_info[:3]), "scene": {}, "algo": {}}
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="om_abl_")
    os.chdir(tmp)
    try:
        with rh.RngControl("det"), contextlib.redirect_stdout(open(os.devnull, "w")):
            # scene-based: one mine, every (deterministic) dispatcher, ten fleet sizes
            for cname, lo, hi, sim_time in (("north_pit_mine_short", 2, 20, None), ("north_pit_mine", 10, 160, 120)):
                cfg = _cfg(cname)
                cfg["dispatcher"]["type"] = DET_DISPATCHERS
                if sim_time:
                    cfg["sim_time"] = sim_time
                path = os.path.join(tmp, cname + ".json")
                with open(path, "w", encoding="utf-8") as fh:
                    json.dump(cfg, fh, ensure_ascii=False)
                ns.run_mod.run_scene_based_fleet_ablation_experiment(path, lo, hi)
                results, orig = captured["scene"]
                out["scene"][cname] = {"min_truck": lo, "max_truck": hi, "sim_time": cfg["sim_time"], "original_fleet_size": orig,
                                       "results": results}
            # algo-based: a directory of scenes, two dispatchers
            sdir = os.path.join(tmp, "scenes")
            os.makedirs(sdir)
            for cname, sim_time in (("north_pit_mine_short", None), ("ragged_mine", 90)):
                cfg = _cfg(cname)
                if sim_time:
                    cfg["sim_time"] = sim_time
                with open(os.path.join(sdir, cname + ".json"), "w", encoding="utf-8") as fh:
                    json.dump(cfg, fh, ensure_ascii=False)
            ns.run_mod.run_algo_based_fleet_ablation_experiment(sdir, "NaiveDispatcher", "SQDispatcher", 1, 40)
            out["algo"] = {"baseline": "NaiveDispatcher", "target": "SQDispatcher", "min_truck": 1, "max_truck": 40,
                           "sim_times": {"north_pit_mine_short": _cfg("north_pit_mine_short")["sim_time"], "ragged_mine": 90},
                           "scenes": captured["algo"]}
    finally:
        os.chdir(cwd)
    with open(os.path.join(HERE, "ablation_det.json"), "w", encoding="utf-8") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    for cname, v in out["scene"].items():
        for d, r in v["results"].items():
            print(cname, d, r["fleet_sizes"], [round(x, 2) for x in r["productions"]])
    for sc, v in out["algo"]["scenes"].items():
        print(sc, v["fleet_sizes"], [round(x, 4) for x in v["ratios"]])


if __name__ == "__main__":
    main()
