"""How small can the jam pool be?  Runs the C3 shard (north_pit, SQ + SPTF, live RNG) with several pool capacities and
prints the number of instances that overflowed (OM_IERR_JAM_OVERFLOW) and the episode-kernel time.  GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import openmines_b200 as om
from bench import load_cfg, workload

cfg = load_cfg()
disp, seeds, mode, desc = workload("c3", int(os.environ.get("N", "16384")), 0)
for J in [int(x) for x in os.environ.get("JS", "8,12,16,24,32,48,64").split(",")]:
    mine = om.BatchedMine(cfg, len(disp), device=0, jam_capacity=J)
    info = mine.launch_info()
    ms = []
    for rep in range(3):
        mine.reset(disp, seeds=np.asarray(seeds) + rep * 1000003, rng_mode=mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = mine.run(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        s = res.summary_np()
        nerr = int((s["err"] != 0).sum())
        print("J=%d rep=%d warps/CTA=%d state=%d B  errs=%d / %d  jams/episode mean %.1f max %d  kernel %.2f ms" % (
            J, rep, info["warps_per_cta"], mine.state_bytes, nerr, len(disp), s["n_jams"].mean(), s["n_jams"].max(), ms[-1]), flush=True)
    del mine
