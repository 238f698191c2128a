"""openmines_b200 -- batched, B200-native discrete-event engine for the OpenMines hot path.

Only the hot path of 370025263/openmines is here: stepping many independent mine instances
(seed x dispatcher x config) as hand-written sm_100a CUDA behind the reference's own
dispatcher / Mine.start / gym-step surface.  See DESIGN.md and INTEGRATION.md.
"""
from .config import pack_config, load_config, builtin_config, PackedConfig          # noqa: F401
from .dispatcher import (BaseDispatcher, NaiveDispatcher, RandomDispatcher, NearestDispatcher,   # noqa: F401
                         FixedGroupDispatcher, SPTFDispatcher, SQDispatcher, ShortestTripDispatcher, by_name)
from .engine import BatchedMine, RunResult, DISPATCHER_IDS          # noqa: F401
from .external import run_external, build_mine_view                 # noqa: F401
from .pipeline import EpisodePipeline                               # noqa: F401

__version__ = "0.1.0"
