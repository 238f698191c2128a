"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the third-party package SimPy.

The reference (370025263/openmines) pins ``simpy==4.0.2`` (openmines/requirements.txt:8)
but SimPy is neither vendored under /root/reference nor installed in this image, and
there is no network.  This package restates the *published* SimPy 4.0.2 semantics that
the reference relies on (call sites: mine.py:31,359-391; truck.py:8,38-51,148,210,228,
242,329,339,350,370,411,419; load_site.py:128; dump_site.py:23):

* ``Environment.schedule(event, priority, delay)`` pushes ``(now+delay, priority, eid, event)``
  on a binary heap; URGENT=0 pops before NORMAL=1 at equal time; ``eid`` is a global
  monotonically increasing counter so equal (time, priority) pops FIFO by schedule order.
* ``step()`` pops the minimum, sets ``now``, detaches ``event.callbacks`` and calls them.
* ``Timeout(delay)`` schedules itself NORMAL at creation.
* ``Process(generator)`` schedules an ``Initialize`` event URGENT with delay 0; ``_resume``
  sends the triggering event's value into the generator; a yielded event that is already
  processed resumes immediately (loop), otherwise ``_resume`` is appended to its callbacks;
  on ``StopIteration`` the Process event itself is scheduled NORMAL, delay 0.
* ``run(until=T)`` schedules a plain event URGENT at T whose callback stops the loop, so
  no NORMAL event with time == T is ever processed.
* ``Resource(capacity)``: see ``resources/resource.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import anything under ``oracle/``.  The product (openmines_b200) never does.
If a real ``simpy`` is importable it is preferred by ``oracle/ref_harness.py`` and the
stand-in is only asserted equal against it.
"""
from .core import Environment, Event, Timeout, Process, Initialize, Interrupt, \
    StopSimulation, EmptySchedule, URGENT, NORMAL
from .resources.resource import Resource, Request, Release
from . import resources

__version__ = "4.0.2+standin"
__all__ = ["Environment", "Event", "Timeout", "Process", "Initialize", "Interrupt",
           "Resource", "Request", "Release", "URGENT", "NORMAL"]
