"""Capacity-limited resource with a FIFO put-queue (SimPy 4.0.2 semantics, restated).

TEST INFRASTRUCTURE ONLY -- see package docstring.

Semantics the openmines reference depends on:

* ``Request.__init__`` appends itself to ``put_queue``, registers ``_trigger_get`` as a
  callback and calls ``_trigger_put``.  ``_trigger_put`` examines the queue **head only**
  (``_do_put`` returns ``None`` so the scan stops after one element): if ``len(users) <
  capacity`` the head joins ``users``, is removed from the queue and ``succeed()``s
  (scheduled NORMAL at the current time).  The head need not be the newcomer.
* ``res.queue is res.put_queue``; ``res.count == len(res.users)``.
* ``release(req)`` creates a ``Release``: ``_do_get`` removes ``req`` from ``users``
  **immediately** and the Release event is scheduled NORMAL now; only when that event is
  *processed* does its ``_trigger_put`` callback hand the unit to the queue head.
* ``with req:`` exit = ``cancel()`` (drops an un-granted request from the queue) followed
  by ``resource.release(req)`` unless the generator is being closed.
"""
from ..core import Event


class Put(Event):
    def __init__(self, resource):
        super().__init__(resource._env)
        self.resource = resource
        self.proc = self.env.active_process
        resource.put_queue.append(self)
        self.callbacks.append(resource._trigger_get)
        resource._trigger_put(None)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.cancel()
        return None

    def cancel(self):
        if not self.triggered:
            self.resource.put_queue.remove(self)


class Get(Event):
    def __init__(self, resource):
        super().__init__(resource._env)
        self.resource = resource
        self.proc = self.env.active_process
        resource.get_queue.append(self)
        self.callbacks.append(resource._trigger_put)
        resource._trigger_get(None)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.cancel()
        return None

    def cancel(self):
        if not self.triggered:
            self.resource.get_queue.remove(self)


class Request(Put):
    def __exit__(self, exc_type, exc_value, traceback):
        super().__exit__(exc_type, exc_value, traceback)
        if exc_type is not GeneratorExit:
            self.resource.release(self)
        return None


class Release(Get):
    def __init__(self, resource, request):
        self.request = request
        super().__init__(resource)


class Resource:
    def __init__(self, env, capacity=1):
        if capacity <= 0:
            raise ValueError('"capacity" must be > 0.')
        self._env = env
        self._capacity = capacity
        self.put_queue = []
        self.get_queue = []
        self.users = []
        self.queue = self.put_queue

    @property
    def capacity(self):
        return self._capacity

    @property
    def count(self):
        return len(self.users)

    def request(self):
        return Request(self)

    def release(self, request):
        return Release(self, request)

    def _do_put(self, event):
        if len(self.users) < self._capacity:
            self.users.append(event)
            event.usage_since = self._env.now
            event.succeed()

    def _do_get(self, event):
        try:
            self.users.remove(event.request)
        except ValueError:
            pass
        event.succeed()

    def _trigger_put(self, get_event):
        idx = 0
        while idx < len(self.put_queue):
            put_event = self.put_queue[idx]
            proceed = self._do_put(put_event)
            if not put_event.triggered:
                idx += 1
            elif self.put_queue.pop(idx) != put_event:
                raise RuntimeError("Put queue invariant violated")
            if not proceed:
                break

    def _trigger_get(self, put_event):
        idx = 0
        while idx < len(self.get_queue):
            get_event = self.get_queue[idx]
            proceed = self._do_get(get_event)
            if not get_event.triggered:
                idx += 1
            elif self.get_queue.pop(idx) != get_event:
                raise RuntimeError("Get queue invariant violated")
            if not proceed:
                break
