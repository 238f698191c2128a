"""TEST INFRASTRUCTURE ONLY.

``oracle/`` holds the CPU checker for the batched mine-simulation hot path:

* ``simpy_standin/``  -- stand-in for the un-vendored third-party SimPy 4.0.2 (semantics restated)
* ``ref_harness.py``  -- runs the UNMODIFIED reference from /root/reference (authoring container only)
* ``des_oracle.py``   -- CPU restatement of the reference's algorithm (travels to the GPU box)
* ``philox.py``       -- pure-Python Philox4x32-10 mirror of the device generator
* ``trace_format.py`` -- the canonical 32-byte event record shared by all legs

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product package ``openmines_b200`` never imports it and has no
CPU fallback: it fails loudly when the CUDA library is missing.
"""
