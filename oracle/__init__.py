"""CPU oracle for the GeneWalk hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``genewalk_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.
See ``gw_oracle.c`` for what is restated and the "parity unpinned" note on the gensim step.
"""
