"""CPU oracle for the karma k-mer front end.

TEST INFRASTRUCTURE ONLY.  Nothing under ``karma_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker
or as the timed CPU baseline -- never as the thing shipped.
"""
