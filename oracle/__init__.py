"""CPU oracle for the selective-scan hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``medical_image_analysis_b200/`` may
import this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker / CPU baseline, never as the thing that is shipped.
"""
