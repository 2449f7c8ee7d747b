"""CPU oracle for the event-representation hot path - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package, and only as the checker / the CPU baseline being reported.  The product package
(event_representation_study_b200) never imports it and has no CPU fallback.
"""
