"""oracle/ -- CPU restatements of the reference algorithms (TEST INFRASTRUCTURE ONLY).

Nothing under yolov6_b200/ imports this package.  Allowed importers: tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  See DESIGN.md section "Oracle".
"""
