"""CPU oracle for the ntSynt hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (ntsynt_b200/) never does.
"""
