"""CPU oracle of the KLT hot path -- TEST INFRASTRUCTURE ONLY (see oracle/klt_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
