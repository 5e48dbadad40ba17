"""TEST INFRASTRUCTURE ONLY: CPU oracle for the NB-LDPC decode path (see nbldpc_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
