"""TEST INFRASTRUCTURE: CPU oracle for the PA-Star hot path (see pastar_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (mpi_pastar_msa_b200)
never does.
"""
