"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the GetHI hot path (see gethi_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package.  The product (crime_b200, libgh_cuda.so, host/) never does.
"""
