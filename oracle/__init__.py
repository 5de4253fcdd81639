"""CPU oracle for the VB update loop of VBMatrixFactorization.jl.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (vbmatrixfactorization.jl_b200/) imports this.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
"""
