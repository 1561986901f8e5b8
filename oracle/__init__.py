"""CPU oracle for the NeRF ray-render hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from this package; the product path (nerf-and-dietnerf_b200/) never does and fails loudly
when its CUDA library is missing.  See oracle/nerf_oracle.py for the parity-pin statement.
"""
