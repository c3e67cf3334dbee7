"""CPU/GPU *checkers* for the MEMC-Net motion-compensation hot path.

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.  The product
(memc-net_b200/) never does.

  oracle.cpu      ctypes bindings of oracle/memc_oracle.c (our restatement, f32 + f64 builds)
  oracle.ref      ctypes bindings of oracle/_ref/*.so = the reference's OWN sources
                  (my_lib.c against th_stub/TH.h; my_lib_kernel.cu for sm_100a), built by
                  oracle/Makefile from /root/reference where present
  oracle.pyloop   literal pure-Python/torch loops for tiny cases (BASELINE.json configs[0])
"""
