"""TEST INFRASTRUCTURE ONLY -- the parity oracles of the strided map/permute/reduce path.

  * `oracle.ref`      : ctypes loader of oracle/libstrided_ref.so, the C restatement of the reference's
                        planner + threaded scheduler + blocked loop nest (PARITY UNPINNED, see strided_ref.h);
  * `oracle.semantic` : independent NumPy statement of the Base-`Array` semantics the reference's own tests
                        compare against (test/othertests.jl is 100 % differential vs Base).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
