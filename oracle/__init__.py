"""oracle/ — TEST INFRASTRUCTURE ONLY. CPU checkers for the praga_b200 hot path:

  _ref/libpraga_ref.so   the unmodified reference (agrolib sources compiled in place) behind ref_driver.cpp
  liboracle_port.so      plain C++ restatement of the path (oracle_port.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (praga_b200/) never does.
"""
