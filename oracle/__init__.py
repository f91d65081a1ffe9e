"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference (oracle/*.c, oracle/*.py) and the
wrapper around the real reference CUDA sources (oracle/_ref).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this package;
wast3d_b200/ never does."""
