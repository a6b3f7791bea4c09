"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TurtleFFT hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (steganosaurus_b200) never does.
"""
