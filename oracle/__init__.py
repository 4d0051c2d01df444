"""ORACLE — CPU/fp32 restatement of the reference denoiser path.  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package (b200sr)."""
