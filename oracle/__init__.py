"""TEST INFRASTRUCTURE ONLY — CPU restatements used as the parity checker.

Nothing under oracle/ may be imported by the product path (llmlb_b200/): only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
"""
