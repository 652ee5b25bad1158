"""oracle/ — CPU restatement of the reference's algorithm for the MASIC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under masic_b200/ may import this package; it is
imported by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
`--impl reference` legs, always as the checker and never as the thing measured or shipped.
"""
