"""ORACLE -- CPU restatement of the reference's assemble/solve/project path (test infrastructure).

parity unpinned (no reference golden vectors exist; dolfin is not installable here) -- see
DESIGN.md.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (fenics_b200/) never does.
"""
