"""CPU oracle for the CrazyAra hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (crazyara_b200/) never does: it has no CPU fallback.
"""
