"""CPU oracle for the chirpgp filtering/smoothing hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  See chirpgp_oracle.c for the parity status and DESIGN.md section "Oracle".
"""
