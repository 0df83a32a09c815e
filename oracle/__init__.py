"""CPU oracle loader (TEST INFRASTRUCTURE ONLY — see oracle/ghicp_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
from .binding import *  # noqa: F401,F403
