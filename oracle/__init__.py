"""CPU oracle for splpak_b200 -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (splpak_b200) never does.
"""
from .oracle import Oracle, build_oracle, oracle_lib_path  # noqa: F401
