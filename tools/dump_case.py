"""One bond_flag-style run for ncu: L = 32, 4096 chains, dump rule every 64 steps, caps long reached."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.dump_time import t
t(32, 4096, 100000, 64, 8)
t(32, 4096, 100000, 64, 8)
