import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dump_time import t
t(32, 4096, 100000, 0, 0)
t(32, 4096, 400000, 0, 0)
t(32, 4096, 400000, 64, 64)
t(64, 4096, 400000, 64, 64)
