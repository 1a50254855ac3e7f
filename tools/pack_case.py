"""One timed production run: python tools/pack_case.py L n steps algo lanes hist  (also the ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.quick_time import t_philox
L, n, steps, algo, lanes, hist = (int(x) for x in sys.argv[1:7])
t_philox(L, n, steps, algo, lanes, hist=bool(hist), reps=1)
