import sys; sys.path.insert(0,".")
from tools.quick_time import t_philox
t_philox(32, 4096, 200000, 0, 0, hist=True, reps=1)
