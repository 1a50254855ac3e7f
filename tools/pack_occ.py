"""Binary chain at different occupancies (variant builds; development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_time import t_philox
for n in (75776, 94720, 113664, 250000):
    t_philox(32, n, 40000, 1, -2)
t_philox(16, 250000, 40000, 0, -2)
t_philox(16, 250000, 40000, 1, -2)
