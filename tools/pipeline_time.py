"""C3 chain: fused worm_pipeline against the 12-launch chain (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from worm_algorithm_b200 import engine

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for L, n in ((64, 4096), (64, 65536), (128, 16384), (32, 262144)):
    bonds = (torch.rand((n, 2 * L * L), device="cuda") < 0.3).to(torch.uint8)
    def chain():
        img = engine.image(bonds, L); c = [engine.count(img)]
        while img.shape[-1] > 4:
            img = engine.block(img, 0); c.append(engine.count(img))
        return c
    ms_f = t(lambda: engine.pipeline(bonds, L))
    ms_c = t(chain, 5)
    by = 2 * L * L * n
    print("L=%3d n=%6d fused %.4f ms (%.3e configs/s, %.0f GB/s of dump bytes)   chain %.3f ms (%.3e configs/s)" % (
        L, n, ms_f, n / ms_f * 1e3, by / ms_f / 1e6, ms_c, n / ms_c * 1e3), flush=True)
