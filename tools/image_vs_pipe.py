"""worm_image against the fused kernel's image output (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from worm_algorithm_b200 import engine
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for L, n in ((32, 65536), (64, 16384), (128, 4096)):
    bonds = (torch.rand((n, 2 * L * L), device="cuda") < 0.3).to(torch.uint8)
    a = engine.image(bonds, L)
    c, im = engine.pipeline(bonds, L, levels=1, images=[0])
    assert (im[0] == a).all()
    b = n * 18 * L * L
    t1 = timeit(lambda: engine.image(bonds, L)); t2 = timeit(lambda: engine.pipeline(bonds, L, levels=1, images=[0]))
    print("L=%3d image %.3f ms %.0f GB/s   pipeline(images) %.3f ms %.0f GB/s" % (L, t1, b / t1 / 1e6, t2, b / t2 / 1e6))
