"""One pass over every post-processing kernel at sizes larger than L2 (the ncu target for their DRAM traffic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
L, n = 64, 16384
bonds = (torch.rand((n, 2 * L * L), device="cuda") < 0.3).to(torch.uint8)
for _ in range(2):
    img = engine.image(bonds, L)                    # image_kernel_pipe: 128 MiB in, 1 GiB out
    blk = engine.block(img, 0)                      # block_kernel_vec
    cnt = engine.count(img)                         # count_kernel
    c = engine.pipeline(bonds, L)                   # pipeline_kernel
    gi = torch.rand((4096, 64, 64), device="cuda", dtype=torch.float64)
    b = engine.boundaries(gi, np.linspace(0.05, 0.95, 8))   # boundaries_kernel_pair
torch.cuda.synchronize()
print("ok", int(cnt.sum()), int(c[:, 0].sum()))
