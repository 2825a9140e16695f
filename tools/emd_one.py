"""One approx_match call per configuration, for `ncu --metrics gpu__time_duration.sum` launch lists.  python tools/emd_one.py B N [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rfnet_b200 import ops
b, n = int(sys.argv[1]), int(sys.argv[2])
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device="cpu").manual_seed(3)
x1 = (torch.rand((b, n, 3), generator=g) - 0.5).cuda()
x2 = (torch.rand((b, n, 3), generator=g) - 0.5).cuda()
for _ in range(2):
    c, g1, g2 = ops.emd_cost_grad_op(x1, x2, flags)
torch.cuda.synchronize()
print(float(c.sum()))
