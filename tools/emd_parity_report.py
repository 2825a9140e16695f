"""Measured parity of the EMD path against the reference CUDA kernels at the benchmarked shapes.  Development / evidence tool.

    python tools/emd_parity_report.py > profiles/r2_emd_parity.txt        (needs a GPU and oracle/_ref/libref_gpu.so)

For BASELINE.json configs[2] (B=32 n=m=2048; n=m=16384) and two input distributions (independent uniform clouds; GT + N(0, 0.01^2)
noise) it prints, against pc_distance/tf_approxmatch.cu recompiled unchanged for sm_100a:
  * RFNET_EMD_EXACT: max |match - ref| / max ref (0: bit-identical), bitwise-equal fraction, cost and gradient errors;
  * default (reference summation order, flushing exponential, derived exponentials in the final pass): the same numbers;
  * RFNET_EMD_SPLIT_SUMS: the same numbers -- these are what the caps in tests/test_emd_gpu.py are derived from (2 x measured);
  * cost(our match) vs cost(reference match);
  * batch invariance: clouds evaluated in batches of 1, 4, 8, 32 give identical bits.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import ref
from rfnet_b200 import ops, tf_approxmatch

dev = torch.device("cuda:0")
EXACT, SPLIT = 1, 4


def gen(b, n, seed, noisy):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x1 = torch.rand((b, n, 3), generator=g) - 0.5
    x2 = x1 + 0.01 * torch.randn((b, n, 3), generator=g) if noisy else torch.rand((b, n, 3), generator=g) - 0.5
    return x1.to(dev), x2.to(dev)


def relmax(a, w):
    return float((a - w).abs().max() / w.abs().max().clamp_min(1e-30))


def ref_emd(x1, x2):
    b, n, m = x1.shape[0], x1.shape[1], x2.shape[1]
    (match,) = ref.run_gpu("ApproxMatch", [x1, x2], [((b, m, n), torch.float32)])
    (cost,) = ref.run_gpu("MatchCost", [x1, x2, match], [((b,), torch.float32)])
    g1, g2 = ref.run_gpu("MatchCostGrad", [x1, x2, match], [((b, n, 3), torch.float32), ((b, m, 3), torch.float32)])
    return match, cost, g1, g2


print("EMD parity vs the reference CUDA kernels (tf_approxmatch.cu recompiled for sm_100a), %s" % torch.cuda.get_device_name(0))
print("errors: match = max|d| / max(ref match); cost = max over clouds of |d|/ref; grad = max|d| / max|ref grad|")
worst = {"match": 0.0, "cost": 0.0, "grad": 0.0}
for (b, n) in ((32, 2048), (4, 16384), (1, 16384), (32, 1024), (64, 64)):
    for noisy in (False, True):
        x1, x2 = gen(b, n, 9000 + n + b, noisy)
        want, wcost, wg1, wg2 = ref_emd(x1, x2)
        tag = "B=%d n=m=%d %s" % (b, n, "GT+noise(0.01)" if noisy else "independent uniform")
        for flags, name in ((EXACT, "exact  "), (0, "default"), (SPLIT, "split  ")):
            got = ops.approx_match_op(x1, x2, flags)
            cost, g1, g2 = ops.emd_cost_grad_op(x1, x2, flags)
            e_m = relmax(got, want)
            eq = float((got == want).float().mean())
            e_c = float(((cost - wcost).abs() / wcost.abs()).max())
            e_g = max(relmax(g1, wg1), relmax(g2, wg2))
            c_of_ours = tf_approxmatch.match_cost(x1, x2, got)
            e_cc = float(((c_of_ours - wcost).abs() / wcost.abs()).max())
            print("%-38s %s: match %.3e (bitwise equal %.5f)  cost(fused) %.3e  cost(match_cost of our match) %.3e  grad %.3e"
                  % (tag, name, e_m, eq, e_c, e_cc, e_g), flush=True)
            if flags == SPLIT:
                worst["match"] = max(worst["match"], e_m)
                worst["cost"] = max(worst["cost"], e_c, e_cc)
                worst["grad"] = max(worst["grad"], e_g)
            del got
        del want
print("worst split-sums errors: match %.3e  cost %.3e  grad %.3e   (test caps = 2 x these)" % (worst["match"], worst["cost"], worst["grad"]))

# batch invariance at 2048 and 16384: same bits for every batch size
for (n, batches) in ((2048, (1, 4, 8, 32)), (16384, (1, 4, 8))):
    bmax = max(batches)
    x1, x2 = gen(bmax, n, 31 + n, False)
    for flags, name in ((0, "default"), (SPLIT, "split sums")):
        full = ops.emd_cost_grad_op(x1, x2, flags)
        ok = True
        for b in batches:
            for lo in (0, bmax - b):
                part = ops.emd_cost_grad_op(x1[lo:lo + b].contiguous(), x2[lo:lo + b].contiguous(), flags)
                ok = ok and all(torch.equal(p, f[lo:lo + b]) for p, f in zip(part, full))
        print("batch invariance n=m=%d (%s), batches %s of the same clouds: cost and gradients bitwise equal = %s" % (n, name, batches, ok))
