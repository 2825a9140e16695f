"""GPU parity of approx_match / match_cost / match_cost_grad against the CPU oracle and the reference CUDA kernels.

Tolerances.  BASELINE.json asks for 1e-4 relative on EMD.  match_cost, match_cost_grad (for a GIVEN match) and the final
loss meet it against every oracle.  The match MATRIX itself is a different matter: approx_match is an ill-conditioned
float32 iteration -- perturbing each exp() by 1 ulp, or summing a row in a different order, moves entries of `match` by up
to ~1e-4 of the largest entry (measured: float32 oracle vs the same code in float64 differs by 9e-5 at n=512; the
reference's own CPU and GPU kernels differ by 6e-4).  So:
  * against the reference CUDA kernel, in the configuration where our sums run in the reference's order (no candidate
    split), `match` must agree to 1e-5 of its largest entry -- tighter than asked;
  * against the CPU oracle (libm expf instead of MUFU ex2) and in split mode, `match` is held to
    max(1e-4, 8 x noise band) of its largest entry, where the noise band is measured on the same inputs as
    |float32 oracle - float64 oracle| (oracle.port.approx_match_noise_band: 1e-6 on most clouds, up to 4e-3 on some), and
    never looser than MATCH_RTOL = 5e-2;
    the well-conditioned quantities derived from it (cost, marginals) are held to 1e-4."""
import numpy as np
import pytest
import torch

from conftest import cloud
from oracle import port, ref

pytestmark = pytest.mark.gpu
RTOL = 1e-4
MATCH_RTOL = 5e-2


def match_tol(x1, x2, want):
    return port.approx_match_tolerance(x1, x2, want, cap=MATCH_RTOL)


def close(got, want, rtol=RTOL):
    scale = max(float(np.abs(want).max()), 1e-30)
    err = float(np.abs(got - want).max())
    assert err <= rtol * scale, "max abs err %.3e vs scale %.3e (rel %.3e)" % (err, scale, err / scale)


def t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# n = dataset points (xyz1), m = query points (xyz2); includes n != m (integer-division multipliers) and ragged tiles
SHAPES = [(1, 1, 1), (2, 16, 16), (2, 100, 100), (1, 257, 130), (1, 64, 200), (2, 512, 512), (1, 1030, 1030)]


@pytest.mark.parametrize("b,n,m", SHAPES)
def test_approx_match_vs_oracle(cuda, rng, b, n, m):
    from rfnet_b200 import tf_approxmatch
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    want = port.approx_match(x1, x2)
    got = tf_approxmatch.approx_match(t(x1, cuda), t(x2, cuda)).cpu().numpy()
    assert got.shape == (b, m, n)
    close(got, want, match_tol(x1, x2, want))
    # the well-conditioned functional of the plan: its cost, at the asked 1e-4
    assert np.allclose(port.match_cost(x1, x2, got), port.match_cost(x1, x2, want), rtol=RTOL, atol=1e-7)
    # transport-plan properties: non-negative, row/column mass bounded by the multipliers (tf_approxmatch.cu:4-10)
    assert (got >= 0).all()
    multiL, multiR = (1, n // m) if n >= m else (m // n, 1)
    assert (got.sum(axis=1) <= multiL * (1 + 1e-3)).all() and (got.sum(axis=2) <= multiR * (1 + 1e-3)).all()


def test_approx_match_noisy_copy(cuda, rng):
    # second input distribution of SURVEY.md 8(d): GT + N(0, 0.01^2) noise (stresses the high levels)
    from rfnet_b200 import tf_approxmatch
    x1 = cloud(rng, 2, 400)
    x2 = (x1 + rng.normal(0, 0.01, x1.shape)).astype(np.float32)
    want = port.approx_match(x1, x2)
    got = tf_approxmatch.approx_match(t(x1, cuda), t(x2, cuda)).cpu().numpy()
    close(got, want, match_tol(x1, x2, want))
    assert np.allclose(port.match_cost(x1, x2, got), port.match_cost(x1, x2, want), rtol=RTOL)


@pytest.mark.parametrize("b,n,m", [(2, 100, 100), (1, 257, 130), (2, 512, 512)])
def test_match_cost_and_grad_vs_oracle(cuda, rng, b, n, m):
    from rfnet_b200 import ops, tf_approxmatch
    x1, x2 = cloud(rng, b, n), cloud(rng, b, m)
    match = port.approx_match(x1, x2)
    want_cost = port.match_cost(x1, x2, match)
    got_cost = tf_approxmatch.match_cost(t(x1, cuda), t(x2, cuda), t(match, cuda)).cpu().numpy()
    assert np.allclose(got_cost, want_cost, rtol=RTOL, atol=0)
    w1, w2 = port.match_cost_grad(x1, x2, match)
    g1, g2 = ops.match_cost_grad_op(t(x1, cuda), t(x2, cuda), t(match, cuda))
    close(g1.cpu().numpy(), w1)
    close(g2.cpu().numpy(), w2)


def test_earth_mover_autograd(cuda, rng):
    """earth_mover (vv_recon.py:392-399) forward + backward: gradient = match_cost_grad scaled by 1/(B*n)."""
    from rfnet_b200 import losses
    b, n = 3, 256
    x1n, x2n = cloud(rng, b, n), cloud(rng, b, n)
    x1 = t(x1n, cuda).requires_grad_(True)
    x2 = t(x2n, cuda).requires_grad_(True)
    loss = losses.earth_mover(x1, x2)
    loss.backward()
    match = port.approx_match(x1n, x2n)
    cost = port.match_cost(x1n, x2n, match)
    assert abs(loss.item() - float((cost / n).mean())) <= RTOL * abs(float((cost / n).mean()))
    # gradient for OUR match (the gradient is linear in match; match itself is compared above at MATCH_RTOL)
    from rfnet_b200 import tf_approxmatch
    ours = tf_approxmatch.approx_match(x1.detach(), x2.detach()).cpu().numpy()
    w1, w2 = port.match_cost_grad(x1n, x2n, ours)
    close(x1.grad.cpu().numpy(), w1 / (b * n))
    close(x2.grad.cpu().numpy(), w2 / (b * n))
    o1, o2 = port.match_cost_grad(x1n, x2n, match)
    close(x1.grad.cpu().numpy(), o1 / (b * n), rtol=5e-3)


@pytest.mark.skipif(not ref.available("gpu"), reason="oracle/_ref/libref_gpu.so not built")
@pytest.mark.parametrize("b,n,m,rtol", [(600, 256, 256, 1e-5), (1200, 128, 300, 1e-5), (2, 512, 512, MATCH_RTOL), (1, 2048, 2048, MATCH_RTOL), (1, 1000, 3000, MATCH_RTOL)])
def test_emd_vs_reference_cuda_kernels(cuda, rng, b, n, m, rtol):
    """Against the reference's own CUDA kernels (tf_approxmatch.cu recompiled for sm_100a), same inputs.  The first two
    shapes have enough clouds that no candidate split is used: sums run in the reference's order -> 1e-5."""
    from rfnet_b200 import ops, tf_approxmatch
    x1, x2 = t(cloud(rng, b, n), cuda), t(cloud(rng, b, m), cuda)
    (want,) = ref.run_gpu("ApproxMatch", [x1, x2], [((b, m, n), torch.float32)])
    got = tf_approxmatch.approx_match(x1, x2)
    if rtol >= 1e-4:
        rtol = match_tol(x1.cpu().numpy(), x2.cpu().numpy(), None)
    close(got.cpu().numpy(), want.cpu().numpy(), rtol)
    if rtol < 1e-4:
        frac_equal = float((got == want).float().mean())
        print("bitwise-equal fraction of match entries: %.6f" % frac_equal)
        assert frac_equal > 0.95
    (wcost,) = ref.run_gpu("MatchCost", [x1, x2, want], [((b,), torch.float32)])
    gcost = tf_approxmatch.match_cost(x1, x2, want)
    assert np.allclose(gcost.cpu().numpy(), wcost.cpu().numpy(), rtol=RTOL)
    wg1, wg2 = ref.run_gpu("MatchCostGrad", [x1, x2, want], [((b, n, 3), torch.float32), ((b, m, 3), torch.float32)])
    g1, g2 = ops.match_cost_grad_op(x1, x2, want)
    close(g1.cpu().numpy(), wg1.cpu().numpy())
    close(g2.cpu().numpy(), wg2.cpu().numpy())


def test_emd_full_size_properties(cuda):
    """BASELINE config 3 size (n = m = 16384, one cloud): mass conservation of the plan and cost consistency.
    Every dataset point ends up (almost) fully matched: column sums of match == 1 within 1e-3."""
    from rfnet_b200 import tf_approxmatch
    g = torch.Generator(device="cpu").manual_seed(5)
    n = 16384
    x1 = (torch.rand((1, n, 3), generator=g) - 0.5).to(cuda)
    x2 = (torch.rand((1, n, 3), generator=g) - 0.5).to(cuda)
    match = tf_approxmatch.approx_match(x1, x2)
    assert match.shape == (1, n, n) and bool((match >= 0).all())
    rows, cols = match.sum(dim=2), match.sum(dim=1)
    assert float(rows.max()) <= 1 + 1e-3 and float(cols.max()) <= 1 + 1e-3
    assert float(cols.mean()) > 0.99  # nearly all mass is transported after the level-0 pass
    cost = tf_approxmatch.match_cost(x1, x2, match)
    # cost = <match, dist> computed independently with torch in float64 on a 2048-column slab
    sl = slice(0, 2048)
    d = torch.cdist(x2.double(), x1[:, sl].double())          # (1, m, 2048): match[l,k] pairs xyz2[l] with xyz1[k]
    part = (match[:, :, sl].double() * d).sum()
    part_cost = tf_approxmatch.match_cost(x1[:, sl].contiguous(), x2, match[:, :, sl].contiguous())
    assert abs(part_cost.item() - part.item()) <= 1e-4 * abs(part.item())
    assert 0 < part_cost.item() < cost.item()


@pytest.mark.parametrize("b,n,m", [(1, 1, 1), (2, 100, 100), (1, 257, 130), (1, 64, 200), (3, 512, 512), (1, 1030, 1030), (600, 256, 256)])
def test_emd_cost_fused(cuda, rng, b, n, m):
    """rfnet_emd_cost (approx_match + match_cost in one call, vv_recon.py:396-399): the cost equals the two-op chain at 1e-5;
    with keep_match the matrix written is bit-identical to approx_match's; without it no matrix exists."""
    from rfnet_b200 import ops, tf_approxmatch
    x1n, x2n = cloud(rng, b, n), cloud(rng, b, m)
    x1, x2 = t(x1n, cuda), t(x2n, cuda)
    match = tf_approxmatch.approx_match(x1, x2)
    chain = tf_approxmatch.match_cost(x1, x2, match).cpu().numpy()
    cost0, none = ops.emd_cost_op(x1, x2, False)
    assert none.numel() == 0 and cost0.shape == (b,)
    assert np.allclose(cost0.cpu().numpy(), chain, rtol=1e-5, atol=1e-7)
    cost1, kept = ops.emd_cost_op(x1, x2, True)
    assert torch.equal(kept, match)
    assert torch.equal(cost1, cost0)
    assert np.allclose(tf_approxmatch.emd_cost(x1, x2).cpu().numpy(), chain, rtol=1e-5, atol=1e-7)
    if b * n * m <= 1 << 21:
        want = port.match_cost(x1n, x2n, port.approx_match(x1n, x2n))
        assert np.allclose(cost0.cpu().numpy(), want, rtol=RTOL, atol=1e-7)


def test_emd_cost_fused_edge_and_grad(cuda, rng):
    from rfnet_b200 import ops, tf_approxmatch
    empty = torch.zeros((2, 0, 3), device=cuda)
    pts = t(cloud(rng, 2, 8), cuda)
    cost, _ = ops.emd_cost_op(empty, pts, False)
    assert cost.tolist() == [0.0, 0.0]
    # gradient path keeps the matrix and equals the two-op chain's gradient exactly (same match, same grad kernel)
    x1 = t(cloud(rng, 2, 300), cuda).requires_grad_(True)
    x2 = t(cloud(rng, 2, 300), cuda).requires_grad_(True)
    tf_approxmatch.emd_cost(x1, x2).sum().backward()
    g1, g2 = x1.grad.clone(), x2.grad.clone()
    x1.grad = x2.grad = None
    tf_approxmatch.match_cost(x1, x2, tf_approxmatch.approx_match(x1, x2)).sum().backward()
    assert torch.equal(g1, x1.grad) and torch.equal(g2, x2.grad)
    with pytest.raises(RuntimeError, match="keep_match=False"):
        c, _ = ops.emd_cost_op(x1, x2, False)
        c.sum().backward()


@pytest.mark.parametrize("b,n,m,kind", [(2, 4096, 4096, "cube"), (1, 5000, 4100, "cube"), (1, 4096, 8192, "sphere"), (130, 4096, 4096, "cube"),
                                        (1, 16384, 16384, "sphere")])
def test_pruned_sweeps_are_exact(cuda, b, n, m, kind):
    """From 4096 points per cloud the three sharpest levels run as pruned sweeps (Morton-ordered rows, per-cluster candidate
    masks).  Skipped terms are exact zeros and the surviving ones are added in the same order, so the plan must be
    BIT-IDENTICAL to the dense sweeps' (RFNET_EMD_NO_PRUNE=1), split or not (b = 130 runs unsplit)."""
    import os
    from rfnet_b200 import ops, tf_approxmatch
    g = torch.Generator(device="cpu").manual_seed(1000 + n + m + b)
    def pts(count):
        x = torch.rand((b, count, 3), generator=g) - 0.5
        if kind == "sphere":   # a surface, like real scans, with a clump of exact duplicates
            x = 0.5 * x / x.norm(dim=-1, keepdim=True)
            x[:, : count // 16] = x[:, count // 16: 2 * (count // 16)]
        return x.to(cuda)
    x1, x2 = pts(n), pts(m)
    assert os.environ.get("RFNET_EMD_NO_PRUNE") is None
    pruned = tf_approxmatch.approx_match(x1, x2)
    cost_pruned, _ = ops.emd_cost_op(x1, x2, False)
    os.environ["RFNET_EMD_NO_PRUNE"] = "1"
    try:
        dense = tf_approxmatch.approx_match(x1, x2)
        cost_dense, _ = ops.emd_cost_op(x1, x2, False)
    finally:
        del os.environ["RFNET_EMD_NO_PRUNE"]
    assert torch.equal(pruned, dense)
    assert torch.equal(cost_pruned, cost_dense)
    assert float(dense.sum()) > 0.9 * b * min(n, m)
