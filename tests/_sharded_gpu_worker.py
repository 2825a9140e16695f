"""Worker of tests/test_multigpu_gpu.py: launched with torch.distributed.run, one rank per GPU (NCCL).

Every rank evaluates its slice of a global batch through the public sharded losses (rfnet_b200.losses), backward included;
rank 0 also evaluates the WHOLE batch alone and compares: the losses agree to float32 summation order, and every cloud's
gradient -- gathered from the rank that owns it -- equals the single-GPU gradient (bit for bit for EMD, whose per-cloud
results do not depend on the batch; to 1e-6 for Chamfer, whose scatter uses float atomics)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from rfnet_b200 import losses
    B, n_cd, m_cd, n_emd = 6, 1500, 4100, 2048
    g = torch.Generator(device="cpu").manual_seed(4242)      # the same global batch on every rank
    a = torch.rand((B, n_cd, 3), generator=g) - 0.5
    c = torch.rand((B, m_cd, 3), generator=g) - 0.5
    e1 = torch.rand((B, n_emd, 3), generator=g) - 0.5
    e2 = torch.rand((B, n_emd, 3), generator=g) - 0.5
    lo, hi = losses.shard_bounds(B, rank, world)

    def run(sl, sharded):
        xs = [t[sl].to(dev).requires_grad_(True) for t in (a, c, e1, e2)]
        if sharded:
            cd, _ = losses.sharded_chamfer_big(xs[0], xs[1])
            emd = losses.sharded_earth_mover(xs[2], xs[3])
        else:
            cd, _ = losses.chamfer_big(xs[0], xs[1])
            emd = losses.earth_mover(xs[2], xs[3])
        (cd + emd).backward()
        return cd.detach(), emd.detach(), [x.grad for x in xs]

    cd, emd, grads = run(slice(lo, hi), True)
    # gather every rank's gradients on rank 0 (shards may be uneven: pad to the largest)
    per = -(-B // world)
    gathered = []
    for gr in grads:
        pad = torch.zeros((per,) + tuple(gr.shape[1:]), device=dev)
        pad[: gr.shape[0]] = gr
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad)
        gathered.append(out)
    ok = True
    if rank == 0:
        wcd, wemd, wgrads = run(slice(0, B), False)
        ok &= abs(cd.item() - wcd.item()) <= 1e-6 * abs(wcd.item())
        ok &= abs(emd.item() - wemd.item()) <= 1e-6 * abs(wemd.item())
        for k, wg in enumerate(wgrads):
            for r in range(world):
                rlo, rhi = losses.shard_bounds(B, r, world)
                got = gathered[k][r][: rhi - rlo]
                if k >= 2:
                    ok &= torch.equal(got, wg[rlo:rhi])                        # EMD: batch-invariant kernels, same 1/(B n) factor
                else:
                    ok &= bool(torch.allclose(got, wg[rlo:rhi], rtol=1e-5, atol=1e-9))
        print("SHARDED_OK" if ok else "SHARDED_MISMATCH", "world", world, "cd", cd.item(), wcd.item(), "emd", emd.item(), wemd.item(), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
