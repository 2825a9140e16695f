#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): Chamfer nearest-neighbour search, forward + gradient.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- nn_distance forward + NnDistanceGrad, B=32 clouds per GPU, partial
2048 points vs dense 16384 points, synthetic uniform [-0.5,0.5]^3 clouds.  One step = one pass over one
batch = ONE C-ABI call (rfnet_chamfer_step: search, gradient + key unpack + sqrt partial sums, final reduction: three kernel
launches).  Unit of work = point pair (2*B*N*M per step, two directed searches).  Batches are sharded over ranks with no
data-path collective; the only exchange is the 16-byte all-reduce of the loss partial sums (weak scaling: B per GPU fixed).

One JSON line on stdout (rank 0).  Besides the contract's keys it carries
  roofline      the dominant kernel (nn_filter_kernel) against the FP32 pipe: 6 algorithmic lane-ops per pair (3 sub, 1 mul, 2 fma --
                the reference's operand order admits no fewer; the filtered search issues 3 per pair and evaluates the reference
                expression only where it decides the result), peak = 148 SMs x 128 lanes x sm clock; the direct kernel beside it
  cpu_baseline  the reference's own CPU kernel (oracle/_ref: /root/reference/pc_distance/tf_nndistance.cpp compiled
                unmodified) timed on this host's cores on a bounded sample of the same workload
  e2e           the same metric through the public host-buffer API: every step copies its batch from pinned HOST buffers and reads the
                loss back (copies inside the timed region); e2e.all_outputs: the same with every output of the operator copied back too
  extra         measured after the timed region on ALL ranks (max over ranks, scalar all-reduce included), not part of `value`:
                  sustained        the same step looped >= 2 s, with the clocks sampled during the loop
                  config3          EMD approx_match + match_cost forward + gradient, B=32 TOTAL (strong scaling: 32/N clouds per
                                   GPU), n=m=2048 and 16384: clouds/s and fraction of the MUFU roofline
                  config5          recon loss path: chamfer_big + earth_mover forward + backward on 16384-point clouds, B=64
                                   TOTAL (64/N per GPU)
                  north_star       nn_distance forward at B=32 (per GPU), N=M=16384
                  config4          (rank 0) per-op milliseconds and HBM fraction of FPS / ball query / group / 3-NN / interpolate
                  ref_gpu_kernel   (rank 0, N=1) the reference's own .cu recompiled for sm_100a, timed on the same inputs
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, N, M = 32, 2048, 16384          # BASELINE.json configs[1], per GPU
WORKLOAD = "chamfer_nn_fwd+grad B=32/gpu N=2048(partial) M=16384(dense) fp32 xyz"
METRIC = "chamfer_nn_point_pairs_per_s"
UNIT = "Gpairs/s"
L2_BYTES = 126 * 1024 * 1024
LANE_OPS_PER_PAIR = 6.0
NN_FILTER_DRAM_BYTES = 17079040   # dram__bytes_read.sum + dram__bytes_write.sum of one nn_filter_kernel launch at the bench shape (profiles/r2_nn_filter_full.txt)
HBM_PEAK_FALLBACK = 6544.0         # GB/s, MEASURED_PEAKS.json of this pool (used when the file is absent)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the post-run extra measurements")
    ap.add_argument("--workload", default="chamfer", choices=["chamfer", "recon_loss"],
                    help="chamfer (default, the headline): BASELINE configs[1].  recon_loss: BASELINE configs[4], the recon_test loss "
                         "path -- chamfer_big + earth_mover forward + backward on 16384-point outputs vs GT, B=64 total -- in clouds/s")
    ap.add_argument("--data-dir", default=None, help="recon_loss only: a PCN-style directory with partial/<id>.pcd and complete/<id>.pcd "
                                                      "(recon_test.py:50-51); clouds are read with rfnet_b200.io_util instead of being synthesised")
    ap.add_argument("--list-path", default=None, help="recon_loss --data-dir: file with one model id per line (recon_test.py:46-47)")
    ap.add_argument("--results-dir", default="results", help="recon_loss --data-dir: where results.csv is written (recon_test.py:42-44)")
    return ap.parse_args()


def host_clouds(nclouds, npts, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    return (rng.random((nclouds, npts, 3), dtype=np.float32) - 0.5)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return HBM_PEAK_FALLBACK, "fallback: this pool's measured copy bandwidth (MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU NnDistance + NnDistanceGrad kernels on all host threads
# ------------------------------------------------------------------------------------------------------------------
_CPU_INPUTS = {}


def cpu_reference_rate(nclouds, repeats=1):
    """Gpairs/s of the reference CPU kernels over `nclouds` clouds of the workload shape, one cloud per task, all cores."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    from oracle import port, ref
    kind = "reference" if ref.available("cpu") else "port"
    cores = os.cpu_count() or 1
    if nclouds not in _CPU_INPUTS:
        _CPU_INPUTS[nclouds] = (host_clouds(nclouds, N, 1), host_clouds(nclouds, M, 2))
    x1, x2 = _CPU_INPUTS[nclouds]
    g1, g2 = np.ones((1, N), np.float32), np.ones((1, M), np.float32)

    def one(i):
        a, c = x1[i:i + 1], x2[i:i + 1]
        if kind == "reference":
            d1, i1, d2, i2 = ref.nn_distance(a, c)             # ctypes releases the GIL: tasks run in parallel
            ref.nn_distance_grad(a, c, g1, i1, g2, i2)
        else:
            d1, i1, d2, i2 = port.nn_distance(a, c, fused=False)
            port.nn_distance_grad(a, c, g1, i1, g2, i2)
        return float(d1[0, 0])

    best = None
    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(one, range(min(cores, nclouds))))          # warm-up (page in, spin up threads)
        for _ in range(repeats):
            t0 = time.perf_counter()
            list(ex.map(one, range(nclouds)))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    pairs = 2.0 * nclouds * N * M
    return pairs / best / 1e9, best, cores, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    nclouds = max(cores, 8)                                    # bounded sample per step: ~0.15 core-seconds per cloud
    rates, times = [], []
    for s in range(args.warmup + args.steps):
        rate, dt, cores, kind = cpu_reference_rate(nclouds)
        if s >= args.warmup:
            rates.append(rate)
            times.append(dt)
    value = 2.0 * nclouds * N * M * len(times) / sum(times) / 1e9
    sample = "%d clouds of the workload shape per step (of %d per GPU batch), one cloud per task on %d host threads" % (nclouds, B, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs during the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag, self.ok = index, [], set(), None, threading.Event(), False
        self.period, self.power = period, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag.is_set():
            self.sample()
            time.sleep(self.period)

    def stop(self):
        self.stop_flag.set()
        self.join()

    def result(self):
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}
        if self.power:
            out["power_w_max"] = max(self.power)
        return out


# ------------------------------------------------------------------------------------------------------------------
# helpers shared by both workloads
# ------------------------------------------------------------------------------------------------------------------
class Dist:
    """Rank bookkeeping + the two collectives the bench itself needs (barrier, max of a timing)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        if self.world == 1:
            return float(ms)
        t = self.torch.tensor([ms], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, iters, warm=1):
        """ms per call of fn: `warm` untimed calls, barrier + synchronize, `iters` calls between two CUDA events on the
        current stream, barrier + synchronize, MAX over ranks."""
        torch = self.torch
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1)) / iters

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def count_rfnet_launches(torch, fn):
    """number of OUR kernels (namespace rfnet::) one call of fn launches, from CUPTI activity records"""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        return sum(1 for e in prof.events() if "rfnet" in e.name)
    except Exception:
        return None


def mufu_peak(sm_max_hz):
    return 148 * 16 * sm_max_hz


def emd_shard_step(D, n, clouds_total, seed, flags=0):
    """EMD forward + gradient on this rank's slice of `clouds_total` clouds of n points (cost + both MatchCostGrad gradients,
    no match matrix) followed by the path's only collective: the all-reduce of [sum cost / n, count] (earth_mover,
    vv_recon.py:392-399).  Returns (callable, clouds on this rank)."""
    torch, dist = D.torch, D.dist
    from rfnet_b200 import losses, ops
    lo, hi = losses.shard_bounds(clouds_total, D.rank, D.world)
    nb = hi - lo
    g = torch.Generator(device="cpu").manual_seed(seed)
    x1 = (torch.rand((clouds_total, n, 3), generator=g) - 0.5)[lo:hi].contiguous().to(D.dev)
    x2 = (torch.rand((clouds_total, n, 3), generator=g) - 0.5)[lo:hi].contiguous().to(D.dev)
    res = torch.zeros(2, device=D.dev)

    def step():
        if nb:
            cost, g1, g2 = ops.emd_cost_grad_op(x1, x2, flags)
            part = torch.stack([(cost / float(n)).sum(), cost.new_tensor(float(nb))])
        else:
            part = torch.zeros(2, device=D.dev)
        if D.world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        res.copy_(part)
    return step, nb


def recon_shard_step(D, n, clouds_total, seed, host_inputs=None):
    """BASELINE configs[4]: chamfer_big + earth_mover of 16384-point outputs vs GT, forward + backward w.r.t. the output cloud,
    on this rank's slice of `clouds_total` clouds, then the all-reduce of the six partial sums (vv_recon.py:381-399,484-493)."""
    torch, dist = D.torch, D.dist
    from rfnet_b200 import losses, ops
    lo, hi = losses.shard_bounds(clouds_total, D.rank, D.world)
    nb = hi - lo
    g = torch.Generator(device="cpu").manual_seed(seed)
    if host_inputs is None:
        out = (torch.rand((clouds_total, n, 3), generator=g) - 0.5)[lo:hi].contiguous().to(D.dev)
        gt = (torch.rand((clouds_total, n, 3), generator=g) - 0.5)[lo:hi].contiguous().to(D.dev)
    else:   # device buffers the caller refills from pinned host memory every step
        out, gt = torch.empty_like(host_inputs[0], device=D.dev), torch.empty_like(host_inputs[1], device=D.dev)
    res = torch.zeros(6, device=D.dev)
    f32, i32 = torch.float32, torch.int32
    nbm = max(nb, 1)
    o = dict(d1=torch.empty((nbm, n), dtype=f32, device=D.dev), i1=torch.empty((nbm, n), dtype=i32, device=D.dev),
             d2=torch.empty((nbm, n), dtype=f32, device=D.dev), i2=torch.empty((nbm, n), dtype=i32, device=D.dev),
             g1=torch.empty((nbm, n, 3), dtype=f32, device=D.dev), g2=torch.empty((nbm, n, 3), dtype=f32, device=D.dev), s=torch.zeros(4, device=D.dev))
    ws = torch.empty(ops.nn_distance_workspace_bytes(nbm, n, n), dtype=torch.uint8, device=D.dev)
    gd = torch.full((nbm, n), 0.25 / (clouds_total * n), device=D.dev)   # upstream of (mean sqrt d1 + mean sqrt d2) / 2 is applied by the caller

    def step():
        if nb:
            ops.raw_chamfer_step(out, gt, gd, gd, o["d1"], o["i1"], o["d2"], o["i2"], o["g1"], o["g2"], o["s"], ws)
            cost, e1, e2 = ops.emd_cost_grad_op(out, gt)
            part = torch.cat([o["s"], torch.stack([(cost / float(n)).sum(), cost.new_tensor(float(nb))])])
        else:
            part = torch.zeros(6, device=D.dev)
        if D.world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        res.copy_(part)
    if host_inputs is not None:
        return step, nb, res, (out, gt)
    return step, nb, res


# ------------------------------------------------------------------------------------------------------------------
# second workload: BASELINE configs[4] (recon_test loss path), EMD-dominated -> clouds/s and the MUFU roofline
# ------------------------------------------------------------------------------------------------------------------
RB_TOTAL, RN = 64, 16384   # clouds in total (sharded over the ranks), points per cloud


def emd_cpu_reference_rate(nclouds, n):
    """clouds/s of the reference's CPU ApproxMatch + MatchCost kernels at n x n on all host threads (one cloud per task)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import port, ref
    kind = "reference" if ref.available("cpu") else "port"
    cores = os.cpu_count() or 1
    x1, x2 = host_clouds(nclouds, n, 11), host_clouds(nclouds, n, 12)

    def one(i):
        a, c = x1[i:i + 1], x2[i:i + 1]
        if kind == "reference":
            return float(ref.match_cost(a, c, ref.approx_match(a, c))[0])
        return float(port.match_cost(a, c, port.approx_match(a, c))[0])

    with ThreadPoolExecutor(max_workers=cores) as ex:
        t0 = time.perf_counter()
        list(ex.map(one, range(nclouds)))
        dt = time.perf_counter() - t0
    return nclouds / dt, dt, cores, kind


def run_recon_files(args):
    """recon_test.py's loss loop on real files (recon_test.py:46-68): per model read partial/complete PCDs, resample the
    partial cloud to 3000 points, evaluate cd = chamfer_big(output, gt) and emd = fidelity_loss(partial, output) as the
    reference does, and write results.csv.  There is no network here, so `output` is <data-dir>/completion/<id>.pcd when it
    exists, else the ground truth with N(0, 0.01^2) noise (said so in the JSON line)."""
    import numpy as np
    import torch
    from rfnet_b200 import io_util, losses
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    ids = [l.strip() for l in open(args.list_path)] if args.list_path else sorted(
        os.path.relpath(os.path.join(r, f), os.path.join(args.data_dir, "complete"))[:-4]
        for r, _, fs in os.walk(os.path.join(args.data_dir, "complete")) for f in fs if f.endswith(".pcd"))
    ids = [i for i in ids if i]
    rng = np.random.default_rng(0)
    res = io_util.ResultsCsv(args.results_dir)
    synthetic_outputs = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for mid in ids:
        partial = io_util.resample_pcd(io_util.read_pcd(os.path.join(args.data_dir, "partial", mid + ".pcd")), 3000, rng)
        complete = io_util.read_pcd(os.path.join(args.data_dir, "complete", mid + ".pcd"))
        cpath = os.path.join(args.data_dir, "completion", mid + ".pcd")
        if os.path.exists(cpath):
            output = io_util.read_pcd(cpath)
        else:
            output = complete + rng.normal(0, 0.01, complete.shape)
            synthetic_outputs += 1
        tp = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.float32))[None].to(dev)
        tg = torch.from_numpy(np.ascontiguousarray(complete, dtype=np.float32))[None].to(dev)
        to = torch.from_numpy(np.ascontiguousarray(output, dtype=np.float32))[None].to(dev)
        cd, _ = losses.chamfer_big(to, tg)                     # recon_test.py:27
        emd = losses.fidelity_loss(tp, to)                     # recon_test.py:28 (the column is named emd there)
        res.add(mid if "/" in mid else "all/" + mid, float(cd), float(emd))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    per_cat = res.close()
    print(json.dumps({"metric": "recon_test_models_per_s", "value": len(ids) / dt if dt > 0 else None, "unit": "models/s", "n_gpus": 1,
                      "higher_is_better": True, "data": "files", "dtype": "f32",
                      "config": {"workload": "recon_test.py loss loop over PCD files", "data_dir": args.data_dir, "models": len(ids),
                                 "results_csv": os.path.join(args.results_dir, "results.csv"), "outputs_synthesised": synthetic_outputs},
                      "per_category_cd_emd": per_cat}), flush=True)
    return 0


def run_recon_loss(args):
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        n_s = 2048   # the reference CPU kernel needs ~40 s and 2 GiB per 16384^2 cloud: time 2048^2 clouds, scale by (2048/16384)^2
        cores = os.cpu_count() or 1
        rates = []
        for s_ in range(args.warmup + args.steps):
            r, dt, cores, kind = emd_cpu_reference_rate(cores, n_s)
            if s_ >= args.warmup:
                rates.append(r)
        value = float(sum(rates) / len(rates)) / 64.0
        sample = "%d clouds of %d^2 per step on %d host threads; clouds/s scaled by 1/64 to 16384^2 (EXTRAPOLATED: work is n*m)" % (cores, n_s, cores)
        print(json.dumps({"impl": "reference", "metric": "recon_loss_clouds_per_s", "value": value, "unit": "clouds/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": "recon_test loss path CD+EMD 16384 pts", "sample": sample},
                          "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return 0
    if args.data_dir:
        return run_recon_files(args)
    from rfnet_b200 import _lib
    _lib.load()
    D = Dist()
    torch = D.torch
    steps = min(args.steps, 20)
    warm = min(max(args.warmup, 3), 5)
    step, nb, res = recon_shard_step(D, RN, RB_TOTAL, 500)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=D.dev)

    def flushed_step():
        flush.zero_()          # inputs (2 x 12.6 MB at 64 clouds) fit L2: every step first overwrites a 256 MiB buffer
        step()

    for _ in range(warm):
        flushed_step()
    launches = count_rfnet_launches(torch, step)
    sampler = ClockSampler(D.local)
    sampler.sample()
    sampler.start()
    ms = D.timed(flushed_step, steps, warm=0)
    sampler.stop()
    value = RB_TOTAL / (ms * 1e-3)
    clocks = sampler.result()
    # e2e: the same step with this rank's clouds coming from pinned HOST memory every step and the six loss sums read back
    from rfnet_b200 import losses as _losses
    lo, hi = _losses.shard_bounds(RB_TOTAL, D.rank, D.world)
    g = torch.Generator(device="cpu").manual_seed(500)
    h_out = (torch.rand((RB_TOTAL, RN, 3), generator=g) - 0.5)[lo:hi].contiguous().pin_memory()
    h_gt = (torch.rand((RB_TOTAL, RN, 3), generator=g) - 0.5)[lo:hi].contiguous().pin_memory()
    host_res = torch.empty(6).pin_memory()
    e_step, _, e_res, e_bufs = recon_shard_step(D, RN, RB_TOTAL, 500, host_inputs=(h_out, h_gt))

    def e2e_step():
        e_bufs[0].copy_(h_out, non_blocking=True)
        e_bufs[1].copy_(h_gt, non_blocking=True)
        e_step()
        host_res.copy_(e_res, non_blocking=True)
    k_e2e = max(2, min(steps, 5))
    ms_e2e = D.timed(e2e_step, k_e2e, warm=1)
    e2e_value = RB_TOTAL / (ms_e2e * 1e-3)
    sm_max = (clocks["sm_max_mhz"] or 1965) * 1e6
    achieved = value / D.world * 30.0 * RN * RN
    line = {"metric": "recon_loss_clouds_per_s", "value": value, "unit": "clouds/s", "n_gpus": D.world, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "recon_test loss path: chamfer_big + earth_mover forward + backward, %d clouds TOTAL of %d points vs GT (BASELINE configs[4])" % (RB_TOTAL, RN),
                       "parallelism": "batch-sharded x%d (%d clouds on rank 0), six loss partial sums all-reduced" % (D.world, nb),
                       "l2": "flushed: every step begins by overwriting a 256 MiB buffer (2 x the 126 MB L2), inside the timed region",
                       "emd": "rfnet_emd_cost_grad: cost and both gradients without any (b, m, n) match tensor"},
            "clocks": clocks, "gpu_launches": (launches or 0) * steps, "gpu_launches_per_step": launches,
            "e2e": {"value": e2e_value, "unit": "clouds/s", "h2d_bytes_per_step": 2 * (hi - lo) * RN * 12, "d2h_bytes_per_step": 24, "steps": k_e2e,
                    "api": "per step: this rank's output and GT clouds copied from pinned host memory, rfnet_chamfer_step + rfnet_emd_cost_grad, six loss sums read back"},
            "roofline": {"bound": "mufu_ex2_pipe (co-limited with the FP32 pipe)", "kernel": "rfnet::emd_sweep_kernel (21 sweeps) + emd_pair_kernel x2, nn_search",
                         "achieved": achieved / 1e12, "peak": mufu_peak(sm_max) / 1e12, "unit": "Tex2/s", "frac": achieved / mufu_peak(sm_max),
                         "peak_source": "148 SMs x 16 MUFU lanes x %.0f MHz (architectural)" % (sm_max / 1e6), "algorithmic_ex2_per_cloud": 30.0 * RN * RN,
                         "traffic": None}}
    if D.rank == 0 and D.world == 1:
        try:
            r, dt, cores, kind = emd_cpu_reference_rate(os.cpu_count() or 1, 2048)
            line["cpu_baseline"] = {"value": r / 64.0, "unit": "clouds/s", "cores": cores, "kind": kind,
                                    "sample": "%d clouds of 2048^2 (approx_match + match_cost CPU kernels), one per thread, %.1f s; scaled by 1/64 to 16384^2 (EXTRAPOLATED)" % (cores, dt)}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "clouds/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}
    if D.rank == 0:
        print(json.dumps(line), flush=True)
    D.close()
    return 0


# ------------------------------------------------------------------------------------------------------------------
# extras
# ------------------------------------------------------------------------------------------------------------------
def config4_block(D, sm_max, use_graph=True):
    """BASELINE configs[3] (the sampling / grouping / interpolation chain at B=32) on rank 0: ms per op and, for the HBM-bound
    ones, algorithmic bytes / time against the measured copy bandwidth (SURVEY.md 8d byte counts)."""
    torch = D.torch
    from rfnet_b200 import ops, tf_grouping, tf_interpolate, tf_sampling
    peak, peak_src = hbm_peak()
    b, n, m, ns, c = 32, 16384, 2048, 32, 64
    g = torch.Generator(device="cpu").manual_seed(6)
    x = (torch.rand((b, n, 3), generator=g) - 0.5).to(D.dev)

    def t(fn, iters=10):
        """ms per call on the device: `iters` calls captured into one CUDA graph and replayed (several of these ops take 30-60 us, about
        what a Python custom-op dispatch costs, so an eager loop would time the host); eager back-to-back loop if capture fails."""
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        try:
            if not use_graph or os.environ.get("RFNET_BENCH_EAGER"):      # e.g. under ncu, or next to live NCCL communicators
                raise RuntimeError("eager timing requested")
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(iters):
                    fn()
            graph.replay()
            torch.cuda.synchronize()
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            del graph
            return e0.elapsed_time(e1) / iters
        except Exception:
            torch.cuda.synchronize()
            eager_timed.append(True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    eager_timed = []
    out = {"shape": "B=32, 16384 -> 2048 FPS points, r=0.1, nsample=32, c=3 and c=64", "hbm_peak_gbs": peak, "hbm_peak_source": peak_src}
    ms = t(lambda: tf_sampling.farthest_point_sample(m, x), 5)
    out["farthest_point_sample"] = {"ms": ms, "us_per_pick": ms * 1e3 / m}
    idx = tf_sampling.farthest_point_sample(m, x)
    q = tf_sampling.gather_point(x, idx)
    radius = torch.tensor([0.1], dtype=torch.float32, device=D.dev)   # a tensor input in the reference too (tf_grouping.cpp:93-95)
    ms = t(lambda: tf_grouping.query_ball_point(radius, ns, x, q))
    out["query_ball_point"] = {"ms": ms}
    gi, _ = tf_grouping.query_ball_point(radius, ns, x, q)
    for cc in (3, c):
        pts = torch.randn((b, n, cc), generator=g).to(D.dev)
        ms = t(lambda: tf_grouping.group_point(pts, gi))
        byts = 4.0 * b * m * ns * (cc + 1) + 4.0 * b * n * cc
        out["group_point_c%d" % cc] = {"ms": ms, "GB_per_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak}
        go = torch.randn((b, m, ns, cc), generator=g).to(D.dev)
        ms = t(lambda: ops.group_point_grad_op(pts, gi, go))
        out["group_point_grad_c%d" % cc] = {"ms": ms, "GB_per_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak,
                                            "note": "standalone call: builds the inverted index (6 small launches) and sums"}
        plan = ops.scatter_plan_op(gi.reshape(b, -1), n)
        ms = t(lambda: ops.group_point_grad_planned_op(go, plan, n))
        out["group_point_grad_c%d_planned" % cc] = {"ms": ms, "GB_per_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak,
                                                    "note": "what autograd runs: the inverted index was built at forward time"}
        del pts, go, plan
    ms = t(lambda: tf_interpolate.three_nn(x, q))
    out["three_nn"] = {"ms": ms}
    d3, i3 = tf_interpolate.three_nn(x, q)
    w = 1.0 / torch.clamp(d3, min=1e-10)
    w = w / w.sum(-1, keepdim=True)
    feats = torch.randn((b, m, c), generator=g).to(D.dev)
    ms = t(lambda: tf_interpolate.three_interpolate(feats, i3, w))
    byts = 4.0 * b * n * (c + 6) + 4.0 * b * m * c
    out["three_interpolate_c%d" % c] = {"ms": ms, "GB_per_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak}
    go = torch.randn((b, n, c), generator=g).to(D.dev)
    ms = t(lambda: ops.three_interpolate_grad_op(feats, i3, w, go))
    out["three_interpolate_grad_c%d" % c] = {"ms": ms, "GB_per_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak}
    plan = ops.scatter_plan_op(i3.reshape(b, -1), m)
    ms = t(lambda: ops.three_interpolate_grad_planned_op(go, w, plan, m))
    out["three_interpolate_grad_c%d_planned" % c] = {"ms": ms, "GB_per_s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peak}
    if use_graph and not os.environ.get("RFNET_BENCH_EAGER"):
        out["timing"] = "CUDA-graph replay of 10 calls per op (device time)" + (", %d op(s) timed as an eager loop" % len(eager_timed) if eager_timed else "")
    else:
        out["timing"] = "eager loop of 10 calls per op (ops under ~60 us are bounded by the Python dispatch, not the device; the 1-GPU line replays CUDA graphs)"
    return out


def ref_gpu_kernel_block(D, d1s, d2s):
    """The kernel to beat: the reference's own tf_nndistance .cu recompiled unchanged for sm_100a
    (oracle/_ref/libref_gpu.so, the parity checker), timed on the bench's inputs after the timed region."""
    torch = D.torch
    try:
        from oracle import ref
        if not ref.available("gpu"):
            return {"unavailable": "oracle/_ref/libref_gpu.so not built"}
        x1, x2 = d1s[0], d2s[0]
        spec = [((B, N), torch.float32), ((B, N), torch.int32), ((B, M), torch.float32), ((B, M), torch.int32)]
        for _ in range(2):
            ref.run_gpu("NnDistance", [x1, x2], spec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ref.run_gpu("NnDistance", [x1, x2], spec)
        e1.record()
        torch.cuda.synchronize()
        return {"nn_distance_fwd_ms": e0.elapsed_time(e1) / 5, "what": "NmDistanceKernel (tf_ops/CD/tf_nndistance_g.cu:4-130) recompiled for sm_100a, B=32 2048 vs 16384"}
    except Exception as ex:
        return {"unavailable": repr(ex)}


def main():
    args = parse()
    if args.workload == "recon_loss":
        return run_recon_loss(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- rfnet_b200 has no CPU path (use --impl reference for the CPU arm)")
    from rfnet_b200 import _lib, losses, ops
    lib = _lib.load()
    D = Dist()
    dist, world, rank, dev = D.dist, D.world, D.rank, D.dev

    # ---- inputs: enough distinct batches that consecutive steps never find their inputs in L2 (rotating sets > L2)
    set_bytes = B * (N + M) * 12
    nsets = max(2, -(-int(1.5 * L2_BYTES) // set_bytes))
    h1 = torch.from_numpy(host_clouds(nsets * B, N, 1000 * 2 + rank)).reshape(nsets, B, N, 3).pin_memory()
    h2 = torch.from_numpy(host_clouds(nsets * B, M, 7000 * 2 + rank)).reshape(nsets, B, M, 3).pin_memory()
    d1s, d2s = h1.to(dev), h2.to(dev)
    gd1 = torch.full((B, N), 0.5 / (B * N * world), device=dev)   # upstream grads of a mean-type loss
    gd2 = torch.full((B, M), 0.5 / (B * M * world), device=dev)
    sums = torch.zeros(4, device=dev)
    f32, i32 = torch.float32, torch.int32
    o = dict(d1=torch.empty((B, N), dtype=f32, device=dev), i1=torch.empty((B, N), dtype=i32, device=dev),
             d2=torch.empty((B, M), dtype=f32, device=dev), i2=torch.empty((B, M), dtype=i32, device=dev),
             g1=torch.empty((B, N, 3), dtype=f32, device=dev), g2=torch.empty((B, M, 3), dtype=f32, device=dev))
    parts = [torch.zeros(4, device=dev) for _ in range(2)]
    ws = torch.empty(ops.nn_distance_workspace_bytes(B, N, M), dtype=torch.uint8, device=dev)
    pending = []

    def step(i):
        part = parts[i & 1]
        # one C-ABI call: search, then gradient + unpack + sqrt partial sums, then the fixed-order reduction (4 launches with the preparation of the filtered search)
        ops.raw_chamfer_step(d1s[i % nsets], d2s[i % nsets], gd1, gd2, o["d1"], o["i1"], o["d2"], o["i2"], o["g1"], o["g2"], part, ws)
        if world > 1:
            # the path's only collective: 16 bytes over NCCL/NVLink, on NCCL's own stream.  Nothing on the device consumes
            # the reduced loss, so the compute stream only joins it one step later (it overlaps the next step's kernels).
            pending.append((dist.all_reduce(part, op=dist.ReduceOp.SUM, async_op=True), part))
            if len(pending) > 1:
                work, done = pending.pop(0)
                work.wait()
                sums.copy_(done)
        else:
            sums.copy_(part)

    def flush():
        while pending:
            work, done = pending.pop(0)
            work.wait()
            sums.copy_(done)

    for i in range(args.warmup):
        step(i)
    flush()
    D.barrier()
    launches_per_step = count_rfnet_launches(torch, lambda: (step(0), flush()))
    D.barrier()

    sampler = ClockSampler(D.local)
    sampler.sample()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    flush()                                                         # every step's all-reduce has joined the compute stream
    e1.record()
    D.barrier()
    sampler.stop()
    sampler.sample()
    ms = D.max_ms(e0.elapsed_time(e1))
    pairs_per_step = 2.0 * B * N * M * world
    value = pairs_per_step * args.steps / (ms * 1e-3) / 1e9
    clocks = sampler.result()
    sm_max = (clocks["sm_max_mhz"] or 1965) * 1e6

    # ---- dominant kernel against its roofline: the forward search alone, CUDA events on its stream -- the default (filtered
    # search: nn_prepare_kernel + nn_filter_kernel + key unpack) and, beside it, the direct kernel (the reference expression for
    # every pair: nn_search_kernel, r1's kernel, RFNET_NN_DIRECT)
    def time_forward(direct):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            ops.raw_nn_distance(d1s[i % nsets], d2s[i % nsets], o["d1"], o["i1"], o["d2"], o["i2"], ws, direct=direct)
        torch.cuda.synchronize()
        f0.record()
        for i in range(args.steps):
            ops.raw_nn_distance(d1s[i % nsets], d2s[i % nsets], o["d1"], o["i1"], o["d2"], o["i2"], ws, direct=direct)
        f1.record()
        torch.cuda.synchronize()
        return f0.elapsed_time(f1) / args.steps
    fwd_ms = time_forward(False)
    fwd_direct_ms = time_forward(True)
    peak = 148 * 128 * sm_max / 1e12                               # T lane-ops/s at max clock
    achieved = 2.0 * B * N * M * LANE_OPS_PER_PAIR / (fwd_ms * 1e-3) / 1e12
    # measured FP32 pipe peak: a dependency-free FFMA2 stream on every SM (rfnet_probe_fp32)
    import ctypes
    sink = torch.zeros(4, device=dev)
    lane_ops = ctypes.c_ulonglong(0)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.rfnet_probe_fp32(2000, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(lane_ops), stream)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    lib.rfnet_probe_fp32(20000, ctypes.c_void_p(sink.data_ptr()), ctypes.byref(lane_ops), stream)
    p1.record()
    torch.cuda.synchronize()
    peak_measured = lane_ops.value / (p0.elapsed_time(p1) * 1e-3) / 1e12
    roofline = {"bound": "fp32_fma_pipe",
                "kernel": "rfnet::nn_filter_kernel (timed as the forward call: + nn_prepare_kernel, key memset, key unpack, ~4%)",
                "achieved": achieved, "peak": peak, "unit": "Tlaneop/s", "frac": achieved / peak,
                "frac_note": "achieved counts the ALGORITHMIC 6 FP32 lane-ops per pair of the reference expression (SURVEY 8d).  The filtered search "
                             "visits every pair with the 3-lane-op expanded form and evaluates the reference expression only where it decides the "
                             "result (bit-identical outputs), so the fraction can exceed what a kernel that executes all 6 can reach; "
                             "frac_executed is the FP32 lane-ops it actually issues per pair (3 + tails) over the same peak",
                "executed_laneops_per_pair": 3.0, "frac_executed": achieved / peak * 3.0 / LANE_OPS_PER_PAIR,
                "peak_source": "148 SMs x 128 FP32 lanes x %.0f MHz (architectural; MEASURED_PEAKS.json has no FP32-pipe figure)" % (sm_max / 1e6),
                "peak_measured_ffma2_stream": peak_measured, "frac_of_measured": achieved / peak_measured,
                "algorithmic_laneops_per_pair": LANE_OPS_PER_PAIR, "fwd_ms": fwd_ms, "step_ms": ms / args.steps,
                "frac_whole_step": pairs_per_step / world * LANE_OPS_PER_PAIR / (ms / args.steps * 1e-3) / 1e12 / peak,
                "direct_kernel": {"kernel": "rfnet::nn_search_kernel (RFNET_NN_DIRECT: all 6 lane-ops for every pair; r1's kernel)", "fwd_ms": fwd_direct_ms,
                                  "frac": 2.0 * B * N * M * LANE_OPS_PER_PAIR / (fwd_direct_ms * 1e-3) / 1e12 / peak},
                "traffic": NN_FILTER_DRAM_BYTES, "traffic_note": "dram__bytes_read+write of one nn_filter_kernel launch at this config, ncu --set full (profiles/r2_nn_filter_full.txt); inputs are 7.08 MB + 9.4 MB of prepared rows, compute-bound"}

    # ---- e2e: the same step through the public host-buffer API (rfnet_b200.host.ChamferHostPipeline): every step copies
    # its inputs from pinned host memory and EVERY output of the operator back; copies of neighbouring steps overlap the
    # kernels on separate streams, the cross-rank all-reduce runs on a side stream
    from rfnet_b200.host import ChamferHostPipeline
    k_e2e = max(3, min(args.steps, 100))

    def run_e2e(outputs):
        pipe = ChamferHostPipeline(B, N, M, dev, depth=3, grad_scale1=0.5 / (B * N * world), grad_scale2=0.5 / (B * M * world), outputs=outputs)
        for i in range(3):
            pipe.submit(h1[i % nsets], h2[i % nsets], losses.all_reduce_scalars)
        pipe.drain()
        D.barrier()
        e0.record()
        for i in range(k_e2e):
            pipe.submit(h1[i % nsets], h2[i % nsets], losses.all_reduce_scalars)
        torch.cuda.current_stream().wait_stream(pipe.s_out)
        e1.record()
        D.barrier()
        last = pipe.wait((pipe.count - 1) % pipe.depth)
        loss = float((last["sums"][0] / last["sums"][1] + last["sums"][2] / last["sums"][3]) / 2)   # chamfer_big read back on the host
        v = pairs_per_step * k_e2e / (D.max_ms(e0.elapsed_time(e1)) * 1e-3) / 1e9
        return v, pipe.h2d_bytes, pipe.d2h_bytes, loss

    v_loop, h2d, d2h_loop, e2e_loss = run_e2e(("sums",))
    v_all, _, d2h_all, _ = run_e2e(ChamferHostPipeline.ALL_OUTPUTS)
    # e2e.value: what the reference's training / test loop does per step (vv_recon.py:424-428, recon_test.py:60) -- the batch goes
    # host -> device, the loss comes back.  e2e.all_outputs: the same call with EVERY output of NnDistance + NnDistanceGrad copied
    # back as well (dist, idx, gradients: 11.8 MB per step and GPU): an operator-level e2e, bound by the host's memory bandwidth
    # once several GPUs share it (18.9 MB per 0.5 ms step and GPU).
    e2e = {"value": v_loop, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_loop, "steps": k_e2e, "loss_read_back": e2e_loss,
           "api": "rfnet_b200.host.ChamferHostPipeline.submit(pinned xyz1, xyz2) -> pinned loss sums (chamfer_big read back on the host); 3-slot ring, "
                  "copies overlap compute, the cross-rank all-reduce runs on a side stream",
           "all_outputs": {"value": v_all, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_all,
                           "note": "same call, dist1, idx1, dist2, idx2, grad_xyz1, grad_xyz2 and the loss sums all copied to pinned host memory every step"}}

    # ---- extras (all ranks take part; rank 0 reports)
    extra = {}
    if not args.no_extra:
        # sustained: the same step for >= 2 s (the timed region above is a burst of tens of milliseconds)
        per = ms / args.steps
        k_sus = max(args.steps, int(2200.0 / per))
        s2 = ClockSampler(D.local, period=0.02)
        s2.start()
        D.barrier()
        e0.record()
        for i in range(k_sus):
            step(i)
        flush()
        e1.record()
        D.barrier()
        s2.stop()
        ms_sus = D.max_ms(e0.elapsed_time(e1))
        extra["sustained"] = {"steps": k_sus, "seconds": ms_sus / 1e3, "value": pairs_per_step * k_sus / (ms_sus * 1e-3) / 1e9, "unit": UNIT,
                              "ms_per_step": ms_sus / k_sus, "clocks": s2.result()}
        # north-star shape: nn_distance forward, B=32 per GPU, 16384 x 16384
        g = torch.Generator(device="cpu").manual_seed(5)
        y1 = (torch.rand((B, M, 3), generator=g) - 0.5).to(dev)
        y2 = (torch.rand((B, M, 3), generator=g) - 0.5).to(dev)
        ms_ns = D.timed(lambda: ops.nn_distance_op(y1, y2), 10)
        pr = 2.0 * B * M * M
        extra["north_star_chamfer_nn_fwd_B32_16384x16384"] = {"ms": ms_ns, "Gpairs_per_s_per_gpu": pr / ms_ns / 1e6, "frac_fp32_peak": pr * LANE_OPS_PER_PAIR / (ms_ns * 1e-3) / 1e12 / peak}
        del y1, y2
        # config 3: EMD forward + gradient, B=32 TOTAL sharded over the ranks (strong scaling), with the scalar all-reduce
        c3 = {"what": "approx_match + match_cost + match_cost_grad as one matrix-free call per rank (rfnet_emd_cost_grad) on 32/N clouds, "
                      "then the all-reduce of the loss partial sums; max over ranks", "scaling": "strong", "clouds_total": 32}
        for en, iters in ((2048, 10), (16384, 3)):
            for flags, tag in ((0, ""), (4, "_split_sums")):
                stp, nb = emd_shard_step(D, en, 32, 40 + en, flags)
                ms_e = D.timed(stp, iters, warm=2)
                cps = 32 / (ms_e * 1e-3)
                c3["n%d%s" % (en, tag)] = {"ms": ms_e, "clouds_per_s": cps, "clouds_per_gpu": nb, "launches": count_rfnet_launches(torch, stp),
                                           "frac_mufu_peak_per_gpu": cps / world * 30.0 * en * en / mufu_peak(sm_max)}
                del stp
        c3["modes"] = ("default: every sum one chain in the reference kernel's order (match within 1e-5 of the reference CUDA kernel, bit-identical with "
                       "RFNET_EMD_EXACT); *_split_sums: RFNET_EMD_SPLIT_SUMS, each sum cut into pieces (more parallelism per cloud, different "
                       "rounding order: up to 1e-3 of the largest match entry)")
        extra["config3_emd_fwd_grad_B32_total"] = c3
        # config 5: recon loss path, B=64 TOTAL
        stp, nb, _ = recon_shard_step(D, 16384, 64, 500)
        ms_r = D.timed(stp, 2, warm=1)
        extra["config5_recon_loss_fwd_bwd_B64_total"] = {"ms": ms_r, "clouds_per_s": 64 / (ms_r * 1e-3), "clouds_per_gpu": nb, "scaling": "strong",
                                                         "frac_mufu_peak_per_gpu": 64 / (ms_r * 1e-3) / world * 30.0 * 16384 * 16384 / mufu_peak(sm_max),
                                                         "what": "chamfer_big + earth_mover forward + backward on 16384-point clouds, six partial sums all-reduced"}
        del stp
        if rank == 0:
            try:
                extra["config4"] = config4_block(D, sm_max, use_graph=(world == 1))   # no stream capture next to live NCCL communicators
            except Exception as ex:
                extra["config4"] = {"error": repr(ex)}
            if world == 1:
                extra["ref_gpu_kernel"] = ref_gpu_kernel_block(D, d1s, d2s)
                if "nn_distance_fwd_ms" in extra["ref_gpu_kernel"]:
                    extra["ref_gpu_kernel"]["speedup_of_our_forward"] = extra["ref_gpu_kernel"]["nn_distance_fwd_ms"] / fwd_ms
        D.barrier()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": pairs_per_step, "clouds_per_gpu": B, "parallelism": "batch-sharded x%d, 16-byte loss all-reduce" % world,
                       "l2": "rotating %d input batches (%.0f MB > 126 MB L2), one per step" % (nsets, nsets * set_bytes / 1e6),
                       "step": "one rfnet_chamfer_step call: nn_prepare_kernel + nn_filter_kernel + chamfer_epilogue_kernel + chamfer_epilogue_final_kernel (+3 memsets)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": (launches_per_step or 0) * args.steps, "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "extra": extra}

    if rank == 0 and world == 1:
        try:
            cores = os.cpu_count() or 1
            rate, dt, cores, kind = cpu_reference_rate(max(2 * cores, 16), repeats=2)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "%d clouds of the workload shape (N=%d, M=%d), fwd+grad, one cloud per task on %d threads, best of 2 (%.1f s)" % (max(2 * cores, 16), N, M, cores, dt)}
        except Exception as ex:  # the oracle is test infrastructure; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    D.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
