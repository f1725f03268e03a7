#!/usr/bin/env python
"""bench.py — coverage fwd+bwd point x pose evals/s (BASELINE.json metric) on N B200s of one node.

Workload (config.workload): BASELINE config 4 — trajectory optimisation forward+backward,
64 body waypoints x 5 cameras = 320 evaluated poses, 100M-point synthetic box cloud.  The cloud is
point-sharded over the ranks (STRONG scaling: the total cloud is fixed, a rank holds N/G points);
per step every rank runs pass A (per-pose min/max) -> NCCL MIN/MAX all-reduce -> pass B (fused
log-odds fusion + gradient accumulators) -> NCCL SUM all-reduce -> O(W) epilogue -> torch chain
rule through the multi-camera front end to the 64x4 body parameters.

One JSON line on rank 0:
  value       evals/s with the cloud resident in HBM in the Morton order ModelTraj gives it once per cloud (CUDA
              events, barrier + sync both sides, max over ranks); evals = N points x W poses per objective+gradient
              evaluation (dense-equivalent: the exact tile pruning skips pairs that provably cannot matter)
  dense       the same step with pruning off (every pair fully evaluated, bit-identical rewards)
  e2e         same metric through the public API with HOST (pinned) inputs: every step copies this rank's cloud
              shard host->device, orders it (cov_spatial_sort), evaluates objective + gradient and reads loss +
              gradients back
  roofline    the dominant kernel of the timed region (pruned pass B, HBM-bound: 16 B/point) against the measured HBM
              peak, plus the dense kernels against the FP32 peak (FMA probe measured live; nominal alongside)
  cpu_baseline  oracle/torch_port.py (torch CPU autograd port of the reference) on a bounded sample, N=1 only
--impl reference times that CPU port alone (all host threads) on the same config/metric.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# frozen algorithmic work constants
FLOP_FWD, FLOP_BWD = 64, 86
BYTES_PASS_A, BYTES_PASS_B = 12, 16          # per point: xyz read; xyz read + rewards written
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
N_WAYPOINTS, N_CAMS = 64, 5
N_POINTS_C4 = 100_000_000
CHUNKS = 64                                   # cloud = 64 seeded chunks, so 1/2/4/8-rank clouds are identical
BOX_LO, BOX_HI = (-10.0, -10.0, -1.0), (30.0, 30.0, 4.0)


def body_waypoints(n=N_WAYPOINTS, length=35.0):
    """Gentle S-curve, spacing >= 0.5 m (so wps_step = 1), yaw = path tangent."""
    xs = np.linspace(0.0, length, n)
    ys = 0.5 * xs + 0.3 * np.sin(xs) - 8.0
    yaw = np.arctan2(0.5 + 0.3 * np.cos(xs), 1.0)
    return torch.tensor(np.stack([xs - 5.0, ys, np.zeros(n), yaw], 1), dtype=torch.float32)


def make_cloud_shard(n_total, rank, world, device):
    per_chunk = n_total // CHUNKS
    assert per_chunk * CHUNKS == n_total and CHUNKS % world == 0
    lo = torch.tensor(BOX_LO, device=device)
    hi = torch.tensor(BOX_HI, device=device)
    own = range(rank * CHUNKS // world, (rank + 1) * CHUNKS // world)
    out = torch.empty(per_chunk * len(own), 3, device=device)
    g = torch.Generator(device=device)
    for i, c in enumerate(own):
        g.manual_seed(1000 + c)
        out[i * per_chunk:(i + 1) * per_chunk] = torch.rand(per_chunk, 3, generator=g, device=device) * (hi - lo) + lo
    return out


class ClockSampler:
    """nvidia-smi SM clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# fixed sample of the workload for the CPU arm: 20 000 points x all 320 poses per step (the env override exists for the
# contract test in tests/, which only checks the line's shape)
CPU_SAMPLE_POINTS = int(os.environ.get("COV_BENCH_REF_POINTS", "20000"))


def cpu_reference_rate(body, rig, steps, warmup, n_points=CPU_SAMPLE_POINTS):
    """evals/s of the torch-CPU port on a FIXED, stated sample of the same workload: all 320 poses, the first
    `n_points` points of the seeded cloud generator (torch autograd keeps ~240 B per point per pose, which bounds the
    sample).  Fixed so that the figure is reproducible from run to run."""
    from oracle import torch_port, coverage_oracle as orc
    from trajectory_optimization_b200 import multicam
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    K = torch.from_numpy(orc.K_DEFAULT.copy())
    with torch.no_grad():
        t, q = multicam.camera_poses_from_body(body, rig)
    P, Q = t.reshape(-1, 3).contiguous(), q.reshape(-1, 4).contiguous()
    W = P.shape[0]
    g = torch.Generator().manual_seed(1000)
    lo, hi = torch.tensor(BOX_LO), torch.tensor(BOX_HI)
    pts = torch.rand(n_points, 3, generator=g) * (hi - lo) + lo
    for _ in range(warmup):
        torch_port.traj_step(pts, P, Q, K, orc.IMG_WIDTH, orc.IMG_HEIGHT)
    t0 = time.perf_counter()
    for _ in range(steps):
        torch_port.traj_step(pts, P, Q, K, orc.IMG_WIDTH, orc.IMG_HEIGHT)
    dt = time.perf_counter() - t0
    return dict(value=n_points * W * steps / dt, unit="point*pose evals/s", cores=threads, kind="port",
                sample=f"{n_points} points x {W} poses (fixed sample of the c4 cloud generator), {steps} fwd+bwd steps of "
                       f"oracle/torch_port.py (torch {torch.__version__} CPU autograd, {threads} threads), "
                       f"{dt / steps * 1e3:.0f} ms/step"), dt / steps


N_POINTS_C5, N_TRAJ_C5, N_WPS_C5 = 50_000_000, 1024, 32


def c5_trajectories(n_traj=N_TRAJ_C5, n_wps=N_WPS_C5):
    """1024 candidate trajectories = the base S-curve + N(0, 1 m) lateral offsets and N(0, 0.2 rad) yaw jitter, seed 2
    (SURVEY.md 8d); one camera per waypoint.  Returns poses (T, P, 3), quats (T, P, 4) as fp32 torch tensors."""
    gen = np.random.default_rng(2)
    base = body_waypoints(n_wps, 20.0).numpy()
    poses = np.repeat(base[None, :, :3], n_traj, 0) + gen.normal(0, 1.0, (n_traj, 1, 3)) * np.array([1, 1, 0])
    yaw = base[None, :, 3] + gen.normal(0, 0.2, (n_traj, n_wps))
    quats = np.stack([np.cos(yaw / 2), 0 * yaw, 0 * yaw, np.sin(yaw / 2)], -1)
    return torch.tensor(poses, dtype=torch.float32), torch.tensor(quats, dtype=torch.float32)


def run_c5(args):
    """BASELINE config 5: batched candidate-trajectory sweep, 1024 trajectories x 32 waypoints on a 50M-point cloud,
    forward only (two passes: per-pose normalisers, then per-trajectory log-odds fusion), at 1/2/4/8 GPUs.
    Sharding: the cloud fits one GPU (600 MB), so every rank holds the WHOLE cloud and evaluates its own
    block of trajectories: no collective on the data path, one all-gather of 1024 doubles.  Strong scaling."""
    from oracle import coverage_oracle as orc
    n_total, T, Pn = args.points, N_TRAJ_C5, N_WPS_C5
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = {"workload": "c5: candidate-trajectory sweep, %d trajectories x %d waypoints (one camera each), forward only, "
                       "%gM-point synthetic box cloud; trajectories sharded over %d GPU(s), whole cloud on every GPU"
                       % (T, Pn, n_total / 1e6, world),
           "n_points": n_total, "n_trajectories": T, "poses_per_trajectory": Pn, "parallelism": f"trajectories/{world}",
           "l2_policy": "inputs larger than L2 (600 MB of points streamed by every pass)"}
    poses, quats = c5_trajectories()
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import torch_port
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        K = torch.from_numpy(orc.K_DEFAULT.copy())
        g = torch.Generator().manual_seed(1000)
        n_s, t_s = CPU_SAMPLE_POINTS, 8
        pts = torch.rand(n_s, 3, generator=g) * (torch.tensor(BOX_HI) - torch.tensor(BOX_LO)) + torch.tensor(BOX_LO)

        def cpu_step():
            with torch.no_grad():
                return [torch_port.traj_vis_loss(pts, poses[t], quats[t], K, orc.IMG_WIDTH, orc.IMG_HEIGHT)[0] for t in range(t_s)]

        for _ in range(args.warmup):
            cpu_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_step()
        dt = (time.perf_counter() - t0) / args.steps
        val = n_s * t_s * Pn / dt
        base = dict(value=val, unit="point*pose evals/s", cores=threads, kind="port",
                    sample=f"{n_s} points x {t_s} of the {T} trajectories x {Pn} poses per step, forward only, "
                           f"oracle/torch_port.py on {threads} threads, {dt * 1e3:.0f} ms/step")
        print(json.dumps({"impl": "reference", "metric": "candidate sweep fwd point*pose evals/s", "value": val,
                          "unit": "point*pose evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": dict(cfg, same_config=False), "cpu_baseline": base,
                          "e2e": {"value": val, "unit": "point*pose evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    import torch.distributed as dist
    from trajectory_optimization_b200 import _lib, ops, tools
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    _lib.lib()
    K, img_w, img_h = tools.load_intrinsics(dev)
    pts = make_cloud_shard(n_total, 0, 1, dev)             # the WHOLE cloud on every rank (identical seeds)
    spts, _ = ops.spatial_sort(pts)
    del pts
    boxes = ops.tile_boxes(spts)
    P, Q = poses.to(dev), quats.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def sweep(Pd, Qd):
        return ops.sweep_rewards(spts, Pd, Qd, K, img_w, img_h, boxes=boxes, presorted=True, group=group, shard="trajectories")

    for _ in range(args.warmup):
        res = sweep(P, Q)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: sweep(P, Q), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    pairs = float(n_total) * T * Pn
    value = pairs * args.steps / (ms_total * 1e-3)
    # end to end: the candidate poses come from pinned host memory every step, the T means go back to the host
    host_P = torch.empty(poses.shape, dtype=torch.float32, pin_memory=True).copy_(poses)
    host_Q = torch.empty(quats.shape, dtype=torch.float32, pin_memory=True).copy_(quats)
    host_out = torch.empty(T, dtype=torch.float64, pin_memory=True)

    def e2e_step():
        P.copy_(host_P, non_blocking=True)
        Q.copy_(host_Q, non_blocking=True)
        host_out.copy_(sweep(P, Q), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    # every pair evaluated (pruning off) on a sample of 32 trajectories; same means required
    t_s = 32
    with ops.evaluation(dense=True):
        res_d = ops.sweep_rewards(spts, P[:t_s], Q[:t_s], K, img_w, img_h, boxes=boxes, presorted=True)
        ms_d = timed(lambda: ops.sweep_rewards(spts, P[:t_s], Q[:t_s], K, img_w, img_h, boxes=boxes, presorted=True), 1)
    rel = float(((res[:t_s] - res_d).abs() / res_d.abs()).max())
    if rel > 1e-9:
        raise SystemExit("bench.py c5: pruned sweep differs from the dense sweep by %.2e" % rel)
    dense_tf = 2.0 * n_total * t_s * Pn * FLOP_FWD / (ms_d * 1e-3) / 1e12   # two forward passes per pair
    if rank != 0:
        sys.stdout.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)
        return
    line = {"metric": "candidate sweep fwd point*pose evals/s", "value": value, "unit": "point*pose evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "value_is": "DENSE-EQUIVALENT rate: N x T x P (point, pose) pairs per sweep / sweep time; the exact pruning skips "
                        "pairs that provably cannot matter (see `dense`)",
            "config": cfg, "clocks": clocks,
            "dense": {"ms_for_sample": ms_d, "sample": "%d of the %d trajectories, every pair evaluated in both passes" % (t_s, T),
                      "value": float(n_total) * t_s * Pn / (ms_d * 1e-3), "unit": "point*pose evals/s",
                      "extrapolated_full_sweep_s": ms_d * 1e-3 * T / t_s},
            "parity_check": {"max_rel_diff_pruned_vs_dense_means": rel, "trajectories_compared": t_s},
            "e2e": {"value": pairs * args.steps / (ms_e2e * 1e-3), "unit": "point*pose evals/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": (host_P.numel() + host_Q.numel()) * 4 * world, "d2h_bytes_per_step": T * 8 * world,
                    "note": "candidate poses from pinned host memory in, per-trajectory means out, every step; the cloud is resident"},
            "gpu_launches": None,
            "roofline": {"kernel": "dense sweep kernels on the 32-trajectory sample (cov_traj_minmax_kernel + cov_sweep_kernel)",
                         "bound": "fp32", "unit": "TFLOP/s", "achieved": dense_tf, "peak": FP32_NOMINAL_TFLOPS,
                         "frac": dense_tf / FP32_NOMINAL_TFLOPS, "traffic": None,
                         "note": "FP32 CUDA-core bound (no tensor-core work: the transform is 3x4); peak = nominal FP32 "
                                 "148 SM x 128 lanes x 2 x 1.965 GHz; 64 flop per forward evaluation, two passes"},
            "cpu_baseline": None, "mean_reward_range": [float(res.min()), float(res.max())]}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


ORIGINAL_AFFINITY = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else set()


def pin_to_gpu_numa_node(local):
    """Bind this process (and the pinned host buffers it allocates from now on) to the CPUs of the NUMA node the GPU
    hangs off, so that the host->device copies of several ranks do not all cross the same socket interconnect.
    Best effort: returns a short description, or the reason nothing was done."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bus is None:
            out = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bus = out[-12:].lower() if out else None
        else:
            bus = "%04x:%02x:%02x.0" % (torch.cuda.get_device_properties(local).pci_domain_id, bus,
                                         torch.cuda.get_device_properties(local).pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "GPU reports no NUMA node"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"no allowed CPU on NUMA node {node}"
        os.sched_setaffinity(0, allowed)
        return f"bound to NUMA node {node} ({len(allowed)} CPUs) of GPU {bus}"
    except Exception as exc:
        return "not bound (%s)" % (str(exc).splitlines()[0][:80],)


def config_dict(n_total, world):
    return {"workload": "c4: trajectory optimisation fwd+bwd, 64 waypoints x 5 cams = 320 poses, "
                        f"{n_total / 1e6:g}M-point synthetic box cloud, point-sharded over {world} GPU(s)",
            "n_points": n_total, "n_poses": N_WAYPOINTS * N_CAMS, "parallelism": f"points/{world}",
            "l2_policy": "inputs larger than L2 (>=150 MB of points per rank streamed twice per step)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from trajectory_optimization_b200 import multicam
    body, rig = body_waypoints(), multicam.ring_rig(N_CAMS)
    base, ms = cpu_reference_rate(body, rig, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "coverage fwd+bwd point*pose evals/s", "value": base["value"],
            "unit": "point*pose evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(args.points, args.gpus), same_config=False, n_points_timed=CPU_SAMPLE_POINTS,
                           note="the CPU arm times a fixed 20 000-point sample of this workload per step and reports a RATE; "
                                "a full c4 step at this rate would take hours"),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=None, help="total cloud size (default: the workload's)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4 (default, the headline): trajectory optimisation fwd+bwd; c5: candidate-trajectory sweep")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): --points is the TOTAL cloud, sharded over the ranks; weak: --points PER RANK")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.points is None:
        args.points = N_POINTS_C4 if args.workload == "c4" else N_POINTS_C5
    if args.workload == "c5":
        return run_c5(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from trajectory_optimization_b200 import _lib, multicam, ops, tools
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = pin_to_gpu_numa_node(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    L = _lib.lib()
    n_total = args.points * (world if args.scaling == "weak" else 1)
    W = N_WAYPOINTS * N_CAMS
    K, img_w, img_h = tools.load_intrinsics(dev)
    rig = multicam.ring_rig(N_CAMS)
    body0 = body_waypoints()
    body = body0.to(dev).requires_grad_(True)
    pts = make_cloud_shard(n_total, rank, world, dev)
    n_local = pts.shape[0]

    rig7 = multicam.rig_tensor(rig, dev)

    ws_keep = torch.empty(L.cov_traj_workspace_bytes(n_local, W), dtype=torch.uint8, device=dev)  # one per cloud, as ModelTraj
    def step(points, perm, boxes, keep=None):
        body.grad = None
        t, q = multicam.camera_poses_fused(body, rig7)          # (x, y, z, yaw) x rig -> 320 camera poses
        rewards, mean = ops.coverage_traj(points, t, q, K, img_w, img_h, n_total=n_total, group=group, reward_index=perm,
                                          boxes=boxes, workspace=ws_keep)
        loss = 1.0 / (mean + 1e-6)
        loss.backward()
        if keep is not None:
            keep["rewards"] = rewards
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput (value): the product's default path ----
    # ModelTraj orders its cloud once at construction (the reference builds one model per cloud and iterates the
    # optimiser on it); that one-off sort is timed separately below and is inside the e2e figure.
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pts_sorted, perm = ops.spatial_sort(pts)
    ev0.record()
    pts_sorted, perm = ops.spatial_sort(pts)
    boxes = ops.tile_boxes(pts_sorted)
    ev1.record()
    torch.cuda.synchronize()
    ms_sort = ev0.elapsed_time(ev1)
    boxes = ops.tile_boxes(pts_sorted)
    for _ in range(args.warmup):
        step(pts_sorted, perm, boxes)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_eager = timed(lambda: step(pts_sorted, perm, boxes), args.steps)
    # The same step replayed from one CUDA graph (trajectory_optimization_b200.graphs: every kernel of the library and
    # the NCCL all-reduces are capturable): this is the product's way to run a launch-bound step, and `value`.
    graph_note = None
    run_step = lambda: step(pts_sorted, perm, boxes)  # noqa: E731
    try:
        from trajectory_optimization_b200.graphs import GraphedCall
        loss_e = step(pts_sorted, perm, boxes).detach().clone()
        grad_e = body.grad.detach().clone()
        gcall = GraphedCall(lambda: (step(pts_sorted, perm, boxes), body.grad), warmup=2)
        loss_g, grad_g = gcall()
        torch.cuda.synchronize()
        ok = (abs(float(loss_g) - float(loss_e)) <= 1e-6 * abs(float(loss_e))
              and float((grad_g - grad_e).abs().max()) <= 1e-5 * float(grad_e.abs().max()))
        if not ok:
            raise RuntimeError("graph replay does not reproduce the eager step")
        run_step = gcall
        graph_note = "step replayed from one CUDA graph (graphs.GraphedCall), checked against the eager step"
    except Exception as exc:  # keep the eager path if capture is not possible on this box
        graph_note = "CUDA graph capture unavailable (%s): eager step" % (str(exc).splitlines()[0][:120],)
    ms_total = timed(run_step, args.steps)
    value = n_total * W * args.steps / (ms_total * 1e-3)
    # The timed region lasts a few ms to a few tens of ms, shorter than one nvidia-smi sampling period, so the same
    # step keeps running (untimed) until the sampler has seen ~1.5 s of this exact load.
    t_end = time.perf_counter() + 1.5
    n_cont = 0
    while time.perf_counter() < t_end:
        for _ in range(20):
            run_step()
        torch.cuda.synchronize()
        n_cont += 20
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = ("sampled every 50 ms from the start of the timed region through %d further untimed repetitions of "
                          "the same step (the timed region itself lasts %.1f ms)" % (n_cont, ms_total))
    # same step with pruning switched off: every (point, pose) pair fully evaluated.  PARITY CHECK on the bench
    # configuration itself, in this run: the pruned step's per-point rewards and loss must equal the dense step's bit for
    # bit, its gradients to fp32 summation order.
    last = {}
    loss_p = step(pts_sorted, perm, boxes, last).detach().clone()
    rewards_p, grad_p = last.pop("rewards"), body.grad.detach().clone()
    with ops.evaluation(dense=True):
        for _ in range(2):
            loss_d = step(pts_sorted, perm, boxes, last).detach().clone()
        rewards_d, grad_d = last.pop("rewards"), body.grad.detach().clone()
        parity = {"rewards_bit_equal_dense": bool(torch.equal(rewards_p, rewards_d)),
                  "loss_rel_diff_vs_dense": abs(float(loss_p) - float(loss_d)) / abs(float(loss_d)),
                  "grad_rel_diff_vs_dense": float((grad_p - grad_d).abs().max() / grad_d.abs().max()),
                  "note": "pruned vs dense step on THIS configuration and shard, same run; rewards compared with torch.equal"}
        if world > 1:
            flags = torch.tensor([1.0 if parity["rewards_bit_equal_dense"] else 0.0, -parity["grad_rel_diff_vs_dense"]],
                                 device=dev, dtype=torch.float64)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            parity["rewards_bit_equal_dense"], parity["grad_rel_diff_vs_dense"] = bool(flags[0] > 0.5), float(-flags[1])
        if not parity["rewards_bit_equal_dense"] or parity["grad_rel_diff_vs_dense"] > 1e-4 or parity["loss_rel_diff_vs_dense"] > 1e-6:
            raise SystemExit("bench.py: pruned step does not reproduce the dense step: %r" % (parity,))
        del rewards_p, rewards_d
        dense_steps = max(2, min(args.steps, 5))
        ms_dense = timed(lambda: step(pts_sorted, perm, boxes), dense_steps) / dense_steps

    # ---- end to end: host (pinned) inputs in, loss + gradients out, every step ----
    # Every step consumes a NEW cloud from pinned host memory (1.2 GB over PCIe at N=1) plus the 64x4 body parameters:
    # host->device copy, spatial ordering, objective + gradient, loss + gradients back to the host.  The copy of
    # step i+1's cloud is issued on a copy stream while step i computes (double-buffered device cloud, as a streaming
    # consumer of PointCloud2 messages would do); one cloud copy and one sort per step are inside the timed region.
    host_pts = torch.empty(pts.shape, dtype=torch.float32, pin_memory=True)
    host_pts.copy_(pts)
    host_body = torch.empty(body0.shape, dtype=torch.float32, pin_memory=True)
    host_body.copy_(body0)
    host_out = torch.empty(1 + body0.numel(), dtype=torch.float32, pin_memory=True)
    bufs = [pts, torch.empty_like(pts)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()
    for ev in ready + free:
        ev.record(main)
    counter = [0]

    def e2e_step():
        i = counter[0]
        counter[0] += 1
        cur, nxt = i % 2, (i + 1) % 2
        with torch.cuda.stream(copy_stream):          # prefetch the next step's cloud
            copy_stream.wait_event(free[nxt])
            bufs[nxt].copy_(host_pts, non_blocking=True)
            ready[nxt].record(copy_stream)
        main.wait_event(ready[cur])
        with torch.no_grad():
            body.copy_(host_body, non_blocking=True)
        sorted_cur, perm_cur = ops.spatial_sort(bufs[cur])
        free[cur].record(main)
        loss = step(sorted_cur, perm_cur, ops.tile_boxes(sorted_cur))
        host_out[:1].copy_(loss.detach().reshape(1), non_blocking=True)
        host_out[1:].copy_(body.grad.reshape(-1), non_blocking=True)
        main.synchronize()

    with torch.cuda.stream(copy_stream):
        bufs[0].copy_(host_pts, non_blocking=True)
        ready[0].record(copy_stream)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = n_total * W * args.steps / (ms_e2e * 1e-3)
    h2d = host_pts.numel() * 4 * world + host_body.numel() * 4 * world
    d2h = host_out.numel() * 4 * world
    pts = bufs[0]
    del bufs, host_pts
    pts_sorted, perm = ops.spatial_sort(pts)
    boxes = ops.tile_boxes(pts_sorted)

    # The reference's own usage (one cloud per model, many optimiser steps): the cloud stays on the device, every step
    # copies the 64x4 body parameters host->device, evaluates, and reads loss + gradients back, synchronously.
    def params_only_step():
        with torch.no_grad():
            body.copy_(host_body, non_blocking=True)
        loss = step(pts_sorted, perm, boxes)
        host_out[:1].copy_(loss.detach().reshape(1), non_blocking=True)
        host_out[1:].copy_(body.grad.reshape(-1), non_blocking=True)
        main.synchronize()

    for _ in range(2):
        params_only_step()
    ms_params = timed(params_only_step, args.steps)

    # ---- per-call timing for the roofline (pass A = cov_traj_minmax, pass B = cov_traj_fused) ----
    import ctypes
    with torch.no_grad():
        t, q = multicam.camera_poses_fused(body, rig7)
    P, Q = t.contiguous(), q.contiguous()
    cam = _lib.camera(img_w, img_h, 1.0, 5.0, 1e-6)
    minmax = torch.empty(2 * W, device=dev)
    acc = torch.empty(W * _lib.ACC_STRIDE + 1, dtype=torch.float64, device=dev)
    rewards = torch.empty(n_local, device=dev)
    wsb = L.cov_traj_workspace_bytes(n_local, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    stats_dev = torch.zeros(8, dtype=torch.int64, device=dev)
    # as the product path calls them (ops.CoverageTrajFn): pass A pre-fills pass B's rewards under its idle bandwidth
    call_opts = {"a": _lib.traj_opts(stats=stats_dev, prefill=rewards), "b": _lib.traj_opts(stats=stats_dev, rewards_prefilled=True)}

    def pass_a():
        _lib.check(L.cov_traj_minmax(pts_sorted.data_ptr(), n_local, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(),
                                     ctypes.byref(cam), boxes.data_ptr(), minmax.data_ptr(), ctypes.byref(call_opts["a"]),
                                     ws.data_ptr(), wsb, stream), "cov_traj_minmax")

    def pass_b():
        _lib.check(L.cov_traj_fused(pts_sorted.data_ptr(), n_local, P.data_ptr(), Q.data_ptr(), W, K.data_ptr(),
                                    ctypes.byref(cam), boxes.data_ptr(), minmax.data_ptr(), None, perm.data_ptr(),
                                    rewards.data_ptr(), acc.data_ptr(), ctypes.byref(call_opts["b"]), ws.data_ptr(), wsb,
                                    stream), "cov_traj_fused")

    def global_minmax():
        pass_a()
        ops._all_reduce_minmax(minmax, W, group)

    global_minmax()
    pass_b()
    reps = max(3, min(args.steps, 10))
    stats_dev.zero_()
    pass_a()
    pass_b()
    torch.cuda.synchronize()
    st = [int(x) for x in stats_dev.tolist()]        # work counters of exactly one pass A + one pass B
    call_opts["a"], call_opts["b"] = _lib.traj_opts(prefill=rewards), _lib.traj_opts(rewards_prefilled=True)  # no counters
    ms_a = timed(pass_a, reps) / reps
    global_minmax()
    ms_b = timed(pass_b, reps) / reps
    call_opts["a"] = call_opts["b"] = _lib.traj_opts(dense=True)
    global_minmax()
    pass_b()
    ms_a_dense = timed(pass_a, reps) / reps
    global_minmax()
    ms_b_dense = timed(pass_b, reps) / reps
    call_opts["a"], call_opts["b"] = _lib.traj_opts(prefill=rewards), _lib.traj_opts(rewards_prefilled=True)
    global_minmax()
    pass_b()
    gated = float((rewards != 0.5).float().mean().item())  # fraction of points with at least one gated pose

    # FP32 / MUFU probes (measured peak for the dense kernels' roofline denominator)
    sink = torch.zeros(1, device=dev)
    iters = 4096
    L.cov_probe_fma(iters, sink.data_ptr(), stream)
    L.cov_probe_ex2(iters, sink.data_ptr(), stream)
    n_fma = [0]
    n_ex2 = [0]
    ms_fma = timed(lambda: n_fma.__setitem__(0, L.cov_probe_fma(iters, sink.data_ptr(), stream)), 3) / 3
    ms_ex2 = timed(lambda: n_ex2.__setitem__(0, L.cov_probe_ex2(iters, sink.data_ptr(), stream)), 3) / 3
    fp32_meas = 2.0 * n_fma[0] / (ms_fma * 1e-3) / 1e12
    mufu_meas = n_ex2[0] / (ms_ex2 * 1e-3) / 1e12

    def finish():
        # Multi-rank runs leave without tearing NCCL down: destroying a communicator whose collectives live in a captured
        # CUDA graph can block forever; the numbers are out, the processes simply end.
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    evals_local = n_local * W
    gbs_b = n_local * BYTES_PASS_B / (ms_b * 1e-3) / 1e9
    gbs_a = n_local * BYTES_PASS_A / (ms_a * 1e-3) / 1e9
    dense_tf_b = evals_local * FLOP_FWD / (ms_b_dense * 1e-3) / 1e12
    dense_tf_a = evals_local * FLOP_FWD / (ms_a_dense * 1e-3) / 1e12

    def frac(a, b):
        return a / max(b, 1)

    # The product path's dominant call is pass B on the ordered cloud.  With the tile pruning the arithmetic left is
    # ~0.3 % of the pairs; the floor of the call is then streaming the cloud once: 12 B/point read + 4 B/point of rewards
    # written = 16 B/point, which is what `achieved` counts (algorithmic bytes / call time).  The call
    # actually moves less (`traffic`: tiles no pose can reach are never read) and spends its time on irregular per-item
    # arithmetic, so `executed` reports the flops it really performs against the FP32 peak as well.  The dense kernels
    # (pruning off: every pair gets the 64-flop forward) are FP32-issue bound and are reported under `dense`.
    traffic, traffic_src = None, "no ncu summary found under profiles/"
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_pass_b_traffic.json")))
        traffic = float(tr["dram_bytes_per_call_at_1e8_points"]) * n_local / 1e8
        traffic_src = tr["source"] + "; scaled linearly to this shard's %d points" % n_local
    except Exception:
        pass
    item_pts = 128                                   # one (warp, pose) pair = 128 point x pose evaluations
    flops_b = (st[1] * FLOP_FWD + st[4] * (FLOP_FWD + FLOP_BWD)) * item_pts
    flops_a = st[3] * FLOP_FWD * 256
    roofline = {
        "kernel": "cov_traj_fused call = pass B on the Morton-ordered cloud, pruning on: cov_traj_table_kernel, "
                  "cov_cull_kernel (+ work list), cov_traj_fused_tiles_kernel (persistent warps; ~90 % of the call); the "
                  "rewards were pre-filled with 1/2 by the pass-A call before it (its 4 B/point are in BOTH calls' byte counts "
                  "below: bytes_per_point 16 here, 12 for pass A, as SURVEY 8d defines them)",
        "bound": "hbm", "achieved": gbs_b, "peak": hbm_peak, "unit": "GB/s", "frac": gbs_b / hbm_peak,
        "peak_source": hbm_src, "bytes_per_point": BYTES_PASS_B, "ms_per_launch": ms_b,
        "traffic": traffic, "traffic_source": traffic_src,
        "executed": {
            "note": "arithmetic the pruned call really performs, from the kernels' own work counters (one counted call): "
                    "(warp, pose) pairs evaluated x 128 points x 64 flop + pairs differentiated x 128 x (64 + 86) flop, over the "
                    "CALL time, against the nominal FP32 peak",
            "flop_per_call": flops_b, "tflops": flops_b / (ms_b * 1e-3) / 1e12,
            "frac_of_fp32_nominal": flops_b / (ms_b * 1e-3) / 1e12 / FP32_NOMINAL_TFLOPS,
            "pairs_evaluated": st[1] * item_pts, "pairs_differentiated": st[4] * item_pts,
            "pass_a_tflops": flops_a / (ms_a * 1e-3) / 1e12},
        "pass_a": {"kernel": "cov_traj_minmax call, pruning on: memset, cov_traj_prepare_kernel (pose table + seed), "
                             "cov_cull_kernel (+ work list), cov_traj_minmax_tiles_kernel<8,2>", "bound": "hbm",
                   "bytes_per_point": BYTES_PASS_A, "ms_per_launch": ms_a, "achieved": gbs_a, "frac": gbs_a / hbm_peak},
        "work_executed": {
            "note": "(warp, pose) pairs, as fractions of all pairs: listed by the cull / evaluated / (pass B) differentiated",
            "pass_b": {"tile_listed": frac(st[6], st[0]), "evaluated": frac(st[1], st[0]), "differentiated": frac(st[4], st[0])},
            "pass_a": {"tile_listed": frac(st[7], st[2]), "prefiltered": frac(st[5], st[2]), "evaluated": frac(st[3], st[2])}},
        "dense": {
            "note": "pruning off: every (point, pose) pair fully evaluated; FP32-issue bound; 64 flop per forward evaluation",
            "pass_b": {"kernel": "cov_traj_fused_kernel<4,0,2>", "bound": "fp32", "ms_per_launch": ms_b_dense,
                       "achieved": dense_tf_b, "peak": fp32_meas, "unit": "TFLOP/s", "frac": dense_tf_b / fp32_meas,
                       "frac_of_nominal": dense_tf_b / FP32_NOMINAL_TFLOPS},
            "pass_a": {"kernel": "cov_traj_minmax_kernel<8,2,2>", "bound": "fp32", "ms_per_launch": ms_a_dense,
                       "achieved": dense_tf_a, "peak": fp32_meas, "unit": "TFLOP/s", "frac": dense_tf_a / fp32_meas,
                       "frac_of_nominal": dense_tf_a / FP32_NOMINAL_TFLOPS},
            "peak_source": "FP32 FMA probe measured live in this run (cov_probe_fma); MEASURED_PEAKS.json has no FP32 "
                           "entry", "peak_nominal": FP32_NOMINAL_TFLOPS, "flop_per_eval": FLOP_FWD,
            "mufu_T_per_s_measured": mufu_meas, "mufu_per_eval": 3},
        "points_with_gated_pose_frac": gated,
    }
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, ORIGINAL_AFFINITY)   # the CPU arm gets every host core back
        except Exception:
            pass
        cpu_baseline, _ = cpu_reference_rate(body0, rig, steps=3, warmup=1)
    line = {"metric": "coverage fwd+bwd point*pose evals/s", "value": value, "unit": "point*pose evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "value_is": "DENSE-EQUIVALENT rate: every one of the N x W (point, pose) pairs gets its exact result per step; the "
                        "exact pruning arithmetically evaluates only `pairs_evaluated_frac` of them (see `dense` for the rate "
                        "with every pair evaluated)",
            "pairs_evaluated_frac": frac(st[1], st[0]),
            "config": dict(config_dict(n_total, world),
                           pruning="exact tile-level distance-bound pruning on a Morton-ordered cloud (default); see `dense`",
                           launch=graph_note,
                           collectives=("none (one GPU)" if world == 1 else
                                        "in-kernel NVLink exchange: cov_peer_allreduce (peer stores + flags, csrc/cov_peer.cu), "
                                        "one launch per exchange step" if ops.peer_exchange(group, dev, W) is not None else
                                        "NCCL all_reduce (MAX of 2W floats, SUM of 22W+1 doubles)"),
                           cloud_order="ordered once per cloud by cov_spatial_sort (%.2f ms for this rank's shard, outside "
                                       "`value`, inside `e2e`)" % ms_sort),
            "clocks": clocks,
            "eager": {"value": n_total * W * args.steps / (ms_eager * 1e-3), "unit": "point*pose evals/s",
                      "ms_per_step": ms_eager / args.steps, "note": "same step launched kernel by kernel from Python"},
            "dense": {"value": n_total * W / (ms_dense * 1e-3), "unit": "point*pose evals/s", "ms_per_step": ms_dense,
                      "note": "same step with cov_traj_opts.dense: every pair fully evaluated"},
            "parity_check": parity,
            "e2e": {"value": e2e_value, "unit": "point*pose evals/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "a NEW cloud from pinned host memory every step (copy + spatial ordering + objective + gradient + "
                            "read-back); bound by the PCIe copy",
                    "host_affinity": numa_note,
                    "resident_cloud": {"value": n_total * W * args.steps / (ms_params * 1e-3), "unit": "point*pose evals/s",
                                       "ms_per_step": ms_params / args.steps,
                                       "h2d_bytes_per_step": host_body.numel() * 4 * world, "d2h_bytes_per_step": d2h,
                                       "note": "the reference's usage: one cloud per model; per step only the body parameters "
                                               "go in and loss + gradients come back, eager launches, synchronous"}},
            # own kernels per step: rig poses; pass A: prepare, cull, tiles; pass B: table, cull, tiles; epilogue; rig backward
            "gpu_launches": 9 * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
