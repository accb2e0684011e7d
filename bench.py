#!/usr/bin/env python
"""Benchmark of the B200-native feature-field hot path (BASELINE.json: train rays/s, config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU arm (oracle port of run())
    torchrun --nproc-per-node N bench.py --gpus N ...              # ray-sharded data parallel (weak scaling)

One step = one training iteration of autolabel's SimpleTrainer on a 4096-ray batch per GPU:
march -> field -> composite -> loss -> backward -> Adam (+ the occupancy refresh every 16 steps,
amortised inside the timed region), EVERY marched sample composited as the reference's training kernels do
(train_t_thresh = 0; `early_termination` reports the opt-in 1e-4 transmittance cut next to it).
`value` = rays/s with batches resident in HBM; `e2e` = the same through trainer.train_one_step() with HOST
(pinned) batches: the H2D copy of the batch and a D2H read of the loss inside the timed region.
Every N starts the timed region from the SAME model state: all ranks run the identical single-GPU pre-training
(rank-0 seeds, no exchange), rank 0's state is broadcast, and only then do the ranks draw their own rays and
exchange gradients.  Prints ONE JSON line (rank 0), after the process group is gone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_rays_per_s"
RAYS = 4096


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--pretrain", type=int, default=3000,
                    help="untimed training steps (fresh rays every step) before warm-up, so that the occupancy grid and the "
                         "samples/ray are those of a trained scene")
    ap.add_argument("--feature-dim", type=int, default=64)
    ap.add_argument("--rays", type=int, default=RAYS, help="rays per GPU and step (C2/C3: 4096; C5: 1024 with --feature-dim 512)")
    ap.add_argument("--density-thresh", type=float, default=10.0,
                    help="occupancy threshold of the marched path; 10 is what the reference's own cuda_ray entry point "
                         "passes (torch_ngp/main_nerf.py:47,91); NeRFRenderer's constructor default is 0.01 (renderer.py:76)")
    ap.add_argument("--render-frames", type=int, default=4,
                    help="full frames rendered per rank for the render leg (frames/s, rgb+depth+semantic+features); 0 = skip")
    ap.add_argument("--train-t-thresh", type=float, default=0.0,
                    help="0 (default) = composite every marched sample, the reference's training semantics "
                         "(raymarching.cu:547-740).  > 0 opts into training-time early termination: samples behind the "
                         "point where a ray's transmittance drops below this value skip the heads / compositing / "
                         "backward (the constant of the reference's marched inference kernel, raymarching.cu:929-935)")
    ap.add_argument("--no-early-leg", action="store_true", help="skip the secondary leg with train_t_thresh = 1e-4")
    ap.add_argument("--c5-steps", type=int, default=-1,
                    help="timed steps of the C5 leg (1296x968-shaped scene at train factor 2, 512-d features, 1024 rays/step) "
                         "reported inside the same JSON line; -1 = min(steps, 50) at N = 1, 0 = skip")
    ap.add_argument("--c5-pretrain", type=int, default=1500)
    ap.add_argument("--ncu-c5", type=int, default=0,
                    help="profiling aid: like --ncu-range, but the profiled range is this many training steps of the C5 configuration")
    ap.add_argument("--grad-exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = one kernel over NVLink peer memory (reduce-scatter + sharded Adam + all-gather, "
                         "csrc/peer.cu); 'nccl' = all_reduce(param.grad) + Adam on every rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rays", type=int, default=0, help="rays per CPU step; 0 = the workload's full batch (--rays)")
    ap.add_argument("--cpu-budget-s", type=float, default=420.0,
                    help="--impl reference: if the first step projects the K + W steps beyond this many seconds, fewer "
                         "timed steps are run (and reported)")
    ap.add_argument("--ncu-render", type=int, default=0,
                    help="profiling aid: like --ncu-range, but the profiled range renders this many full frames (inference)")
    ap.add_argument("--ncu-range", type=int, default=0,
                    help="profiling aid: after pretrain+warm-up run this many steps inside cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench line")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"{'C2' if (args.rays, args.feature_dim) == (4096, 64) else 'C5' if args.feature_dim == 512 else 'custom'}: synthetic {args.frames}x({args.width}x{args.height}) RGB-D scene, hg+freq encoder, 128-wide "
                    f"density/colour MLPs, {args.feature_dim}-d feature head, 2 classes, {args.rays} rays/GPU/step",
        "rays_per_gpu": args.rays, "frames": args.frames, "resolution": [args.width, args.height],
        "encoding": "hg+freq", "feature_dim": args.feature_dim, "n_classes": 2, "density_thresh": args.density_thresh,
        "train_t_thresh": args.train_t_thresh,
        "parallelism": f"dp{world} (ray-sharded, gradient all-reduce)" if world > 1 else "single GPU",
    }


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_port_rate(args, steps, warmup, threads=None, budget_s=None):
    """rays/s of the reference-shaped CPU path (oracle/run_path.py: NeRFRenderer.run() + SimpleTrainer.train_step + Adam,
    fp32 torch, all host threads) on the SAME workload: full batches (--rays rays x 256 uniform samples, the reference's
    sampling, renderer.py:190) drawn from the same synthetic scene (frames rendered lazily on the host), same model
    shape and bound."""
    from oracle import run_path
    from scene_synth import SyntheticScene
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    n = args.cpu_rays or args.rays
    scene = SyntheticScene(args.frames, args.height, args.width, args.feature_dim, n_classes=2, seed=0, device='cpu', lazy=True)
    scene.gen.manual_seed(1000)
    field = run_path.OracleField('hg+freq', 128, 128, args.feature_dim, 2, bound=scene.bound(), seed=0)
    opt = field.optimizer()

    def batch():
        b = scene.next_train(n)
        b['direction_norms'] = b['direction_norms'].reshape(-1, 1)
        return b
    done_w, t_first = 0, None
    for _ in range(warmup):
        t0 = time.perf_counter()
        run_path.train_step(field, opt, batch())
        t_first = time.perf_counter() - t0
        done_w += 1
    if budget_s is not None and t_first is not None and t_first * (steps + max(warmup - 1, 0)) > budget_s:
        steps = max(1, int(budget_s / t_first) - max(warmup - 1, 0))
    pool = [batch() for _ in range(steps)]        # batch synthesis (lazy frame rendering) stays outside the timed region
    t0 = time.perf_counter()
    for b in pool:
        run_path.train_step(field, opt, b)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n / dt, dt, threads, steps, (f"{n} rays x 256 uniform samples per step drawn from the same synthetic scene (run() path: "
                                        "field fwd+bwd on every sample, compositing, loss, Adam over 14.3M params), fp32 torch CPU")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    rate, dt, threads, steps, sample = cpu_port_rate(args, args.steps, max(args.warmup, 1), budget_s=args.cpu_budget_s)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": max(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": dict(workload_config(args, world), train_t_thresh=None,
                                                          sampling="run(): 256 uniform samples per ray (cuda_ray=False)"),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def build_trainer(args, device, height=None, width=None, feature_dim=None, feature_hw=None):
    from autolabel_b200.models import ALNetwork
    from autolabel_b200.trainer import SimpleTrainer
    from scene_synth import SyntheticScene
    torch.manual_seed(0)                                   # identical on every rank (occupancy refresh RNG)
    F = feature_dim or args.feature_dim
    scene = SyntheticScene(args.frames, height or args.height, width or args.width, F, feature_hw=feature_hw, n_classes=2,
                           seed=0, device=device)
    scene.gen.manual_seed(1000)                            # pre-training: rank 0's ray stream on EVERY rank
    model = ALNetwork(encoding='hg+freq', num_layers=2, hidden_dim=128, geo_feat_dim=15, num_layers_color=2,
                      hidden_dim_color=128, hidden_dim_semantic=F, semantic_classes=2,
                      bound=scene.bound(), cuda_ray=True, density_scale=1, density_thresh=args.density_thresh)
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=5e-3,
                          train_t_thresh=args.train_t_thresh)
    trainer = SimpleTrainer('bench', opt, model, device=device, fp16=True, workspace=None, log_interval=0)
    model.train()
    model.mark_untrained_grid(scene.poses, scene.intrinsics)
    return scene, model, trainer


def opt_lr(trainer):
    return float(trainer.optimizer.param_groups[0]['lr'])


class Legs:
    """Timed loops shared by the training legs: barrier + synchronize on both sides, CUDA events on the launching stream,
    max over ranks."""

    def __init__(self, device, world):
        self.device, self.world = device, world
        self.host_ms = 0.0

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        self.host_ms = (time.perf_counter() - t0) * 1e3 / max(steps, 1)   # host time to ENQUEUE a step (no sync)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.device)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()


def align_refresh(trainer, scene, rays, offset=4):
    """Untimed steps until global_step % update_interval == offset, so that every timed leg sees the occupancy refresh at
    the same phase (K = 20 steps from offset 4 contain exactly one refresh = 1/20 per step against the true 1/16)."""
    from autolabel_b200.trainer import PackedBatch
    while trainer.global_step % trainer.update_interval != offset % trainer.update_interval:
        trainer.train_one_step(PackedBatch.pack(scene.next_train(rays)))


def train_legs(args, scene, model, trainer, device, rank, world, rays, steps, warmup, sampler=None):
    """value (device-resident packed batches) and e2e (pinned host packed batches in, loss out) of one workload."""
    from autolabel_b200 import _lib
    from autolabel_b200.trainer import PackedBatch
    legs = Legs(device, world)
    n_pool = min(max(steps + warmup, 32), 600)

    def fresh(n=n_pool, host=False):
        # a distinct batch for every warm-up / timed step (recycling a small pool lets the field overfit those rays and the
        # step gets ~20 % faster, profiles/r1g_diag_step.txt); packed = one flat buffer per batch = ONE copy per step
        return [PackedBatch.pack(scene.next_train(rays), device='cpu' if host else None, pin=host) for _ in range(n)]
    pool = fresh()
    host_pool = fresh(host=True)                            # never seen before the e2e leg
    for i in range(warmup):
        trainer.train_one_step(pool[(len(pool) - 1 - i) % len(pool)])
    align_refresh(trainer, scene, rays)
    if sampler is not None:
        sampler.start()
    launches0 = _lib.lib.al_launch_count() + trainer.graph_kernel_launches
    step0 = trainer.global_step
    ms = legs.timed(lambda i: trainer.train_one_step(pool[i % len(pool)]), steps)
    launches = _lib.lib.al_launch_count() + trainer.graph_kernel_launches - launches0
    refreshes = sum(1 for g in range(step0, step0 + steps) if g % trainer.update_interval == 0)
    res = {"ms": ms, "host_enqueue_ms": legs.host_ms, "launches": int(launches), "refreshes_in_timed_region": refreshes,
           "samples_per_ray": float(model.last_meta[1].item()) / rays,
           "alive_samples_per_ray": float(model.last_alive_meta[0].item()) / rays, "loss": float(trainer.last_loss.item())}

    def e2e_step(i):
        trainer.train_one_step(host_pool[i % len(host_pool)]).item()
    for b in fresh(min(warmup, 5), host=True):
        trainer.train_one_step(b).item()
    align_refresh(trainer, scene, rays)
    res["ms_e2e"] = legs.timed(e2e_step, steps)
    res["h2d_bytes"] = int(host_pool[0].flat.numel())
    res["clocks"] = sampler.stop() if sampler is not None else None
    res["legs"] = legs
    res["fresh"] = fresh
    return res


def main():
    global RAYS
    args = parse()
    RAYS = args.rays
    if args.impl == "reference":
        return run_reference(args)
    from autolabel_b200 import parallel
    from autolabel_b200.trainer import PackedBatch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    rank, world, local_rank = parallel.init_distributed()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    import torch.distributed as dist

    if args.ncu_c5 > 0:
        print(json.dumps(c5_leg(args, device, rank, world, 0)))
        return
    scene, model, trainer = build_trainer(args, device)

    # ---- untimed: the SAME single-GPU pre-training on every rank (rank-0 ray stream, no exchange), so every N starts its
    # timed region from the same occupancy grid / samples per ray; fresh rays every step
    for _ in range(args.pretrain):
        trainer.train_one_step(PackedBatch.pack(scene.next_train(RAYS)))
    grad_exchange = "none (single GPU)"
    if world > 1:
        # replicas drift apart in the last bits (float atomics in the hash-grid scatter): make them identical, then start the
        # exchange from the replicated optimiser state and give every rank its own ray stream
        parallel.broadcast_parameters(model)
        mc = torch.tensor([model.mean_count, model.local_step], device=device)
        dist.broadcast(mc, src=0)
        model.mean_count, model.local_step = int(mc[0]), int(mc[1])
        grad_exchange = None
        if args.grad_exchange == "peer":
            try:
                peer = parallel.PeerShardedAdam(model, lr=opt_lr(trainer), init_from=trainer.optimizer)
                trainer.set_optimizer(peer)
                trainer.grad_sync = None
                grad_exchange = ("peer memory kernel (al_peer_adam_step), " +
                                 ("multimem.ld_reduce / multimem.st (NVLS)" if peer.multicast else "peer loads / stores"))
            except Exception as e:  # symmetric memory unavailable on this box: keep training over NCCL, say so
                grad_exchange = f"nccl all_reduce + replicated Adam (peer path unavailable: {e!r})"
        if grad_exchange is None or grad_exchange.startswith("nccl"):
            trainer.grad_sync = parallel.GradientAllReduce(model.parameters(), trainer.optimizer)
            trainer._graph_state = None
            grad_exchange = grad_exchange or "nccl all_reduce + replicated Adam"
        scene.gen.manual_seed(1000 + rank)                 # from here on each rank samples its own rays

    if args.ncu_render > 0:
        model.eval()
        b = scene.get_test(0)
        o, d, nrm = b['rays_o'].view(1, -1, 3), b['rays_d'].view(1, -1, 3), b['direction_norms']
        with torch.no_grad():
            model.render(o, d, nrm, staged=True, perturb=False)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            for _ in range(args.ncu_render):
                model.render(o, d, nrm, staged=True, perturb=False)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        print(json.dumps({"ncu_render_frames": args.ncu_render, "samples_per_ray": float(model.last_meta[1].item()) / (scene.h * scene.w)}))
        return
    if args.ncu_range > 0:
        pool = [PackedBatch.pack(scene.next_train(RAYS)) for _ in range(max(args.warmup, args.ncu_range))]
        for b in pool[:args.warmup]:
            trainer.train_one_step(b)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(args.ncu_range):
            trainer.train_one_step(pool[i % len(pool)])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_range_steps": args.ncu_range, "samples_per_ray": float(model.last_meta[1].item()) / RAYS}))
        return

    sampler = ClockSampler(local_rank) if rank == 0 else None
    r = train_legs(args, scene, model, trainer, device, rank, world, RAYS, args.steps, args.warmup, sampler)
    legs, fresh = r["legs"], r["fresh"]
    value = RAYS * world * args.steps / (r["ms"] * 1e-3)
    e2e_value = RAYS * world * args.steps / (r["ms_e2e"] * 1e-3)

    # ---- secondary: the opt-in training-time early termination (train_t_thresh = 1e-4), reported next to the headline
    early = None
    if not args.no_early_leg and args.train_t_thresh == 0.0:
        try:
            model.train_t_thresh = 1e-4
            n_e = max(args.steps // 2, 1)
            for b in fresh(max(min(args.warmup, 5), 3)):
                trainer.train_one_step(b)
            align_refresh(trainer, scene, RAYS)
            pool = fresh(n_e)
            ms_e = legs.timed(lambda i: trainer.train_one_step(pool[i % len(pool)]), n_e)
            early = {"value": RAYS * world * n_e / (ms_e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e / n_e, "steps": n_e,
                     "train_t_thresh": 1e-4, "alive_samples_per_ray": float(model.last_alive_meta[0].item()) / RAYS,
                     "note": "opt-in: samples behind T < 1e-4 skip heads / compositing / backward; NOT the reference's training "
                             "semantics, not the headline"}
        except Exception as e:
            early = {"error": repr(e)}
        finally:
            model.train_t_thresh = args.train_t_thresh
            for b in fresh(3):
                trainer.train_one_step(b)

    # ---- f1: the batch SAMPLER inside the timed region too (the reference draws a batch per step on the host,
    # dataset.py:182-242): DeviceSceneDataset keeps the scene in HBM, draws on the device without host synchronisation and
    # writes a PackedBatch with one kernel
    dev_ds = None
    try:
        dev_ds = device_dataset_leg(args, scene, trainer, device, rank, world, legs)
    except Exception as e:
        dev_ds = {"error": repr(e)}

    detail = phase_detail(args, scene, model, trainer, device) if rank == 0 else None
    render = None
    if args.render_frames > 0:
        try:
            render = render_leg(args, scene, model, device, rank, world, legs.timed)
        except Exception as e:  # the training line must survive a failure of the second leg
            render = {"metric": "render_frames_per_s", "error": repr(e)}
            model.train()

    # ---- config C5 (512-d LSeg-shaped feature head, 1024 rays/step) inside the same line, N = 1 only
    c5 = None
    c5_steps = (min(args.steps, 50) if world == 1 else 0) if args.c5_steps < 0 else args.c5_steps
    if c5_steps > 0 and args.feature_dim != 512:
        try:
            c5 = c5_leg(args, device, rank, world, c5_steps)
        except Exception as e:
            c5 = {"error": repr(e)}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            rate, dt, threads, n_cpu, sample = cpu_port_rate(args, steps=1, warmup=1)
            cpu = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample + " (1 warm-up + 1 timed step)",
                   "ms_per_step": dt * 1e3}
            if render is not None and "error" not in render:
                render["cpu_baseline"] = cpu_render_rate(args)
        except Exception as e:
            cpu = {"error": repr(e)}

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms"] / args.steps, "host_enqueue_ms_per_step": r["host_enqueue_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "config": dict(workload_config(args, world), pretrain_steps=args.pretrain,
                                                pretrain="identical single-GPU pre-training on every rank (rank-0 rays, no "
                                                         "exchange), then rank 0's state broadcast",
                                                grad_exchange=grad_exchange,
                                                samples_per_ray=r["samples_per_ray"], alive_samples_per_ray=r["alive_samples_per_ray"],
                                                refreshes_in_timed_region=r["refreshes_in_timed_region"],
                                                final_loss=r["loss"],
                                                l2="per-step working set (57 MB table + 57 MB gradients + 114 MB Adam moments "
                                                   "+ per-sample buffers) exceeds the 126 MB L2; no explicit flush"),
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": r["ms_e2e"] / args.steps,
                    "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": 4,
                    "api": "SimpleTrainer.train_one_step(PackedBatch in pinned host memory) -> loss.item()"},
            "gpu_launches": r["launches"], "clocks": r["clocks"], "roofline": detail["roofline"] if detail else None,
            "cpu_baseline": cpu,
        }
        if dev_ds:
            line["e2e_device_dataset"] = dev_ds
        if early:
            line["early_termination"] = early
        if detail:
            line["phases_ms"] = detail["phases_ms"]
            if detail.get("rooflines"):
                line["rooflines"] = detail["rooflines"]
        if render:
            line["render"] = render
        if c5:
            line["c5"] = c5
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            # NCCL_DEBUG=INFO (the driver's rank check) makes every rank print teardown lines to stdout: let the other
            # ranks finish theirs, so that the JSON line is the LAST line of the job's stdout
            sys.stdout.flush()
            time.sleep(2.0)
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL prints more teardown lines from its library destructors at interpreter exit ("Closing env plugin ..."): leave
        # without running them, so that nothing follows the JSON line on stdout
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def device_dataset_leg(args, scene, trainer, device, rank, world, legs):
    """rays/s through `for batch in DeviceSceneDataset: trainer.train_one_step(batch); loss.item()` -- batch sampling (the
    reference's `_next_train`: image / pixel draws with the labelled-pixel policy, ray generation, target gathers) inside
    the timed region, on the device."""
    from autolabel_b200.dataset import DeviceSceneDataset
    depth_mm = (scene.depths * 1000.0).round().clamp(0, 65535).to(torch.int32).cpu().numpy().astype('uint16')
    sem = (scene.semantics + 1).clamp(min=0).to(torch.uint8)                     # 0 = unlabeled, class c -> c + 1
    ds = DeviceSceneDataset(scene.images, depth_mm, sem, scene.poses, scene.intrinsics, (scene.w, scene.h),
                            features=scene.features, feature_size=(scene.fw, scene.fh), batch_size=args.rays,
                            n_classes=2, device=device, seed=2000 + rank)
    del depth_mm, sem
    it = iter(ds)
    for _ in range(max(min(args.warmup, 5), 3)):
        trainer.train_one_step(next(it)).item()
    align_refresh(trainer, scene, args.rays)
    ms = legs.timed(lambda i: trainer.train_one_step(next(it)).item(), args.steps)
    del ds
    return {"value": args.rays * world * args.steps / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / args.steps,
            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
            "api": "for batch in DeviceSceneDataset(...): SimpleTrainer.train_one_step(batch) -> loss.item()  (sampling in the "
                   "timed region, scene resident in HBM)"}


def c5_leg(args, device, rank, world, steps):
    """BASELINE.json config 5: ScanNet-shaped 1296x968 scene at train factor 2 (648x484), 512-d LSeg-shaped feature head
    (feature maps 121x162), 1024 rays/step — value, e2e and the roofline of the wide-head GEMM kernel."""
    import copy
    a = copy.copy(args)
    a.height, a.width, a.feature_dim, a.rays = 484, 648, 512, 1024
    a.frames = min(args.frames, 100)                      # 100 x (121 x 162 x 512) fp16 feature maps = 2 GB of HBM
    a.pretrain = args.c5_pretrain
    from autolabel_b200.trainer import PackedBatch
    scene, model, trainer = build_trainer(a, device, feature_hw=(121, 162))
    for _ in range(a.pretrain):
        trainer.train_one_step(PackedBatch.pack(scene.next_train(a.rays)))
    if args.ncu_c5 > 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(args.ncu_c5):
            trainer.train_one_step(PackedBatch.pack(scene.next_train(a.rays)))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return {"ncu_c5_steps": args.ncu_c5, "samples_per_ray": float(model.last_meta[1].item()) / a.rays}
    r = train_legs(a, scene, model, trainer, device, rank, world, a.rays, steps, min(args.warmup, 10))
    out = {"metric": METRIC, "config": dict(workload_config(a, world), pretrain_steps=a.pretrain,
                                            samples_per_ray=r["samples_per_ray"]),
           "value": a.rays * world * steps / (r["ms"] * 1e-3), "unit": "rays/s", "steps": steps,
           "ms_per_step": r["ms"] / steps,
           "e2e": {"value": a.rays * world * steps / (r["ms_e2e"] * 1e-3), "unit": "rays/s", "ms_per_step": r["ms_e2e"] / steps,
                   "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": 4},
           "gpu_launches": r["launches"]}
    try:
        from bench_detail import measure_wide
        out.update(measure_wide(a, scene, model, trainer, device, a.rays))
    except Exception as e:
        out["roofline"] = {"error": repr(e)}
    del scene, model, trainer
    torch.cuda.empty_cache()
    return out


def cpu_render_rate(args):
    """CPU baseline of the render half of the metric: the reference's staged render() (renderer.py:685-744: rows of rays in
    chunks of 4096 through run() with 256 uniform samples) as ported in oracle/run_path.py, on a bounded sample of a
    640x480 frame (rows of one frame), all host threads."""
    from oracle import run_path
    from scene_synth import SyntheticScene
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    scene = SyntheticScene(args.frames, args.height, args.width, args.feature_dim, n_classes=2, seed=0, device='cpu', lazy=True)
    field = run_path.OracleField('hg+freq', 128, 128, args.feature_dim, 2, bound=scene.bound(), seed=0)
    b = scene.get_test(0)
    rows = 8                                               # 8 of the 480 rows: 5120 rays x 256 samples
    o = b['rays_o'][:rows].reshape(-1, 3)
    d = b['rays_d'][:rows].reshape(-1, 3)
    nrm = b['direction_norms'].reshape(args.height, args.width)[:rows].reshape(-1)
    with torch.no_grad():
        run_path.run(field, o[:args.width], d[:args.width], nrm[:args.width])          # warm-up: one row
        t0 = time.perf_counter()
        for head in range(0, o.shape[0], 4096):
            run_path.run(field, o[head:head + 4096], d[head:head + 4096], nrm[head:head + 4096])
        dt = time.perf_counter() - t0
    frame_s = dt * args.height / rows
    return {"value": 1.0 / frame_s, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{rows} of {args.height} rows of one {args.width}x{args.height} frame ({o.shape[0]} rays x 256 uniform samples, "
                      "staged render() of the run() path, six maps), scaled to a frame", "s_per_frame": frame_s}


def render_leg(args, scene, model, device, rank, world, timed):
    """Second half of BASELINE.json's metric: full-frame render (rgb + depth + semantic logits + features for every
    pixel) in frames/s, frame-sharded over the ranks with no communication (export.py / render.py shape).
    `value`: rays resident in HBM; `e2e`: host rays in (pinned), the six output maps back to the host."""
    from autolabel_b200.parallel import shard_frames
    model.eval()
    n = args.render_frames
    frames = shard_frames(min(scene.n, n * world), rank, world)[:n]
    batches = [scene.get_test(i) for i in frames]
    H, W = scene.h, scene.w
    dev_in = [(b['rays_o'].view(1, -1, 3), b['rays_d'].view(1, -1, 3), b['direction_norms']) for b in batches]
    host_in = [tuple(t.cpu().pin_memory() for t in d) for d in dev_in]
    spr = []

    def render_dev(i):
        o, d, nrm = dev_in[i % n]
        with torch.no_grad():
            model.render(o, d, nrm, staged=True, perturb=False)

    host_out = {}

    def render_e2e(i):
        o, d, nrm = (t.to(device, non_blocking=True) for t in host_in[i % n])
        with torch.no_grad():
            out = model.render(o, d, nrm, staged=True, perturb=False)
        for k, v in out.items():                       # the six maps go back to pinned host memory
            if k not in host_out:
                host_out[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            host_out[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        return host_out

    render_dev(0)
    spr.append(float(model.last_meta[1].item()) / (H * W))
    ms = timed(render_dev, n)
    render_e2e(0)
    ms_e2e = timed(render_e2e, n)
    model.train()
    out_bytes = H * W * 4 * (1 + 1 + 3 + model.semantic_classes + model.hidden_dim_semantic + 3)
    roof = None
    if rank == 0:
        try:
            roof = render_roofline(model, device, int(spr[0] * H * W))
        except Exception as e:
            roof = {"error": repr(e)}
    return {"metric": "render_frames_per_s", "roofline": roof, "value": n * world / (ms * 1e-3), "unit": "frames/s",
            "resolution": [W, H], "frames_per_rank": n, "ms_per_frame": ms / n, "samples_per_ray": spr[0],
            "early_termination": bool(model.early_termination),
            "outputs": "image, depth, depth_variance, semantic logits, semantic_features, coordinates_map",
            "e2e": {"value": n * world / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_frame": ms_e2e / n,
                    "h2d_bytes_per_frame": H * W * 4 * 7, "d2h_bytes_per_frame": out_bytes}}


def render_roofline(model, device, samples_per_frame):
    """The render leg's dominant kernel (the position encoder: 16 levels x 8 corner gathers per sample) timed standalone
    with CUDA events on a frame-sized batch of in-bound positions; algorithmic bytes per sample as DESIGN.md section 4."""
    import ctypes
    from bench_detail import _time, peaks
    from autolabel_b200._lib import call, ptr, stream_ptr
    M = int(min(max(samples_per_frame, 1 << 20), 1 << 23))
    desc = model.field_desc()
    xyz = ((torch.rand(M, 3, device=device) * 2 - 1) * float(model.bound)).contiguous()
    x_enc = torch.empty(M, int(desc.in_pad), dtype=torch.float16, device=device)
    st = stream_ptr(device)

    def k():
        call("al_encode_position", ptr(xyz), M, None, float(model.bound), desc.encoding, desc.table, desc.offsets, desc.L, desc.S,
             desc.H, 0, ptr(x_enc), int(desc.in_pad), st)
    ms = _time(k, reps=10, warm=2)
    pk = peaks()
    gb = M * (12 + int(desc.L) * 8 * 8 + int(desc.in_pad) * 2) / 1e9
    return {"kernel": "encode_position", "bound": "hbm", "limiter": "L1/L2 gather throughput (uniformly random positions: the "
            "worst case for table locality)", "achieved": gb / (ms * 1e-3), "peak": pk["hbm"], "unit": "GB/s",
            "frac": gb / (ms * 1e-3) / pk["hbm"], "traffic": None, "samples": M, "avg_launch_ms": ms, "peak_source": pk["source"]}


def phase_detail(args, scene, model, trainer, device):
    """CUDA-event timing of the dominant kernel of a step, run standalone through its C-ABI entry point
    on the buffers of a real step (achieved = algorithmic work / average launch time)."""
    try:
        from bench_detail import measure
    except Exception as e:  # pragma: no cover
        return {"roofline": None, "phases_ms": {"error": repr(e)}}
    try:
        return measure(args, scene, model, trainer, device, RAYS)
    except Exception as e:  # the training line must survive a failure of the per-kernel leg
        return {"roofline": None, "phases_ms": {"error": repr(e)}}


if __name__ == "__main__":
    main()
