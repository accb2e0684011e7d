#!/usr/bin/env python
"""Benchmark of the B200-native feature-field hot path (BASELINE.json: train rays/s, config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU arm (oracle port of run())
    torchrun --nproc-per-node N bench.py --gpus N ...              # ray-sharded data parallel (weak scaling)

One step = one training iteration of autolabel's SimpleTrainer on a 4096-ray batch per GPU:
march -> field -> composite -> loss -> backward -> Adam (+ the occupancy refresh every 16 steps,
amortised inside the timed region).  `value` = rays/s with batches resident in HBM; `e2e` = the same
through trainer.train_one_step() with HOST (pinned) batches: H2D copies of the batch and a D2H read
of the loss inside the timed region.  Prints ONE JSON line (rank 0).
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
    ap.add_argument("--train-t-thresh", type=float, default=1e-4,
                    help="training-time early termination: samples behind the point where a ray's transmittance drops "
                         "below this value skip the heads / compositing / backward (the constant of the reference's "
                         "marched inference kernel, raymarching.cu:929-935); 0 = composite every marched sample")
    ap.add_argument("--grad-exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = one kernel over NVLink peer memory (reduce-scatter + sharded Adam + all-gather, "
                         "csrc/peer.cu); 'nccl' = all_reduce(param.grad) + Adam on every rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rays", type=int, default=256, help="rays of the bounded CPU sample")
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
def cpu_port_rate(args, steps, warmup, threads=None):
    """rays/s of the reference-shaped CPU path (oracle/run_path.py) on a bounded sample of the workload:
    `cpu_rays` rays x 256 uniform samples, forward + backward + Adam, fp32, all host threads."""
    from oracle import run_path
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    n = args.cpu_rays
    field = run_path.OracleField('hg+freq', 128, 128, args.feature_dim, 2, bound=5.0, seed=0)
    opt = field.optimizer()
    d = torch.randn(n, 3, generator=g)
    data = {
        'rays_o': (torch.rand(n, 3, generator=g) - 0.5) * 2.0, 'rays_d': d / d.norm(dim=1, keepdim=True),
        'direction_norms': torch.ones(n, 1) + 0.2 * torch.rand(n, 1, generator=g), 'pixels': torch.rand(n, 3, generator=g),
        'depth': torch.rand(n, generator=g) * 3, 'semantic': torch.randint(-1, 2, (n,), generator=g),
        'features': torch.rand(n, args.feature_dim, generator=g),
    }
    for _ in range(warmup):
        run_path.train_step(field, opt, data)
    t0 = time.perf_counter()
    for _ in range(steps):
        run_path.train_step(field, opt, data)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n / dt, dt, threads, f"{n} rays x 256 uniform samples per step (run() path: field fwd+bwd, compositing, Adam over 14.3M params), fp32 torch CPU"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warmup = min(args.warmup, 1)
    rate, dt, threads, sample = cpu_port_rate(args, steps, warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------- our arm
def build_trainer(args, device, rank):
    from autolabel_b200.models import ALNetwork
    from autolabel_b200.trainer import SimpleTrainer
    from scene_synth import SyntheticScene
    torch.manual_seed(0)                                   # identical on every rank (occupancy refresh RNG)
    scene = SyntheticScene(args.frames, args.height, args.width, args.feature_dim, n_classes=2, seed=0, device=device)
    scene.gen.manual_seed(1000 + rank)                     # each rank samples its own rays
    model = ALNetwork(encoding='hg+freq', num_layers=2, hidden_dim=128, geo_feat_dim=15, num_layers_color=2,
                      hidden_dim_color=128, hidden_dim_semantic=args.feature_dim, semantic_classes=2,
                      bound=scene.bound(), cuda_ray=True, density_scale=1, density_thresh=args.density_thresh)
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=5e-3)
    trainer = SimpleTrainer('bench', opt, model, device=device, fp16=True, workspace=None, log_interval=0)
    model.train()
    model.train_t_thresh = args.train_t_thresh
    model.mark_untrained_grid(scene.poses, scene.intrinsics)
    return scene, model, trainer


def opt_lr(trainer):
    return float(trainer.optimizer.param_groups[0]['lr'])


def batch_bytes(b):
    return sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v))


def main():
    global RAYS
    args = parse()
    RAYS = args.rays
    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION prints to stdout) out of it
    os.environ["NCCL_DEBUG"] = os.environ.get("AL_NCCL_DEBUG", "WARN")
    if args.impl == "reference":
        return run_reference(args)
    from autolabel_b200 import _lib
    from autolabel_b200 import parallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    rank, world, local_rank = parallel.init_distributed()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    import torch.distributed as dist

    scene, model, trainer = build_trainer(args, device, rank)
    grad_exchange = "none (single GPU)"
    if world > 1:
        parallel.broadcast_parameters(model)
        grad_exchange = None
        if args.grad_exchange == "peer":
            try:
                peer = parallel.PeerShardedAdam(model, lr=opt_lr(trainer))
                trainer.optimizer = peer
                trainer.optimizers = [peer]
                trainer.grad_sync = None
                grad_exchange = ("peer memory kernel (al_peer_adam_step), " +
                                 ("multimem.ld_reduce / multimem.st (NVLS)" if peer.multicast else "peer loads / stores"))
            except Exception as e:  # symmetric memory unavailable on this box: keep training over NCCL, say so
                grad_exchange = f"nccl all_reduce + replicated Adam (peer path unavailable: {e!r})"
        if grad_exchange is None or grad_exchange.startswith("nccl"):
            trainer.grad_sync = parallel.GradientAllReduce(model.parameters(), trainer.optimizer)
            grad_exchange = grad_exchange or "nccl all_reduce + replicated Adam"

    # ---- untimed: converge the occupancy grid, then W warm-up steps
    for _ in range(args.pretrain):
        trainer.train_one_step(scene.next_train(RAYS))
    # A distinct batch for every warm-up / timed step of a leg (resident in HBM, resp. pinned host memory, before the
    # timed region starts).  Recycling a small pool lets the field overfit those rays within a few hundred steps: densities
    # sharpen along them, fewer samples stay alive and the step gets ~20 % faster than on fresh rays
    # (profiles/r1g_diag_step.txt was taken that way).
    n_pool = min(max(args.steps + args.warmup, 32), 600)

    def fresh(n=n_pool):
        return [scene.next_train(RAYS) for _ in range(n)]
    pool = fresh()
    host_pool = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in fresh()]      # never seen before the e2e leg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        host_ms[0] = (time.perf_counter() - t0) * 1e3 / max(steps, 1)   # host time to ENQUEUE a step (no sync)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(args.warmup):
        trainer.train_one_step(pool[(len(pool) - 1 - i) % len(pool)])
    if args.ncu_range > 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for i in range(args.ncu_range):
            trainer.train_one_step(pool[i % len(pool)])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_range_steps": args.ncu_range, "samples_per_ray": float(model.last_meta[1].item()) / RAYS}))
        return
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # kernels of this library executed in the timed region: direct enqueues (al_launch_count) + the kernels inside every
    # graph replay (counted when the graph was captured)
    launches0 = _lib.lib.al_launch_count() + trainer.graph_kernel_launches
    ms = timed(lambda i: trainer.train_one_step(pool[i % len(pool)]), args.steps)
    launches = _lib.lib.al_launch_count() + trainer.graph_kernel_launches - launches0
    host_enqueue_ms = host_ms[0]
    samples_per_ray = float(model.last_meta[1].item()) / RAYS
    alive_per_ray = float(model.last_alive_meta[0].item()) / RAYS
    loss_val = float(trainer.last_loss.item())

    # ---- end to end: host (pinned) batches in, loss out, every step
    def e2e_step(i):
        loss = trainer.train_one_step(host_pool[i % len(host_pool)])
        loss.item()
    for b in fresh(min(args.warmup, 5)):
        trainer.train_one_step({k: v.cpu().pin_memory() for k, v in b.items()}).item()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # the device-resident loop once more after the end-to-end leg (fresh rays again): shows how far the two legs drift apart
    # through continued training alone
    n_rep = max(args.steps // 2, 1)
    pool = fresh(n_rep)
    ms_rep = timed(lambda i: trainer.train_one_step(pool[i % len(pool)]), n_rep)
    repeat = {"value": RAYS * world * n_rep / (ms_rep * 1e-3), "unit": "rays/s", "ms_per_step": ms_rep / n_rep, "steps": n_rep,
              "alive_samples_per_ray": float(model.last_alive_meta[0].item()) / RAYS}

    # ---- the same step with every marched sample composited (train_t_thresh = 0), reported next to the headline
    exact = None
    if args.train_t_thresh > 0:
        model.train_t_thresh = 0.0
        n_exact = max(args.steps // 2, 1)
        for b in fresh(max(args.warmup, 3)):
            trainer.train_one_step(b)
        pool = fresh(n_exact)
        ms_exact = timed(lambda i: trainer.train_one_step(pool[i % len(pool)]), n_exact)
        exact = {"value": RAYS * world * n_exact / (ms_exact * 1e-3), "unit": "rays/s", "ms_per_step": ms_exact / n_exact,
                 "steps": n_exact, "train_t_thresh": 0.0}
        model.train_t_thresh = args.train_t_thresh
        for b in fresh(3):
            trainer.train_one_step(b)

    value = RAYS * world * args.steps / (ms * 1e-3)
    e2e_value = RAYS * world * args.steps / (ms_e2e * 1e-3)

    detail = phase_detail(args, scene, model, trainer, device) if rank == 0 else None
    render = None
    if args.render_frames > 0:
        try:
            render = render_leg(args, scene, model, device, rank, world, timed)
        except Exception as e:  # the training line must survive a failure of the second leg
            render = {"metric": "render_frames_per_s", "error": repr(e)}
            model.train()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        rate, dt, threads, sample = cpu_port_rate(args, steps=2, warmup=1)
        cpu = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample, "ms_per_step": dt * 1e3}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "host_enqueue_ms_per_step": host_enqueue_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "config": dict(workload_config(args, world), pretrain_steps=args.pretrain,
                                                grad_exchange=grad_exchange,
                                                samples_per_ray=samples_per_ray, alive_samples_per_ray=alive_per_ray,
                                                final_loss=loss_val,
                                                l2="per-step working set (57 MB table + 57 MB gradients + 114 MB Adam moments "
                                                   "+ per-sample buffers) exceeds the 126 MB L2; no explicit flush"),
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": batch_bytes(host_pool[0]), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": detail["roofline"] if detail else None,
            "cpu_baseline": cpu,
        }
        line["value_repeat_after_e2e"] = repeat
        if exact:
            line["exact_compositing"] = exact
        if detail:
            line["phases_ms"] = detail["phases_ms"]
        if render:
            line["render"] = render
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    return {"metric": "render_frames_per_s", "value": n * world / (ms * 1e-3), "unit": "frames/s",
            "resolution": [W, H], "frames_per_rank": n, "ms_per_frame": ms / n, "samples_per_ray": spr[0],
            "early_termination": bool(model.early_termination),
            "outputs": "image, depth, depth_variance, semantic logits, semantic_features, coordinates_map",
            "e2e": {"value": n * world / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_frame": ms_e2e / n,
                    "h2d_bytes_per_frame": H * W * 4 * 7, "d2h_bytes_per_frame": out_bytes}}


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
