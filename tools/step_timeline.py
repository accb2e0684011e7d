#!/usr/bin/env python
"""In-situ timeline of one training step: CUDA events around every C-ABI call of the eager (kernel-by-kernel) step,
averaged over the steps of a steady-state window.  Unlike bench_detail.py (each kernel re-launched back to back with
warm caches) these durations are what the kernels take INSIDE the step, after the optimiser has streamed 400 MB
through L2.  Usage: python tools/step_timeline.py [--steps 48] [--pretrain 2000] > profiles/<round>_step_timeline.md"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--pretrain", type=int, default=2000)
    a = ap.parse_args()
    import bench
    from autolabel_b200 import _lib, renderer, trainer as trainer_mod, optim
    sys.argv = [sys.argv[0]]
    args = bench.parse()
    dev = torch.device("cuda", 0)
    scene, model, trainer = bench.build_trainer(args, dev, 0)
    for _ in range(a.pretrain):
        trainer.train_one_step(scene.next_train(bench.RAYS))
    trainer.use_graph = False
    pool = [scene.next_train(bench.RAYS) for _ in range(16)]
    for i in range(16):
        trainer.train_one_step(pool[i % 16])
    torch.cuda.synchronize()

    records = []          # (name, start_event, end_event)
    orig_call = _lib.call

    def timed_call(name, *args_):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_call(name, *args_)
        e1.record()
        records.append((name, e0, e1))
        return r

    for mod in (_lib, renderer, trainer_mod, optim):
        if hasattr(mod, "call"):
            mod.call = timed_call
    from autolabel_b200 import raymarching, models
    raymarching.call = timed_call
    models.call = timed_call

    step_marks = []
    for i in range(a.steps):
        s0 = torch.cuda.Event(enable_timing=True)
        s0.record()
        n0 = len(records)
        trainer.train_one_step(pool[i % 16])
        s1 = torch.cuda.Event(enable_timing=True)
        s1.record()
        step_marks.append((s0, s1, n0, len(records), trainer.global_step - 1))
    torch.cuda.synchronize()

    agg = collections.OrderedDict()
    refresh_steps, plain = [], []
    for s0, s1, n0, n1, gs in step_marks:
        total = s0.elapsed_time(s1)
        (refresh_steps if gs % trainer.update_interval == 0 else plain).append(total)
        if gs % trainer.update_interval == 0:
            continue
        seen = collections.Counter()
        prev_end = s0
        for name, e0, e1 in records[n0:n1]:
            seen[name] += 1
            key = f"{name}#{seen[name]}" if name == "al_adam_step" else name
            d = agg.setdefault(key, [0.0, 0.0, 0])
            d[0] += e0.elapsed_time(e1)
            d[1] += prev_end.elapsed_time(e0)
            d[2] += 1
            prev_end = e1
        d = agg.setdefault("(tail: after the last call)", [0.0, 0.0, 0])
        d[1] += prev_end.elapsed_time(s1)
        d[2] += 1
    live = int(model.last_meta[0].item())
    print(f"# In-situ step timeline (eager step, {len(plain)} plain steps, {live} live samples, "
          f"{live / bench.RAYS:.1f} samples/ray)\n")
    print(f"plain step: {sum(plain) / len(plain):.3f} ms;  step with occupancy refresh: "
          f"{(sum(refresh_steps) / len(refresh_steps)) if refresh_steps else float('nan'):.3f} ms "
          f"({len(refresh_steps)} of {len(step_marks)})\n")
    print("| C-ABI call | in-call ms | gap before ms |\n|---|---|---|")
    tot_in = tot_gap = 0.0
    for k, (t_in, t_gap, n) in agg.items():
        print(f"| `{k}` | {t_in / n:.4f} | {t_gap / n:.4f} |")
        tot_in += t_in / n
        tot_gap += t_gap / n
    print(f"| **sum** | {tot_in:.4f} | {tot_gap:.4f} |")


if __name__ == "__main__":
    main()
