#!/usr/bin/env python
"""Where does a training step's wall time go?  (a) drift: ms/step and alive samples/ray over consecutive 100-step
blocks after the bench's pre-training; (b) host: time the host needs to enqueue each part of train_one_step
(perf_counter, no synchronisation inside the block) next to the device time of the block; (c) the same block
with a synchronisation after every step (what `e2e` does).  Usage: python tools/diag_step.py > gpurun_out/diag.txt"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    sys.argv = [sys.argv[0]]
    args = bench.parse()
    dev = torch.device("cuda", 0)
    scene, model, trainer = bench.build_trainer(args, dev, 0)
    R = bench.RAYS
    for _ in range(args.pretrain):
        trainer.train_one_step(scene.next_train(R))
    pool = [scene.next_train(R) for _ in range(32)]

    def block(n, sync_each=False):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(n):
            loss = trainer.train_one_step(pool[i % 32])
            if sync_each:
                loss.item()
        host = (time.perf_counter() - t0) * 1e3 / n
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, host

    print("block  ms/step  host_enqueue_ms  marched/ray  alive/ray  mean_count")
    for b in range(8):
        ms, host = block(100)
        print(f"{b:3d}  {ms:7.3f}  {host:7.3f}  {float(model.last_meta[1].item()) / R:8.1f}  "
              f"{float(model.last_alive_meta[0].item()) / R:8.1f}  {model.mean_count}")
    ms, host = block(100, sync_each=True)
    print(f"sync-each-step block: {ms:.3f} ms/step (host {host:.3f})")
    # 16-step windows without the occupancy refresh
    keep = trainer.update_interval
    trainer.update_interval = 10 ** 9
    ms, host = block(96)
    print(f"no occupancy refresh: {ms:.3f} ms/step (host {host:.3f})")
    trainer.update_interval = keep

    # host time of the parts of one step
    import collections
    acc = collections.OrderedDict()

    def tick(name, fn):
        t0 = time.perf_counter()
        r = fn()
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    n = 64
    torch.cuda.synchronize()
    for i in range(n):
        data = pool[i % 32]
        if trainer.global_step % trainer.update_interval == 0:
            tick("update_extra_state", model.update_extra_state)
        tick("zero_grad", lambda: [o.zero_grad() for o in trainer.optimizers])
        tick("graph_step", lambda: trainer._graph_train_step(data))
        tick("optimizer.step", lambda: [o.step() for o in trainer.optimizers])
        trainer.global_step += 1
        if i % 16 == 15:
            tick("sync", torch.cuda.synchronize)
    torch.cuda.synchronize()
    print("host ms per step by part:", {k: round(v / n, 4) for k, v in acc.items()})


if __name__ == "__main__":
    main()
