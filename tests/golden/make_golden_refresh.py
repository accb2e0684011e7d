"""Freeze what the reference's OWN `NeRFRenderer.update_extra_state` (torch_ngp/nerf/renderer.py:563-683) computes into
tests/golden/ref_update_extra_state.npz.   Runs in the dev container (CPU; needs /root/reference):

    python tests/golden/make_golden_refresh.py

The reference function is executed UNMODIFIED on CPU.  Replaced around it (none of it is the code under test):
  * `raymarching.morton3D / morton3D_invert / packbits` (CUDA-only in the reference) -> the C oracle
    (oracle/ngp_oracle.c, pinned bit-exactly on the reference's own kernels, tests/test_oracle_pinned.py);
  * `self.density` -> tests/helpers.checker_density, a density field that is exact in fp32 on every device;
  * the random draws `torch.rand_like / torch.randint` -> tests/helpers.TorchRngTape (a seeded numpy stream; torch's
    CPU and CUDA generators differ, the tape makes the draws a function of call order and shapes only).
Scenario: grid marked by `mark_untrained_grid` (golden poses), one FULL refresh (iter_density < 16, :575-619), then a
PARTIAL refresh (iter_density >= 16, :623-654) of the field shifted by 0.125 (so the sampled cells change value), with a populated step counter for the mean_count rule (:677-680).
The density grid after each refresh is stored as fp64 sums over blocks of 4096 consecutive (Morton-ordered) cells, the
bitfield in full.  Why not a hash of the grid: torch evaluates `2 * coords / (H - 1)` as a true division on the CPU and as
a multiplication by the fp32 reciprocal on CUDA, so query positions differ in the last bit between the device this golden
is made on (CPU) and the device the product runs on; a handful of the 4.2 M jittered queries then fall on the other side
of a checker boundary.  Block sums localise such cells (the test allows a few blocks to differ by one checker step) while
any difference in control flow — decay, update rule, draw order, Morton mapping — changes essentially every block.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
OUT = os.path.join(ROOT, "tests", "golden", "ref_update_extra_state.npz")

BOUND, DENSITY_THRESH, SEED = 2.0, 10.0, 20261018
STEP_COUNTS = [41000, 39500, 40210, 38777, 40001]          # local_step = 5 marched-sample totals


def main():
    from make_golden_run import import_reference
    from oracle import ngp
    from tests.helpers import TorchRngTape, checker_density
    models, raymarching = import_reference()
    import torch_ngp.nerf.renderer as ref_renderer

    def morton_cpu(coords):
        return torch.from_numpy(ngp.morton3D(coords.numpy().astype(np.int32)).astype(np.int32))

    def morton_inv_cpu(indices):
        return torch.from_numpy(ngp.morton3D_invert(indices.numpy().astype(np.int32)).astype(np.int32))

    def packbits_cpu(grid, thresh, bitfield=None):
        return torch.from_numpy(ngp.packbits(grid.numpy(), float(thresh)))
    ref_renderer.raymarching.morton3D = morton_cpu
    ref_renderer.raymarching.morton3D_invert = morton_inv_cpu
    ref_renderer.raymarching.packbits = packbits_cpu

    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_run_path.npz"))
    m = models.ALNetwork(encoding='freq', num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=2,
                         hidden_dim_color=64, hidden_dim_semantic=64, semantic_classes=2, bound=BOUND, cuda_ray=True,
                         density_thresh=DENSITY_THRESH)
    shift = [0.0]                                         # the partial refresh sees a shifted field (values really change)
    m.density = lambda x: {'sigma': checker_density(x + shift[0])}
    m.mark_untrained_grid(g["mark_poses"], g["mark_intrinsics"])
    unseen = np.unpackbits(g["mark_unseen_bits"])[:m.density_grid.numel()].astype(bool)
    assert np.array_equal(unseen, (m.density_grid < 0).numpy().reshape(-1))     # the mask the run-path golden froze
    tape = TorchRngTape(torch, SEED)
    ref_renderer.torch = tape
    save = {"bound": np.float32(BOUND), "density_thresh": np.float32(DENSITY_THRESH), "seed": np.int64(SEED),
            "step_counts": np.asarray(STEP_COUNTS, np.int32)}
    try:
        for stage, iter_density, sh in (("full", 0, 0.0), ("partial", 16, 0.125)):
            m.iter_density = iter_density
            shift[0] = sh
            m.step_counter.zero_()
            m.step_counter[:len(STEP_COUNTS), 0] = torch.tensor(STEP_COUNTS, dtype=torch.int32)
            m.local_step = len(STEP_COUNTS)
            m.update_extra_state()
            grid = m.density_grid.numpy()
            save[stage + "_block_sums"] = grid.astype(np.float64).reshape(-1, 4096).sum(axis=1)
            save[stage + "_grid_sum"] = np.float64(grid.astype(np.float64).sum())
            save[stage + "_occupied"] = np.int64((grid > 0).sum())
            save[stage + "_bitfield"] = m.density_bitfield.numpy().copy()
            save[stage + "_mean_density"] = np.float64(m.mean_density)
            save[stage + "_mean_count"] = np.int64(m.mean_count)
            save[stage + "_iter_density"] = np.int64(m.iter_density)
            save[stage + "_local_step"] = np.int64(m.local_step)
            print(stage, "occupied", int((grid > 0).sum()), "unseen", int((grid < 0).sum()), "mean", m.mean_density,
                  "mean_count", m.mean_count, "bits set", int(np.unpackbits(m.density_bitfield.numpy()).sum()))
    finally:
        ref_renderer.torch = torch
    save["n_rng_calls"] = np.int64(len(tape.calls))
    np.savez_compressed(OUT, **save)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(tape.calls), "random draws")


if __name__ == "__main__":
    main()
