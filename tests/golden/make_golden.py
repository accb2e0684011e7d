"""Freeze outputs of the reference's OWN kernels (oracle/_ref = torch_ngp's raymarching.cu /
gridencoder.cu compiled unmodified) into small fixtures that pin the CPU oracle.

Run on a CUDA box from the repo root:  python tests/golden/make_golden.py
Writes tests/golden/ref_*.npz (and a copy under gpurun_out/golden/ so it travels back).
Inputs are regenerated from seeds by tests/helpers.py; the fixtures store them anyway so the CPU
tests do not depend on numpy's RNG stream staying stable.
"""
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ngp  # noqa: E402
from tests.helpers import aabb_of, make_density_grid, make_rays  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BOUND, CASCADE, H = 3.0, 3, 128


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sort_by_ray(rays, arrs, M):
    """Per-ray data in ray-id order (the reference's slot order is atomic-order dependent)."""
    rays = rays.cpu().numpy()
    order = np.argsort(rays[:, 0])
    counts = np.zeros(rays.shape[0], np.int32)
    chunks = [[] for _ in arrs]
    for i in order:
        rid, off, cnt = rays[i]
        counts[rid] = cnt
        if cnt > 0 and off + cnt < M:
            for k, a in enumerate(arrs):
                chunks[k].append(a[off:off + cnt])
    return counts, [np.concatenate(c, axis=0) for c in chunks]


def main():
    rm = build_ref.load_ref("ref_raymarching")
    ge = build_ref.load_ref("ref_gridencoder")
    N = 512
    o, d = make_rays(N, BOUND, seed=21, inside=False)
    grid = make_density_grid(CASCADE, H, seed=22, fill=0.06)
    aabb = aabb_of(BOUND)
    O, D, A, G = dev(o), dev(d), dev(aabb), dev(grid)

    nears, fars = torch.empty(N, device='cuda'), torch.empty(N, device='cuda')
    ni, fi = torch.empty(N, dtype=torch.uint8, device='cuda'), torch.empty(N, dtype=torch.uint8, device='cuda')
    rm.near_far_from_aabb(O, D, A, N, 0.2, nears, fars, ni, fi)
    bits = torch.empty(CASCADE * H ** 3 // 8, dtype=torch.uint8, device='cuda')
    rm.packbits(G, bits.numel(), 0.01, bits)

    coords = np.random.RandomState(23).randint(0, 128, size=(4096, 3)).astype(np.int32)
    mort = torch.empty(4096, dtype=torch.int32, device='cuda')
    rm.morton3D(dev(coords), 4096, mort)

    fix = dict(rays_o=o, rays_d=d, aabb=aabb, nears=nears.cpu().numpy(), fars=fars.cpu().numpy(),
               near_idx=ni.cpu().numpy(), far_idx=fi.cpu().numpy(), bits=bits.cpu().numpy(),
               grid_seed=np.array([22]), grid_fill=np.array([0.06]), coords=coords, morton=mort.cpu().numpy())

    M = N * 1024
    for tag, perturb, dtg in [("p", 1, 0.0), ("np", 0, 0.0), ("pg", 1, 1.0 / 256)]:
        xyzs = torch.zeros(M, 3, device='cuda'); dirs = torch.zeros(M, 3, device='cuda')
        deltas = torch.zeros(M, 2, device='cuda'); ts = torch.zeros(M, 1, device='cuda')
        rays = torch.empty(N, 3, dtype=torch.int32, device='cuda')
        counter = torch.zeros(2, dtype=torch.int32, device='cuda')
        rm.march_rays_train(O, D, bits, BOUND, dtg, 1024, N, CASCADE, H, M, nears, fars, xyzs, dirs, deltas, ts, rays,
                            counter, perturb)
        torch.cuda.synchronize()
        counts, (x, dl, t) = sort_by_ray(rays, [xyzs.cpu().numpy(), deltas.cpu().numpy(), ts.cpu().numpy()], M)
        fix[f"march_{tag}_counts"] = counts
        fix[f"march_{tag}_xyzs"] = x
        fix[f"march_{tag}_deltas"] = dl
        fix[f"march_{tag}_ts"] = t[:, 0]

    # inference marching: one iteration of 4 steps from the near plane
    alive = torch.arange(N, dtype=torch.int32, device='cuda')
    x = torch.zeros(N * 4, 3, device='cuda'); dd = torch.zeros(N * 4, 3, device='cuda'); dl = torch.zeros(N * 4, 2, device='cuda')
    rm.march_rays(N, 4, alive, nears.clone(), O, D, BOUND, 0.0, 1024, CASCADE, H, bits, nears, fars, x, dd, dl, 0)
    fix["infer_xyzs"], fix["infer_deltas"] = x.cpu().numpy(), dl.cpu().numpy()

    # compositing (3 channels): forward + backward of the reference kernels
    counts = fix["march_p_counts"]
    tot = int(counts.sum())
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays_seq = np.stack([np.arange(N, dtype=np.int32), offs, counts], axis=1).astype(np.int32)
    Mc = tot + 128 - tot % 128
    g = torch.Generator().manual_seed(24)
    sig = (torch.rand(Mc, generator=g) * 8).cuda()
    rgb = torch.rand(Mc, 3, generator=g).cuda()
    dlt = torch.zeros(Mc, 2, device='cuda'); dlt[:tot] = dev(fix["march_p_deltas"])
    ws, dep, img = torch.empty(N, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, 3, device='cuda')
    R = dev(rays_seq)
    rm.composite_rays_train_forward(sig, rgb, dlt, R, Mc, N, ws, dep, img)
    gws, gim = torch.randn(N, generator=g).cuda(), torch.randn(N, 3, generator=g).cuda()
    gs, gr = torch.zeros(Mc, device='cuda'), torch.zeros(Mc, 3, device='cuda')
    rm.composite_rays_train_backward(gws, gim, sig, rgb, dlt, R, ws, img, Mc, N, gs, gr)
    fix.update(comp_rays=rays_seq, comp_sigmas=sig.cpu().numpy(), comp_rgbs=rgb.cpu().numpy(), comp_deltas=dlt.cpu().numpy(),
               comp_ws=ws.cpu().numpy(), comp_depth=dep.cpu().numpy(), comp_image=img.cpu().numpy(),
               comp_gws=gws.cpu().numpy(), comp_gimage=gim.cpu().numpy(), comp_gsigmas=gs.cpu().numpy(),
               comp_grgbs=gr.cpu().numpy())
    np.savez_compressed(os.path.join(OUT, "ref_raymarching.npz"), **fix)

    # hash grid: hg+freq hyper-parameters, probe table (entry index in channel 0) + random table values
    offsets = ngp.grid_offsets(16, 16, 2.0, 19, 3)
    L, B = 16, 2048
    rng = np.random.RandomState(25)
    xin = rng.uniform(0, 1, size=(B, 3)).astype(np.float32)
    xin[:16] = rng.uniform(-0.1, 1.1, size=(16, 3)).astype(np.float32)
    xin[16:20] = np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1, 0, 0.25]], np.float32)
    tseed = 26
    table = np.random.RandomState(tseed).uniform(-0.1, 0.1, size=(int(offsets[-1]), 2)).astype(np.float32)
    out = torch.empty(L, B, 2, device='cuda')
    ge.grid_encode_forward(dev(xin), dev(table), dev(offsets), out, B, 3, 2, L, 1.0, 16, False, torch.empty(1, device='cuda'), 0)
    gfix = dict(inputs=xin, offsets=offsets, table_seed=np.array([tseed]), outputs=out.cpu().numpy())
    # corner-aligned probe at every level: the kernel output IS the entry index of corner 0
    probe = np.zeros_like(table)
    for l in range(L):
        probe[offsets[l]:offsets[l + 1], 0] = np.arange(offsets[l + 1] - offsets[l])
    probe_idx = np.zeros((L, 256), np.float32)
    probe_x = np.zeros((L, 256, 3), np.float32)
    for l in range(L):
        scale = np.float32(2.0 ** l * 16 - 1)
        res = int(np.ceil(scale)) + 1
        cells = rng.randint(1, res - 1, size=(256, 3))
        px = ((cells.astype(np.float64) - 0.5) / np.float64(scale))
        px = np.clip(px, 0, 1).astype(np.float32)
        po = torch.empty(L, 256, 2, device='cuda')
        ge.grid_encode_forward(dev(px), dev(probe), dev(offsets), po, 256, 3, 2, L, 1.0, 16, False, torch.empty(1, device='cuda'), 0)
        probe_idx[l] = po[l, :, 0].cpu().numpy()
        probe_x[l] = px
    gfix.update(probe_x=probe_x, probe_out=probe_idx)
    gl = torch.randn(L, B, 2, generator=g).cuda()
    gt = torch.zeros(int(offsets[-1]), 2, device='cuda')
    ge.grid_encode_backward(gl, dev(xin), dev(table), dev(offsets), gt, B, 3, 2, L, 1.0, 16, False, torch.empty(1, device='cuda'),
                            torch.empty(1, device='cuda'), 0)
    nz = torch.nonzero(gt.abs().sum(1)).view(-1)
    gfix.update(bwd_grad=gl.cpu().numpy(), bwd_rows=nz.cpu().numpy().astype(np.int32), bwd_vals=gt[nz].cpu().numpy())
    np.savez_compressed(os.path.join(OUT, "ref_gridencoder.npz"), **gfix)

    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    for f in ("ref_raymarching.npz", "ref_gridencoder.npz"):
        shutil.copy(os.path.join(OUT, f), os.path.join(dst, f))
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
