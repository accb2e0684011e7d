"""al_compact_alive (training-time early termination): alive-prefix counts against oracle/field_oracle.alive_prefix,
the packed rows against the source rows, and t_thresh = 0 as the identity."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _segments(N, seed, max_len=300):
    rng = np.random.RandomState(seed)
    counts = rng.randint(0, max_len, size=N).astype(np.int32)
    counts[rng.rand(N) < 0.1] = 0
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays = np.stack([np.arange(N, dtype=np.int32), offsets, counts], axis=1)
    return rays, int(counts.sum())


def _run(thresh, N=700, seed=3, in_pad=48, sigma_hi=400.0, drop_last=False):
    from autolabel_b200._lib import call, ptr, stream_ptr
    rays_np, total = _segments(N, seed)
    M = total + 64 if not drop_last else total     # offset + count >= M drops the last non-empty ray (raymarching.cu:459)
    g = torch.Generator().manual_seed(seed)
    dev = 'cuda'
    sigma = (torch.rand(M, generator=g) ** 4 * sigma_hi).to(dev)
    deltas = torch.stack([torch.full((M,), 0.0034), torch.rand(M, generator=g) * 0.02], 1).to(dev)
    xyzs, tpos = torch.randn(M, 3, generator=g).to(dev), torch.rand(M, generator=g).to(dev)
    sray = torch.randint(0, N, (M,), generator=g, dtype=torch.int32).to(dev)
    x_enc = torch.randn(M, in_pad, generator=g).half().to(dev)
    h16 = torch.randn(M, 16, generator=g).to(dev)
    rays = torch.from_numpy(rays_np).to(dev)
    ldv = 7
    o = dict(rays_c=torch.full((N, 3), -7, dtype=torch.int32, device=dev), meta_c=torch.zeros(2, dtype=torch.int32, device=dev),
             xyzs=torch.zeros(M, 3, device=dev), deltas=torch.zeros(M, 2, device=dev), tpos=torch.zeros(M, device=dev),
             sray=torch.zeros(M, dtype=torch.int32, device=dev), x_enc=torch.zeros(M, in_pad, dtype=torch.float16, device=dev),
             h16=torch.zeros(M, 16, device=dev), vals=torch.zeros(M, ldv, device=dev))
    ws = torch.empty(N, dtype=torch.int32, device=dev)
    call("al_compact_alive", ptr(sigma), ptr(deltas), ptr(rays), M, N, 1.5, thresh, ptr(xyzs), ptr(tpos), ptr(sray),
         ptr(x_enc), in_pad, ptr(h16), ptr(o['rays_c']), ptr(o['meta_c']), ptr(o['xyzs']), ptr(o['deltas']), ptr(o['tpos']),
         ptr(o['sray']), ptr(o['x_enc']), ptr(o['h16']), ptr(o['vals']), ldv, ptr(ws), stream_ptr(torch.device(dev)))
    torch.cuda.synchronize()
    src = dict(sigma=sigma, deltas=deltas, xyzs=xyzs, tpos=tpos, sray=sray, x_enc=x_enc, h16=h16, rays=rays, M=M)
    return src, o


@pytest.mark.parametrize("in_pad", [48, 64])
def test_alive_prefix_counts_and_rows(in_pad):
    from oracle import field_oracle as fo
    thresh = 1e-4
    src, o = _run(thresh, in_pad=in_pad)
    counts, T_before = fo.alive_prefix(src['sigma'], src['deltas'], src['rays'], src['M'], 1.5, thresh)
    rc = o['rays_c'].cpu().numpy()
    rays = src['rays'].cpu().numpy()
    assert np.array_equal(rc[:, 0], rays[:, 0])
    got = rc[:, 2].astype(np.int64)
    want = counts.numpy()
    # __expf / fp32 products vs float64: counts may differ by one where T sits within 1e-3 (relative) of the threshold
    diff = np.nonzero(got != want)[0]
    for n in diff:
        off = rays[n, 1]
        k = min(got[n], want[n])
        assert abs(got[n] - want[n]) == 1 and abs(T_before[off + k].item() / thresh - 1) < 1e-3, (n, got[n], want[n])
    assert len(diff) <= 3
    assert (got < rays[:, 2]).mean() > 0.3 and (got == rays[:, 2]).mean() > 0.1, "the case must mix cut and uncut rays"
    assert np.array_equal(rc[:, 1], np.concatenate([[0], np.cumsum(got)[:-1]]))
    assert o['meta_c'].cpu().tolist() == [int(got.sum())] * 2
    for n in range(rays.shape[0]):
        a, s0, d0 = got[n], rays[n, 1], rc[n, 1]
        if a == 0:
            continue
        for k in ('xyzs', 'deltas', 'tpos', 'sray', 'x_enc', 'h16'):
            assert torch.equal(o[k][d0:d0 + a], src[k][s0:s0 + a]), (k, n)
        assert torch.equal(o['vals'][d0:d0 + a, 0], src['sigma'][s0:s0 + a])
    assert float(o['vals'][:, 1:].abs().sum()) == 0          # only column 0 of the vals matrix is touched


def test_zero_threshold_is_the_identity_and_dropped_rays_stay_empty():
    src, o = _run(0.0, drop_last=True)
    rays = src['rays'].cpu().numpy()
    rc = o['rays_c'].cpu().numpy()
    valid = (rays[:, 2] > 0) & (rays[:, 1] + rays[:, 2] < src['M'])
    assert (~valid & (rays[:, 2] > 0)).sum() == 1              # exactly the last non-empty ray is over budget
    assert np.array_equal(rc[:, 2], np.where(valid, rays[:, 2], 0))
    assert np.array_equal(rc[valid][:, 1], rays[valid][:, 1])   # nothing before the dropped ray moves
    tot = int(rc[:, 2].sum())
    assert torch.equal(o['xyzs'][:tot], src['xyzs'][:tot]) and torch.equal(o['x_enc'][:tot], src['x_enc'][:tot])
