"""Device-resident dataset (autolabel_b200/dataset.py, csrc/dataset.cu) against the reference's own sampler output
(tests/golden/ref_dataset.npz, generated from autolabel/dataset.py by tests/golden/make_golden_dataset.py) and the
numpy oracle on identical draws."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_dataset.npz"))


def _dataset(batch_size=2048, seed=0):
    from autolabel_b200.dataset import DeviceSceneDataset
    w, h, fw, fh, F = (int(v) for v in G['meta'])
    return DeviceSceneDataset(G['images'], G['depths'], G['semantics'], G['poses'], tuple(G['intrinsics']), (w, h),
                              features=G['features'], feature_size=(fw, fh), batch_size=batch_size, seed=seed)


def test_train_batch_matches_reference_on_its_draws():
    ds = _dataset()
    out = ds.sample_batch(torch.from_numpy(G['image_index']).cuda(), torch.from_numpy(G['ray_indices']).cuda(), None)
    for k, gk in [('rays_o', 'train_rays_o'), ('direction_norms', 'train_norms'), ('pixels', 'train_pixels'),
                  ('depth', 'train_depth'), ('semantic', 'train_semantic'), ('features', 'train_features')]:
        a = out[k].cpu().numpy()
        assert a.dtype == G[gk].dtype and a.shape == G[gk].shape, k
        assert np.array_equal(a, G[gk]), k                                              # bit-exact
    assert np.abs(out['rays_d'].cpu().numpy() - G['train_rays_d']).max() <= 1e-6          # BLAS mat-vec order: 1 ulp


def test_full_frame_rays_match_reference():
    ds = _dataset()
    t = ds._get_test(2)
    w, h = ds.w, ds.h
    assert np.array_equal(t['rays_o'].cpu().numpy(), G['test_rays_o'])
    assert np.array_equal(t['direction_norms'].cpu().numpy(), G['test_norms'])
    assert np.abs(t['rays_d'].cpu().numpy() - G['test_rays_d']).max() <= 1e-6
    assert np.array_equal(t['depth'].cpu().numpy(), G['test_depth'])
    assert np.array_equal(t['semantic'].cpu().numpy(), G['test_semantic'])
    assert t['pixels'].shape == (h, w, 3) and t['H'] == h and t['W'] == w


def test_jittered_batch_matches_oracle():
    from oracle import dataset_oracle as do
    ds = _dataset()
    img, idx, jit = ds.draw()
    out = ds.sample_batch(img, idx, jit)
    w, h, fw, fh, F = (int(v) for v in G['meta'])
    R = np.ascontiguousarray(G['poses'][:, :3, :3])
    ref = do.next_train(G['images'], G['depths'], G['semantics'], G['features'], R, G['poses'][:, :3, 3], w, fw, fh, h,
                        tuple(float(v) for v in G['intrinsics']), img.cpu().numpy(), idx.cpu().numpy(), jit.cpu().numpy())
    for k in ('rays_o', 'rays_d', 'direction_norms', 'pixels', 'depth', 'semantic', 'features'):
        assert np.array_equal(out[k].cpu().numpy(), ref[k]), k      # the kernel follows the oracle's operation order exactly


def test_sampling_policy():
    """Half of the chunks (in expectation) come from labelled pixels of one class (dataset.py:204-213)."""
    ds = _dataset(batch_size=4096, seed=5)
    labelled_chunks, total = 0, 0
    for _ in range(40):
        b = ds._next_train()
        assert b['rays_o'].shape == (4096, 3) and b['features'].shape == (4096, ds.feature_dim)
        sem = b['semantic'].view(-1, 512)
        same = (sem == sem[:, :1]).all(dim=1) & (sem[:, 0] >= 0)
        labelled_chunks += int(same.sum())
        total += sem.shape[0]
        assert torch.allclose(b['rays_d'].norm(dim=1), torch.ones(4096, device='cuda'), atol=1e-5)
    assert 0.3 < labelled_chunks / total < 0.7
    it = iter(ds)
    assert set(next(it)) == {'rays_o', 'rays_d', 'direction_norms', 'pixels', 'depth', 'semantic', 'features'}


def test_trainer_consumes_device_batches():
    """A batch from the device dataset goes straight into SimpleTrainer.train_one_step (no host copies)."""
    from types import SimpleNamespace
    from autolabel_b200.models import ALNetwork
    from autolabel_b200.trainer import SimpleTrainer
    ds = _dataset(batch_size=1024, seed=1)
    torch.manual_seed(0)
    m = ALNetwork(encoding='hg+freq', num_layers=2, hidden_dim=128, num_layers_color=2, hidden_dim_color=128,
                  hidden_dim_semantic=64, semantic_classes=2, bound=2.0, cuda_ray=True).cuda()
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=5e-3)
    tr = SimpleTrainer('ds', opt, m, device='cuda', workspace=None, log_interval=0)
    m.train()
    for _ in range(3):
        loss = tr.train_one_step(ds._next_train())
    assert torch.isfinite(loss).item()
