"""al_render_epilogue against oracle/postprocess_oracle.py on the same seeded frame: labels
bit-exact wherever the two best candidates are separated by more than fp32 summation noise, uint8 colours within 1."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F,C,T", [(64, 2, 5), (512, 40, 150), (64, 606, 33)])
def test_render_epilogue_matches_oracle(F, C, T):
    from autolabel_b200.postprocess import FeatureTransformer, compute_semantics, render_epilogue
    from oracle import postprocess_oracle as po
    rng = np.random.RandomState(F + C)
    H, W = 30, 41
    N = H * W
    K = 3 + C + F
    out = rng.normal(size=(N, K)).astype(np.float32)                # the compositing buffer (rgb | logits | features)
    out[:, :3] = rng.uniform(0, 1, size=(N, 3))
    out[5, 3:3 + C] = 0.25                                          # an exact tie: first index wins
    text = rng.normal(size=(T, F)).astype(np.float32)
    mean = rng.normal(size=F).astype(np.float32)
    comp = np.linalg.qr(rng.normal(size=(F, 3)))[0].T.astype(np.float32)
    proj = po.pca_project(out[:, 3 + C:], mean, comp)
    fmin, frange = proj.min(0) * 0.8, (proj.max(0) - proj.min(0)) * 0.7     # some values clip on both sides
    d = torch.from_numpy(out).cuda()
    outputs = {'image': d[:, :3].view(H, W, 3), 'semantic': d[:, 3:3 + C].view(H, W, C),
               'semantic_features': d[:, 3 + C:].view(H, W, F)}
    ft = FeatureTransformer(mean, comp, fmin, frange, text)
    r = render_epilogue(outputs, feature_transform=ft)
    torch.cuda.synchronize()
    assert r['rgb8'].shape == (H, W, 3) and r['label'].shape == (H, W) and r['pca8'].dtype == torch.uint8
    assert np.array_equal(r['rgb8'].cpu().numpy().reshape(N, 3), po.rgb_u8(out[:, :3]))
    lab = r['label'].cpu().numpy().reshape(-1)
    assert np.array_equal(lab, po.semantic_argmax(out[:, 3:3 + C])) and lab[5] == 0
    sims = po.text_similarities(out[:, 3 + C:].astype(np.float64), text.astype(np.float64))
    srt = np.sort(sims, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-5
    tl = r['text_label'].cpu().numpy().reshape(-1)
    assert clear.mean() > 0.99 and np.array_equal(tl[clear], np.argmax(sims, 1)[clear])
    pc = r['pca8'].cpu().numpy().reshape(N, 3).astype(int)
    want = po.pca_colors(out[:, 3 + C:], mean, comp, fmin, frange).astype(int)
    assert np.abs(pc - want).max() <= 1 and (pc == want).mean() > 0.98
    assert (want == 0).any() and (want == 255).any()
    # the script-level helpers
    assert torch.equal(compute_semantics(outputs, None, ft), r['label'])
    assert torch.equal(compute_semantics(outputs, ['a'] * T, ft), r['text_label'])
    assert torch.equal(ft(outputs['semantic_features']), r['pca8'])
