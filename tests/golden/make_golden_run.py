"""Freeze outputs of the reference's OWN model + renderer code on the path it executes today
(`autolabel/models.py` ALNetwork.density / color / semantic, `torch_ngp/nerf/renderer.py` NeRFRenderer.run, :186-320,
`autolabel/trainer.py` SimpleTrainer.train_step loss, :54-94) into tests/golden/ref_run_path.npz.

Runs in the dev container (CPU; needs /root/reference):  python tests/golden/make_golden_run.py

The reference modules are imported UNMODIFIED.  What is absent from the image is replaced as follows:
  * `tinycudann` (a git dependency, not vendored): a small fp32 torch module with the semantics oracle/field_oracle.py
    states (Frequency: sin/cos(2^k pi x) per the oracle's ordering; SphericalHarmonics degree 4 on 2x-1; bias-free
    Linear/ReLU networks whose inputs are padded with ONES to a multiple of 16, flat row-major [out, in] parameters).
    The shim calls the oracle's own encoding / MLP functions, so the golden file pins everything the REFERENCE owns
    on this path (wiring of the heads, raw geo_feat, sigmoid, trunc_exp, sampling, weights, the 1e-4 mask, depth
    normalisation, white background, the loss) and not the tiny-cuda-nn arithmetic, which stays "parity unpinned".
  * the CUDA extensions `_raymarching`, `_gridencoder`, `_shencoder`, `_ffmlp`: never called on this path except
    `near_far_from_aabb`, which is served by the C oracle (oracle/ngp_oracle.c, pinned bit-exactly on the reference's
    own kernel in tests/test_oracle_pinned.py);
  * optional third-party imports (trimesh, mcubes, matplotlib, tensorboardX, torch_ema, torch_scatter, h5py, turtle):
    inert stubs.
Config C1 of BASELINE.json: `encoding='freq'`, 2x64 MLPs, rgb / depth / semantic heads (+ the 64-d feature head).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("AUTOLABEL_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "ref_run_path.npz")
sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__file__ = f"<stub {name}>"
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_tcnn_shim():
    from oracle import field_oracle as fo

    class Encoding(torch.nn.Module):
        def __init__(self, n_input_dims, encoding_config=None, **kw):
            super().__init__()
            self.cfg = dict(encoding_config)
            ot = self.cfg["otype"]
            if ot == "Frequency":
                self.n_output_dims = n_input_dims * 2 * int(self.cfg["n_frequencies"])
            elif ot == "SphericalHarmonics":
                self.n_output_dims = int(self.cfg["degree"]) ** 2
            else:
                raise NotImplementedError(f"tcnn shim: encoding {ot} (config C1 uses Frequency + SphericalHarmonics)")

        def forward(self, x):
            if self.cfg["otype"] == "Frequency":
                return fo.freq_encode(x.float(), int(self.cfg["n_frequencies"]))
            return fo.sh4(x.float())

    class Network(torch.nn.Module):
        def __init__(self, n_input_dims, n_output_dims, network_config, **kw):
            super().__init__()
            self.n_in, self.n_out = n_input_dims, n_output_dims
            self.hidden = int(network_config["n_neurons"])
            self.nh = int(network_config["n_hidden_layers"])
            self.in_pad, self.out_pad = fo.pad16(n_input_dims), fo.pad16(n_output_dims)
            n = self.hidden * self.in_pad + (self.nh - 1) * self.hidden * self.hidden + self.out_pad * self.hidden
            self.params = torch.nn.Parameter(torch.zeros(n))
            self.n_output_dims = n_output_dims

        def forward(self, x):
            return fo.mlp(x.float(), self.params, self.in_pad, self.hidden, self.out_pad, self.nh)[:, :self.n_out]

    _stub("tinycudann", Encoding=Encoding, Network=Network)


def import_reference(tcnn_module=None):
    """Import the reference model / renderer modules UNMODIFIED.  `tcnn_module`: what `import tinycudann` resolves to
    (default: the fp32 torch shim above; tests/test_reference_swap_cpu.py passes autolabel_b200.tcnn)."""
    sys.path.insert(0, REF)
    for name in ("trimesh", "mcubes", "tensorboardX", "torch_ema", "torch_scatter", "h5py", "turtle", "skimage",
                 "skvideo", "skvideo.io", "open3d", "lpips", "imageio", "dearpygui", "dearpygui.dearpygui", "packaging"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        def _cmap(x):
            return np.zeros(np.shape(x) + (4,))
        _cmap.colors = np.zeros((10, 3))
        cm = _stub("matplotlib.cm", tab10=_cmap, inferno=_cmap)
        _stub("matplotlib", cm=cm, pyplot=_stub("matplotlib.pyplot"), patches=_stub("matplotlib.patches"))
    if getattr(sys.modules["turtle"], "__file__", "").startswith("<stub"):
        sys.modules["turtle"].backward = sys.modules["turtle"].forward = None      # junk import of ffmlp/ffmlp.py:2
    if getattr(sys.modules["torch_scatter"], "__file__", "").startswith("<stub"):
        sys.modules["torch_scatter"].segment_csr = None                            # used by the fork's march backward only
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["torch_ema"].ExponentialMovingAverage = object
    # the CUDA extension back ends: import-time names only
    for name in ("_raymarching", "_gridencoder", "_shencoder", "_ffmlp", "_freqencoder"):
        _stub(name)
    if tcnn_module is None:
        install_tcnn_shim()
    else:
        sys.modules["tinycudann"] = tcnn_module
    sys.modules.pop("autolabel.models", None)
    from autolabel import models
    from torch_ngp import raymarching
    return models, raymarching


def main():
    from oracle import field_oracle as fo
    from oracle import ngp
    models, raymarching = import_reference()

    def near_far_cpu(rays_o, rays_d, aabb, min_near=0.2):
        n, f, _, _ = ngp.near_far_from_aabb(rays_o.detach().numpy(), rays_d.detach().numpy(), aabb.numpy(), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)
    raymarching.near_far_from_aabb = near_far_cpu
    import torch_ngp.nerf.renderer as ref_renderer
    ref_renderer.raymarching.near_far_from_aabb = near_far_cpu

    torch.manual_seed(0)
    bound, C, F = 2.0, 2, 64
    m = models.ALNetwork(encoding='freq', num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=2,
                         hidden_dim_color=64, hidden_dim_semantic=F, semantic_classes=C, bound=bound, cuda_ray=False)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for net in (m.sigma_net, m.color_net, m.semantic_features, m.semantic_out):
            fan = net.hidden
            net.params.copy_((torch.rand(net.params.shape, generator=g) * 2 - 1) * (6.0 / (2 * fan)) ** 0.5 * 1.5)
    N, T = 48, 64
    o = (torch.rand(N, 3, generator=g) - 0.5) * 1.2 * bound
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=1)
    norms = torch.rand(N, generator=g) * 0.3 + 1.0
    m.eval()
    with torch.no_grad():
        out = m.run(o, d, norms, num_steps=T, upsample_steps=0, bg_color=None, perturb=False)
    save = {"rays_o": o.numpy(), "rays_d": d.numpy(), "direction_norms": norms.numpy(), "bound": np.float32(bound),
            "num_steps": np.int32(T), "n_classes": np.int32(C), "feat_dim": np.int32(F),
            "w_sigma": m.sigma_net.params.detach().numpy(), "w_color": m.color_net.params.detach().numpy(),
            "w_semf": m.semantic_features.params.detach().numpy(), "w_semo": m.semantic_out.params.detach().numpy()}
    for k, v in out.items():
        save["out_" + k] = v.detach().numpy()

    # render() staging (renderer.py:685-744): B rows of N rays in chunks of max_ray_batch, with per-ray norms [B, N]
    # and with the [B*N, 1] norms `_get_test` returns (dataset.py:248-262), which the slice direction_norms[b:b+1,
    # head:tail] turns into ONE norm (that of flat pixel b) for the whole row (one chunk per row: a second chunk would
    # slice an empty norm and fail in the reference itself)
    B_, N_ = 2, 20
    so, sd = o[:B_ * N_].view(B_, N_, 3), d[:B_ * N_].view(B_, N_, 3)
    with torch.no_grad():
        st2d = m.render(so, sd, norms[:B_ * N_].view(B_, N_), staged=True, max_ray_batch=8, num_steps=T, perturb=False)
        stflat = m.render(so, sd, norms[:B_ * N_].view(-1, 1), staged=True, max_ray_batch=4096, num_steps=T, perturb=False)
    for k in ("image", "depth", "semantic_features"):
        save["staged2d_" + k] = st2d[k].numpy()
        save["stagedflat_" + k] = stflat[k].numpy()

    # the loss of SimpleTrainer.train_step on these outputs (trainer.py:72-92), with the reference's own code path
    from autolabel import trainer as ref_trainer
    opt = types.SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True)
    data = {"rays_o": o, "rays_d": d, "direction_norms": norms.view(-1, 1), "pixels": torch.rand(N, 3, generator=g),
            "depth": torch.rand(N, generator=g) * 3 * (torch.rand(N, generator=g) > 0.3),
            "semantic": torch.randint(-1, C, (N,), generator=g), "features": torch.rand(N, F, generator=g)}
    tr = object.__new__(ref_trainer.SimpleTrainer)
    tr.opt, tr.model, tr.device = opt, m, torch.device("cpu")
    tr.criterion = torch.nn.MSELoss(reduction='none')
    m.train()
    torch.manual_seed(5)
    _, _, loss = tr.train_step(data)
    save["loss_perturbed_seed5"] = np.float32(loss.item())
    for k in ("pixels", "depth", "semantic", "features"):
        save["gt_" + k] = data[k].numpy()
    # mark_untrained_grid (renderer.py:479-561), the reference's own five-loop code on CPU; morton3D from the C oracle
    def morton_cpu(coords):
        return torch.from_numpy(ngp.morton3D(coords.numpy().astype(np.int32)).astype(np.int32))
    ref_renderer.raymarching.morton3D = morton_cpu
    mg = models.ALNetwork(encoding='freq', num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=2,
                          hidden_dim_color=64, hidden_dim_semantic=F, semantic_classes=C, bound=bound, cuda_ray=True)
    rng = np.random.RandomState(7)
    n_pose = 6
    poses = np.zeros((n_pose, 4, 4), np.float32)
    for i in range(n_pose):                      # cameras on a circle looking roughly inwards (camera-to-world)
        a = 2 * np.pi * i / n_pose + rng.uniform(-0.2, 0.2)
        pos = np.array([1.4 * np.cos(a), 1.4 * np.sin(a), rng.uniform(-0.3, 0.3)], np.float32)
        z = -pos / np.linalg.norm(pos) + rng.uniform(-0.1, 0.1, 3)
        z /= np.linalg.norm(z)
        x = np.cross(z, [0, 0, 1.0]); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        poses[i, :3, 0], poses[i, :3, 1], poses[i, :3, 2], poses[i, :3, 3], poses[i, 3, 3] = x, y, z, pos, 1
    intrinsics = np.array([51.2, 51.2, 32.0, 24.0], np.float32)       # fx, fy, cx, cy of a 64x48 image
    mg.density_grid.fill_(0.5)
    mg.mark_untrained_grid(poses, intrinsics)
    save["mark_poses"], save["mark_intrinsics"] = poses, intrinsics
    save["mark_unseen_bits"] = np.packbits((mg.density_grid < 0).numpy().reshape(-1))
    save["mark_cascade"] = np.int32(mg.cascade)
    print("mark_untrained_grid: unseen fraction", float((mg.density_grid < 0).float().mean()))
    np.savez_compressed(OUT, **save)
    print("wrote", OUT, {k: v.shape for k, v in save.items() if hasattr(v, "shape") and v.ndim}, "loss", loss.item())


if __name__ == "__main__":
    main()
