"""Device-resident scene dataset — the batch contract of the reference's ``autolabel/dataset.py``
(``BaseDataset`` :152-300, ``SceneDataset`` :303-449) with the arrays in HBM and the per-batch work in
one kernel (``al_dataset_sample``, csrc/dataset.cu) instead of numpy fancy indexing + numba + seven
host-to-device copies per step.

Same attribute names as the reference dataset (``images, depths, semantics, features, poses, rotations,
origins, intrinsics, w, h, n_examples, batch_size, sample_chunk_size, feature_width/height/dim,
min_bounds, max_bounds, n_classes, index_sampler, semantic_image_sample_ratio``) and the same batch
dicts from ``_next_train`` (:182-242) and ``_get_test`` (:244-266); iterate it exactly like the
reference (``for batch in dataset`` / ``LenDataset``).  All tensors are CUDA tensors on ``device``.

Sampling policy = the reference's: a batch is ``batch_size // 512`` chunks; each chunk draws one image
— with probability ``semantic_image_sample_ratio`` (when any pixel is labelled) a class is drawn
uniformly, then an image proportionally to its number of pixels of that class, then 512 of those pixels
with replacement; otherwise a uniform image and 512 uniform pixels out of ``pixel_indices`` — and every
ray gets a uniform sub-pixel jitter.  The draws come from a ``torch.Generator`` on the device (the
reference uses Python's / numpy's / numba's global CPU generators, so streams differ; the batch as a
function of the draws is identical and is what tests/test_dataset_gpu.py checks against the reference's
own output, tests/golden/ref_dataset.npz).
"""
import numpy as np
import torch

from ._lib import call, ptr, require_cuda, stream_ptr


class DeviceIndexSampler:
    """``IndexSampler`` (dataset.py:80-149) on the device: for every class (label >= 1; 0 = unlabeled) the
    flat indices of its pixels grouped by image (CSR), so that `sample` is two tiny multinomial draws and
    one gather instead of a Python dict walk."""

    def __init__(self, device):
        self.device = device
        self.classes = torch.empty(0, dtype=torch.long, device=device)
        self.has_semantics = False
        self._per_class = {}
        self._flat = None

    @torch.no_grad()
    def update(self, semantic_maps):
        """semantic_maps: uint8 [n, HW] device tensor."""
        assert semantic_maps.dim() == 2
        n, hw = semantic_maps.shape
        classes = torch.unique(semantic_maps)
        self.classes = classes[classes != 0].long()
        self._per_class = {}
        self.has_semantics = False
        self._flat = None
        for cid in self.classes.tolist():
            where = semantic_maps == cid
            counts = where.sum(dim=1)                       # pixels of this class per image
            total = int(counts.sum().item())
            if total == 0:
                continue
            self.has_semantics = True
            flat = torch.nonzero(where.reshape(-1), as_tuple=False).reshape(-1)   # sorted: image-major
            starts = torch.cumsum(counts, 0) - counts
            self._per_class[cid] = {'pixels': (flat % hw).int(), 'starts': starts, 'counts': counts,
                                    'weights': counts.double() / total}

    def _build_flat(self):
        """All classes in ONE set of device tensors, for draws without host synchronisation: weights [n_cls, n_images],
        counts / starts [n_cls, n_images] (starts already offset into the concatenated pixel list)."""
        cids = [c for c in self.classes.tolist() if c in self._per_class]
        if not cids:
            return None
        w = torch.stack([self._per_class[c]['weights'] for c in cids])
        counts = torch.stack([self._per_class[c]['counts'] for c in cids])
        base, starts, pix = 0, [], []
        for c in cids:
            e = self._per_class[c]
            starts.append(e['starts'] + base)
            pix.append(e['pixels'])
            base += int(e['pixels'].numel())
        self._flat = {'weights': w, 'counts': counts, 'starts': torch.stack(starts), 'pixels': torch.cat(pix)}
        return self._flat

    def sample_chunks(self, n_chunks, count, gen):
        """`n_chunks` independent (class, image, `count` pixels) draws of `sample_class` + `sample` (dataset.py:127-138,
        204-213), entirely on the device: -> image index int32 [n_chunks], pixel indices int32 [n_chunks, count]."""
        f = self._flat if self._flat is not None else self._build_flat()
        n_cls = f['weights'].shape[0]
        cls = torch.randint(0, n_cls, (n_chunks,), generator=gen, device=self.device)
        img = torch.multinomial(f['weights'][cls], 1, generator=gen).view(-1)                  # image ~ pixels of the class
        cnt = f['counts'][cls, img]
        k = (torch.rand(n_chunks, count, generator=gen, device=self.device) * cnt[:, None]).long()
        k = torch.minimum(k, (cnt - 1)[:, None])
        return img.int(), f['pixels'][f['starts'][cls, img][:, None] + k]

    def sample_class(self, gen):
        i = torch.randint(0, self.classes.numel(), (1,), generator=gen, device=self.device)
        return int(self.classes[i].item())

    def sample(self, class_id, count, gen):
        """-> (image index (int), pixel indices int32 [count]) — dataset.py:127-138."""
        e = self._per_class[class_id]
        img = int(torch.multinomial(e['weights'], 1, generator=gen).item())
        k = torch.randint(0, int(e['counts'][img].item()), (count,), generator=gen, device=self.device)
        return img, e['pixels'][e['starts'][img] + k]

    def semantic_indices(self):
        idx = set()
        for e in self._per_class.values():
            idx.update(torch.nonzero(e['counts'] > 0).reshape(-1).tolist())
        return sorted(idx)


class DeviceSceneDataset(torch.utils.data.IterableDataset):
    semantic_image_sample_ratio = 0.5

    def __init__(self, images, depths, semantics, poses, intrinsics, size, features=None, feature_size=None,
                 batch_size=4096, split="train", min_bounds=None, max_bounds=None, n_classes=None, pixel_indices=None,
                 device="cuda", seed=0):
        """images fp32 [n,h*w,3] (or [n,h,w,3]) in [0,1]; depths uint16 millimetres [n,h*w]; semantics uint8 [n,h*w]
        (0 = unlabeled); poses fp32 [n,4,4] camera-to-world in the ngp convention (`_convert_pose`, dataset.py:268-274);
        intrinsics (fx, fy, cx, cy) at `size` = (w, h); features fp16 [n, fh*fw, F] with feature_size = (fw, fh)."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceSceneDataset keeps the scene in GPU memory; there is no CPU fallback")
        self.w, self.h = int(size[0]), int(size[1])
        self.resolution = self.w * self.h
        self.split = split
        self.batch_size = batch_size
        self.sample_chunk_size = 512
        dev = self.device

        def to_dev(a, dtype):
            t = torch.as_tensor(np.ascontiguousarray(a) if isinstance(a, np.ndarray) else a)
            return t.to(device=dev, dtype=dtype).contiguous()

        self.n_examples = int(images.shape[0])
        self.images = to_dev(images, torch.float32).reshape(self.n_examples, self.resolution, 3)
        # uint16 millimetres, kept as the same 16 bits in an int16 tensor (torch's uint16 support is partial)
        d16 = np.ascontiguousarray(depths.cpu().numpy() if torch.is_tensor(depths) else depths).astype(np.uint16)
        self.depths = torch.from_numpy(d16.view(np.int16)).to(dev).reshape(self.n_examples, self.resolution)
        self.semantics = to_dev(semantics, torch.uint8).reshape(self.n_examples, self.resolution)
        self.poses = to_dev(poses, torch.float32)
        self.rotations = self.poses[:, :3, :3].contiguous()
        self.origins = self.poses[:, :3, 3].contiguous()
        self.intrinsics = np.array([float(v) for v in intrinsics], dtype=np.float64)
        self.features = None
        self.feature_width = self.feature_height = self.feature_dim = 0
        if features is not None:
            self.feature_width, self.feature_height = int(feature_size[0]), int(feature_size[1])
            self.features = to_dev(features, torch.float16).reshape(self.n_examples, self.feature_width * self.feature_height, -1)
            self.feature_dim = int(self.features.shape[-1])
        self.min_bounds = None if min_bounds is None else np.asarray(min_bounds, dtype=np.float32)
        self.max_bounds = None if max_bounds is None else np.asarray(max_bounds, dtype=np.float32)
        self.n_classes = n_classes
        self.indices = np.arange(self.n_examples)
        self.pixel_indices = (torch.arange(self.resolution, device=dev, dtype=torch.int32) if pixel_indices is None
                              else to_dev(pixel_indices, torch.int32))
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.index_sampler = DeviceIndexSampler(dev)
        self.index_sampler.update(self.semantics)
        self.error_map = None

    @classmethod
    def from_host(cls, ds, device="cuda", seed=0, **kw):
        """Move a loaded reference dataset (`autolabel.dataset.SceneDataset`, non-lazy) into HBM."""
        feats = getattr(ds, 'features', None)
        return cls(np.asarray(ds.images), np.asarray(ds.depths), np.asarray(ds.semantics), np.asarray(ds.poses),
                   tuple(ds.intrinsics), (ds.w, ds.h), features=feats,
                   feature_size=(ds.feature_width, ds.feature_height) if feats is not None else None,
                   batch_size=ds.batch_size, split=ds.split, min_bounds=getattr(ds, 'min_bounds', None),
                   max_bounds=getattr(ds, 'max_bounds', None), n_classes=getattr(ds, 'n_classes', None),
                   pixel_indices=getattr(ds, 'pixel_indices', None), device=device, seed=seed, **kw)

    def load_features(self, scene_path, name):
        """`_load_features` of the reference (dataset.py:438-449): attach `features/<name>` of `<scene>/features.hdf`
        (h5py when present, else the h5py-free reader hdf5_lite) and keep its pca / min / range attributes."""
        from .hdf5_lite import load_features
        arr, w, h, c, attrs = load_features(scene_path, name)
        self.features = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device, torch.float16).contiguous()
        self.feature_width, self.feature_height, self.feature_dim = int(w), int(h), int(c)
        self.feature_attrs = attrs
        return self

    # ------------------------------------------------------------ iteration (dataset.py:174-180)
    def __iter__(self):
        if self.split == "train":
            while True:
                yield self._next_train()
        else:
            for i in range(self.n_examples):
                yield self._get_test(i)

    def __len__(self):
        return self.n_examples

    # ------------------------------------------------------------ kernels
    def _launch(self, n, image_index, image0, ray_indices, jitter, want_targets, want_features):
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        feats = self.features if want_features else None
        if want_targets:
            # one flat buffer per batch (trainer.PackedBatch): SimpleTrainer moves it into its step graph with ONE copy
            from .trainer import PackedBatch
            out = PackedBatch(n, self.feature_dim if feats is not None else 0, dev)
        else:
            out = {'rays_o': torch.empty(n, 3, **f32), 'rays_d': torch.empty(n, 3, **f32),
                   'direction_norms': torch.empty(n, 1, **f32)}
        fx, fy, cx, cy = (float(v) for v in self.intrinsics)
        call("al_dataset_sample", ptr(self.images), ptr(self.depths), ptr(self.semantics), ptr(feats), ptr(self.rotations),
             ptr(self.origins), self.w, self.h, max(self.feature_width, 1), max(self.feature_height, 1), self.feature_dim,
             fx, fy, cx, cy, ptr(image_index), int(image0), ptr(ray_indices), ptr(jitter), n, self.sample_chunk_size,
             ptr(out['rays_o']), ptr(out['rays_d']), ptr(out['direction_norms']), ptr(out.get('pixels')), ptr(out.get('depth')),
             ptr(out.get('semantic')), ptr(out.get('features')), stream_ptr(dev))
        return out

    @torch.no_grad()
    def sample_batch(self, image_index, ray_indices, jitter=None):
        """The batch of `_next_train` for GIVEN draws (image per 512-ray chunk, flat pixel index and jitter per ray)."""
        require_cuda(image_index, ray_indices, jitter)
        image_index = image_index.to(torch.int32).contiguous()
        ray_indices = ray_indices.to(torch.int32).contiguous()
        n = ray_indices.numel()
        if image_index.numel() * self.sample_chunk_size < n:
            raise ValueError("one image index per chunk of 512 rays is required")
        if jitter is not None:
            jitter = jitter.float().contiguous()
        return self._launch(n, image_index, 0, ray_indices, jitter, True, True)

    @torch.no_grad()
    def draw(self):
        """The random draws of one `_next_train` call (dataset.py:204-213), on the device and WITHOUT host
        synchronisation: per chunk a coin decides between a labelled-pixel draw (class uniform, image proportional to
        its pixels of that class, 512 of those pixels with replacement) and a uniform one; both are drawn for every
        chunk and selected by the coin."""
        chunks = self.batch_size // self.sample_chunk_size
        c, dev, gen = self.sample_chunk_size, self.device, self.gen
        image_index = torch.randint(0, self.n_examples, (chunks,), generator=gen, device=dev, dtype=torch.int32)
        pick = torch.randint(0, self.pixel_indices.numel(), (chunks, c), generator=gen, device=dev)
        ray_indices = self.pixel_indices[pick]
        if self.index_sampler.has_semantics:
            coin = torch.rand(chunks, generator=gen, device=dev) < self.semantic_image_sample_ratio
            img_l, pix_l = self.index_sampler.sample_chunks(chunks, c, gen)
            image_index = torch.where(coin, img_l, image_index)
            ray_indices = torch.where(coin[:, None], pix_l, ray_indices)
        jitter = torch.rand(chunks * c, 2, generator=gen, device=dev)
        return image_index, ray_indices.reshape(-1), jitter

    def _next_train(self):
        return self.sample_batch(*self.draw())

    @torch.no_grad()
    def _get_test(self, image_index):
        """dataset.py:244-266: full-frame rays through pixel centres + views of the stored targets."""
        h, w = self.h, self.w
        r = self._launch(self.resolution, None, int(image_index), None, None, False, False)
        out = {
            'pixels': self.images[image_index].view(h, w, 3),
            'rays_o': r['rays_o'].view(h, w, 3), 'rays_d': r['rays_d'].view(h, w, 3),
            # a true fp64 division (torch multiplies by the reciprocal when the divisor is a host scalar)
            'depth': torch.div((self.depths[image_index].to(torch.int32) & 0xFFFF).double(),
                               torch.full((), 1000.0, dtype=torch.float64, device=self.device)).view(h, w),
            'semantic': (self.semantics[image_index].long() - 1).view(h, w),
            'H': h, 'W': w, 'direction_norms': r['direction_norms'],
        }
        if self.features is not None:
            out['features'] = self.features[image_index]
        return out

    # ------------------------------------------------------------ annotation updates (dataset.py:420-436)
    @torch.no_grad()
    def set_semantic_map(self, image_index, semantic):
        self.semantics[image_index] = torch.as_tensor(semantic).to(self.device, torch.uint8).reshape(self.resolution)
        self.update_sampler()

    def update_sampler(self):
        self.index_sampler.update(self.semantics)
