"""f4: the h5py-free `features.hdf` reader (autolabel_b200/hdf5_lite.py) on a file of the layout
scripts/compute_feature_maps.py:82-118 produces (group `features`, chunked float16 [N, H, W, C] with the lzf filter,
attributes pca / min / range), written by the independent test writer tests/hdf5_writer.py.  h5py is absent from this
image: the reader is pinned on the format specification and this writer only (stated in its header)."""
import os
import pickle

import numpy as np
import pytest

from autolabel_b200 import hdf5_lite
from tests.hdf5_writer import lzf_compress, write_features_hdf


def test_lzf_round_trip():
    rng = np.random.RandomState(0)
    for data in (b"", b"a", b"abcabcabcabcabcabc" * 40, bytes(rng.randint(0, 4, 5000).astype(np.uint8)),
                 bytes(rng.randint(0, 256, 3000).astype(np.uint8)), b"\0" * 10000):
        z = lzf_compress(data)
        assert hdf5_lite.lzf_decompress(z, len(data)) == data
    assert len(lzf_compress(b"\0" * 10000)) < 200             # back references really are emitted


@pytest.mark.parametrize("compress", [True, False])
def test_features_hdf_round_trip(tmp_path, compress):
    rng = np.random.RandomState(1)
    N, H, W, C = 5, 9, 12, 64                                # DINO-shaped (90 x 120 x 64 in the real file), edge chunks
    feats = np.maximum(rng.randn(N, H, W, C), 0).astype(np.float16)     # ReLU codes: many zeros, compressible
    pca = pickle.dumps({"components": rng.randn(3, C)})
    attrs = {"pca": pca, "min": rng.randn(3), "range": rng.rand(3).astype(np.float64)}
    path = os.path.join(tmp_path, "features.hdf")
    write_features_hdf(path, "dino", feats, chunks=(1, 4, 5, 64), attrs=attrs, compress=compress)
    with hdf5_lite.File(path) as hdf:
        assert hdf.keys() == ["features"] and hdf["features"].keys() == ["dino"]
        ds = hdf["features/dino"]
        assert ds.shape == (N, H, W, C) and ds.dtype == np.float16
        assert np.array_equal(ds[:], feats)
        assert np.array_equal(ds[2, :, 3], feats[2, :, 3])
        assert np.allclose(ds.attrs["min"], attrs["min"]) and np.allclose(ds.attrs["range"], attrs["range"])
        assert pickle.loads(ds.attrs["pca"].tobytes())["components"].shape == (3, C)      # scripts/compute_feature_maps.py:122
        with pytest.raises(KeyError):
            hdf["features/lseg"]
    arr, w, h, c, a = hdf5_lite.load_features(str(tmp_path), "dino")                     # autolabel/dataset.py:438-449
    assert arr.shape == (N, H * W, C) and (w, h, c) == (W, H, C) and np.array_equal(arr.reshape(N, H, W, C), feats)


def test_rejects_non_hdf5(tmp_path):
    p = os.path.join(tmp_path, "x.hdf")
    open(p, "wb").write(b"not hdf5 at all")
    with pytest.raises(ValueError):
        hdf5_lite.File(p)
