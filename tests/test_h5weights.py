"""Keras .h5 checkpoint reader/writer (src/UtilsFiles.py:153-164, src/NeRF.py:343-351) -- CPU only."""
import importlib
import os

import numpy as np
import pytest

from oracle import nerf_oracle as O
from oracle.h5lite import H5File, load_keras_nerf_weights

h5w = importlib.import_module("nerf-and-dietnerf_b200.h5weights")

REF_H5 = "/root/reference/Results/50px_alexander_71pics_sphere_nerf_save_dir_4/saved_weights/NeRF_model_epoch_095.h5"


def _random_flat(cfg, seed):
    return O.glorot_params(cfg.shapes, seed, bias_scale=0.3).numpy()


@pytest.mark.parametrize("n_angles", [2, 0])
def test_write_read_round_trip_is_bit_exact(tmp_path, n_angles):
    cfg = O.NetCfg(5, 4, n_angles)
    pc, pf = _random_flat(cfg, 1), _random_flat(cfg, 2)
    path = str(tmp_path / "NeRF_model_epoch_007.h5")
    h5w.save_flat_params(path, pc, pf, cfg.shapes)
    rc, rf = h5w.load_flat_params(path)
    assert np.array_equal(rc, pc) and np.array_equal(rf, pf)
    # the independent reader of the oracle (developed against the reference's own file) parses it the same way
    oc, of, shapes = load_keras_nerf_weights(path)
    assert np.array_equal(oc, pc) and np.array_equal(of, pf)
    assert [tuple(shapes[i]) for i in sorted(shapes)] == list(cfg.shapes) * 2
    # coarse-only model
    h5w.save_flat_params(path, pc, None, cfg.shapes)
    rc, rf = h5w.load_flat_params(path)
    assert np.array_equal(rc, pc) and rf is None


def test_written_file_has_the_layout_keras_writes(tmp_path):
    cfg = O.NetCfg()
    path = str(tmp_path / "w.h5")
    h5w.save_flat_params(path, _random_flat(cfg, 1), _random_flat(cfg, 2), cfg.shapes)
    names = sorted(h5w.H5Reader(path).datasets())
    assert len(names) == 44
    assert names[0] == "/model/dense/bias:0" and "/model/dense_10/kernel:0" in names
    assert "/model_1/dense_11/kernel:0" in names and "/model_1/dense_21/bias:0" in names
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and int.from_bytes(raw[40:48], "little") == len(raw)      # EOF address
    for attr in (b"layer_names", b"weight_names", b"backend", b"keras_version", b"2.7.0", b"tensorflow",
                 b"dense_10/kernel:0"):
        assert attr in raw


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="needs the reference checkout (build container only)")
def test_reads_the_checkpoint_the_reference_trained(tmp_path):
    pc, pf = h5w.load_flat_params(REF_H5)
    oc, of, _ = load_keras_nerf_weights(REF_H5)
    assert pc.size == 514332 and np.array_equal(pc, oc) and np.array_equal(pf, of)
    # same dataset names, shapes and values after a rewrite
    path = str(tmp_path / "rewrite.h5")
    h5w.save_flat_params(path, pc, pf, O.NetCfg().shapes)
    ref, new = H5File(REF_H5).datasets(), h5w.H5Reader(path).datasets()
    assert sorted(ref) == sorted(new)
    assert all(ref[k].shape == new[k].shape and np.array_equal(ref[k], new[k]) for k in ref)
    # and the committed golden fixture was made from the same numbers
    from conftest import ROOT
    pin = np.load(os.path.join(ROOT, "tests", "golden", "alexander50_pin.npz"))
    assert np.array_equal(pin["params_coarse"], pc) and np.array_equal(pin["params_fine"], pf)


def test_written_checkpoint_opens_with_libhdf5(tmp_path):
    """Opt-in cross-check of the hand-written HDF5 writer with the real library (h5py / libhdf5 are not in this image, so
    the test skips here; anywhere they exist it opens a written file and compares names, shapes and values)."""
    h5py = pytest.importorskip("h5py")
    import importlib
    h5w = importlib.import_module("nerf-and-dietnerf_b200.h5weights")
    net = importlib.import_module("nerf-and-dietnerf_b200.network")
    ncfg = importlib.import_module("nerf-and-dietnerf_b200._lib").NetCfg(5, 4, 2, 256, 128, 0.05)
    shapes = net.layer_shapes(ncfg)
    n = sum(i * o + o for i, o in shapes)
    rng = np.random.default_rng(0)
    pc, pf = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    path = tmp_path / "NeRF_model_epoch_001.h5"
    h5w.save_flat_params(str(path), pc, pf, shapes)
    found = {}
    with h5py.File(path, "r") as f:
        f.visititems(lambda name, obj: found.__setitem__(name, np.asarray(obj)) if isinstance(obj, h5py.Dataset) else None)
        assert "layer_names" in f.attrs
    kernels = sorted(k for k in found if k.endswith("kernel:0"))
    assert len(kernels) == 2 * len(shapes)
    back_c, back_f = h5w.load_flat_params(str(path))
    assert np.array_equal(back_c, pc) and np.array_equal(back_f, pf)
    total = sum(v.size for v in found.values())
    assert total == 2 * n
