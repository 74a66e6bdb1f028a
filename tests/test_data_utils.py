"""CPU tests of the host-side readers (scanobjectnn_b200/data_utils.py) against the reference's file layouts."""
import numpy as np
import pytest

from scanobjectnn_b200 import data_utils as du


def _write_bin(path, pts):
    np.concatenate([[np.float32(len(pts))], pts.astype(np.float32).ravel()]).astype(np.float32).tofile(path)


def test_load_pc_file_layout_and_background_filter(tmp_path):
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((50, 11)).astype(np.float32)
    pts[:, -1] = np.array([0, 1, 2, 7, 7, 7, 9, 9, 7, 0] * 5)        # semantic label: 0/1/2 = background classes
    f = tmp_path / "obj.bin"
    _write_bin(f, pts)
    xyz = du.load_pc_file(str(f))
    assert xyz.shape == (50, 3) and np.array_equal(xyz, pts[:, :3])
    fg = du.load_pc_file(str(f), with_bg=False)
    assert np.array_equal(fg, pts[pts[:, -1] == 7][:, :3])           # most frequent non-background label
    s = tmp_path / "suncg.bin"
    _write_bin(s, pts[:, :3])
    assert np.array_equal(du.load_pc_file(str(s), suncg=True), pts[:, :3])
    bad = tmp_path / "bad.bin"
    np.array([5.0, 1, 2, 3], np.float32).tofile(bad)
    with pytest.raises(ValueError):
        du.load_pc_file(str(bad))


def test_masks_and_epoch_subset():
    m = np.array([[-1, 3, 3, -1], [5, 5, -1, 5]])
    assert np.array_equal(du.convert_to_binary_mask(m), [[0, 1, 1, 0], [1, 1, 0, 1]])
    pcs = np.arange(4 * 10 * 3, dtype=np.float32).reshape(4, 10, 3)
    labels = np.arange(4)
    sampled, lab, idx_pts, order = du.get_current_data_h5(pcs, labels, 6, np.random.default_rng(1), return_indices=True)
    assert sampled.shape == (4, 6, 3) and np.array_equal(lab, labels[order])
    assert np.array_equal(sampled, pcs[order][:, idx_pts, :])        # the same subset for every cloud


def test_h5_reader_reports_missing_dependency():
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            du.load_h5("whatever.h5")
