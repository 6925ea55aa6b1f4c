"""Pins the numpy restatements used by the full-size GPU tests (tests/np_restatement.py) against the C oracle on CPU."""
import numpy as np
import pytest

from np_restatement import np_first_per_voxel, np_iqr, np_keys, np_map_insert


@pytest.mark.parametrize("vox,cap,spread", [(1.0, 10, 8.0), (0.5, 20, 3.0), (1.0, 1, 5.0)])
def test_numpy_restatements_match_the_oracle(port, rng, vox, cap, spread):
    xyz = rng.normal(size=(30000, 3)) * spread
    assert np.array_equal(np_keys(xyz, vox).astype(np.int32), port.vox_index(xyz, vox))
    assert np.array_equal(xyz[np_first_per_voxel(xyz, 0.5 * vox)], port.voxel_downsample(xyz, 0.5 * vox))
    for n in (1, 2, 3, 30, 31, 30000):
        assert np.array_equal(np_iqr(xyz[:n]), port.iqr(xyz[:n]))
    m = port.Map(vox, 1.0e4, cap)
    m.insert(xyz)
    k, c, p = m.dump()
    nk, nc, npts = np_map_insert(xyz, vox, cap)
    assert np.array_equal(k, nk) and np.array_equal(c, nc) and np.array_equal(p, npts)
