"""The reference's real-data fixture (2540 wall matches with outliers), read in
place from the reference checkout (it is not copied into this repository, so
these tests only run where /root/reference exists)."""
import os

import numpy as np
import pytest

FIXTURE = "/root/reference/C++ Codes/Runtime Test/CPU_Runtime Test/orig_pts_wall.txt"
pytestmark = pytest.mark.skipif(not os.path.exists(FIXTURE), reason="reference fixture not present")


def test_reader_matches_reference_format():
    from sks_homography_b200.io import read_points
    pool = read_points(FIXTURE)
    assert pool.shape == (2540, 4) and pool.dtype == np.float32
    assert np.allclose(pool[0], [356.39, 218.91, 439.44, 251.08])
    assert pool.min() > 0 and pool[:, [0, 2]].max() < 800 and pool[:, [1, 3]].max() < 640


def test_ransac_oracle_on_real_matches(oracle):
    """~63 % of the matches fit one homography at 3 px."""
    from sks_homography_b200.io import read_points
    corr = read_points(FIXTURE)[None]
    keys = oracle.ransac(corr, 2000, seed=11, thr2=9.0)
    frac = float(keys[0] >> np.uint64(32)) / corr.shape[1]
    assert 0.55 < frac < 0.70
