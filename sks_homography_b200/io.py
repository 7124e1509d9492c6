"""Correspondence-file reader: the text format of the reference's fixture
(`CPU_Runtime Test/orig_pts_wall.txt`, read by `read_points`, CPU/utils.cpp:6-21 and
GPU.cu:31-46): first line = number of matches, then one `x1 y1 x2 y2` per line
(CRLF tolerated).  Returns the interleaved (x, y, X, Y) pool that
sks_cuda_gather_samples_* and the RANSAC kernel consume."""
from __future__ import annotations

import numpy as np


def read_points(filename: str, dtype=np.float32) -> np.ndarray:
    with open(filename, "r", newline=None) as f:
        first = f.readline()
        try:
            n = int(first.split()[0])
        except (IndexError, ValueError) as e:
            raise ValueError(f"{filename}: first line must hold the match count") from e
        pool = np.empty((n, 4), dtype=dtype)
        for i in range(n):
            parts = f.readline().split()
            if len(parts) < 4:
                raise ValueError(f"{filename}: line {i + 2} does not hold 'x1 y1 x2 y2'")
            pool[i] = [float(p) for p in parts[:4]]
    return pool
