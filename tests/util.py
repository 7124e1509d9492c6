import numpy as np


def bits(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def same_bits(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Element-wise: identical bit pattern, or NaN on both sides (x86 and the GPU
    produce different NaN payloads/signs for 0/0)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    return (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))


def assert_same_bits(a, b, what=""):
    ok = same_bits(a, b)
    if not ok.all():
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {(~ok).sum()} of {ok.size} elements differ; first at {i}: "
                             f"{np.asarray(a)[i]!r} vs {np.asarray(b)[i]!r}")


def reproject_error(H, src, tar):
    """max over the 4 points of |proj(H, src_k) - tar_k| in pixels; H [n,9]."""
    H = np.asarray(H, dtype=np.float64).reshape(-1, 3, 3)
    s = np.asarray(src, dtype=np.float64).reshape(-1, 4, 2)
    t = np.asarray(tar, dtype=np.float64).reshape(-1, 4, 2)
    p = np.concatenate([s, np.ones_like(s[..., :1])], axis=-1)      # [n,4,3]
    q = np.einsum("nij,nkj->nki", H, p)
    q = q[..., :2] / q[..., 2:3]
    return np.sqrt(((q - t) ** 2).sum(-1)).max(-1)
