"""torch.library operator binding: the solvers as registered
custom ops, forward only, with fake (meta) kernels so they trace under
torch.compile / FakeTensorMode.  The ops only dispatch to libsks_cuda through
sks_homography_b200.api; there is still no CPU compute path.

    torch.ops.sks_b200.solve(src, tar, "aca" | "sks", normalize)      -> H [n, 9]
    torch.ops.sks_b200.aca_rect(tar, M, mx, my, width, ratio, normalize) -> H [n, 9]
    torch.ops.sks_b200.ransac(corr, n_hyp, seed, thr2)                -> (H [P,9], count [P], hyp [P])
    torch.ops.sks_b200.rect_warp_grid(tar, M, mx, my, width, ratio, gw, gh, x0, y0, dx, dy)
                                                                      -> grid [n, gh, gw, 2]
    torch.ops.sks_b200.tensor_aca_rect(src34, tar34, scale, div)      -> H [bs, 3, 3]
        the reference's TensorACA_rect(bs, src, tar, scale, div) (PY.py:286-309) on its own tensors:
        [bs,3,4] homogeneous corners, scale / div one-element tensors on the same device -- read by
        the kernel, so the op neither synchronises nor breaks a torch.compile graph
"solve" also takes "ge" (the competitor RHO-GE).
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import api


@torch.library.custom_op("sks_b200::solve", mutates_args=())
def solve(src: Tensor, tar: Tensor, solver: str, normalize: bool) -> Tensor:
    return api.solve(solver, src.reshape(-1, 8), tar.reshape(-1, 8), normalize=normalize)


@solve.register_fake
def _(src, tar, solver, normalize):
    return src.new_empty((src.numel() // 8, 9))


@torch.library.custom_op("sks_b200::aca_rect", mutates_args=())
def aca_rect(tar: Tensor, M: Tensor | None, mx: float, my: float, width: float, ratio: float,
             normalize: bool) -> Tensor:
    return api.aca_rect(tar.reshape(-1, 8), width, ratio, mx, my, M=M, normalize=normalize)


@aca_rect.register_fake
def _(tar, M, mx, my, width, ratio, normalize):
    return tar.new_empty((tar.numel() // 8, 9))


@torch.library.custom_op("sks_b200::ransac", mutates_args=())
def ransac(corr: Tensor, n_hyp: int, seed: int, thr2: float) -> tuple[Tensor, Tensor, Tensor]:
    keys = api.ransac_keys(corr, n_hyp, seed, thr2)
    H, cnt, _ = api.ransac_finalize(corr, n_hyp, seed, thr2, keys)
    _, hyp = api.decode_keys(keys)
    return H, cnt, hyp


@ransac.register_fake
def _(corr, n_hyp, seed, thr2):
    P = corr.shape[0]
    return (corr.new_empty((P, 9)), corr.new_empty((P,), dtype=torch.int32),
            corr.new_empty((P,), dtype=torch.int64))


@torch.library.custom_op("sks_b200::rect_warp_grid", mutates_args=())
def rect_warp_grid(tar: Tensor, M: Tensor | None, mx: float, my: float, width: float, ratio: float,
                   gw: int, gh: int, x0: float, y0: float, dx: float, dy: float) -> Tensor:
    return api.aca_rect_warp_grid(tar.reshape(-1, 8), width, ratio, gw, gh, M_x=mx, M_y=my, M=M,
                                  x0=x0, y0=y0, dx=dx, dy=dy)


@rect_warp_grid.register_fake
def _(tar, M, mx, my, width, ratio, gw, gh, x0, y0, dx, dy):
    return tar.new_empty((tar.numel() // 8, gh, gw, 2))


@torch.library.custom_op("sks_b200::tensor_aca_rect", mutates_args=())
def tensor_aca_rect(src34: Tensor, tar34: Tensor, scale: Tensor, div: Tensor) -> Tensor:
    return api.TensorACA_rect(tar34.shape[0], src34, tar34, scale, div)


@tensor_aca_rect.register_fake
def _(src34, tar34, scale, div):
    return tar34.new_empty((tar34.shape[0], 3, 3))
