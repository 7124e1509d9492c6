"""torch.library binding: registration + shape propagation on CPU (fake tensors),
numerics on the GPU."""
import numpy as np
import pytest
import torch

from util import assert_same_bits


def test_ops_are_registered_and_trace_with_fake_tensors(sks):
    from torch._subclasses.fake_tensor import FakeTensorMode
    import sks_homography_b200.torch_ops  # noqa: F401
    assert hasattr(torch.ops.sks_b200, "solve") and hasattr(torch.ops.sks_b200, "aca_rect")
    with FakeTensorMode():
        s = torch.empty(100, 4, 2)
        t = torch.empty(100, 4, 2)
        H = torch.ops.sks_b200.solve(s, t, "aca", True)
        assert H.shape == (100, 9) and H.dtype == torch.float32
        R = torch.ops.sks_b200.aca_rect(torch.empty(7, 8, dtype=torch.float64), None, 1.0, 2.0, 50.0, 1.25, False)
        assert R.shape == (7, 9) and R.dtype == torch.float64
        Hb, cnt, hyp = torch.ops.sks_b200.ransac(torch.empty(5, 64, 4), 128, 1, 4.0)
        assert Hb.shape == (5, 9) and cnt.shape == (5,) and hyp.dtype == torch.int64
        G = torch.ops.sks_b200.rect_warp_grid(torch.empty(6, 8), None, 0.0, 0.0, 128.0, 1.0, 16, 12,
                                              0.0, 0.0, 1.0, 1.0)
        assert G.shape == (6, 12, 16, 2)
        T = torch.ops.sks_b200.tensor_aca_rect(torch.empty(9, 3, 4), torch.empty(9, 3, 4), torch.empty(1), torch.empty(1))
        assert T.shape == (9, 3, 3) and T.dtype == torch.float32


@pytest.mark.gpu
def test_ops_match_oracle_on_gpu(sks, oracle, cuda):
    import sks_homography_b200.torch_ops  # noqa: F401
    s, t = oracle.synth_quads(0, 10_000, 3, 1, np.float32)
    H = torch.ops.sks_b200.solve(torch.from_numpy(s).to(cuda), torch.from_numpy(t).to(cuda), "sks", True)
    assert_same_bits(H.cpu().numpy(), oracle.solve("sks", s, t), "torch op sks")
    R = torch.ops.sks_b200.aca_rect(torch.from_numpy(t).to(cuda), None, 15.0, 12.0, 128.0, 1.0, True)
    assert_same_bits(R.cpu().numpy(), oracle.aca_rect(t, 15.0, 12.0, 128.0, 1.0), "torch op rect")
    G = torch.ops.sks_b200.solve(torch.from_numpy(s).to(cuda), torch.from_numpy(t).to(cuda), "ge", True)
    assert_same_bits(G.cpu().numpy(), oracle.solve("ge", s, t), "torch op ge")
    W = torch.ops.sks_b200.rect_warp_grid(torch.from_numpy(t).to(cuda), None, 15.0, 12.0, 128.0, 1.0, 9, 7,
                                          15.0, 12.0, 16.0, 21.0)
    Hu = oracle.aca_rect(t, 15.0, 12.0, 128.0, 1.0, normalize=False)
    assert_same_bits(W.cpu().numpy(), oracle.warp_grid(Hu, 9, 7, 15.0, 12.0, 16.0, 21.0), "torch op warp grid")
    torch.library.opcheck(torch.ops.sks_b200.solve.default,
                          (torch.from_numpy(s[:64]).to(cuda), torch.from_numpy(t[:64]).to(cuda), "aca", False),
                          test_utils=("test_schema", "test_faketensor"))


@pytest.mark.gpu
def test_tensor_aca_rect_op_matches_the_reference_golden(sks, golden, cuda):
    """The reference's TensorACA_rect signature as a registered op: same bits as the golden H obtained by
    executing the reference's torch statements, scale / div passed as device tensors (no sync), opcheck clean."""
    import sks_homography_b200.torch_ops  # noqa: F401
    g = golden["ref_torch"]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    args = (dev(g["src_new"]), dev(g["tar_new"]), dev(g["scale"]), dev(g["div"]))
    H = torch.ops.sks_b200.tensor_aca_rect(*args)
    assert_same_bits(H.cpu().numpy(), g["H_rect"], "tensor_aca_rect op")
    torch.library.opcheck(torch.ops.sks_b200.tensor_aca_rect.default, args, test_utils=("test_schema", "test_faketensor"))
