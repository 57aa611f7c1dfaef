"""Host side of the learned-dynamics Jacobian: BatchNorm folding, zero padding and the closed form
dy/dx = (W3 diag(m2)) W2 diag(m1) W1 against autograd done the reference's way (nn_ilqr.py:265-274).  CPU only."""
import numpy as np
import pytest
import torch

from curiousrl_b200.algorithm.ilqr_solver.nn_ilqr import (FoldedMLP, LargeNetwork, SmallNetwork, SmallResidualNetwork,
                                                          gaussian_filter_matrix)


def _trained_like(cls, d, n, seed):
    torch.manual_seed(seed)
    net = cls(d, n).double()
    net.train()
    with torch.no_grad():
        for _ in range(3):
            net(torch.randn(256, d, dtype=torch.float64) * 2.0)      # BatchNorm running statistics away from (0, 1)
    return net.eval()


def _closed_form(f: FoldedMLP, x):
    w1, b1, w2, b2, w3, b3 = (t.double() for t in (f.w1, f.b1, f.w2, f.b2, f.w3, f.b3))
    z1 = x @ w1.T + b1
    m1 = (z1 > 0).double()
    z2 = (z1 * m1) @ w2.T + b2
    m2 = (z2 > 0).double()
    y = (z2 * m2) @ w3.T + b3
    jac = torch.einsum("ok,sk,kh,sh,hj->soj", w3, m2, w2, m1, w1)
    return y, jac


@pytest.mark.parametrize("cls,d,n", [(SmallNetwork, 8, 6), (LargeNetwork, 11, 8), (SmallNetwork, 6, 4)])
def test_folded_network_equals_the_module_and_its_autograd_jacobian(cls, d, n):
    net = _trained_like(cls, d, n, seed=d)
    f = FoldedMLP(net, "cpu")
    assert f.h1 % 32 == 0 and f.h2 % 32 == 0 and f.w2t.shape == (f.h1, f.h2)
    assert torch.equal(f.w2t, f.w2.t())
    x = torch.randn(20, d, dtype=torch.float64) * 2.0
    # the folded tensors are float32 (what the device takes): evaluate the closed form from them in float64
    y, jac = _closed_form(f, x)
    with torch.no_grad():
        y_ref = net(x)
    assert float((y - y_ref).abs().max()) <= 1e-5 * float(y_ref.abs().max())
    eye = torch.eye(n, dtype=torch.float64)
    for t in range(x.shape[0]):                                           # nn_ilqr.py:265-274
        r = x[t].repeat(n, 1).requires_grad_(True)
        net(r).backward(eye)
        assert float((jac[t] - r.grad).abs().max()) <= 1e-5 * float(r.grad.abs().max())


def test_other_network_shapes_are_refused():
    with pytest.raises(NotImplementedError):
        FoldedMLP(SmallResidualNetwork(8, 6).eval(), "cpu")


def test_filter_matrix_is_scipys_filter():
    from scipy.ndimage import gaussian_filter1d
    rng = np.random.default_rng(0)
    for T, sigma in ((100, 5), (30, 5), (7, 2.5)):
        F = rng.normal(size=(T, 6, 8))
        G = gaussian_filter_matrix(T, sigma)
        assert np.abs(np.einsum("tu,uij->tij", G, F) - gaussian_filter1d(F, sigma=sigma, axis=0)).max() <= 1e-14
        assert np.allclose(G.sum(axis=1), 1.0)
