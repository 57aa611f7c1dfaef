"""CPU restatement of ONE NNiLQR iteration, given the network.  TEST INFRASTRUCTURE - never imported by the product.

Reference: CuriousRL/algorithm/ilqr_solver/nn_ilqr.py
  :265-274  eval_grad_dynamic_model - per time step, n copies of the row through the network, backward with the identity
  :376-384  one iteration of `solve`: C, c at the trajectory; F = network Jacobians, Gaussian-filtered along T
            (scipy.ndimage.gaussian_filter1d, axis 0); backward pass on them; forward pass (line search) on the REAL system.
The reference's class only runs on a CUDA device (every tensor is `.cuda()`) and imports `torch_optimizer`, which this image
does not have, so it cannot produce golden vectors in the build container: **parity unpinned against reference runs** for
this row.  The restatement follows the cited lines on the CPU with the reference's own arithmetic types (float32 network,
float64 iLQR) and the already pinned `oracle/ilqr_numpy.py` for everything that is not the network.
"""
import numpy as np
import torch
from scipy.ndimage import gaussian_filter1d

from . import ilqr_numpy as orc


def network_jacobians(network: torch.nn.Module, X: np.ndarray, n: int) -> np.ndarray:
    """nn_ilqr.py:265-274 on the CPU: X [T, n+m] float64 -> F [T, n, n+m] float64 (computed in float32)."""
    network = network.cpu().float().eval()
    T, d = X.shape
    F = np.zeros((T, n, d))
    eye = torch.eye(n)
    rows = torch.from_numpy(X).float()
    for tau in range(T):
        x = rows[tau].repeat(n, 1).requires_grad_(True)
        network(x).backward(eye)
        F[tau] = x.grad.double().numpy()
    return F


def iteration(prob: orc.Problem, network, X, J_last, target=None, sigma=5, max_line_search=10, gamma=0.5,
              criterion=1e-6, stopping_method="vanilla"):
    """One pass of the loop body at nn_ilqr.py:376-384.  Returns dict(F, K, k, X, J, alpha_index, stop)."""
    P = prob.params(target)
    _, c, C = orc.derivs(prob, X, target)
    F = network_jacobians(network, X, prob.n)
    F = gaussian_filter1d(F, sigma=sigma, axis=0) if sigma > 0 else F
    K, k = orc.backward(prob, F, c, C)
    Xn, Jn, aidx, _, _ = orc.line_search(prob, X, K, k, J_last, target, max_line_search, gamma)
    return dict(F=F, K=K, k=k, X=Xn, J=Jn, alpha_index=aidx, stop=orc.is_stop(Jn, J_last, criterion, stopping_method))
