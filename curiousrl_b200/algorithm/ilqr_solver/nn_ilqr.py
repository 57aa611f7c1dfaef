"""NNiLQR: iLQR with a learned dynamics Jacobian (reference: CuriousRL/algorithm/ilqr_solver/nn_ilqr.py:22-420).

The reference fits a network  x_{t+1} = net(x_t, u_t)  to random roll-outs of the scenario and then runs iLQR where
the F matrices of the backward pass are the network's Jacobians, smoothed along the horizon with a Gaussian filter
(:376-380), while the line search rolls out the REAL system (`iLQRWrapper.forward_pass` uses `self.dynamic_model`,
ilqr_wrapper.py:192-213); whenever the iteration stops the network is re-trained on the trajectories seen so far
(:393-404).  Same class names, constructor arguments and methods here.  What runs where:

* the networks (`SmallNetwork`, `LargeNetwork`, `SmallResidualNetwork`) are torch modules and are trained with torch
  (`torch.optim.RAdam` stands in for `torch_optimizer.RAdam`, which is not in the image) - plumbing, as in the reference;
* `NNiLQRDynamicModel.eval_grad_dynamic_model` - per time step n copies of the row pushed forward and backward through
  autograd in the reference (:265-274) - is ONE call of `crl_mlp_jacobian` for all points of all trajectories:
  BatchNorm folded into the weights, the (points x n) x H2 x H1 product on the tensor cores (csrc/nn_jacobian.cu);
* smoothing = `crl_smooth_time` with the matrix of scipy's `gaussian_filter1d`; cost derivatives, the Riccati pass on the
  learned F (`crl_backward_dense`) and the line search on the real system (`crl_forward_pass`) are the entry points
  `BasiciLQR` uses.

`solve()` keeps the reference's single-problem loop (with re-training); `solve_batch` iterates B problems that share the
learned model in lock step on the device (no re-training inside), which is where the Jacobian batch gets large.
"""
from __future__ import annotations

import ctypes
import os
import time as tm
from typing import Optional

import numpy as np
import torch
from torch import nn

from ... import _native
from ...data import Data, Dataset
from ...utils.Logger import logger
from .device import DeviceModel
from .solver import iLQRDynamicModel, iLQRObjectiveFunction, iLQRWrapper, _np

VALI_DATASET_SIZE = 10


# ------------------------------------------------------------------------------------------------ networks
class Residual(nn.Module):
    """The residual block of `SmallResidualNetwork` (nn_ilqr.py:22-45)."""

    def __init__(self, action_channels, output_channels, is_shorcut=True):
        super().__init__()
        self.is_shorcut = is_shorcut
        self.bn1 = nn.BatchNorm1d(action_channels)
        self.relu = nn.ReLU()
        self.linear1 = nn.Linear(action_channels, output_channels)
        self.bn2 = nn.BatchNorm1d(output_channels)
        self.linear2 = nn.Linear(output_channels, output_channels)
        self.shorcut = nn.Linear(action_channels, output_channels)
        self.main_track = nn.Sequential(self.bn1, self.relu, self.linear1, self.bn2, self.relu, self.linear2)

    def forward(self, X):
        Y = self.main_track(X) + (self.shorcut(X) if self.is_shorcut else X)
        return torch.nn.functional.relu(Y)


class SmallResidualNetwork(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.layer = nn.Sequential(Residual(in_dim, 64), Residual(64, 32), Residual(32, 16), Residual(16, 8),
                                   nn.Linear(8, out_dim))

    def forward(self, x):
        return self.layer(x)


class _ThreeLayer(nn.Module):
    """Linear - BatchNorm - ReLU - Linear - BatchNorm - ReLU - Linear: the form `crl_mlp_jacobian` evaluates."""
    hidden = (128, 64)

    def __init__(self, in_dim, out_dim):
        super().__init__()
        h1, h2 = self.hidden
        self.layer = nn.Sequential(nn.Linear(in_dim, h1), nn.BatchNorm1d(h1), nn.ReLU(),
                                   nn.Linear(h1, h2), nn.BatchNorm1d(h2), nn.ReLU(), nn.Linear(h2, out_dim))

    def forward(self, x):
        return self.layer(x)


class SmallNetwork(_ThreeLayer):
    """nn_ilqr.py:66-82"""
    hidden = (128, 64)


class LargeNetwork(_ThreeLayer):
    """nn_ilqr.py:84-100"""
    hidden = (800, 400)


def _ceil32(v):
    return (int(v) + 31) // 32 * 32


class FoldedMLP:
    """A `_ThreeLayer` network in eval mode as the plain affine-relu chain the device kernels take: BatchNorm (running
    statistics) folded into the preceding Linear, hidden widths zero-padded to multiples of 32, W2 in both layouts."""

    def __init__(self, network: nn.Module, device):
        seq = getattr(network, "layer", None)
        ok = (isinstance(seq, nn.Sequential) and len(seq) == 7 and isinstance(seq[0], nn.Linear) and
              isinstance(seq[1], nn.BatchNorm1d) and isinstance(seq[3], nn.Linear) and isinstance(seq[4], nn.BatchNorm1d) and
              isinstance(seq[6], nn.Linear))
        if not ok:
            raise NotImplementedError("the device Jacobian handles Linear-BatchNorm-ReLU-Linear-BatchNorm-ReLU-Linear networks "
                                      "(SmallNetwork, LargeNetwork); %s is not one" % type(network).__name__)

        def fold(lin, bn):
            W, b = lin.weight.detach().double(), lin.bias.detach().double()
            s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            return s[:, None] * W, s * (b - bn.running_mean.detach().double()) + bn.bias.detach().double()

        W1, b1 = fold(seq[0], seq[1])
        W2, b2 = fold(seq[3], seq[4])
        W3, b3 = seq[6].weight.detach().double(), seq[6].bias.detach().double()
        self.d, self.n = int(W1.shape[1]), int(W3.shape[0])
        self.h1, self.h2 = _ceil32(W1.shape[0]), _ceil32(W2.shape[0])
        if self.h2 > 512 or self.d > 16 or self.n > 16:
            raise NotImplementedError("the device Jacobian needs n + m <= 16, n <= 16 and a second hidden layer <= 512 wide")
        f32 = dict(dtype=torch.float32, device=device)
        self.w1 = torch.zeros((self.h1, self.d), **f32)
        self.w1[:W1.shape[0]] = W1
        self.b1 = torch.zeros(self.h1, **f32)
        self.b1[:b1.shape[0]] = b1
        self.w2 = torch.zeros((self.h2, self.h1), **f32)
        self.w2[:W2.shape[0], :W2.shape[1]] = W2
        self.w2t = self.w2.t().contiguous()
        self.b2 = torch.zeros(self.h2, **f32)
        self.b2[:b2.shape[0]] = b2
        self.w3 = torch.zeros((self.n, self.h2), **f32)
        self.w3[:, :W3.shape[1]] = W3
        self.b3 = b3.to(**f32).contiguous()
        self._ws = None

    def jacobian(self, x: torch.Tensor, impl: int = 1):
        """x [S, d] float32 on the device -> (y [S, n], jac [S, n, d]) float32."""
        lib = _native.load()
        x = x.to(torch.float32).contiguous()
        S = int(x.shape[0])
        y = torch.empty((S, self.n), dtype=torch.float32, device=x.device)
        jac = torch.empty((S, self.n, self.d), dtype=torch.float32, device=x.device)
        need = lib.crl_mlp_jacobian_workspace_bytes(self.d, self.h1, self.h2, self.n, S, impl)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _native.check(lib.crl_mlp_jacobian(self.d, self.h1, self.h2, self.n, S, p(x), p(self.w1), p(self.b1), p(self.w2),
                                           p(self.w2t), p(self.b2), p(self.w3), p(self.b3), p(y), p(jac), p(self._ws),
                                           self._ws.numel(), impl,
                                           ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                      "crl_mlp_jacobian")
        return y, jac


_FILTER_CACHE = {}


def gaussian_filter_matrix(T: int, sigma: float) -> np.ndarray:
    """G with  gaussian_filter1d(F, sigma, axis=0) == einsum('tu,u...->t...', G, F)  (the filter is linear; scipy's default
    reflecting boundary and 4-sigma truncation come with it)."""
    key = (int(T), float(sigma))
    if key not in _FILTER_CACHE:
        from scipy.ndimage import gaussian_filter1d
        _FILTER_CACHE[key] = np.ascontiguousarray(gaussian_filter1d(np.eye(T), sigma=sigma, axis=0))
    return _FILTER_CACHE[key]


def smooth_time(F: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """F [B, T, ...] float64 on the device, G [T, T] float64 on the device -> filtered copy (crl_smooth_time)."""
    F = F.contiguous()
    B, T = int(F.shape[0]), int(F.shape[1])
    W = int(F.numel() // (B * T))
    out = torch.empty_like(F)
    _native.check(_native.load().crl_smooth_time(B, T, W, ctypes.c_void_p(G.data_ptr()), ctypes.c_void_p(F.data_ptr()),
                                                 ctypes.c_void_p(out.data_ptr()),
                                                 ctypes.c_void_p(torch.cuda.current_stream(F.device).cuda_stream)),
                  "crl_smooth_time")
    return out


# ------------------------------------------------------------------------------------------------ learned dynamics
class NNiLQRDynamicModel(iLQRDynamicModel):
    """A network fitted to the system's transitions (nn_ilqr.py:102-274).  CUDA only, like the reference's."""

    def __init__(self, network, init_state, init_action, device=None, jacobian_impl=1):
        self._init_action = init_action
        self._init_state = init_state
        self._n = int(init_state.shape[0])
        self._m = int(init_action.shape[1])
        self._T = int(init_action.shape[0])
        if not torch.cuda.is_available():
            raise RuntimeError("NNiLQRDynamicModel needs a CUDA device (there is no CPU path)")
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._model = network.to(self._device)
        self._jacobian_impl = jacobian_impl
        self._folded: Optional[FoldedMLP] = None

    # -- data ----------------------------------------------------------------------------------
    def _process_data(self, dataset):
        """(x_t, u_t) -> x_{t+1} pairs of the stored trajectories (rows that end a trajectory have no successor)."""
        traj = dataset.fetch_all_data().state
        idx = torch.arange(len(traj), device=traj.device)
        keep = idx[(idx + 1) % self._T != 0]
        return traj[keep], traj[keep + 1, 0:self._n]

    # -- training (nn_ilqr.py:130-230) -----------------------------------------------------------
    def _fit(self, X, Y, max_epoch, stopping_criterion, lr, on_check=None):
        loss_fun = nn.MSELoss()
        optimizer = torch.optim.RAdam(self._model.parameters(), lr=lr, weight_decay=1e-4)
        losses = np.zeros(max_epoch)
        for epoch in range(max_epoch):
            self._model.train()
            optimizer.zero_grad()
            obj = loss_fun(self._model(X), Y)
            obj.backward()
            optimizer.step()
            value = obj.item()
            losses[epoch] = value
            if value < stopping_criterion or epoch % 100 == 0:
                if on_check is not None:
                    on_check(epoch, value)
                if value < stopping_criterion:
                    self._model.eval()
                    self._folded = None
                    return losses[:epoch + 1]
        raise Exception("Maximum epoch in the training is reached!")

    def pretrain(self, dataset_train: Dataset, dataset_vali: Dataset, max_epoch=50000, stopping_criterion=1e-4, lr=1e-3,
                 model_name="NeuralDynamic.model"):
        """Fit the network to `dataset_train` until the training loss falls below `stopping_criterion`; the fitted model
        is saved under the logger's folder and loaded from there when it already exists (nn_ilqr.py:130-207)."""
        model_path = logger.logger_path
        if model_path is None:
            raise Exception("Please call logger.set_folder_name(folder_name).set_is_use_logger(True) to enable the logger function!")
        os.makedirs(model_path, exist_ok=True)
        path = os.path.join(model_path, model_name)
        if os.path.exists(path):
            logger.info("[+ +] Model file \"" + model_name + "\" exists. Loading....")
            self._model.load_state_dict(torch.load(path, map_location=self._device))
            self._model.eval()
            self._folded = None
            return
        logger.info("[+ +] Model file \"" + model_name + "\" does NOT exist. Pre-traning starts...")
        X_train, Y_train = self._process_data(dataset_train)
        X_vali, Y_vali = self._process_data(dataset_vali)
        vali = []

        def check(epoch, value):
            self._model.eval()
            with torch.no_grad():
                obj_vali = nn.functional.mse_loss(self._model(X_vali), Y_vali).item()
            vali.append(obj_vali)
            logger.info("[+ +] Epoch: %5d     Train Loss: %.5e     Vali Loss:%.5e" % (epoch + 1, value, obj_vali))

        start = tm.time()
        losses = self._fit(X_train, Y_train, max_epoch, stopping_criterion, lr, check)
        logger.info("[+ +] Pretraining finished! Model file \"" + model_name + "\" is saved!")
        logger.info("[+ +] Pretraining time: %.8f" % (tm.time() - start))
        torch.save(self._model.state_dict(), path)
        np.savez(os.path.join(model_path, model_name + "_training.npz"), Train_loss=losses, Vali_loss=np.asarray(vali))

    def retrain(self, dataset, max_epoch=10000, stopping_criterion=1e-3, lr=1e-3):
        logger.info("[+ +] Re-traning starts...")
        X_train, Y_train = self._process_data(dataset)
        self._fit(X_train, Y_train, max_epoch, stopping_criterion, lr,
                  lambda epoch, value: logger.info("[+ +] Epoch: %5d   Train Obj: %.5e" % (epoch + 1, value)))
        logger.info("[+ +] Re-training finished!")

    # -- evaluation through the network ------------------------------------------------------------
    def _predict(self, rows: torch.Tensor) -> torch.Tensor:
        self._model.eval()
        with torch.no_grad():
            return self._model(rows)

    def eval_traj(self, init_state=None, action_traj=None):
        """Open-loop rollout through the network (nn_ilqr.py:232-246): [T, n+m, 1] float64."""
        x0 = self._init_state if init_state is None else init_state
        U = self._init_action if action_traj is None else action_traj
        U = torch.from_numpy(np.asarray(U, dtype=np.float64).reshape(self._T, self._m)).float().to(self._device)
        traj = torch.zeros((self._T, self._n + self._m), device=self._device)
        traj[0, :self._n] = torch.from_numpy(np.asarray(x0, dtype=np.float64).reshape(-1)).float().to(self._device)
        traj[0, self._n:] = U[0]
        for tau in range(self._T - 1):
            traj[tau + 1, :self._n] = self._predict(traj[tau].reshape(1, -1))[0]
            traj[tau + 1, self._n:] = U[tau + 1]
        return traj.cpu().double().numpy().reshape(self._T, self._m + self._n, 1)

    def update_traj(self, old_traj, K_matrix, k_vector, alpha):
        """Closed-loop rollout through the network (nn_ilqr.py:248-263)."""
        n, m, T = self._n, self._m, self._T
        old = np.asarray(old_traj, dtype=np.float64).reshape(T, n + m, 1)
        new = np.zeros((T, n + m, 1))
        new[0] = old[0]
        for tau in range(T - 1):
            delta_x = new[tau, :n] - old[tau, :n]
            new[tau, n:] = old[tau, n:n + m] + (K_matrix[tau] @ delta_x + alpha * k_vector[tau])
            row = torch.from_numpy(new[tau].T).float().to(self._device)
            new[tau + 1, 0:n, 0] = self._predict(row)[0].cpu().double().numpy()
        return new

    def folded(self) -> FoldedMLP:
        if self._folded is None:
            self._model.eval()
            self._folded = FoldedMLP(self._model, self._device)
        return self._folded

    def eval_grad_dynamic_model(self, trajectory):
        """Jacobians of the network at every row: NumPy [T, n+m, 1] -> [T, n, n+m] float64 (nn_ilqr.py:265-274);
        torch [B, T, n+m] -> [B, T, n, n+m] float64 on the device."""
        d = self._n + self._m
        if isinstance(trajectory, torch.Tensor):
            B = trajectory.reshape(-1, self._T, d).shape[0]
            _, jac = self.folded().jacobian(trajectory.reshape(-1, d), self._jacobian_impl)
            return jac.double().reshape(B, self._T, self._n, d)
        rows = torch.from_numpy(np.asarray(trajectory, dtype=np.float64).reshape(self._T, d)).to(self._device)
        _, jac = self.folded().jacobian(rows, self._jacobian_impl)
        return jac.double().cpu().numpy()


# ------------------------------------------------------------------------------------------------ the solver
class NNiLQR(iLQRWrapper):
    """Neural-network iLQR (nn_ilqr.py:277-420): same constructor arguments; `device`, `seed`, `pretrain_max_epoch` and
    `jacobian_impl` (1 = tensor cores, 0 = FP32 CUDA cores) are defaulted extensions."""

    def __init__(self, iLQR_stopping_criterion=1e-6, max_line_search=10, gamma=0.5, line_search_method="vanilla",
                 stopping_method="vanilla", network_class=LargeNetwork, trial_no=100, training_stopping_criterion=1e-4,
                 iLQR_max_iter=1000, decay_rate=0.98, decay_rate_max_iters=200, gaussian_filter_sigma=5,
                 gaussian_noise_sigma=1, device=None, seed=None, pretrain_max_epoch=50000, jacobian_impl=1):
        super().__init__(stopping_criterion=iLQR_stopping_criterion, max_line_search=max_line_search, gamma=gamma,
                         line_search_method=line_search_method, stopping_method=stopping_method,
                         network_class=network_class, trial_no=trial_no,
                         training_stopping_criterion=training_stopping_criterion, iLQR_max_iter=iLQR_max_iter,
                         decay_rate=decay_rate, decay_rate_max_iters=decay_rate_max_iters,
                         gaussian_filter_sigma=gaussian_filter_sigma, gaussian_noise_sigma=gaussian_noise_sigma)
        self._network_class = network_class
        self._trial_no = trial_no
        self._training_stopping_criterion = training_stopping_criterion
        self._iLQR_max_iter = iLQR_max_iter
        self._decay_rate = decay_rate
        self._decay_rate_max_iters = decay_rate_max_iters
        self._gaussian_filter_sigma = gaussian_filter_sigma
        self._gaussian_noise_sigma = gaussian_noise_sigma
        self._device = device
        self._rng = np.random.default_rng(seed) if seed is not None else np.random
        self._seed = seed
        self._pretrain_max_epoch = pretrain_max_epoch
        self._jacobian_impl = jacobian_impl
        self._trajectory = None

    def _real_device_model(self, scenario) -> DeviceModel:
        constr = np.asarray(scenario.constr, dtype=np.float64)
        n = int(scenario.init_state.shape[0])
        T = int(scenario.init_action.shape[0])
        if scenario.name in _native.GENERIC_MODEL_IDS:
            compiled = np.asarray(_native.model_bounds(_native.GENERIC_MODEL_IDS[scenario.name]), dtype=np.float64)
            return DeviceModel(scenario.name, T, 0.0, 0.0, dtype="float64", device=self._device,
                               bounds=None if np.array_equal(compiled, constr) else constr)
        return DeviceModel(scenario.name, T, constr[n, 0], constr[n, 1], dtype="float64", device=self._device)

    def init(self, scenario) -> "NNiLQR":
        if not scenario.with_model() or scenario.is_action_discrete() or scenario.is_output_image():
            raise Exception("Scenario \"" + scenario.name + "\" cannot learn with NNiLQR")
        self._example_name = scenario.name
        self._dev = self._real_device_model(scenario)
        # the REAL system (rollouts of the line search, data generation) and the scenario's cost
        self._dynamic_model = iLQRDynamicModel(self._dev, scenario.init_state, scenario.init_action, self)
        self._obj_fun = iLQRObjectiveFunction(self._dev, scenario.add_param, self)
        if self._seed is not None:
            torch.manual_seed(self._seed)
        network = self._network_class(scenario.n + scenario.m, scenario.n)
        action_constr = np.asarray(scenario.constr, dtype=np.float64)[scenario.n:]
        if np.any(np.isinf(action_constr)):
            raise Exception("The constraint of dynamic system action must not be inf!")

        def random_trajectories(count):
            """`count` roll-outs of the real system under uniformly random actions - one batched device call."""
            U = self._rng.uniform(action_constr[:, 0], action_constr[:, 1], size=[count, scenario.T, len(action_constr)])
            x0 = np.repeat(np.asarray(scenario.init_state, dtype=np.float64).reshape(1, -1), count, axis=0)
            X, _ = self._dev.rollout(x0, self._params_for(1), U)
            return X.reshape(count * scenario.T, -1).float()

        self._dataset_train = Dataset(buffer_size=self._trial_no * scenario.T, state_dim=scenario.n + scenario.m, action_dim=1)
        self._dataset_train.update_dataset(Data(state=random_trajectories(self._trial_no)))
        dataset_vali = Dataset(buffer_size=VALI_DATASET_SIZE * scenario.T, state_dim=scenario.n + scenario.m, action_dim=1)
        dataset_vali.update_dataset(Data(state=random_trajectories(VALI_DATASET_SIZE)))
        self._nn_dynamic_model = NNiLQRDynamicModel(network, scenario.init_state, scenario.init_action, device=self._dev.device,
                                                    jacobian_impl=self._jacobian_impl)
        self._nn_dynamic_model.pretrain(self._dataset_train, dataset_vali, max_epoch=self._pretrain_max_epoch,
                                        stopping_criterion=self._training_stopping_criterion)
        return self

    nn_dynamic_model = property(lambda self: self._nn_dynamic_model)

    def _filter(self) -> torch.Tensor:
        key = (self._dev.T, float(self._gaussian_filter_sigma))
        if getattr(self, "_G_key", None) != key:
            self._G = torch.from_numpy(gaussian_filter_matrix(*key)).to(self._dev.device)
            self._G_key = key
        return self._G

    def learned_F(self, X: torch.Tensor) -> torch.Tensor:
        """Smoothed network Jacobians at the rows of X [B, T, n+m] (float64, device) -> [B, T, n, n+m] float64."""
        F = self._nn_dynamic_model.eval_grad_dynamic_model(X)
        return smooth_time(F, self._filter()) if self._gaussian_filter_sigma > 0 else F

    # ------------------------------------------------------------------ solve (one problem, nn_ilqr.py:363-411)
    def solve(self):
        dev = self._dev
        trajectory = self._dynamic_model.eval_traj()                      # initial feasible trajectory (real system)
        init_obj = self._obj_fun.eval_obj_fun(trajectory)
        logger.info("[+ +] Initial Obj.Val.: %.5e" % init_obj)
        self.set_obj_fun_value(init_obj)
        new_data = []
        result_obj_val = np.zeros(self._iLQR_max_iter)
        result_iter_time = np.zeros(self._iLQR_max_iter)
        re_train_stopping_criterion = self._training_stopping_criterion
        start_time = tm.time()
        for i in range(int(self._iLQR_max_iter)):
            if i == 1:
                start_time = tm.time()
            iter_start_time = tm.time()
            X = dev.real(trajectory.reshape(1, dev.T, dev.d))
            _, c, C = dev.derivs(X, self._params_for(1), want=("c", "C"))
            F = self.learned_F(X)
            K, k = dev.backward_dense(F, c, C)
            trajectory, _C, _c, _F, obj_val, isStop = self.forward_pass(trajectory, _np(K)[0], _np(k)[0][:, :, None])
            if i < self._decay_rate_max_iters:
                re_train_stopping_criterion = re_train_stopping_criterion * self._decay_rate
            iter_time = tm.time() - iter_start_time
            logger.info("[+ +] Iter.No.:%3d  Iter.Time:%.3e   Obj.Val.:%.5e" % (i, iter_time, obj_val))
            result_obj_val[i] = obj_val
            result_iter_time[i] = iter_time
            if isStop:
                if len(new_data) != 0:                                    # the optimal trajectory is in the data set already
                    trajectory_noisy = trajectory
                else:
                    noise = self._rng.normal(0, self._gaussian_noise_sigma, [dev.T, dev.m, 1])
                    trajectory_noisy = self._dynamic_model.eval_traj(action_traj=trajectory[:, dev.n:] + noise)
                new_data += [trajectory_noisy]
                data = np.concatenate(new_data[-int(self._trial_no / 5):])[:, :, 0]
                self._dataset_train.update_dataset(Data(state=torch.from_numpy(data).float().to(dev.device)))
                logger.save_to_json(trajectory=trajectory.tolist(), trajectroy_noisy=trajectory_noisy.tolist())
                self._nn_dynamic_model.retrain(self._dataset_train, max_epoch=100000,
                                               stopping_criterion=re_train_stopping_criterion)
                new_data = []
            else:
                new_data += [trajectory]
        self._trajectory = trajectory
        if logger.logger_path is not None:
            np.savez(os.path.join(logger.logger_path, "_result.npz"), obj_val=result_obj_val, iter_time=result_iter_time)
        logger.info("[+ +] Completed! All Time:%.5e" % (tm.time() - start_time))
        self.result_obj_val = result_obj_val
        return None

    # ------------------------------------------------------------------ batched extension
    def solve_batch(self, init_state, add_param=None, max_iter=None):
        """B problems (initial states [B, n], targets add_param [B, p] / [B, T, p] / None = the scenario's) that share the
        learned model, iterated in lock step on the device: per iteration one cost-derivative call, ONE Jacobian call for
        all B * T rows, one smoothing, one Riccati call and one line-search call on the real system.  Problems that meet
        the stopping test keep their trajectory; no re-training inside.  Returns (trajectory [B, T, n+m], cost [B],
        iterations [B] int32, stopped [B] bool) on the device."""
        dev = self._dev
        x0 = dev.real(init_state).reshape(-1, dev.n)
        B = x0.shape[0]
        prm = self._params_for(B) if add_param is None else dev.real(add_param)
        u0 = np.asarray(self._dynamic_model._init_action, dtype=np.float64)
        X, J = dev.rollout(x0, prm, u0 if u0.any() else None)
        J_last = J.clone()
        iters = torch.zeros(B, dtype=torch.int32, device=dev.device)
        stopped = torch.zeros(B, dtype=torch.bool, device=dev.device)
        for _ in range(int(self._iLQR_max_iter if max_iter is None else max_iter)):
            _, c, C = dev.derivs(X, prm, want=("c", "C"))
            F = self.learned_F(X)
            K, k = dev.backward_dense(F, c, C)
            Xn, Jn, _aidx, stop = dev.forward_pass(self._opts(), X, K, k, J_last, prm)
            live = ~stopped
            X = torch.where(live.reshape(B, 1, 1), Xn, X)
            J_last = torch.where(live, Jn, J_last)
            iters += live.to(torch.int32)
            stopped |= stop.bool()
            if bool(stopped.all()):
                break
        return X, J_last, iters, stopped
