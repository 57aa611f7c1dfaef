from .device import BatchResult, DeviceModel
from .solver import iLQRWrapper, iLQRObjectiveFunction, iLQRDynamicModel, BasiciLQR
from .log_barrier import LogBarrieriLQR, LogBarrierBatchResult
from .nn_ilqr import (NNiLQR, NNiLQRDynamicModel, SmallNetwork, LargeNetwork, SmallResidualNetwork, Residual, FoldedMLP,
                      gaussian_filter_matrix)
