from .device import BatchResult, DeviceModel
from .solver import iLQRWrapper, iLQRObjectiveFunction, iLQRDynamicModel, BasiciLQR
from .log_barrier import LogBarrieriLQR, LogBarrierBatchResult
