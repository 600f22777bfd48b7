"""Device buffers for the reference-layout API.

The reference moves its arrays with `numba.cuda.to_device` and reads them back
with `.copy_to_host()` (main/simulation.py:161-190, 295-299).  This module
offers the same two verbs on top of torch CUDA tensors, which are used purely
as device memory: the compute is done by libgof_b200.so on the raw pointers.

State arrays are fp32 on the device (float64 host arrays are rounded once at
upload); index arrays are int32.
"""
from __future__ import annotations

import numpy as np
import torch


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("game_of_flies_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device(device if device is not None else "cuda")


class DeviceArray:
    """A 1-D device array with numba's DeviceNDArray verbs."""

    __slots__ = ("tensor", "host_dtype")

    def __init__(self, tensor: torch.Tensor, host_dtype=None):
        assert tensor.is_cuda and tensor.is_contiguous()
        self.tensor = tensor
        self.host_dtype = np.dtype(host_dtype) if host_dtype is not None else None

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    @property
    def size(self) -> int:
        return self.tensor.numel()

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return self.tensor.dtype

    def copy_to_host(self, ary: np.ndarray | None = None) -> np.ndarray:
        """Like numba's: fill `ary` in place if given (any float/int dtype), else return a new array."""
        host = self.tensor.detach().cpu().numpy()
        if ary is None:
            return host.astype(self.host_dtype) if self.host_dtype is not None else host
        np.copyto(ary, host.reshape(ary.shape), casting="unsafe")
        return ary

    def copy_from_host(self, ary: np.ndarray) -> None:
        src = np.ascontiguousarray(ary).reshape(self.shape)
        self.tensor.copy_(torch.from_numpy(src.astype(_np_dtype(self.tensor.dtype), copy=False)))

    def __len__(self):
        return self.size


def _np_dtype(dt: torch.dtype):
    return {torch.float32: np.float32, torch.int32: np.int32, torch.float64: np.float64,
            torch.uint8: np.uint8}[dt]


def to_device(ary, device=None) -> DeviceArray:
    """cuda.to_device: floats become fp32, integers int32."""
    dev = _require_cuda(device)
    a = np.ascontiguousarray(ary)
    if np.issubdtype(a.dtype, np.floating):
        t = torch.from_numpy(a.astype(np.float32)).to(dev)
    elif np.issubdtype(a.dtype, np.integer) or a.dtype == np.bool_:
        t = torch.from_numpy(a.astype(np.int32)).to(dev)
    else:
        raise TypeError(f"unsupported dtype {a.dtype}")
    return DeviceArray(t, host_dtype=a.dtype)


def device_array(n: int, dtype=torch.float32, device=None) -> DeviceArray:
    dev = _require_cuda(device)
    return DeviceArray(torch.zeros(int(n), dtype=dtype, device=dev))


def as_ptr(x) -> int:
    """Raw device pointer of a DeviceArray / CUDA tensor, 0 for None."""
    if x is None:
        return 0
    if isinstance(x, DeviceArray):
        return x.ptr
    if isinstance(x, torch.Tensor):
        assert x.is_cuda and x.is_contiguous()
        return x.data_ptr()
    raise TypeError(f"expected a device array, got {type(x)}")


def current_stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream
