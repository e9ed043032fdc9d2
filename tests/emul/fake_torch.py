"""fake_torch -- TEST STUB (tests/emul): the dozen torch calls the gpu-marked tests use to hold
"device" buffers, backed by numpy.  tools/emul_gpu_suite.py installs it as `torch` in a CHILD
process so that the tests that drive the device API (`mtz_dev_*`, `mtz_k_lz4_*`) with
torch.cuda tensors can run against the emulated library, where device memory is host memory.
Never imported by the product or by the normal test session."""
import numpy as np

import time as _time

uint8 = np.uint8
int64 = np.int64
int32 = np.int32
float64 = np.float64


class Tensor(object):
    def __init__(self, a):
        self.a = a

    def cuda(self):
        return Tensor(self.a.copy())                      # H2D: a device copy

    def cpu(self):
        return Tensor(self.a.copy())

    def numpy(self):
        return self.a

    def data_ptr(self):
        return self.a.ctypes.data

    def numel(self):
        return int(self.a.size)

    def __getitem__(self, k):
        return Tensor(self.a[k])

    def copy_(self, other):
        self.a[...] = other.a
        return self

    def item(self):
        return self.a.reshape(-1)[0].item()

    def tolist(self):
        return self.a.tolist()

    def __len__(self):
        return len(self.a)


def from_numpy(a):
    return Tensor(np.ascontiguousarray(a))


def zeros(n, dtype=uint8, device=None):
    return Tensor(np.zeros(n, dtype=dtype))


def empty(n, dtype=uint8, device=None):
    return Tensor(np.full(n, 0x5A, dtype=dtype) if dtype == uint8 else np.zeros(n, dtype=dtype))


def equal(x, y):
    return bool(np.array_equal(x.a, y.a))


def tensor(data, dtype=None, device=None):
    return Tensor(np.array(data, dtype=dtype))


def device(*a):
    return None


class _Stream(object):
    cuda_stream = 0


class _Event(object):
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = _time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


class _Cuda(object):
    Stream = _Stream
    Event = _Event

    @staticmethod
    def synchronize():
        return None

    @staticmethod
    def set_device(d):
        return None

    @staticmethod
    def empty_cache():
        return None

    @staticmethod
    def is_available():
        return True


cuda = _Cuda()
