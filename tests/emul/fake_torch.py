"""fake_torch -- TEST STUB (tests/emul): the dozen torch calls the gpu-marked tests use to hold
"device" buffers, backed by numpy.  tools/emul_gpu_suite.py installs it as `torch` in a CHILD
process so that the tests that drive the device API (`mtz_dev_*`, `mtz_k_lz4_*`) with
torch.cuda tensors can run against the emulated library, where device memory is host memory.
Never imported by the product or by the normal test session."""
import numpy as np

uint8 = np.uint8
int64 = np.int64
int32 = np.int32


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


class _Cuda(object):
    @staticmethod
    def synchronize():
        return None

    @staticmethod
    def is_available():
        return True


cuda = _Cuda()
