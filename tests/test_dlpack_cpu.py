"""Host logic of the tensor boundary (tfrpn/_tensor.py) without a GPU: a DLPack producer that is neither
torch nor TensorFlow is ingested zero-copy, and unsupported objects are rejected."""
import numpy as np
import pytest
import torch

from tfrpn import _tensor


class FakeProducer:
    """Minimal third-party array: only the two DLPack protocol methods, backed by a NumPy buffer."""

    def __init__(self, a):
        self._a = a
        self.calls = []

    def __dlpack__(self, *args, **kw):
        self.calls.append(kw)
        kw.pop("stream", None)      # a CPU producer has no stream to synchronise
        return self._a.__dlpack__(**{k: v for k, v in kw.items() if k == "max_version" and v is not None})

    def __dlpack_device__(self):
        return self._a.__dlpack_device__()


def test_fake_dlpack_producer_is_ingested_zero_copy():
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    prod = FakeProducer(a)
    o = _tensor.Origin()
    t = _tensor.ingest(prod, o, "boxes")
    assert o.kind == "dlpack" and isinstance(t, torch.Tensor) and t.dtype == torch.float32
    assert tuple(t.shape) == (2, 3, 4) and t.data_ptr() == a.ctypes.data      # same memory, no copy
    assert len(prod.calls) == 1
    a[0, 0, 0] = 99.0
    assert float(t[0, 0, 0]) == 99.0


def test_ingest_rejects_unknown_objects_and_wrong_dtypes():
    with pytest.raises(TypeError):
        _tensor.ingest(object(), _tensor.Origin(), "x")
    prod = FakeProducer(np.zeros((3, 4), np.float64))
    if not torch.cuda.is_available():
        # float64 boxes are refused before any device work (float32 tensors are never cast silently)
        with pytest.raises(ValueError):
            _tensor.to_device(prod, torch.float32, _tensor.Origin(), "boxes")


def test_no_cuda_means_no_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        _tensor.to_device(np.zeros((3, 4), np.float32), torch.float32, _tensor.Origin(), "boxes")
