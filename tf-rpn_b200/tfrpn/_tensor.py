"""Tensor plumbing for the drop-in functions: anything in, same kind out.

Accepted inputs: torch CUDA tensors (zero copy), any object exposing ``__dlpack__`` on a CUDA
device (e.g. TensorFlow >= 2.2 eager tensors through ``tf.experimental.dlpack``; zero copy), and
HOST buffers (NumPy arrays, torch CPU tensors), which are copied to the GPU, processed there and
copied back.  The arithmetic always runs in libtfrpn_cuda.so on the GPU; PyTorch is used only to
own device memory and streams.
"""
import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("tfrpn needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def default_device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


class Origin:
    """Remembers where the first tensor argument came from, to return results the same way."""
    __slots__ = ("kind", "device")

    def __init__(self):
        self.kind = None     # "torch" | "numpy" | "torch_cpu" | "tf" | "dlpack"
        self.device = None

    def note(self, kind, device):
        if self.kind is None:
            self.kind = kind
        if self.device is None and device is not None and device.type == "cuda":
            self.device = device


def _is_tf(x):
    return type(x).__module__.split(".")[0] == "tensorflow"


def ingest(x, origin, name="tensor"):
    """-> a torch view of `x` wherever it lives (no copy for torch tensors and DLPack producers).

    DLPack stream contract: ``torch.from_dlpack`` calls ``x.__dlpack__(stream=s)`` with ``s`` = torch's CURRENT
    stream on the producer's device (1 for the legacy default stream), and the producer must make the tensor's
    pending writes visible to work later enqueued on ``s`` -- the stream every kernel of this call is launched on
    (``stream_ptr``).  So a producer that computes on its own stream (TensorFlow) is ordered before our kernels
    by the protocol itself, not by an assumption about default streams."""
    if isinstance(x, torch.Tensor):
        t = x
        origin.note("torch" if t.is_cuda else "torch_cpu", t.device)
    elif isinstance(x, np.ndarray) or isinstance(x, (list, tuple, float, int)):
        a = np.asarray(x)
        if a.dtype == np.float64 and not isinstance(x, np.ndarray):
            a = a.astype(np.float32)   # Python lists of floats, like TF's convert_to_tensor
        t = torch.from_numpy(np.ascontiguousarray(a))
        origin.note("numpy", None)
    elif _is_tf(x):
        import tensorflow as tf  # only when the caller already uses it
        t = torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(x))
        origin.note("tf", t.device)
    elif hasattr(x, "__dlpack__"):
        t = torch.from_dlpack(x)
        origin.note("dlpack", t.device)
    else:
        raise TypeError("%s: unsupported tensor type %r" % (name, type(x)))
    return t


def to_device(x, dtype, origin, name="tensor"):
    """-> contiguous torch CUDA tensor of `dtype` (float32 tensors are never cast silently)."""
    t = ingest(x, origin, name)
    if dtype == torch.float32 and t.dtype != torch.float32:
        raise ValueError("%s must be float32, got %s" % (name, t.dtype))
    if dtype == torch.int32 and t.dtype in (torch.int64, torch.int16, torch.int8, torch.uint8):
        t = t.to(torch.int32)
    if dtype == torch.uint8 and t.dtype == torch.bool:
        t = t.to(torch.uint8)
    if t.dtype != dtype:
        raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_cuda:
        dev = origin.device or default_device()
        t = t.to(dev, non_blocking=False)
    elif origin.device is not None and t.device != origin.device:
        raise ValueError("%s is on %s but earlier arguments are on %s" % (name, t.device, origin.device))
    origin.note(origin.kind, t.device)
    return t.contiguous()


def from_device(t, origin):
    """Return a result in the framework the inputs came from."""
    if origin.kind == "numpy":
        return t.cpu().numpy()
    if origin.kind == "torch_cpu":
        return t.cpu()
    if origin.kind == "tf":
        import tensorflow as tf
        return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))
    return t


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
