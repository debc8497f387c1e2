"""NumPy stand-in for the ~45 ``tf.*`` primitives that /root/reference/utils/{bbox,train}_utils.py
and predictor.py:52-60 call.  TEST INFRASTRUCTURE (part of ``oracle/``), never shipped.

Purpose: TensorFlow cannot be installed in this image, but the reference is pure Python
over TF eager ops.  Putting this directory first on ``sys.path`` lets
``tests/golden/make_golden.py`` import and run the reference's UNMODIFIED source, so the
op composition, broadcasting, op order and dtype walk in the golden vectors are the
reference's own; only the primitive kernels below are restated (each is a single IEEE
float32 NumPy op, or a [TF-internal] rule stated in its docstring).

Type rules mirrored from TF eager: binary ops between tensors require equal dtypes;
Python scalars / lists adopt the tensor's dtype; int / int ``truediv`` promotes to float64;
Python floats become float32 tensors, Python ints int32.
"""
import heapq

import numpy as np

float32 = np.dtype("float32")
float64 = np.dtype("float64")
int32 = np.dtype("int32")
int64 = np.dtype("int64")
bool = np.dtype("bool")  # noqa: A001  (mirrors tf.bool)

_pybool = type(True)


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a):
        self.a = np.asarray(a)

    # -- numpy / python interop
    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def numpy(self):
        return self.a

    @property
    def dtype(self):
        return self.a.dtype

    @property
    def shape(self):
        return self.a.shape

    def __int__(self):
        return int(self.a)

    __index__ = __int__

    def __float__(self):
        return float(self.a)

    def __len__(self):
        return len(self.a)

    def __iter__(self):
        return (Tensor(x) for x in self.a)

    def __repr__(self):
        return "shim.Tensor(%r)" % (self.a,)

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(i.a if isinstance(i, Tensor) else i for i in idx)
        elif isinstance(idx, Tensor):
            idx = idx.a
        return Tensor(self.a[idx])

    # -- arithmetic with TF's dtype discipline
    def _coerce(self, other):
        if isinstance(other, Tensor):
            assert other.a.dtype == self.a.dtype, (
                "TF would raise InvalidArgumentError: %s vs %s" % (self.a.dtype, other.a.dtype))
            return other.a
        return np.asarray(_to_numpy(other), dtype=self.a.dtype)

    def _bin(self, other, fn, rev=False):
        o = self._coerce(other)
        with np.errstate(all="ignore"):
            r = fn(o, self.a) if rev else fn(self.a, o)
        return Tensor(r)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __neg__(self): return Tensor(-self.a)

    def __truediv__(self, o, rev=False):
        if self.a.dtype.kind in "iu":           # [TF-internal] int truediv -> float64
            s = Tensor(self.a.astype(np.float64))
            if isinstance(o, Tensor):
                o = Tensor(o.a.astype(np.float64))
            return s.__truediv__(o, rev)
        return self._bin(o, np.true_divide, rev)

    def __rtruediv__(self, o): return self.__truediv__(o, True)


def _to_numpy(x):
    """Nested lists of Tensors / scalars -> ndarray (what tf.convert_to_tensor accepts)."""
    if isinstance(x, Tensor):
        return x.a
    if isinstance(x, (list, tuple)):
        return np.asarray([_to_numpy(v) for v in x])
    return np.asarray(x)


def _t(x, dtype=None):
    """convert_to_tensor: Python float -> f32, Python int -> i32, bool -> bool."""
    if isinstance(x, Tensor):
        return x
    a = _to_numpy(x)
    if dtype is not None:
        return Tensor(a.astype(dtype))
    if a.dtype == np.float64:
        a = a.astype(np.float32)
    elif a.dtype == np.int64:
        a = a.astype(np.int32)
    return Tensor(a)


def _like(x, ref):
    return x.a if isinstance(x, Tensor) else np.asarray(_to_numpy(x), dtype=ref.dtype)


# ---- construction / shape ----------------------------------------------------------
def constant(value, dtype=None):
    return _t(value, dtype)


def cast(x, dtype):
    return Tensor(_to_numpy(x).astype(dtype))


def range(start, limit=None, delta=1):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return Tensor(np.arange(start, limit, delta, dtype=np.int32))


def shape(x):
    return Tensor(np.asarray(_t(x).a.shape, dtype=np.int32))


def _shape_tuple(s):
    if isinstance(s, Tensor):
        return tuple(int(v) for v in s.a.reshape(-1))
    return tuple(int(v) for v in s)


def reshape(x, s):
    return Tensor(_t(x).a.reshape(_shape_tuple(s)))


def stack(values, axis=0):
    return Tensor(np.stack([_t(v).a for v in values], axis=axis))


def split(x, n, axis=0):
    return [Tensor(p) for p in np.split(_t(x).a, n, axis=axis)]


def squeeze(x, axis=None):
    return Tensor(np.squeeze(_t(x).a, axis=axis))


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_t(x).a, axis))


def transpose(x, perm=None):
    return Tensor(np.transpose(_t(x).a, perm))


def meshgrid(*xs):
    return [Tensor(m) for m in np.meshgrid(*[_t(x).a for x in xs])]  # default indexing="xy"


def zeros_like(x, dtype=None):
    return Tensor(np.zeros_like(_t(x).a, dtype=dtype))


def ones_like(x, dtype=None):
    return Tensor(np.ones_like(_t(x).a, dtype=dtype))


def fill(dims, value):
    return Tensor(np.full(_shape_tuple(dims), value))


# ---- elementwise -------------------------------------------------------------------
def sqrt(x):
    return Tensor(np.sqrt(_t(x).a))


def exp(x):
    with np.errstate(all="ignore"):
        return Tensor(np.exp(_t(x).a))


def round(x):  # noqa: A001   [TF-internal] round half to even
    return Tensor(np.rint(_t(x).a))


def maximum(x, y):
    x = _t(x) if not isinstance(y, Tensor) else x
    if isinstance(x, Tensor):
        return x._bin(y, np.maximum)
    return y._bin(x, np.maximum, True)


def minimum(x, y):
    x = _t(x) if not isinstance(y, Tensor) else x
    if isinstance(x, Tensor):
        return x._bin(y, np.minimum)
    return y._bin(x, np.minimum, True)


def add(x, y):
    return _t(x) + y


def truediv(x, y):
    return _t(x) / y


def clip_by_value(x, lo, hi):
    x = _t(x)
    return Tensor(np.minimum(np.maximum(x.a, _like(lo, x.a)), _like(hi, x.a)))


def equal(x, y):
    x = _t(x)
    return Tensor(x.a == x._coerce(y))


def not_equal(x, y):
    x = _t(x)
    return Tensor(x.a != x._coerce(y))


def greater(x, y):
    x = _t(x)
    return Tensor(x.a > x._coerce(y))


def less(x, y):
    x = _t(x)
    return Tensor(x.a < x._coerce(y))


def logical_and(x, y):
    return Tensor(np.logical_and(_t(x).a, _t(y).a))


def logical_or(x, y):
    return Tensor(np.logical_or(_t(x).a, _t(y).a))


def logical_not(x):
    return Tensor(np.logical_not(_t(x).a))


def where(cond, x=None, y=None):
    c = _t(cond).a
    if x is None:
        return Tensor(np.argwhere(c).astype(np.int64))      # (n, rank) int64, row-major order
    if isinstance(x, Tensor):
        xa = x.a
        ya = _like(y, xa)
    else:
        ya = _t(y).a
        xa = _like(x, ya)
    return Tensor(np.where(c, xa, ya))


# ---- reductions / indexing ---------------------------------------------------------
def reduce_max(x, axis=None):
    return Tensor(np.max(_t(x).a, axis=axis))


def reduce_sum(x, axis=None):
    a = _t(x).a
    return Tensor(np.sum(a, axis=axis, dtype=a.dtype))


def reduce_any(x, axis=None):
    return Tensor(np.any(_t(x).a, axis=axis))


def argmax(x, axis=None, output_type=int64):
    """[TF-internal] first maximal index; NaN never wins Eigen's '>' compare."""
    a = _t(x).a
    a = np.where(np.isnan(a), -np.inf, a) if a.dtype.kind == "f" else a
    return Tensor(np.argmax(a, axis=axis).astype(output_type))


def argsort(values, axis=-1, direction="ASCENDING", stable=False):
    """[TF-internal] DESCENDING = top_k(values, n).indices: equal values keep ascending index."""
    a = _t(values).a
    assert axis == -1
    if direction == "DESCENDING":
        key = -a.astype(np.int64) if a.dtype.kind in "iu" else -a.astype(np.float64)
    else:
        key = a
    return Tensor(np.argsort(key, axis=-1, kind="stable").astype(np.int32))


def gather(params, indices, batch_dims=0, axis=None):
    p, i = _t(params).a, _t(indices).a
    if batch_dims == 0:
        return Tensor(np.take(p, i, axis=0 if axis is None else axis))
    assert batch_dims == 1
    return Tensor(np.stack([p[b][i[b]] for b in np.arange(p.shape[0])], axis=0))


def gather_nd(params, indices):
    p, i = _t(params).a, _t(indices).a
    return Tensor(p[tuple(i[..., d] for d in np.arange(i.shape[-1]))])


def scatter_nd(indices, updates, shape):  # noqa: A002
    i, u = _t(indices).a, _t(updates).a
    out = np.zeros(_shape_tuple(shape), dtype=u.dtype)
    idx = tuple(i[..., d] for d in np.arange(i.shape[-1]))
    if u.dtype == np.bool_:
        out[idx] = np.logical_or(out[idx], u)
    else:
        np.add.at(out, idx, u)
    return Tensor(out)


class nn:  # noqa: N801
    @staticmethod
    def top_k(x, k=1, sorted=True):  # noqa: A002
        """[TF-internal] values descending, equal values -> lower index first."""
        a = _t(x).a
        idx = np.argsort(-a.astype(np.float64), axis=-1, kind="stable")[..., :int(k)]
        return Tensor(np.take_along_axis(a, idx, axis=-1)), Tensor(idx.astype(np.int32))


class math:  # noqa: N801
    @staticmethod
    def log(x):
        with np.errstate(all="ignore"):
            return Tensor(np.log(_t(x).a))


class _Random:
    def __init__(self):
        self._rng = np.random.default_rng(0)

    def set_seed(self, seed):
        self._rng = np.random.default_rng(seed)

    def uniform(self, shape, minval=0, maxval=None, dtype=float32, seed=None):  # noqa: A002
        s = _shape_tuple(shape)
        if np.dtype(dtype).kind in "iu":
            return Tensor(self._rng.integers(int(minval), int(maxval), size=s).astype(dtype))
        hi = 1.0 if maxval is None else float(maxval)
        return Tensor(self._rng.uniform(float(minval), hi, size=s).astype(dtype))


random = _Random()


# ---- tf.image.combined_non_max_suppression -------------------------------------------
def _iou_scalar(bx, i, j):
    """Scalar transcription of the IOU helper of TF's CombinedNonMaxSuppression kernel
    [TF-internal]; every intermediate is rounded to float32 like the C++ floats."""
    f = np.float32
    ymin_i, xmin_i = min(bx[i][0], bx[i][2]), min(bx[i][1], bx[i][3])
    ymax_i, xmax_i = max(bx[i][0], bx[i][2]), max(bx[i][1], bx[i][3])
    ymin_j, xmin_j = min(bx[j][0], bx[j][2]), min(bx[j][1], bx[j][3])
    ymax_j, xmax_j = max(bx[j][0], bx[j][2]), max(bx[j][1], bx[j][3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0)
    iymin, ixmin = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iymax, ixmax = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = f(max(f(iymax - iymin), f(0)) * max(f(ixmax - ixmin), f(0)))
    return f(inter / f(f(area_i + area_j) - inter))


class image:  # noqa: N801
    @staticmethod
    def flip_left_right(img):
        """Reverse the width axis (axis -2 of (..., H, W, C)); used by utils/data_utils.py:65."""
        return Tensor(np.flip(_t(img).a, axis=-2).copy())

    @staticmethod
    def combined_non_max_suppression(boxes, scores, max_output_size_per_class, max_total_size,
                                     iou_threshold=0.5, score_threshold=float("-inf"),
                                     pad_per_class=False, clip_boxes=True, name=None):
        """Scalar restatement of TF's CPU kernel for q = num_classes = 1 [TF-internal]:
        max-heap on score (ties: lower index first, see oracle docstring), pop, test against
        the already selected boxes newest-first, suppress iff IoU > iou_threshold, stop at
        max_output_size_per_class; pad to max_total_size with zeros; clip outputs."""
        bxs, scs = _t(boxes).a, _t(scores).a
        assert bxs.shape[2] == 1 and scs.shape[2] == 1, "shim covers the RPN case (1 class)"
        B = bxs.shape[0]
        per_class, total = int(max_output_size_per_class), int(max_total_size)
        out_n = min(total, per_class) if pad_per_class else total
        nb = np.zeros((B, out_n, 4), np.float32)
        ns = np.zeros((B, out_n), np.float32)
        nc = np.zeros((B, out_n), np.float32)
        nv = np.zeros((B,), np.int32)
        thr = np.float32(iou_threshold)
        for b in np.arange(B):
            bx = bxs[b, :, 0, :]
            sc = scs[b, :, 0]
            heap = [(-float(s), int(i)) for i, s in enumerate(sc) if s > np.float32(score_threshold)]
            heapq.heapify(heap)
            selected = []
            while len(selected) < per_class and heap:
                _, i = heapq.heappop(heap)
                keep = True
                for j in reversed(selected):
                    if _iou_scalar(bx, i, j) > thr:
                        keep = False
                        break
                if keep:
                    selected.append(i)
            selected = selected[:out_n]
            n = len(selected)
            nv[b] = n
            for r, i in enumerate(selected):
                nb[b, r] = np.clip(bx[i], 0, 1) if clip_boxes else bx[i]
                ns[b, r] = sc[i]
        return Tensor(nb), Tensor(ns), Tensor(nc), Tensor(nv)


class _Huber:
    def __init__(self, reduction=None, delta=1.0):
        self.delta = np.float32(delta)

    def __call__(self, y_true, y_pred):
        """Keras Huber, Reduction.NONE, as in the pinned TF 2.0.0 (environment.yml:49-52):
        purely elementwise -- the mean over the last axis only appeared in TF >= 2.1
        [TF-internal]; reg_loss's own reduce_sum(axis=-1) (train_utils.py:178) relies on it."""
        e = np.subtract(_t(y_pred).a, _t(y_true).a)
        ae = np.abs(e)
        q = np.minimum(ae, self.delta)
        lin = ae - q
        return Tensor(np.float32(0.5) * (q * q) + self.delta * lin)


class _BCE:
    def __call__(self, y_true, y_pred):
        """Keras BinaryCrossentropy(from_logits=False), mean reduction, eps 1e-7 [TF-internal]."""
        eps = np.float32(1e-7)
        p = np.clip(_t(y_pred).a, eps, np.float32(1) - eps)
        t = _t(y_true).a
        bce = -(t * np.log(p + eps) + (np.float32(1) - t) * np.log(np.float32(1) - p + eps))
        return Tensor(np.mean(bce, dtype=np.float32))


class losses:  # noqa: N801
    class Reduction:
        NONE = "none"

    Huber = _Huber
    BinaryCrossentropy = _BCE
