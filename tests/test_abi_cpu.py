"""CPU-side checks of the boundary: the shared library loads, exports exactly the symbols that
include/tfrpn.h declares, and refuses to run without a CUDA device (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tfrpn.h")).read()
    return sorted(set(re.findall(r"^TFRPN_API [\w \*]+?\b(tfrpn_\w+)\(", text, flags=re.M)))


def test_header_symbols_are_bound_and_exported():
    from tfrpn import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.PROTOTYPES) == syms
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (tfrpn_\w+)", nm)))
    assert exported == syms


def test_version_and_error_string():
    from tfrpn import _lib
    lib = _lib.load()
    assert lib.tfrpn_version() == 100
    assert isinstance(lib.tfrpn_last_error(), bytes)


def test_base_anchors_host_matches_oracle():
    """The only host-side arithmetic in the library (utils/bbox_utils.py:3-21)."""
    import numpy as np
    from oracle import rpn_oracle as O
    from tfrpn import _lib
    from tfrpn.utils import bbox_utils
    for hp in (O.get_hyper_params("vgg16"), O.get_hyper_params("mobilenet_v2"),
               dict(O.get_hyper_params("vgg16"), img_size=(800, 1333), feature_map_shape=(50, 84))):
        cfg = bbox_utils._anchor_cfg(hp)
        host = (C.c_float * 36)()
        assert _lib.load().tfrpn_base_anchors_host(C.byref(cfg), host) == 0
        got = np.asarray(list(host), np.float32).reshape(9, 4)
        assert np.array_equal(got.view(np.uint32), O.generate_base_anchors(hp).view(np.uint32))


def test_bad_arguments_are_reported():
    from tfrpn import _lib
    lib = _lib.load()
    assert lib.tfrpn_base_anchors_host(None, None) == -1
    assert b"null" in lib.tfrpn_last_error()
    with pytest.raises(ValueError):
        _lib.check(lib.tfrpn_iou_map(None, 0, None, 1, 1, 1, None, None))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tfrpn import _lib
    from tfrpn.utils import bbox_utils, train_utils
    out = C.c_void_p()
    assert _lib.load().tfrpn_create(C.byref(out), -1) == -3      # TFRPN_ERR_CUDA
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bbox_utils.generate_anchors(train_utils.get_hyper_params("vgg16"))
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bbox_utils.generate_iou_map(np.zeros((4, 4), np.float32), np.zeros((1, 2, 4), np.float32))


def test_hyper_params_quirks_match_reference():
    """utils/train_utils.py:20-38: only existing keys with truthy values are overridden."""
    from tfrpn.utils import train_utils
    hp = dict(train_utils.get_hyper_params("vgg16", total_pos_bboxes=64, bogus=3, total_neg_bboxes=0))
    assert hp["total_pos_bboxes"] == 64 and "bogus" not in hp and hp["total_neg_bboxes"] == 128
    assert hp["anchor_count"] == 9 and hp["test_nms_topn"] == 300
    train_utils.get_hyper_params("vgg16", total_pos_bboxes=128)
    assert train_utils.get_step_size(10, 4) == 3


def test_expand_targets_host_matches_numpy_scatter():
    """tfrpn_expand_targets_host is host-only code: the dense bbox_deltas from its compact form, both from
    scratch and incrementally on top of the previous step's rows."""
    import numpy as np
    from tfrpn import _lib
    lib = _lib.load()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rng = np.random.default_rng(3)
    B, N = 5, 700
    deltas = rng.normal(size=(B, N, 4)).astype(np.float32)        # garbage: the first call must zero it
    prev, prev_tp = None, 0
    for step, TP in enumerate([16, 16, 40, 8]):
        idx = np.full((B, TP), -1, np.int32)
        rows = rng.normal(size=(B, TP, 4)).astype(np.float32)
        want = np.zeros((B, N, 4), np.float32)
        for b in range(B):
            k = int(rng.integers(0, TP + 1))
            idx[b, :k] = rng.choice(N, k, replace=False)
            want[b, idx[b, :k]] = rows[b, :k]
        rc = lib.tfrpn_expand_targets_host(vp(idx), vp(rows), B, N, TP, vp(prev) if prev is not None else None, prev_tp,
                                           vp(deltas))
        assert rc == 0 and np.array_equal(deltas, want), step
        prev, prev_tp = idx, TP
    assert lib.tfrpn_expand_targets_host(None, None, 1, 1, 1, None, 0, None) == -1


def test_expand_labels_host_matches_numpy():
    """tfrpn_expand_labels_host (the host side of tfrpn_rpn_targets_sparse): codes 2 * anchor + label -> dense labels,
    from scratch (every other entry -1) and incrementally (only the previous step's entries are reset)"""
    import numpy as np
    from tfrpn import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    B, N, Q = 3, 500, 40
    def codes_for(seed):
        r = np.random.default_rng(seed)
        c = np.full((B, Q), -1, np.int32)
        for b in range(B):
            n = int(r.integers(0, Q + 1))
            idx = r.choice(N, size=n, replace=False)
            c[b, :n] = 2 * idx + r.integers(0, 2, size=n)
        return c
    def dense(c):
        out = np.full((B, N), -1.0, np.float32)
        for b in range(B):
            for v in c[b][c[b] >= 0]:
                out[b, v >> 1] = float(v & 1)
        return out
    c1, c2 = codes_for(1), codes_for(2)
    labels = rng.normal(size=(B, N)).astype(np.float32)          # garbage: a full rebuild must not depend on it
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert lib.tfrpn_expand_labels_host(vp(c1), B, N, Q, None, 0, vp(labels)) == 0
    assert np.array_equal(labels, dense(c1))
    assert lib.tfrpn_expand_labels_host(vp(c2), B, N, Q, vp(c1), Q, vp(labels)) == 0
    assert np.array_equal(labels, dense(c2))
    assert lib.tfrpn_expand_labels_host(None, B, N, Q, None, 0, vp(labels)) != 0


def test_ctypes_structs_match_header_layout(tmp_path):
    """Every struct that crosses the C ABI: size and field offsets of the ctypes mirror (tfrpn/_lib.py) equal
    what a C compiler makes of include/tfrpn.h (the header is plain C: gcc compiles it without CUDA)."""
    from tfrpn import _lib
    pairs = {"tfrpn_anchor_cfg": _lib.AnchorCfg, "tfrpn_target_cfg": _lib.TargetCfg,
             "tfrpn_target_debug": _lib.TargetDebug, "tfrpn_nms_cfg": _lib.NmsCfg,
             "tfrpn_proposal_cfg": _lib.ProposalCfg, "tfrpn_loss_out": _lib.LossOut,
             "tfrpn_step_buffers": _lib.StepBuffers}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tfrpn.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out.splitlines()}
    for cname, cls in pairs.items():
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
