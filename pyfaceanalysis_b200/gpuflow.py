"""GPU flow object: drop-in for the ``networks[i]`` entries of the reference.

The reference calls ``networks[i].execute(subimages_arr, benchmark=benchmark)``
(``FaceDetectUpdated.py:699``, ``face_analysis.py:1064,1257``) on an ``mdp.Flow`` patched by cuicuilco to
accept ``benchmark=``; ``x`` is a C-contiguous ``float64 (N, input_dim)`` array and the result a new
``float64 (N, F)`` array the caller slices and boolean-indexes.  ``GpuFlow`` is
built from the unpickled flow object and exposes the same call; iterating it yields node facades with
``execute(x)``.

No CPU path exists here: every ``execute`` goes through ``libhgsfa.so``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, plan


class GpuFlow(object):
    def __init__(self, flow, device=0, input_dim=None, igsfa_mode="auto"):
        self._flow_obj = flow
        self.device = int(device)
        self.spec = plan.compile_flow(flow, input_dim=input_dim, igsfa_mode=igsfa_mode)
        self._blob = plan.serialize(self.spec)
        self._handle = C.c_void_p()
        lib = _lib.load()
        _lib.check(lib.hgsfa_plan_create(self._blob, len(self._blob), self.device, C.byref(self._handle)))
        self.input_dim = self.spec.input_dim
        self.output_dim = self.spec.output_dim
        self._node_cache = {}
        # layers 0-2 fused into one lane-resident kernel for uint8 inputs (front.py); why not, otherwise
        self.fused_front = getattr(self.spec, "front", None) is not None
        self.front_reason = getattr(self.spec, "front_reason", "")

    # ---- mdp.Flow surface -------------------------------------------------------------------
    @property
    def flow(self):
        return [self[i] for i in range(len(self))]

    def __len__(self):
        return len(plan.flow_nodes(self._flow_obj))

    def __getitem__(self, i):
        nodes = plan.flow_nodes(self._flow_obj)
        if isinstance(i, slice):
            return GpuFlow(nodes[i], device=self.device)
        if i < 0:
            i += len(nodes)
        if i not in self._node_cache:
            self._node_cache[i] = GpuNode(self, i)
        return self._node_cache[i]

    def __iter__(self):
        return iter(self.flow)

    def __call__(self, x, nodenr=None):
        return self.execute(x, nodenr=nodenr)

    def execute(self, x, benchmark=None, nodenr=None, out_dtype=np.float64, n_features=None, out=None):
        """``flow.execute(x, benchmark=...)``.

        x : (N, input_dim) array of uint8 / float32 / float64 (anything else is converted to float64).
        Returns a new (N, F) ``out_dtype`` array (float64 by default, like the reference);
        ``n_features`` keeps only the first features (the caller's ``sl[:, 0:D]``).
        ``benchmark`` is accepted for call compatibility; per-stage times are reported by ``stats()``.
        ``out`` (optional) is a preallocated C-contiguous (N, F) float32/float64 array to fill, e.g. in
        pinned memory, instead of allocating a fresh result like the reference does.
        """
        if nodenr is not None and nodenr != len(self) - 1:
            return self[:nodenr + 1].execute(x, out_dtype=out_dtype, n_features=n_features)
        x = np.asarray(x)
        if x.ndim != 2:
            raise ValueError("x has rank %d, should be 2" % x.ndim)
        if x.shape[1] != self.input_dim:
            # MDP's _pre_execution_checks
            raise ValueError("x has dimension %d, should be %d" % (x.shape[1], self.input_dim))
        if x.dtype not in (np.uint8, np.float32, np.float64):
            x = x.astype(np.float64)
        if x.strides[1] != x.itemsize or x.strides[0] % x.itemsize:
            x = np.ascontiguousarray(x)
        ld = x.strides[0] // x.itemsize if x.shape[0] > 1 else x.shape[1]
        f = self.output_dim if n_features is None else int(n_features)
        if out is not None:
            if out.shape != (x.shape[0], f) or not out.flags.c_contiguous or out.dtype not in (np.float32, np.float64):
                raise ValueError("out must be a C-contiguous (%d, %d) float32/float64 array" % (x.shape[0], f))
            y = out
        else:
            y = np.empty((x.shape[0], f), dtype=out_dtype)
        if x.shape[0] == 0:
            return y
        lib = _lib.load()
        _lib.check(lib.hgsfa_plan_execute(self._handle, _lib.ptr(x), _lib.dtype_code(x.dtype), x.shape[0], ld,
                                          _lib.ptr(y), _lib.dtype_code(y.dtype), f, None))
        return y

    # ---- device-resident fast path ----------------------------------------------------------
    def execute_device(self, d_x, n, x_dtype, layout=_lib.ROWMAJOR, ld=None, d_y=None, y_dtype=np.float32,
                       n_features=None, stream=None):
        """Raw-pointer variant: ``d_x`` / ``d_y`` are device addresses (ints) on this flow's device."""
        f = self.output_dim if n_features is None else int(n_features)
        lib = _lib.load()
        _lib.check(lib.hgsfa_plan_execute_device(self._handle, C.c_void_p(d_x), _lib.dtype_code(x_dtype), layout,
                                                 int(n), int(ld if ld is not None else self.input_dim),
                                                 C.c_void_p(d_y), _lib.dtype_code(y_dtype), f,
                                                 C.c_void_p(stream) if stream else None))

    def execute_torch(self, x, layout=_lib.ROWMAJOR, n=None, n_features=None, out=None, out_dtype=None):
        """``x``: CUDA torch tensor, (N, input_dim) row-major uint8/float32/float64, or a TILED
        uint8/float32 buffer with ``n`` given.  Returns a CUDA float32 (N, F) tensor; runs on torch's
        current stream."""
        import torch
        if x.device.type != "cuda" or x.device.index != self.device:
            raise ValueError("tensor lives on %s, flow on cuda:%d" % (x.device, self.device))
        if layout == _lib.ROWMAJOR:
            if x.dim() != 2 or x.shape[1] != self.input_dim:
                raise ValueError("x has dimension %s, should be %d" % (tuple(x.shape[1:]), self.input_dim))
            n = x.shape[0]
            if x.stride(1) != 1:
                x = x.contiguous()
            ld = x.stride(0) if n > 1 else self.input_dim
        else:
            if n is None:
                raise ValueError("tiled input needs n")
            ld = self.input_dim
        np_dtype = {torch.uint8: np.uint8, torch.float32: np.float32, torch.float64: np.float64}[x.dtype]
        f = self.output_dim if n_features is None else int(n_features)
        if out is None:
            out = torch.empty((n, f), dtype=out_dtype or torch.float32, device=x.device)
        y_np = {torch.float32: np.float32, torch.float64: np.float64}[out.dtype]
        self.execute_device(x.data_ptr(), n, np_dtype, layout, ld, out.data_ptr(), y_np, f,
                            stream=torch.cuda.current_stream(x.device).cuda_stream)
        return out

    # ---- accounting -------------------------------------------------------------------------
    def flops(self, n, x_dtype=np.uint8):
        a, e, b = C.c_double(), C.c_double(), C.c_double()
        _lib.check(_lib.load().hgsfa_plan_flops(self._handle, int(n), _lib.dtype_code(x_dtype), C.byref(a),
                                                C.byref(e), C.byref(b)))
        return dict(algorithmic=a.value, executed=e.value, min_bytes=b.value)

    def stats(self):
        launches, ms = C.c_int64(), C.c_double()
        _lib.check(_lib.load().hgsfa_plan_stats(self._handle, C.byref(launches), C.byref(ms)))
        return dict(launches=launches.value, last_ms=ms.value)

    def profile(self, enable=True):
        """Bracket every layer launch of the following executes with CUDA events (see ``op_stats``)."""
        _lib.check(_lib.load().hgsfa_plan_profile(self._handle, int(bool(enable))))

    def op_stats(self):
        """Per-op totals since ``profile(True)``: list of dict(ms, engine, alg_flops, exe_flops) (flops per window)."""
        n = len(self.spec.ops)
        ms, alg, exe = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        eng = (C.c_int32 * n)()
        _lib.check(_lib.load().hgsfa_plan_op_stats(self._handle, n, ms, C.cast(eng, C.c_void_p), alg, exe))
        # "front": layers 0-2 fused in one kernel (time booked on op 0); "f16": single-layer FP16-split tcgen05 kernel
        names = {0: "ffma", 1: "tc", 2: "front", 3: "f16"}
        return [dict(ms=ms[i], engine=names.get(eng[i], "?"), alg_flops=alg[i], exe_flops=exe[i]) for i in range(n)]

    def set_chunks(self, front=0, back=0):
        _lib.check(_lib.load().hgsfa_plan_set_chunks(self._handle, int(front), int(back)))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _lib.load().hgsfa_plan_destroy(self._handle)
            self._handle = C.c_void_p()
        for nd in getattr(self, "_node_cache", {}).values():
            nd.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuNode(object):
    """``flow[i]``: one node of the flow with ``execute(x)`` (compiled lazily as a one-node plan)."""

    def __init__(self, parent, index):
        self._parent = parent
        self.index = index
        self.node = plan.flow_nodes(parent._flow_obj)[index]
        self._gpu = None

    @property
    def input_dim(self):
        return plan.node_input_dim(self.node)

    def execute(self, x, **kw):
        if self._gpu is None:
            self._gpu = GpuFlow([self.node], device=self._parent.device, input_dim=np.asarray(x).shape[1])
        return self._gpu.execute(x, **kw)

    __call__ = execute

    def close(self):
        if self._gpu is not None:
            self._gpu.close()
            self._gpu = None

    def __repr__(self):
        return "<GpuNode %d %s>" % (self.index, type(self.node).__name__)
