"""ctypes binding of ``libflexdm_mfp.so`` (the C ABI in ``include/flexdm_mfp.h``).

PyTorch is used for device memory and streams only: every tensor handed to the library is a raw device pointer,
every call runs on ``torch.cuda.current_stream()``.  There is no CPU fallback: if the library is missing or a
call fails this module raises.
"""
import ctypes
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .spec import get_attribute_groups, get_dataset_name, get_valid_input_columns

MAX_FIELDS = 16
NAME_LEN = 96
SORT_KEYS = ["type", "left", "top", "width", "height"]  # tensor_utils.py:11

# FLEXDM_MFP_LIB: development aid for A/B runs of two builds of the same extension inside one GPU call (file name next to this module)
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.environ.get("FLEXDM_MFP_LIB", "libflexdm_mfp.so"))
_lib = None


MAX_CANVAS = 8


class FieldDesc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 48), ("kind", ctypes.c_int32), ("C", ctypes.c_int32), ("input_dim", ctypes.c_int32),
                ("task_id", ctypes.c_int32), ("has_cond", ctypes.c_int32), ("reserved", ctypes.c_int32), ("cond_mask", ctypes.c_uint64)]


class Config(ctypes.Structure):
    _fields_ = [("num_fields", ctypes.c_int32), ("type_field", ctypes.c_int32), ("latent_dim", ctypes.c_int32), ("num_blocks", ctypes.c_int32),
                ("sort_pos", ctypes.c_int32), ("pos_task_id", ctypes.c_int32), ("total_columns", ctypes.c_int32),
                ("sort_fields", ctypes.c_int32 * 5), ("dropout", ctypes.c_float), ("l2", ctypes.c_float), ("input_dtype", ctypes.c_int32),
                ("length_input_dim", ctypes.c_int32), ("block_type", ctypes.c_int32), ("context", ctypes.c_int32), ("context_rows", ctypes.c_int32),
                ("n_canvas", ctypes.c_int32), ("canvas_input_dim", ctypes.c_int32 * MAX_CANVAS), ("canvas_names", (ctypes.c_char * 48) * MAX_CANVAS)]


class Variable(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * NAME_LEN), ("offset", ctypes.c_int64), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
                ("ld", ctypes.c_int32), ("l2", ctypes.c_int32)]


class Batch(ctypes.Structure):
    _fields_ = [("length", ctypes.c_void_p), ("cols", ctypes.c_void_p * MAX_FIELDS)]


_PTR_ARRAY = ctypes.c_void_p * MAX_FIELDS
GATHER_MAX_COLUMNS = 24


class GatherDesc(ctypes.Structure):
    _fields_ = [("n_columns", ctypes.c_int32), ("words", ctypes.c_int32 * GATHER_MAX_COLUMNS), ("pad_word", ctypes.c_uint32 * GATHER_MAX_COLUMNS),
                ("src", ctypes.c_void_p * GATHER_MAX_COLUMNS), ("dst", ctypes.c_void_p * GATHER_MAX_COLUMNS)]


_SIGNATURES = {
    "mfp_last_error": (ctypes.c_char_p, []),
    "mfp_version": (ctypes.c_int, []),
    "mfp_create": (ctypes.c_int, [ctypes.POINTER(Config), ctypes.POINTER(FieldDesc), ctypes.POINTER(ctypes.c_void_p)]),
    "mfp_destroy": (None, [ctypes.c_void_p]),
    "mfp_param_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "mfp_num_variables": (ctypes.c_int32, [ctypes.c_void_p]),
    "mfp_get_variable": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(Variable)]),
    "mfp_logit_width": (ctypes.c_int32, [ctypes.c_void_p]),
    "mfp_field_logit_offset": (ctypes.c_int32, [ctypes.c_void_p, ctypes.c_int32]),
    "mfp_workspace_bytes": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]),
    "mfp_bind": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mfp_sample_tasks": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "mfp_mask_corrupt": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Batch), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    "mfp_shuffle_inputs": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Batch), ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "mfp_mask_for_test": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Batch), ctypes.POINTER(ctypes.c_void_p),
                                         ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    "mfp_set_context_ids": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "mfp_set_canvas_columns": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "mfp_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Batch), ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                   ctypes.c_void_p]),
    "mfp_loss": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Batch), ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_float, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "mfp_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Batch), ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]),
    "mfp_backward_num_stages": (ctypes.c_int32, [ctypes.c_void_p]),
    "mfp_backward_stage_range": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "mfp_backward_stages": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Batch), ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_void_p]),
    "mfp_optimizer_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "mfp_regularization_loss": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "mfp_merge_prediction": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p]),
    "mfp_gather_documents": (ctypes.c_int, [ctypes.POINTER(GatherDesc), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_void_p]),
    "mfp_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "mfp_set_gemm_impl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "mfp_set_deterministic": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "mfp_allreduce_gradients_nvls": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                    ctypes.c_uint32, ctypes.c_void_p]),
    "mfp_set_packed_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "mfp_set_doc_offset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "mfp_profile_begin": (ctypes.c_int, [ctypes.c_void_p]),
    "mfp_profile_end": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double)]),
    "mfp_debug_attention": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_int32, ctypes.c_void_p]),
    "mfp_debug_attention_bwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "mfp_debug_gemm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                      ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_void_p]),
}


def exported_symbols() -> List[str]:
    return list(_SIGNATURES.keys())


def load_library():
    """Load the CUDA extension; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError("flex_dm_b200: %s is missing -- build the CUDA extension first (there is no CPU fallback)" % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def _check(lib, rc, what):
    if rc != 0:
        raise EngineError("%s failed (%d): %s" % (what, rc, lib.mfp_last_error().decode("utf-8", "replace")))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class Engine:
    """One ``mfp_engine`` handle plus the device buffers it is bound to."""

    def __init__(self, input_columns: Dict, num_blocks: int = 4, latent_dim: int = 256, dropout: float = 0.1, l2: Optional[float] = 1e-2,
                 device: Optional[torch.device] = None, block_type: str = "deepsvg", input_dtype: str = "set", context: Optional[str] = None):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("flex_dm_b200: a CUDA device is required (there is no CPU fallback)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.all_columns = input_columns
        self.columns = get_valid_input_columns(input_columns)
        self.keys = list(self.columns.keys())
        if len(self.keys) > MAX_FIELDS:
            raise ValueError("too many sequence columns")
        from .masking import get_task_names  # local import: masking imports engine types

        self.task_names = get_task_names(input_columns)
        groups = get_attribute_groups(input_columns.keys())
        task_of = {}
        for gi, (gname, members) in enumerate(groups.items()):
            for m in members:
                task_of[m] = 2 + gi
        fields = (FieldDesc * len(self.keys))()
        for i, key in enumerate(self.keys):
            c = self.columns[key]
            fields[i].name = key.encode()
            fields[i].kind = 0 if c["type"] == "categorical" else 1
            fields[i].C = int(c["shape"][-1])
            fields[i].input_dim = int(c.get("input_dim", 0))
            fields[i].task_id = task_of.get(key, -1)
            cond = c.get("loss_condition")
            fields[i].has_cond = 1 if cond else 0
            mask = 0
            if cond:
                if cond["key"] != "type":
                    raise ValueError("loss_condition key must be 'type'")
                for j, flag in enumerate(cond["mask"]):
                    if flag:
                        mask |= 1 << j
            fields[i].cond_mask = mask
        cfg = Config()
        cfg.num_fields = len(self.keys)
        cfg.type_field = self.keys.index("type")
        cfg.latent_dim = latent_dim
        cfg.num_blocks = num_blocks
        cfg.sort_pos = 1 if get_dataset_name(input_columns.keys()) == "rico" else 0
        cfg.pos_task_id = self.task_names.index("pos")
        cfg.total_columns = len(input_columns)
        for i, k in enumerate(SORT_KEYS):
            cfg.sort_fields[i] = self.keys.index(k)
        cfg.dropout = float(dropout)
        cfg.l2 = -1.0 if l2 is None else float(l2)
        cfg.block_type = {"deepsvg": 0, "transformer": 1}[block_type]  # transformer.py:232-236
        cfg.input_dtype = {"set": 0, "shuffled_set": 1, "sorted_set": 2}[input_dtype]  # mfp.py:104-105, encoder.py:41
        cfg.length_input_dim = int(input_columns["length"]["input_dim"])
        cfg.context = {None: 0, "id": 1, "length": 2, "canvas": 3, "canvas_add": 4}[context]  # encoder.py:11,96-110,228-249
        cfg.context_rows = {"id": len(self.task_names), "length": int(input_columns["length"]["input_dim"])}.get(context, 0)
        # canvas contexts: the non-sequence columns of get_valid_input_columns(input_columns, use_canvas=True) (encoder.py:34-37)
        self.canvas_keys = [k for k, c in get_valid_input_columns(input_columns, True).items() if not c["is_sequence"]] if context in ("canvas", "canvas_add") else []
        if context in ("canvas", "canvas_add"):
            assert len(self.canvas_keys) > 0, (self.keys, self.canvas_keys)  # encoder.py:205-206
            if len(self.canvas_keys) > MAX_CANVAS:
                raise ValueError("too many canvas columns")
            cfg.n_canvas = len(self.canvas_keys)
            for i, k in enumerate(self.canvas_keys):
                if input_columns[k]["type"] != "categorical" or tuple(input_columns[k]["shape"]) != (1,):
                    raise NotImplementedError("canvas column %s: only categorical columns of shape (1,) are supported" % k)
                cfg.canvas_input_dim[i] = int(input_columns[k]["input_dim"])
                cfg.canvas_names[i].value = k.encode()
        self.context = context
        self._context_ids = None  # keeps the tensors alive while the engine holds their pointers
        self._canvas_cols = None
        self.input_dtype = input_dtype
        self.cfg = cfg
        handle = ctypes.c_void_p()
        _check(self.lib, self.lib.mfp_create(ctypes.byref(cfg), fields, ctypes.byref(handle)), "mfp_create")
        self.handle = handle
        self.num_fields = len(self.keys)
        self.param_count = int(self.lib.mfp_param_count(handle))
        self.logit_width = int(self.lib.mfp_logit_width(handle))
        self.logit_offsets = [int(self.lib.mfp_field_logit_offset(handle, f)) for f in range(self.num_fields)]
        self.variables = OrderedDict()
        v = Variable()
        for i in range(int(self.lib.mfp_num_variables(handle))):
            _check(self.lib, self.lib.mfp_get_variable(handle, i, ctypes.byref(v)), "mfp_get_variable")
            self.variables[v.name.decode()] = (int(v.offset), int(v.rows), int(v.cols), int(v.ld), bool(v.l2))
        with torch.cuda.device(self.device):
            self.params = torch.zeros(self.param_count, dtype=torch.float32, device=self.device)
            self.grads = torch.zeros_like(self.params)
            self.adam_m = torch.zeros_like(self.params)
            self.adam_v = torch.zeros_like(self.params)
        self.B = self.S = 0
        self.workspace = None
        self.metrics_width = 3 * self.num_fields + 2  # per field loss/score_num/score_den, data loss, l2 loss

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mfp_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def variable_view(self, name: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, rows, cols, ld, _ = self.variables[name]
        buf = self.params if buf is None else buf
        return buf.as_strided((rows, cols), (ld, 1), off)

    def variable_shape(self, name: str):
        """Shape of the reference variable (1-D for biases / LayerNorm parameters)."""
        _, rows, cols, _, _ = self.variables[name]
        last = name.rsplit("/", 1)[-1]
        return (cols,) if last in ("bias", "gamma", "beta") else (rows, cols)

    def set_weights(self, weights: Dict[str, np.ndarray]):
        for name in self.variables:
            if name not in weights:
                raise KeyError("missing variable %s" % name)
            w = torch.as_tensor(np.asarray(weights[name], dtype=np.float32)).reshape(self.variable_view(name).shape)
            self.variable_view(name).copy_(w.to(self.device))

    def get_weights(self, buf: Optional[torch.Tensor] = None) -> "OrderedDict[str, np.ndarray]":
        out = OrderedDict()
        for name in self.variables:
            out[name] = self.variable_view(name, buf).detach().cpu().numpy().reshape(self.variable_shape(name)).copy()
        return out

    # ------------------------------------------------------------------ shape binding
    def bind(self, B: int, S: int):
        if (B, S) == (self.B, self.S) and self.workspace is not None:
            return
        nbytes = int(self.lib.mfp_workspace_bytes(self.handle, B, S))
        if nbytes <= 0:
            raise EngineError("mfp_workspace_bytes failed for B=%d S=%d" % (B, S))
        with torch.cuda.device(self.device):
            self.workspace = None
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            self._ws_ptr = (base + 255) // 256 * 256
            T = B * S
            self.modified = []
            self.masks = []
            for key in self.keys:
                c = self.columns[key]
                dt = torch.int32 if c["type"] == "categorical" else torch.float32
                self.modified.append(torch.zeros((B, S, int(c["shape"][-1])), dtype=dt, device=self.device))
                self.masks.append(torch.zeros((B, S), dtype=torch.uint8, device=self.device))
            self.tasks = torch.zeros((B,), dtype=torch.int32, device=self.device)
            self.logits = None
            # shuffled copies of the sequence columns (--input_dtype shuffled_set): targets and corruption source of the step
            self.shuffled = [torch.zeros_like(t) for t in self.modified] if self.input_dtype != "set" else None
        _check(self.lib, self.lib.mfp_bind(self.handle, B, S, ctypes.c_void_p(self._ws_ptr), nbytes, _ptr(self.params), _ptr(self.grads),
                                           _ptr(self.adam_m), _ptr(self.adam_v)), "mfp_bind")
        self.B, self.S = B, S
        self._mod_ptrs = _PTR_ARRAY(*[t.data_ptr() for t in self.modified])
        self._mask_ptrs = _PTR_ARRAY(*[t.data_ptr() for t in self.masks])

    def _batch(self, length: torch.Tensor, cols: List[torch.Tensor]) -> Batch:
        b = Batch()
        b.length = length.data_ptr()
        for i, t in enumerate(cols):
            b.cols[i] = t.data_ptr()
        return b

    def modified_batch(self, length: torch.Tensor) -> Batch:
        return self._batch(length, self.modified)

    # ------------------------------------------------------------------ calls
    def sample_tasks(self, allowed: List[int], seed: int, step: int):
        arr = (ctypes.c_int32 * len(allowed))(*allowed)
        _check(self.lib, self.lib.mfp_sample_tasks(self.handle, arr, len(allowed), seed & 0xFFFFFFFF, step & 0xFFFFFFFF, _ptr(self.tasks), _stream()),
               "mfp_sample_tasks")
        return self.tasks

    def mask_corrupt(self, length, cols, tasks, seed: int, step: int):
        b = self._batch(length, cols)
        _check(self.lib, self.lib.mfp_mask_corrupt(self.handle, ctypes.byref(b), _ptr(tasks), seed & 0xFFFFFFFF, step & 0xFFFFFFFF,
                                                   ctypes.cast(self._mod_ptrs, ctypes.POINTER(ctypes.c_void_p)),
                                                   ctypes.cast(self._mask_ptrs, ctypes.POINTER(ctypes.c_void_p)), _stream()), "mfp_mask_corrupt")

    def shuffle_inputs(self, length, cols, seed: int, step: int, perm_out: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """shuffle_inputs (tensor_utils.py:47-76): returns the shuffled sequence columns (engine-owned buffers)."""
        b = self._batch(length, cols)
        sp = _PTR_ARRAY(*[t.data_ptr() for t in self.shuffled])
        _check(self.lib, self.lib.mfp_shuffle_inputs(self.handle, ctypes.byref(b), seed & 0xFFFFFFFF, step & 0xFFFFFFFF,
                                                     ctypes.cast(sp, ctypes.POINTER(ctypes.c_void_p)), _ptr(perm_out), _stream()), "mfp_shuffle_inputs")
        return self.shuffled

    def mask_for_test(self, length, cols, masks: List[torch.Tensor]):
        b = self._batch(length, cols)
        mp = _PTR_ARRAY(*[m.data_ptr() for m in masks])
        _check(self.lib, self.lib.mfp_mask_for_test(self.handle, ctypes.byref(b), ctypes.cast(mp, ctypes.POINTER(ctypes.c_void_p)),
                                                    ctypes.cast(self._mod_ptrs, ctypes.POINTER(ctypes.c_void_p)), _stream()), "mfp_mask_for_test")

    def set_context_ids(self, task_ids: torch.Tensor):
        """--context id: the per-document task ids the context token embeds (mfp.py:137; eval.py:100-101)."""
        ids = task_ids.to(device=self.device, dtype=torch.int32).reshape(-1).contiguous()
        if ids.numel() != self.B:
            raise ValueError("expected %d task ids, got %d" % (self.B, ids.numel()))
        self._context_ids = ids
        _check(self.lib, self.lib.mfp_set_context_ids(self.handle, _ptr(ids)), "mfp_set_context_ids")

    def set_canvas_columns(self, columns: List[torch.Tensor]):
        """--context canvas / canvas_add: the batch's canvas columns ``(B, 1)`` int32, in ``canvas_keys`` order."""
        cols = [c.to(device=self.device, dtype=torch.int32).reshape(-1).contiguous() for c in columns]
        if len(cols) != len(self.canvas_keys) or any(c.numel() != self.B for c in cols):
            raise ValueError("expected %d canvas columns of %d documents" % (len(self.canvas_keys), self.B))
        self._canvas_cols = cols
        ptrs = (ctypes.c_void_p * MAX_CANVAS)(*[c.data_ptr() for c in cols])
        _check(self.lib, self.lib.mfp_set_canvas_columns(self.handle, ptrs), "mfp_set_canvas_columns")

    def forward(self, length, cols: Optional[List[torch.Tensor]] = None, training: bool = False, seed: int = 0, step: int = 0,
                logits_out: Optional[torch.Tensor] = None):
        b = self._batch(length, self.modified if cols is None else cols)
        _check(self.lib, self.lib.mfp_forward(self.handle, ctypes.byref(b), 1 if training else 0, seed & 0xFFFFFFFF, step & 0xFFFFFFFF,
                                              _ptr(logits_out), _stream()), "mfp_forward")

    def loss(self, length, target_cols, masks, metrics_out, inv_batch: float, compute_grad: bool, sort_flag=None, sort_tasks=None,
             logits_in=None):
        b = self._batch(length, target_cols)
        mp = _PTR_ARRAY(*[m.data_ptr() for m in masks])
        _check(self.lib, self.lib.mfp_loss(self.handle, ctypes.byref(b), ctypes.cast(mp, ctypes.POINTER(ctypes.c_void_p)), _ptr(sort_flag),
                                           _ptr(sort_tasks), _ptr(logits_in), float(inv_batch), 1 if compute_grad else 0, _ptr(metrics_out),
                                           _stream()), "mfp_loss")

    def backward(self, length, cols: Optional[List[torch.Tensor]] = None, training: bool = True, seed: int = 0, step: int = 0):
        b = self._batch(length, self.modified if cols is None else cols)
        _check(self.lib, self.lib.mfp_backward(self.handle, ctypes.byref(b), 1 if training else 0, seed & 0xFFFFFFFF, step & 0xFFFFFFFF, _stream()),
               "mfp_backward")

    def backward_stages(self, length, first_stage: int, last_stage: int, cols: Optional[List[torch.Tensor]] = None, training: bool = True, seed: int = 0,
                        step: int = 0):
        """Stages ``first_stage..last_stage`` of the backward pass (0 = heads, 1..L = blocks L-1..0, L+1 = encoder), ascending, once per step."""
        b = self._batch(length, self.modified if cols is None else cols)
        _check(self.lib, self.lib.mfp_backward_stages(self.handle, ctypes.byref(b), 1 if training else 0, seed & 0xFFFFFFFF, step & 0xFFFFFFFF,
                                                      int(first_stage), int(last_stage), _stream()), "mfp_backward_stages")

    def backward_stage_ranges(self) -> List[Tuple[int, int]]:
        """``[lo, hi)`` of the flat gradient buffer that is final after each backward stage."""
        out = []
        for s in range(int(self.lib.mfp_backward_num_stages(self.handle))):
            lo, hi = ctypes.c_int64(), ctypes.c_int64()
            _check(self.lib, self.lib.mfp_backward_stage_range(self.handle, s, ctypes.byref(lo), ctypes.byref(hi)), "mfp_backward_stage_range")
            out.append((int(lo.value), int(hi.value)))
        return out

    def optimizer_step(self, t: int, learning_rate: float, clipnorm: Optional[float], l2_out: Optional[torch.Tensor] = None):
        _check(self.lib, self.lib.mfp_optimizer_step(self.handle, int(t), float(learning_rate), float(clipnorm) if clipnorm else 0.0, _ptr(l2_out),
                                                     _stream()), "mfp_optimizer_step")

    def regularization_loss(self, l2_out: torch.Tensor):
        _check(self.lib, self.lib.mfp_regularization_loss(self.handle, _ptr(l2_out), _stream()), "mfp_regularization_loss")

    def merge_prediction(self, field: int, input_col, mask, out, logits_in=None):
        _check(self.lib, self.lib.mfp_merge_prediction(self.handle, field, _ptr(input_col), _ptr(mask), _ptr(logits_in), _ptr(out), _stream()),
               "mfp_merge_prediction")

    def profile_begin(self):
        _check(self.lib, self.lib.mfp_profile_begin(self.handle), "mfp_profile_begin")

    def profile_end(self):
        """-> {"gemm": (ms, launches, algorithmic bytes), "attention": (...)}"""
        ms = (ctypes.c_float * 2)()
        n = (ctypes.c_int32 * 2)()
        nbytes = (ctypes.c_double * 2)()
        _check(self.lib, self.lib.mfp_profile_end(self.handle, ms, n, nbytes), "mfp_profile_end")
        return {"gemm": (float(ms[0]), int(n[0]), float(nbytes[0])), "attention": (float(ms[1]), int(n[1]), float(nbytes[1]))}

    def set_gemm_impl(self, impl: int):
        """0 = tcgen05 TF32 (product path), 1 = fp32 SIMT bring-up GEMM (tests only), 2 = fp32-accurate 3xTF32 tensor-core GEMMs + fp32 attention."""
        _check(self.lib, self.lib.mfp_set_gemm_impl(self.handle, int(impl)), "mfp_set_gemm_impl")

    def set_packed_rows(self, rowmaps: Optional[List[Optional[torch.Tensor]]]):
        """Declares the numerical columns of the following input / target batches packed (``mfp_set_packed_rows``): ``rowmaps[f]`` is a
        device int32 ``[B*S]`` tensor (element -> row of the packed column, -1 = none) or None for a dense column; None clears."""
        if rowmaps is None or all(r is None for r in rowmaps):
            self._rowmaps = None
            _check(self.lib, self.lib.mfp_set_packed_rows(self.handle, None), "mfp_set_packed_rows")
            return
        self._rowmaps = rowmaps  # keeps the tensors alive while the engine holds their pointers
        ptrs = _PTR_ARRAY(*[(r.data_ptr() if r is not None else None) for r in rowmaps])
        _check(self.lib, self.lib.mfp_set_packed_rows(self.handle, ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_void_p))), "mfp_set_packed_rows")

    def set_gradient_buffer(self, buf: torch.Tensor):
        """Rebinds the flat gradient buffer (e.g. to a symmetric-memory tensor for the NVLS all-reduce); takes effect at the next ``bind``."""
        if buf.numel() != self.param_count or buf.dtype != torch.float32 or buf.device != self.device:
            raise ValueError("gradient buffer must be %d float32 on %s" % (self.param_count, self.device))
        self.grads = buf
        self.B = self.S = 0  # force mfp_bind with the new pointer

    def allreduce_gradients_nvls(self, multicast_ptr: int, signal_pads_dev: int, first_slot: int, rank: int, world: int, call: int):
        """``mfp_allreduce_gradients_nvls``: sum of the ranks' gradient buffers through the switch, one kernel on the current stream."""
        _check(self.lib, self.lib.mfp_allreduce_gradients_nvls(self.handle, ctypes.c_void_p(multicast_ptr), ctypes.c_void_p(signal_pads_dev), int(first_slot),
                                                               int(rank), int(world), int(call) & 0xFFFFFFFF, _stream()), "mfp_allreduce_gradients_nvls")

    def set_deterministic(self, on: bool = True):
        """Fixed-order gradient reductions: two runs of the same step are bit-identical (slower than the arrival-order default)."""
        _check(self.lib, self.lib.mfp_set_deterministic(self.handle, 1 if on else 0), "mfp_set_deterministic")

    def set_doc_offset(self, first_document: int):
        """Global index of the bound batch's first document (data-parallel shard): Philox counters follow global indices."""
        _check(self.lib, self.lib.mfp_set_doc_offset(self.handle, int(first_document)), "mfp_set_doc_offset")

    def launch_count(self) -> int:
        return int(self.lib.mfp_launch_count(self.handle))


def gather_documents(sources: List[torch.Tensor], pads: List[int], doc_start: torch.Tensor, doc_len: torch.Tensor, idx: torch.Tensor, S: int) -> List[torch.Tensor]:
    """``mfp_gather_documents``: padded ``[B, S, C]`` batch columns cut out of ragged ``[total_elements, C]`` device arrays."""
    lib = load_library()
    B = int(idx.numel())
    d = GatherDesc()
    d.n_columns = len(sources)
    outs = []
    for c, (src, pad) in enumerate(zip(sources, pads)):
        assert src.is_cuda and src.dim() == 2 and src.element_size() == 4 and src.is_contiguous()
        out = torch.empty((B, S, src.shape[1]), dtype=src.dtype, device=src.device)
        d.words[c] = int(src.shape[1])
        d.pad_word[c] = int(pad) & 0xFFFFFFFF
        d.src[c] = src.data_ptr()
        d.dst[c] = out.data_ptr()
        outs.append(out)
    _check(lib, lib.mfp_gather_documents(ctypes.byref(d), _ptr(doc_start), _ptr(doc_len), _ptr(idx), B, int(S), _stream()), "mfp_gather_documents")
    return outs


def debug_attention(qkv: torch.Tensor, length: torch.Tensor, B: int, S: int, impl: int = 0):
    """(out [B*S, 256], lse [B, 8, S]) of the attention core on a [B*S, 768] QKV activation (bring-up / unit tests)."""
    lib = load_library()
    out = torch.zeros((B * S, 256), dtype=torch.float32, device=qkv.device)
    lse = torch.zeros((B, 8, S), dtype=torch.float32, device=qkv.device)
    _check(lib, lib.mfp_debug_attention(_ptr(qkv), _ptr(length), B, S, _ptr(out), _ptr(lse), impl, _stream()), "mfp_debug_attention")
    return out, lse


def debug_attention_bwd(qkv, length, B: int, S: int, out, lse, dout, impl: int = 0):
    """dqkv [B*S, 768] of the attention core (bring-up / unit tests)."""
    lib = load_library()
    dqkv = torch.full((B * S, 768), float("nan"), dtype=torch.float32, device=qkv.device)
    _check(lib, lib.mfp_debug_attention_bwd(_ptr(qkv), _ptr(length), B, S, _ptr(out), _ptr(lse), _ptr(dout), _ptr(dqkv), impl, _stream()),
           "mfp_debug_attention_bwd")
    return dqkv


def debug_gemm(A: torch.Tensor, a_mn: bool, B: torch.Tensor, b_mn: bool, M: int, N: int, K: int, bias=None, relu=False, splits=1, impl=0,
               out: Optional[torch.Tensor] = None, residual=None, relu_src=None, colsum=None) -> torch.Tensor:
    """D[M,N] = epilogue(A . B^T) through the engine's GEMM (bring-up / unit tests); residual / relu_src share D's pitch."""
    lib = load_library()
    D = torch.zeros((M, N), dtype=torch.float32, device=A.device) if out is None else out
    for t in (residual, relu_src):
        assert t is None or t.stride(0) == D.stride(0)
    _check(lib, lib.mfp_debug_gemm(_ptr(A), int(a_mn), A.stride(0), _ptr(B), int(b_mn), B.stride(0), _ptr(D), D.stride(0), M, N, K, _ptr(bias),
                                   int(relu), _ptr(residual), _ptr(relu_src), _ptr(colsum), splits, impl, _stream()), "mfp_debug_gemm")
    return D
