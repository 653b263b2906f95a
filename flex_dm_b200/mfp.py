"""Host-side mirror of ``mfp/models/mfp.py``: the ``MFP`` model behind the Keras-style surface that
``train.py`` / ``eval.py`` / the notebooks drive:

    MFP(input_columns, num_blocks, block_type, masking_method, seq_type, arch_type, context, latent_dim,
        dropout, l2, input_dtype)                                  mfp.py:215-228, train.py:53-65
    model(inputs, training=False, demo_args=None) -> dict          mfp.py:298-347
    model.compile(optimizer=Adam(learning_rate, clipnorm=1.0))     train.py:71-77
    model.fit(...) / model.evaluate(...) / model.metrics_names     train.py:79-92
    model.load_weights(path) / model.save_weights(path)            train.py:67-69,94-97

All tensor work is done by the CUDA engine (``engine.py`` -> ``libflexdm_mfp.so``); this file is argument
handling, buffer staging and bookkeeping.  Tensors in and out are ``torch`` tensors (numpy accepted on input).
"""
import logging
import math
import os
from collections import OrderedDict
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import checkpoint
from .data import ROWS_SUFFIX, unpack_column
from .engine import Engine
from .masking import get_task_names
from .parallel import all_reduce_gradient_slice, all_reduce_gradients, broadcast_parameters, reduce_metric_rows
from .spec import get_dataset_name

logger = logging.getLogger(__name__)


class Adam:
    """``tf.keras.optimizers.Adam(learning_rate, clipnorm)`` configuration (train.py:72-75).  beta_1 = 0.9,
    beta_2 = 0.999, epsilon = 1e-7 are the Keras defaults the reference uses; the update itself is
    ``csrc/optimizer.cu``."""

    def __init__(self, learning_rate: float = 1e-3, clipnorm: Optional[float] = None):
        self.learning_rate = float(learning_rate)
        self.clipnorm = clipnorm
        self.iterations = 0


def get_task_ids(task_names: List[str], masking_method: str) -> List[int]:
    """get_task_cat_dist_sampler (mfp.py:34-43): uniform over the task names listed in ``masking_method``."""
    used_names = masking_method.split("_")
    ids = [i for i, name in enumerate(task_names) if name in used_names]
    assert len(ids) > 0
    return ids


def merge_inputs_and_prediction(inputs: Dict, input_columns: Dict, masks: Dict, prediction: Dict) -> Dict:
    """``mfp.py:46-69`` as the notebooks import it (demo_rico: applied once more to ``model(example, demo_args=...)``'s return value
    with the *full* column dict, which copies the demo-only columns -- ``id``, ``uuid`` -- into the prediction).  Inside ``MFP.__call__``
    the same merge is the engine's ``merge_prediction`` kernel; this is the host-level function for already-materialised outputs:
    prediction where ``masks[key]`` is set, ground truth (one-hot for categorical columns) elsewhere, canvas and demo-only columns copied.
    Like the reference it updates and returns ``prediction``."""
    for key, column in input_columns.items():
        if not column["is_sequence"]:
            prediction[key] = inputs[key]  # keep canvas attributes
        elif key not in masks:
            continue  # demo only attributes
        else:
            pred = prediction[key]
            mask = torch.as_tensor(masks[key]).to(device=pred.device, dtype=torch.bool)
            value = torch.as_tensor(inputs[key]).to(pred.device)
            if column["type"] == "numerical":
                prediction[key] = torch.where(mask[..., None], pred, value.to(pred.dtype))
            else:
                gt = torch.nn.functional.one_hot(value.to(torch.int64), column["input_dim"]).to(pred.dtype)
                prediction[key] = torch.where(mask[..., None, None], pred, gt)
    for key, column in input_columns.items():  # copy unpredicted items for visualization
        if column.get("demo_only", False):
            prediction[key] = inputs[key]
    return prediction


def init_weights(engine: Engine, seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Keras default initialisers: Dense glorot-uniform kernel / zero bias,
    Embedding U(-0.05, 0.05), LayerNorm gamma = 1 / beta = 0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = OrderedDict()
    for name in engine.variables:
        shape = engine.variable_shape(name)
        last = name.rsplit("/", 1)[-1]
        if last == "embeddings":
            w = rng.uniform(-0.05, 0.05, size=shape)
        elif last == "kernel":
            limit = math.sqrt(6.0 / (shape[0] + shape[1]))
            w = rng.uniform(-limit, limit, size=shape)
        elif last == "gamma":
            w = np.ones(shape)
        else:
            w = np.zeros(shape)
        out[name] = w.astype(np.float32)
    return out


class MFP:
    """MFP trainer (mfp.py:210-347) on the B200 engine."""

    def __init__(
        self,
        input_columns: Dict,
        num_blocks: int = 4,
        block_type: str = "deepsvg",
        masking_method: str = "random",
        seq_type: str = "default",
        arch_type: str = "oneshot",
        context: Optional[str] = None,
        input_dtype: str = "set",
        name: str = "mfp",
        use_elemwise_noise: bool = False,
        seed: int = 0,
        device=None,
        **kwargs,  # keys are latent_dim, dropout, l2
    ):
        assert arch_type == "oneshot"  # mfp.py:230
        if block_type not in ("deepsvg", "transformer"):  # get_seq_block, transformer.py:232-236
            raise KeyError(block_type)
        if input_dtype not in ("set", "shuffled_set", "sorted_set"):
            raise ValueError("input_dtype=%r (args.py: set | shuffled_set | sorted_set)" % (input_dtype,))
        if context not in (None, "id", "length", "canvas", "canvas_add"):  # encoder.py:11 CONTEXT_NAMES
            raise AssertionError("context=%r (encoder.py:28: one of None, 'id', 'canvas', 'length', 'canvas_add')" % (context,))
        for flag, value, supported in (("seq_type", seq_type, "default"),
                                       ("use_elemwise_noise", use_elemwise_noise, False)):
            if value != supported:
                raise NotImplementedError("%s=%r is outside the B200 hot path (SURVEY.md section 8f); supported: %r" % (flag, value, supported))
        self.name = name
        self.arch_type = arch_type
        self.context = context
        self._pad_context = True  # False: the caller guarantees length + 1 < S for every document (one free row for the context token)
        self.input_dtype = input_dtype
        self.all_columns = input_columns
        self.input_columns = OrderedDict((k, v) for (k, v) in input_columns.items() if not v.get("demo_only", False))  # mfp.py:235-237
        self.is_autoreg = False
        latent_dim = kwargs.pop("latent_dim", 256)
        dropout = kwargs.pop("dropout", 0.1)
        l2 = kwargs.pop("l2", None)
        if kwargs:
            raise TypeError("unexpected arguments: %s" % sorted(kwargs))
        self.block_type = block_type
        self.engine = Engine(input_columns, num_blocks=num_blocks, latent_dim=latent_dim, dropout=dropout, l2=l2, device=device, block_type=block_type,
                             input_dtype=input_dtype, context=context)
        self.device = self.engine.device
        self.keys = self.engine.keys
        self.task_names = get_task_names(input_columns)
        self.task_ids = get_task_ids(self.task_names, masking_method)
        self.sort_pos = get_dataset_name(input_columns.keys()) == "rico"  # mfp.py:293-296
        self.seed = int(seed)
        self.engine.set_weights(init_weights(self.engine, seed))
        self.optimizer: Optional[Adam] = None
        self._step = 0  # RNG step counter (masks / dropout), advances on every stochastic call
        self._ring = None
        self._ring_pos = 0
        self._world = 1
        self._rank = 0
        self._dist = None
        self._nvls = None
        self._overlap = False
        self.history: List[Dict[str, float]] = []
        self.stop_training = False

    # ------------------------------------------------------------------ distributed (document-sharded DP)
    def set_deterministic(self, on: bool = True):
        """Fixed-order gradient reductions (``mfp_set_deterministic``): identical runs give bit-identical weights, as the reference's
        seeding intends (train.py:18-23).  Off by default: the arrival-order reductions are faster."""
        self.engine.set_deterministic(on)

    NVLS_FIRST_SLOT = 256  # uint32 slot of the symmetric signal pads where this engine's barrier flags start (torch's own ops use the low slots)

    def enable_data_parallel(self, dist_module, world_size: int, overlap: bool = False, rank: Optional[int] = None, transport: str = "auto"):
        """Shard batches over documents; one all-reduce of the flat gradient buffer per step.
        ``overlap=True`` runs the backward in stages and starts each stage's gradient slice as soon as it is final.  Measured on
        2 x B200 it does not pay (2.713 vs 2.70 ms per step): the persistent GEMM kernels hold every SM, so the NCCL kernels only
        get in at kernel boundaries and delay the next GEMM's CTAs by what they would have cost serially; hence off by default."""
        self._overlap = bool(overlap)
        self._dist = dist_module
        self._world = int(world_size)
        # rank r holds documents [r * B_local, (r + 1) * B_local) of the global batch: the engine forms every Philox counter (task ids,
        # mask draws, dropout) from GLOBAL document indices, so the sharded step is the single-process step (train.py:25) on the
        # concatenated batch, not N copies of the same random pattern
        self._rank = int(dist_module.get_rank() if rank is None else rank)
        self._stage_ranges = self.engine.backward_stage_ranges()
        broadcast_parameters(dist_module, self.engine.params, 0)  # every rank starts from rank 0's initialisation
        # Gradient exchange: one NVLS kernel on the step's stream (csrc/allreduce.cu) when the ranks share a multicast-capable NVSwitch domain,
        # otherwise ncclAllReduce.  transport = "auto" | "nvls" | "nccl".
        self._nvls = None
        if transport not in ("auto", "nvls", "nccl"):
            raise ValueError(transport)
        if transport != "nccl" and not overlap and self.device.type == "cuda":
            self._nvls = self._setup_nvls(dist_module, required=(transport == "nvls"))

    def _setup_nvls(self, dist_module, required: bool):
        """Moves the flat gradient buffer into symmetric memory and returns what ``mfp_allreduce_gradients_nvls`` needs, or None.  Collective:
        every rank takes the same decision (a failure on any rank sends all of them to NCCL)."""
        state, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem

            group = dist_module.group.WORLD.group_name
            buf = symm_mem.empty(self.engine.param_count, dtype=torch.float32, device=self.device)
            hdl = symm_mem.rendezvous(buf, group)
            if not hdl.multicast_ptr:
                raise RuntimeError("no multicast (NVLS) support between these GPUs")
            if hdl.signal_pad_size < 4 * (self.NVLS_FIRST_SLOT + self._world) or self.engine.param_count % 4:
                raise RuntimeError("signal pad too small / buffer length not a multiple of 4")
            buf.zero_()
            mc = int(hdl.multicast_ptr) + (buf.data_ptr() - int(hdl.buffer_ptrs[hdl.rank]))
            state = {"buf": buf, "hdl": hdl, "mc": mc, "pads": int(hdl.signal_pad_ptrs_dev), "call": 0}
        except Exception as e:  # noqa: BLE001 -- whatever goes wrong, NCCL still works
            err = e
        ok = torch.tensor([0 if state is None else 1], device=self.device)
        dist_module.all_reduce(ok, op=dist_module.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if required:
                raise RuntimeError("NVLS gradient all-reduce is not available: %s" % (err,))
            logger.info("NVLS gradient all-reduce not available (%s): using ncclAllReduce", err)
            return None
        self.engine.set_gradient_buffer(state["buf"])
        return state

    # ------------------------------------------------------------------ Keras surface
    def compile(self, optimizer=None, run_eagerly=None, **_):
        if optimizer is None or isinstance(optimizer, str):
            optimizer = Adam()
        self.optimizer = optimizer

    @property
    def metrics_names(self) -> List[str]:
        names = ["loss"]
        for key in self.keys:
            names.append(key + "_score")
        for key in self.keys:
            names.append(key + "_loss")
        names.append("total_score")
        return names

    # ------------------------------------------------------------------ staging
    def stage(self, inputs: Dict, non_blocking: bool = True) -> Dict[str, torch.Tensor]:
        """Host -> device copy of the columns the path reads (pinned host tensors copy asynchronously)."""
        out = {}
        for key, column in self.input_columns.items():
            if key not in inputs:
                if key == "length" or column["is_sequence"]:
                    raise KeyError("missing input column %s" % key)
                continue
            x = inputs[key]
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(x)
            want = torch.float32 if column.get("type") == "numerical" else torch.int32
            if x.dtype != want:
                x = x.to(want)
            if x.device != self.device:
                x = x.to(self.device, non_blocking=non_blocking)
            out[key] = x.contiguous()
            if key + ROWS_SUFFIX in inputs:  # packed numerical column (flex_dm_b200.data.pack_batch): its element -> row map travels with it
                r = inputs[key + ROWS_SUFFIX]
                r = torch.from_numpy(r) if isinstance(r, np.ndarray) else r
                out[key + ROWS_SUFFIX] = r.to(device=self.device, dtype=torch.int32, non_blocking=non_blocking).contiguous()
        return out

    def host_columns(self, inputs: Dict) -> Dict[str, torch.Tensor]:
        """The columns the path reads, as host tensors of the engine's dtypes (int32 / float32), ready for an H2D copy."""
        out = {}
        for key, column in self.input_columns.items():
            if key not in inputs:
                if key == "length" or column["is_sequence"]:
                    raise KeyError("missing input column %s" % key)
                continue
            x = inputs[key]
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(x)
            want = torch.float32 if column.get("type") == "numerical" else torch.int32
            out[key] = (x if x.dtype == want else x.to(want)).contiguous()
            if key + ROWS_SUFFIX in inputs:
                r = inputs[key + ROWS_SUFFIX]
                r = torch.from_numpy(r) if isinstance(r, np.ndarray) else r
                out[key + ROWS_SUFFIX] = (r if r.dtype == torch.int32 else r.to(torch.int32)).contiguous()
        return out

    def _bind(self, staged: Dict[str, torch.Tensor], dense: bool = False):
        """Binds the engine to the batch's shape.  Packed numerical columns (``flex_dm_b200.data.pack_batch``) are handed to the engine as
        they are (``mfp_set_packed_rows``); ``dense=True`` -- or a path the engine does not take packed columns on (shuffled / sorted
        inputs) -- expands them on the device first."""
        rows = [staged.get(k + ROWS_SUFFIX) for k in self.keys]
        if any(r is not None for r in rows) and (dense or self.input_dtype != "set"):
            staged = dict(staged)
            for k, r in zip(self.keys, rows):
                if r is not None:
                    staged[k] = unpack_column(staged[k], r)
                    del staged[k + ROWS_SUFFIX]
            rows = [None] * len(self.keys)
        cols = [staged[k] for k in self.keys]
        self._staged = staged
        pad = self.context in ("id", "length", "canvas") and self._pad_context
        if pad:
            # the context token (encoder.py:231-249) takes the row after each document's last element: one more (padding) row so
            # that full-length documents have one too; callers see the caller's S again (``_crop``)
            cols = [c if r is not None else torch.nn.functional.pad(c, (0, 0, 0, 1)) for c, r in zip(cols, rows)]
            rows = [None if r is None else torch.nn.functional.pad(r, (0, 1), value=-1) for r in rows]
        first_dense = next(c for c, r in zip(cols, rows) if r is None)
        B, S = first_dense.shape[:2]
        self.engine.bind(int(B), int(S))
        self.engine.set_packed_rows([None if r is None else r.reshape(-1).contiguous() for r in rows])
        if self._world > 1:
            self.engine.set_doc_offset(self._rank * int(B))
        if self._ring is None:
            self._ring = torch.zeros((256, self.engine.metrics_width), dtype=torch.float32, device=self.device)
        return int(B), int(S), staged["length"].reshape(-1), cols

    def _set_context(self, tasks: torch.Tensor):
        if self.context == "id":  # modified_inputs["task"] (mfp.py:137) -> Encoder input_layer["task"] (encoder.py:234-237)
            self.engine.set_context_ids(tasks)
        elif self.context in ("canvas", "canvas_add"):  # the canvas columns pass through the masking untouched (masking.py:246-248)
            missing = [k for k in self.engine.canvas_keys if k not in self._staged]
            if missing:
                raise KeyError("missing canvas columns %s" % missing)
            self.engine.set_canvas_columns([self._staged[k] for k in self.engine.canvas_keys])

    def _crop(self, x: torch.Tensor) -> torch.Tensor:
        """Drops the extra row ``_bind`` added for the context token."""
        return x[:, :-1] if (self.context in ("id", "length", "canvas") and self._pad_context) else x

    def _next_row(self) -> torch.Tensor:
        row = self._ring[self._ring_pos % self._ring.shape[0]]
        self._ring_pos += 1
        return row

    # ------------------------------------------------------------------ steps
    def train_step(self, inputs: Dict, staged: bool = False) -> torch.Tensor:
        """Keras default train_step: sample tasks, corrupt, forward, loss, backward,
        [all-reduce], clip + Adam.  Returns the device row of raw metrics (see ``metrics_from_row``)."""
        if self.optimizer is None:
            raise RuntimeError("call compile(optimizer=Adam(...)) before training")
        data = inputs if staged else self.stage(inputs)
        B, S, length, cols = self._bind(data)
        eng, seed, step = self.engine, self.seed, self._step
        row = self._next_row()
        tasks = eng.sample_tasks(self.task_ids, seed, step)
        self._set_context(tasks)
        if self.input_dtype != "set":  # mfp.py:104-105: the shuffled batch is what gets corrupted and what the loss targets
            cols = eng.shuffle_inputs(length, cols, seed, step)
        eng.mask_corrupt(length, cols, tasks, seed, step)
        eng.forward(length, None, True, seed, step)
        eng.loss(length, cols, eng.masks, row, 1.0 / (B * self._world), True, sort_tasks=tasks if self.sort_pos else None)
        if self._world > 1 and self._overlap:
            # staged backward: each stage's gradient slice starts its all-reduce as soon as it is final (heads first, encoder last)
            works = []
            for s, (lo, hi) in enumerate(self._stage_ranges):
                eng.backward_stages(length, s, s, None, True, seed, step)
                works.append(all_reduce_gradient_slice(self._dist, eng.grads, lo, hi))
            for w in works:
                if w is not None:
                    w.wait()
        else:
            eng.backward(length, None, True, seed, step)
            if self._world > 1:
                if self._nvls is not None:
                    self._nvls["call"] += 1
                    eng.allreduce_gradients_nvls(self._nvls["mc"], self._nvls["pads"], self.NVLS_FIRST_SLOT, self._rank, self._world, self._nvls["call"])
                else:
                    all_reduce_gradients(self._dist, eng.grads)
        self.optimizer.iterations += 1
        eng.optimizer_step(self.optimizer.iterations, self.optimizer.learning_rate, self.optimizer.clipnorm, row[-1:])
        self._step += 1
        return row

    def test_step(self, inputs: Dict, staged: bool = False) -> torch.Tensor:
        """Keras default test_step: ``self(x, training=False)`` -> the train-time corruption without dropout or update."""
        data = inputs if staged else self.stage(inputs)
        B, S, length, cols = self._bind(data)
        eng, seed, step = self.engine, self.seed, self._step
        row = self._next_row()
        tasks = eng.sample_tasks(self.task_ids, seed, step)
        self._set_context(tasks)
        if self.input_dtype != "set":
            cols = eng.shuffle_inputs(length, cols, seed, step)
        eng.mask_corrupt(length, cols, tasks, seed, step)
        eng.forward(length, None, False, seed, step)
        eng.loss(length, cols, eng.masks, row, 1.0 / (B * self._world), False, sort_tasks=tasks if self.sort_pos else None)
        eng.regularization_loss(row[-1:])
        self._step += 1
        return row

    def metrics_from_row(self, row) -> "OrderedDict[str, float]":
        """Names and formulas of LossLayer's add_loss / add_metric (metrics.py:279-298) from one raw metrics row."""
        r = row.detach().cpu().numpy() if isinstance(row, torch.Tensor) else np.asarray(row)
        F = len(self.keys)
        out = OrderedDict()
        out["loss"] = float(r[3 * F] + r[3 * F + 1])
        total = 0.0
        for f, key in enumerate(self.keys):
            num, den = float(r[3 * f + 1]), float(r[3 * f + 2])
            score = 1.0 if den == 0.0 else num / den
            out[key + "_score"] = score
            total += score
        for f, key in enumerate(self.keys):
            out[key + "_loss"] = float(r[3 * f])
        out["total_score"] = total / len(self.all_columns)
        return out

    def _reduce_rows(self, rows: torch.Tensor) -> torch.Tensor:
        if self._world > 1:
            rows = reduce_metric_rows(self._dist, rows)
        return rows

    def _run_epoch(self, iterator, steps: int, train: bool, staged: bool = False) -> "OrderedDict[str, float]":
        # one metrics row per step of the epoch stays on the device until the epoch ends: the ring must hold them all
        if self._ring is None or steps > self._ring.shape[0]:
            self._ring = torch.zeros((max(256, steps), self.engine.metrics_width), dtype=torch.float32, device=self.device)
            self._ring_pos = 0
        rows = []
        for _ in range(steps):
            batch = next(iterator)
            rows.append(self.train_step(batch, staged=staged) if train else self.test_step(batch, staged=staged))
        stacked = self._reduce_rows(torch.stack(rows)).cpu().numpy()
        per_step = [self.metrics_from_row(r) for r in stacked]
        return OrderedDict((k, float(np.mean([m[k] for m in per_step]))) for k in per_step[0])  # A15: epoch mean of add_metric values

    def fit(self, dataset: Iterable, steps_per_epoch: int, epochs: int = 1, validation_data: Optional[Iterable] = None,
            validation_steps: Optional[int] = None, validation_freq: int = 1, callbacks=None, verbose: int = 2):
        """train.py:79-88.  ``dataset`` yields batch dicts and repeats (train.py:44-46 ``repeat=True``)."""
        from .data import DevicePrefetcher

        if getattr(dataset, "yields_device_batches", False):  # DataSpec.make_dataset(cache="device"): batches are cut out of HBM
            iterator = iter(dataset)
        else:
            iterator = DevicePrefetcher(self, dataset)  # H2D copies of step i+1 run under the compute of step i
        from .callbacks import CallbackList

        hooks = CallbackList(callbacks, self)  # Keras-style objects (flex_dm_b200.callbacks, helpers/callbacks.py:36-66) or callables
        self.stop_training = False
        hooks.on_train_begin()
        for epoch in range(epochs):
            hooks.on_epoch_begin(epoch)
            logs = self._run_epoch(iterator, steps_per_epoch, True, staged=True)
            if validation_data is not None and (epoch + 1) % max(1, validation_freq) == 0:
                val = self._run_epoch(iter(validation_data), validation_steps or 1, False)
                logs.update(("val_" + k, v) for k, v in val.items())
            if not math.isfinite(logs["loss"]):  # TerminateOnNaN (helpers/callbacks.py:57)
                logger.error("loss is not finite, terminating")
                self.history.append(logs)
                break
            self.history.append(logs)
            hooks.on_epoch_end(epoch, logs)
            if verbose:
                logger.info("Epoch %d/%d - %s", epoch + 1, epochs, " - ".join("%s: %.4f" % kv for kv in logs.items()))
            if self.stop_training:  # set by a callback (Keras: model.stop_training)
                break
        hooks.on_train_end()
        return self.history

    def evaluate(self, dataset: Iterable, batch_size=None, steps: Optional[int] = None) -> List[float]:
        """train.py:90: returns values in ``metrics_names`` order."""
        if steps is None:
            dataset = list(dataset)
            steps = len(dataset)
        logs = self._run_epoch(iter(dataset), steps, False)
        return [logs[k] for k in self.metrics_names]

    # ------------------------------------------------------------------ call
    def __call__(self, inputs: Dict, training: bool = False, demo_args: Optional[Dict] = None) -> Dict[str, torch.Tensor]:
        """MFP.call (mfp.py:298-347)."""
        is_demo = True if demo_args else False
        staged = self.stage(inputs)
        B, S, length, cols = self._bind(staged, dense=True)  # merge_inputs_and_prediction copies ground-truth values: dense columns
        eng, seed, step = self.engine, self.seed, self._step
        tasks = eng.sample_tasks(self.task_ids, seed, step)  # mfp.py:301
        if is_demo and "tasks" in demo_args:  # preprocess_for_test(..., demo_args.get("tasks", tasks)) (mfp.py:307-312)
            t = demo_args["tasks"]
            self._set_context((torch.as_tensor(t) if not isinstance(t, torch.Tensor) else t).to(self.device))
        else:
            self._set_context(tasks)
        if is_demo:
            masks = []
            for key in self.keys:
                m = demo_args["masks"][key]
                m = torch.as_tensor(m) if not isinstance(m, torch.Tensor) else m
                m = m.to(self.device).to(torch.uint8)
                if self.context in ("id", "length", "canvas") and self._pad_context:
                    m = torch.nn.functional.pad(m, (0, 1))  # the context token's row is never masked
                masks.append(m.contiguous())
            num_iter = int(demo_args.get("num_iter", 1))
            final_logits = None
            if num_iter > 1:
                final_logits = self._iterative_decode(B, S, length, cols, masks, num_iter, seed, step)  # mfp.py:141-207
            else:
                eng.mask_for_test(length, cols, masks)  # mfp.py:72-92
                eng.forward(length, None, training, seed, step)
            if "tasks" in demo_args:
                t = demo_args["tasks"]
                tasks = (torch.as_tensor(t) if not isinstance(t, torch.Tensor) else t).to(self.device)
        else:
            merge_cols = cols  # merge_inputs_and_prediction receives the caller's (unshuffled) inputs: mfp.py:342-344
            if self.input_dtype != "set":
                cols = eng.shuffle_inputs(length, cols, seed, step)
            eng.mask_corrupt(length, cols, tasks, seed, step)  # mfp.py:95-138
            eng.forward(length, None, training, seed, step)
            row = self._next_row()
            eng.loss(length, cols, eng.masks, row, 1.0 / (B * self._world), False, sort_tasks=tasks if self.sort_pos else None)
            eng.regularization_loss(row[-1:])
            self.last_metrics_row = row
            masks = eng.masks
            cols = merge_cols
        self._step += 1
        # merge_inputs_and_prediction (mfp.py:46-69)
        outputs = {}
        for key, column in self.input_columns.items():
            if not column["is_sequence"]:
                if key in staged:
                    outputs[key] = staged[key]
        for f, key in enumerate(self.keys):
            c = self.engine.columns[key]
            shape = (B, S, c["shape"][-1], c["input_dim"]) if c["type"] == "categorical" else (B, S, c["shape"][-1])
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
            eng.merge_prediction(f, cols[f], masks[f], out, logits_in=final_logits if is_demo else None)
            outputs[key] = self._crop(out)
        for key, column in self.all_columns.items():  # copy unpredicted items for visualization (mfp.py:66-68)
            if column.get("demo_only", False) and key in inputs:
                outputs[key] = inputs[key]
        outputs["tasks"] = tasks.clone()
        return outputs

    def _iterative_decode(self, B: int, S: int, length, cols, masks_u8, num_iter: int, seed: int, step: int) -> torch.Tensor:
        """``iterative_decode`` (mfp.py:141-207), MaskGIT-like: ``num_iter`` forward passes of the engine; after each one the
        categorical predictions whose confidence reaches the document's top-k threshold are written back into the inputs and
        unmasked.  The selection itself is a few small device-tensor operations per pass (control logic around the forward
        passes, which is where the time goes).  Returns the flat ``[B*S, logit_width]`` matrix of final outputs.
        The reference compares a (B, S) confidence with a (B,) threshold (:184), which broadcasts only for B = 1; the threshold
        is applied per document here, which is identical for B = 1."""
        eng = self.engine
        columns = [eng.columns[k] for k in self.keys]
        cat = [f for f, c in enumerate(columns) if c["type"] == "categorical"]
        eng.mask_for_test(length, cols, [torch.zeros_like(m) for m in masks_u8])  # filter_padding alone (:146)
        filtered = [t.clone() for t in eng.modified]
        eng.mask_for_test(length, cols, masks_u8)  # modified_inputs of preprocess_for_test
        masks = [m.reshape(B, S).bool().clone() for m in masks_u8]
        num_masked = sum(masks[f].sum(dim=-1) for f in cat)
        # np.round(num_masked / num_iter) of :152-153, on the device (both round half to even): no host synchronisation in the decode loop
        num_update = torch.round(num_masked.to(torch.float64) / num_iter).to(torch.int64)
        logits = torch.empty((B * S, eng.logit_width), dtype=torch.float32, device=self.device)
        final = None
        rows = torch.arange(B, device=self.device)
        for i in range(num_iter):
            eng.forward(length, None, False, seed, step, logits_out=logits)
            out = self.split_logits(logits, B, S)
            if i == 0:
                final = logits.clone()
            fin = self.split_logits(final, B, S)
            conf = {}
            for f in cat:
                p = torch.softmax(out[self.keys[f]], dim=-1).max(dim=-1).values.mean(dim=-1)  # mean over sub-targets (e.g. RGB) of the max probability
                conf[f] = torch.where(masks[f], p, torch.zeros((), dtype=p.dtype, device=self.device))
            conf_sorted = torch.sort(torch.cat([conf[f] for f in cat], dim=-1), dim=-1, descending=True).values
            threshold = conf_sorted[rows, num_update.clamp(max=conf_sorted.shape[1] - 1)]
            for f in cat:
                key = self.keys[f]
                pred = out[key].argmax(dim=-1).to(torch.int32)
                update = (conf[f] >= threshold[:, None]) & (conf[f] > 0)
                filtered[f] = torch.where(update[:, :, None], pred.reshape(filtered[f].shape), filtered[f])
                masks[f] = torch.where(masks[f] == update, torch.zeros_like(masks[f]), masks[f])
                if i > 0:
                    fin[key].copy_(torch.where(update[:, :, None, None], out[key], fin[key]))
            for f, c in enumerate(columns):  # apply_token(filtered, masks, "masked") (masking.py:68-95) -> the next pass's inputs
                m = masks[f].reshape(B, S, 1)
                if c["type"] == "categorical":
                    eng.modified[f].copy_(torch.where(m, torch.full_like(filtered[f], c["input_dim"]), filtered[f]).reshape(eng.modified[f].shape))
                else:
                    eng.modified[f].copy_(torch.where(m, torch.full_like(filtered[f], 10.0), filtered[f]).reshape(eng.modified[f].shape))
        fin = self.split_logits(final, B, S)
        out = self.split_logits(logits, B, S)
        for f, c in enumerate(columns):  # "use last prediction for numerical fields" (:202-204)
            if c["type"] == "numerical":
                fin[self.keys[f]].copy_(out[self.keys[f]])
        return final

    def model(self, modified_inputs: Dict, training: bool = False, seed: Optional[int] = None, step: int = 0) -> Dict[str, torch.Tensor]:
        """The inner boundary ``self.model(modified_inputs, training)`` = ``Model.call`` (model.py:26-30):
        encoder -> blocks -> decoder on already-corrupted inputs; returns the raw per-field outputs
        (decoder.py:96-110): categorical ``(B,S,C,input_dim)``, numerical ``(B,S,C)``."""
        staged = self.stage(modified_inputs)
        B, S, length, cols = self._bind(staged, dense=True)  # already-corrupted inputs feed the encoder GEMMs directly
        eng = self.engine
        if self.context == "id":
            t = modified_inputs["task"]  # added by preprocess_for_train / preprocess_for_test (mfp.py:91,137)
            self._set_context((torch.as_tensor(np.asarray(t)) if not isinstance(t, torch.Tensor) else t).to(self.device))
        elif self.context is not None:
            self._set_context(None)
        logits = torch.empty((B * S, eng.logit_width), dtype=torch.float32, device=self.device)
        eng.forward(length, cols, training, self.seed if seed is None else seed, step, logits_out=logits)
        return OrderedDict((k, self._crop(v)) for k, v in self.split_logits(logits, B, S).items())

    def split_logits(self, logits: torch.Tensor, B: int, S: int) -> Dict[str, torch.Tensor]:
        out = OrderedDict()
        for f, key in enumerate(self.keys):
            c = self.engine.columns[key]
            off = self.engine.logit_offsets[f]
            if c["type"] == "categorical":
                w = c["shape"][-1] * c["input_dim"]
                out[key] = logits[:, off:off + w].reshape(B, S, c["shape"][-1], c["input_dim"])
            else:
                out[key] = logits[:, off:off + c["shape"][-1]].reshape(B, S, c["shape"][-1])
        return out

    # ------------------------------------------------------------------ weights
    def get_weights(self):
        return self.engine.get_weights()

    def set_weights(self, weights):
        self.engine.set_weights(weights)

    def save_weights(self, path: str):
        """train.py:94-97, helpers/callbacks.py:49-56.  Like Keras, the file format follows the path: anything that is not ``.npz`` is a
        TensorFlow object-based checkpoint (``<path>.index`` + ``<path>.data-00000-of-00001``, variables keyed by the reference's
        attribute paths; written by ``libflexdm_io.so``); ``.npz`` keeps the flat numpy archive."""
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        if path.endswith((".h5", ".hdf5", ".keras")):
            raise NotImplementedError("HDF5 weight files are not supported; the reference writes TensorFlow checkpoints (best.ckpt / final.ckpt)")
        if path.endswith(".npz"):
            np.savez(path, **self.engine.get_weights())
        else:
            checkpoint.save_variables(path, self.engine.get_weights())

    def load_weights(self, path: str):
        """train.py:67-69, eval.py:169-172: a TensorFlow checkpoint prefix (``.../best.ckpt``) or an ``.npz`` archive of ``save_weights``."""
        if checkpoint.is_tf_checkpoint(path):
            wanted = OrderedDict((name, self.engine.variable_shape(name)) for name in self.engine.variables)
            self.engine.set_weights(checkpoint.load_variables(path, wanted))
            return
        npz = path if path.endswith(".npz") else path + ".npz"
        if not os.path.exists(npz):
            raise FileNotFoundError("Neither a TensorFlow checkpoint (%s.index) nor %s exists" % (path, npz))
        with np.load(npz) as data:
            self.engine.set_weights({k: data[k] for k in data.files})
