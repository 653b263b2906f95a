"""The callbacks ``train.py:86`` hands to ``MFP.fit`` (``helpers/callbacks.py:36-66``: TensorBoard, ModelCheckpoint on ``val_total_score``,
TerminateOnNaN, GarbageCollector), without TensorFlow: the epoch-level slice of the Keras callback protocol that list uses.

``MFP.fit(callbacks=[...])`` accepts objects with the Keras method names (``set_model``, ``on_train_begin``, ``on_epoch_begin``,
``on_epoch_end(epoch, logs)``, ``on_train_end``; every one optional, so a maintainer's own ``tf.keras.callbacks.Callback`` subclass that
only needs epoch logs keeps working) and plain callables ``cb(epoch, logs, model)``.  ``model.stop_training = True`` ends training after
the current epoch, as in Keras.  Batch-level hooks are not part of it: the step loop never reads a loss back to the host inside an epoch
(metric rows stay on the device until the epoch's reduction), which is the point of the B200 loop.
"""
import gc
import json
import logging
import math
import os
import shutil
import time
from typing import Dict, Iterable, List, Optional

logger = logging.getLogger(__name__)


class Callback:
    """``tf.keras.callbacks.Callback``: the hooks ``fit`` calls, all no-ops."""

    model = None

    def set_model(self, model) -> None:
        self.model = model

    def on_train_begin(self, logs: Optional[Dict] = None) -> None:
        pass

    def on_epoch_begin(self, epoch: int, logs: Optional[Dict] = None) -> None:
        pass

    def on_epoch_end(self, epoch: int, logs: Optional[Dict] = None) -> None:
        pass

    def on_train_end(self, logs: Optional[Dict] = None) -> None:
        pass


class ModelCheckpoint(Callback):
    """``tf.keras.callbacks.ModelCheckpoint`` as ``helpers/callbacks.py:49-56`` configures it: ``save_weights(filepath)`` at the end of an
    epoch, with ``save_best_only`` only when ``monitor`` improved (``mode`` ``"max"`` / ``"min"``; ``"auto"`` = Keras' rule: max for names
    holding ``acc`` or starting with ``fmeasure``, else min -- the reference always passes the mode).  An epoch whose logs do not
    hold the monitored value (no validation that epoch: ``validation_freq``) is skipped with a warning, as in Keras.  ``filepath`` may use
    ``{epoch}`` (one-based, Keras' ``epoch + 1``) and the names of the logs."""

    def __init__(self, filepath: str, monitor: str = "val_loss", verbose: int = 0, save_best_only: bool = False, save_weights_only: bool = False,
                 mode: str = "auto"):
        if not save_weights_only:
            raise NotImplementedError("save_weights_only=False (a SavedModel of the whole Keras model) has no counterpart here; the reference passes True")
        if mode not in ("auto", "min", "max"):
            raise ValueError("mode=%r" % (mode,))
        if mode == "auto":
            mode = "max" if ("acc" in monitor or monitor.startswith("fmeasure")) else "min"
        self.filepath, self.monitor, self.verbose, self.save_best_only, self.mode = filepath, monitor, verbose, save_best_only, mode
        self.best = -math.inf if mode == "max" else math.inf

    def _improved(self, value: float) -> bool:
        return value > self.best if self.mode == "max" else value < self.best

    def on_epoch_end(self, epoch: int, logs: Optional[Dict] = None) -> None:
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            value = logs.get(self.monitor)
            if value is None:
                logger.warning("Can save best model only with %s available, skipping.", self.monitor)
                return
            if not self._improved(value):
                if self.verbose:
                    logger.info("Epoch %d: %s did not improve from %.5f", epoch + 1, self.monitor, self.best)
                return
            if self.verbose:
                logger.info("Epoch %d: %s improved from %.5f to %.5f, saving model to %s", epoch + 1, self.monitor, self.best, value, path)
            self.best = value
        elif self.verbose:
            logger.info("Epoch %d: saving model to %s", epoch + 1, path)
        directory = os.path.dirname(path)
        if directory:
            os.makedirs(directory, exist_ok=True)
        self.model.save_weights(path)


class TerminateOnNaN(Callback):
    """``tf.keras.callbacks.TerminateOnNaN`` at epoch granularity (the epoch mean of a loss that went NaN / inf in any step is NaN / inf)."""

    def on_epoch_end(self, epoch: int, logs: Optional[Dict] = None) -> None:
        loss = (logs or {}).get("loss")
        if loss is not None and not math.isfinite(loss):
            logger.error("Epoch %d: Invalid loss, terminating training", epoch + 1)
            self.model.stop_training = True


class GarbageCollector(Callback):
    """``helpers/callbacks.py:30-33`` (its ``clear_session`` has nothing to clear here)."""

    def on_epoch_end(self, epoch: int, logs: Optional[Dict] = None) -> None:
        gc.collect()


class ScalarLogger(Callback):
    """What the reference's TensorBoard callback records per epoch (the train and validation logs), as JSON lines in
    ``<log_dir>/scalars.jsonl``: the dependency-free form (``TensorBoard`` below writes real event files where the ``tensorboard`` package
    is installed)."""

    def __init__(self, log_dir: str):
        self.log_dir = log_dir
        self._file = None

    def on_train_begin(self, logs: Optional[Dict] = None) -> None:
        os.makedirs(self.log_dir, exist_ok=True)
        self._file = open(os.path.join(self.log_dir, "scalars.jsonl"), "a")

    def on_epoch_end(self, epoch: int, logs: Optional[Dict] = None) -> None:
        if self._file is None:
            self.on_train_begin()
        record = {"epoch": epoch, "wall_time": time.time()}
        record.update({k: float(v) for k, v in (logs or {}).items()})
        self._file.write(json.dumps(record) + "\n")
        self._file.flush()

    def on_train_end(self, logs: Optional[Dict] = None) -> None:
        if self._file is not None:
            self._file.close()
            self._file = None


class TensorBoard(Callback):
    """``tf.keras.callbacks.TensorBoard(log_dir, write_graph=False, profile_batch=0)`` (``helpers/callbacks.py:44-48``) at its default
    ``update_freq="epoch"``: the epoch's logs as ``epoch_<name>`` scalars, training values into ``<log_dir>/train`` and ``val_*`` values
    into ``<log_dir>/validation``, step = epoch.  Event files are written by the ``tensorboard`` package's own writer
    (``torch.utils.tensorboard``); constructing the callback raises ``ImportError`` where that package is missing (``get_callbacks`` then
    falls back to ``ScalarLogger``).  Graph and profiler output (``write_graph`` / ``profile_batch``) have no counterpart: ncu is the
    profiler of this path."""

    def __init__(self, log_dir: str = "logs", write_graph: bool = False, profile_batch=0, **_ignored):
        from torch.utils.tensorboard import SummaryWriter  # noqa: F401  (fail at construction, not at the first epoch)

        self.log_dir = log_dir
        self._writers: Dict[str, object] = {}

    def _writer(self, name: str):
        if name not in self._writers:
            from torch.utils.tensorboard import SummaryWriter

            self._writers[name] = SummaryWriter(log_dir=os.path.join(self.log_dir, name))
        return self._writers[name]

    def on_epoch_end(self, epoch: int, logs: Optional[Dict] = None) -> None:
        for key, value in (logs or {}).items():
            if key.startswith("val_"):
                self._writer("validation").add_scalar("epoch_" + key[4:], float(value), global_step=epoch)
            else:
                self._writer("train").add_scalar("epoch_" + key, float(value), global_step=epoch)
        for w in self._writers.values():
            w.flush()

    def on_train_end(self, logs: Optional[Dict] = None) -> None:
        for w in self._writers.values():
            w.close()
        self._writers = {}


def get_callbacks(args, dataspec, checkpoint_path: str) -> List[Callback]:
    """``helpers/callbacks.py:36-66``: same arguments (``args.job_dir``), same list order and the same ModelCheckpoint settings."""
    log_dir = os.path.join(args.job_dir, "logs")
    if os.path.exists(log_dir):
        logger.warning("Overwriting log dir: %s" % log_dir)
        shutil.rmtree(log_dir)
    logger.info("checkpoint_path=%s", checkpoint_path)
    logger.info("log_dir=%s", log_dir)
    try:
        tensorboard: Callback = TensorBoard(log_dir=log_dir, write_graph=False, profile_batch=0)
    except ImportError:  # no tensorboard package: keep the scalars as JSON lines
        tensorboard = ScalarLogger(log_dir)
    checkpoint = ModelCheckpoint(checkpoint_path, save_weights_only=True, monitor="val_total_score", mode="max", save_best_only=True, verbose=1)
    return [tensorboard, checkpoint, TerminateOnNaN(), GarbageCollector()]


class CallbackList:
    """Dispatch of ``fit``'s hooks over a mixed list: Keras-style objects (any subset of the hook methods) and plain callables
    ``cb(epoch, logs, model)`` (called at the end of an epoch)."""

    def __init__(self, callbacks: Optional[Iterable], model):
        self.model = model
        self.callbacks = list(callbacks or [])
        for cb in self.callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(model)

    def _call(self, hook: str, *args) -> None:
        for cb in self.callbacks:
            fn = getattr(cb, hook, None)
            if fn is not None:
                fn(*args)

    def on_train_begin(self) -> None:
        self._call("on_train_begin", None)

    def on_epoch_begin(self, epoch: int) -> None:
        self._call("on_epoch_begin", epoch, None)

    def on_epoch_end(self, epoch: int, logs: Dict) -> None:
        for cb in self.callbacks:
            fn = getattr(cb, "on_epoch_end", None)
            if fn is not None:
                fn(epoch, logs)
            elif callable(cb):
                cb(epoch, logs, self.model)

    def on_train_end(self) -> None:
        self._call("on_train_end", None)
