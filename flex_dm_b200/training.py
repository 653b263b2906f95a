"""Everything ``python -m mfp`` does after argument parsing (reference: ``main.py:13-15`` -> ``train.py:16-97``), on the B200 path:
``train(args)`` takes the namespace the reference's ``args.py`` produces and leaves the same job directory behind -- ``args.json``,
``logs/`` (TensorBoard epoch scalars), ``checkpoints/best.ckpt`` (best ``val_total_score``) and ``checkpoints/final.ckpt``.

What differs follows from the engine: the seed keys the model's Philox streams (``MFP(seed=...)``; there is no global TensorFlow
generator to seed), ``cache=True`` is the page cache of the mmapped shards, and an optional ``args.device_cache`` (absent = False) keeps
the training split resident in HBM (``DataSpec.make_dataset(cache="device")``).
"""
import json
import logging
import os
import random

import numpy as np

from .callbacks import get_callbacks
from .dataspec import DataSpec
from .mfp import MFP, Adam

logger = logging.getLogger(__name__)

# the MFP constructor keywords that are command-line flags of the same name (args.py; train.py:53-65)
MODEL_FLAGS = ("num_blocks", "block_type", "masking_method", "seq_type", "arch_type", "context", "latent_dim", "dropout", "l2", "input_dtype")


def train(args) -> dict:
    """Runs the job and returns ``{metric name: value}`` of the final test-split evaluation (the pairs train.py:90-92 prints)."""
    seed = int(getattr(args, "seed", 0))
    np.random.seed(seed)
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)

    checkpoints = os.path.join(args.job_dir, "checkpoints")
    os.makedirs(args.job_dir, exist_ok=True)
    with open(os.path.join(args.job_dir, "args.json"), "w") as f:
        json.dump(vars(args), f, indent=2)

    spec = DataSpec(args.dataset_name, args.data_dir, batch_size=args.batch_size)
    splits = {
        # the training split streams in the packed column format (rows of numerical columns nothing reads are neither parsed nor copied)
        "train": spec.make_dataset("train", shuffle=True, repeat=True, seed=seed, cache="device" if getattr(args, "device_cache", False) else True,
                                   packed=not getattr(args, "device_cache", False)),
        "val": spec.make_dataset("val", cache=True),
        "test": spec.make_dataset("test", cache=True),
    }
    model = MFP(spec.make_input_columns(), seed=seed, **{flag: getattr(args, flag) for flag in MODEL_FLAGS})
    if getattr(args, "weights", None):
        logger.info("Loading %s", args.weights)
        model.load_weights(args.weights)
    model.compile(optimizer=Adam(learning_rate=args.learning_rate, clipnorm=1.0))

    epochs = args.num_epochs
    model.fit(splits["train"], steps_per_epoch=spec.steps_per_epoch("train"), epochs=epochs, validation_data=splits["val"],
              validation_steps=spec.steps_per_epoch("val"), validation_freq=min(args.validation_freq, epochs),
              callbacks=get_callbacks(args, spec, os.path.join(checkpoints, "best.ckpt")), verbose=getattr(args, "verbose", 2))

    results = dict(zip(model.metrics_names, model.evaluate(splits["test"], batch_size=args.batch_size)))
    for name, value in results.items():
        print(name, value)
    final = os.path.join(checkpoints, "final.ckpt")
    logger.info("Saving %s", final)
    model.save_weights(final)
    return results
