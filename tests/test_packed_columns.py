"""Packed numerical columns (``flex_dm_b200.data.pack_batch`` / ``mfp_set_packed_rows``): the rows of the 512-float embedding columns that
``filter_padding`` (masking.py:24-53) would overwrite with <UNUSED> anyway -- padded positions, elements whose type does not carry the
field (data/crello-spec.yml:88-121) -- are not stored, copied or read.  Results must be bit-identical to the dense batch."""
import numpy as np
import pytest
import torch

from flex_dm_b200.data import ROWS_SUFFIX, pack_batch, unpack_column
from flex_dm_b200.spec import make_input_columns, make_synthetic_batch


def test_pack_batch_keeps_exactly_the_rows_the_model_reads():
    cols = make_input_columns("crello")
    batch = make_synthetic_batch(cols, 7, 19, seed=4, lengths="ragged")
    packed = pack_batch(batch, cols)
    valid = np.arange(19)[None, :] <= batch["length"].reshape(-1, 1)
    for key in ("image_embedding", "text_embedding"):
        gate = np.asarray(cols[key]["loss_condition"]["mask"], dtype=bool)[batch["type"][..., 0]]
        rows = packed[key + ROWS_SUFFIX]
        assert rows.dtype == np.int32 and rows.shape == (7, 19)
        assert np.array_equal(rows >= 0, valid & gate)
        assert np.array_equal(np.sort(rows[rows >= 0]), np.arange(packed[key].shape[0]))  # every row referenced once, document by document
        assert np.array_equal(rows[rows >= 0], np.arange(packed[key].shape[0]))
        dense = unpack_column(torch.from_numpy(packed[key]), torch.from_numpy(rows)).numpy()
        assert np.array_equal(dense[rows >= 0], batch[key][rows >= 0]) and not dense[rows < 0].any()
        assert packed[key].nbytes < 0.8 * batch[key].nbytes
    for key in batch:  # everything else passes through untouched
        if key not in ("image_embedding", "text_embedding"):
            assert packed[key] is batch[key]
    rico = make_input_columns("rico")
    rb = make_synthetic_batch(rico, 3, 8, seed=1)
    assert set(pack_batch(rb, rico)) == set(rb)  # no numerical sequence columns: nothing to pack


def _run(cols, method, batches, steps, packed, **kwargs):
    from flex_dm_b200.mfp import MFP, Adam

    m = MFP(cols, num_blocks=2, masking_method=method, latent_dim=256, dropout=0.1, l2=1e-2, seed=5, **kwargs)
    m.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    m.set_deterministic(True)
    rows = []
    for i in range(steps):
        b = batches[i % len(batches)]
        rows.append(m.train_step(pack_batch(b, cols) if packed else b).clone())
    rows.append(m.test_step(pack_batch(batches[0], cols) if packed else batches[0]).clone())
    torch.cuda.synchronize()
    return torch.stack(rows).cpu().numpy(), m.get_weights(), m


@pytest.mark.gpu
@pytest.mark.parametrize("method,kwargs", [("random", {}), ("elem_pos_attr_img_txt", {}), ("random", {"context": "id"}), ("random", {"input_dtype": "sorted_set"})],
                         ids=["random", "multi-task", "context-token", "sorted-set(expanded)"])
def test_packed_batches_train_bit_identically(method, kwargs):
    cols = make_input_columns("crello")
    batches = [make_synthetic_batch(cols, 6, 21, seed=30 + i, lengths="ragged") for i in range(2)]
    r_dense, w_dense, _ = _run(cols, method, batches, 3, False, **kwargs)
    r_packed, w_packed, _ = _run(cols, method, batches, 3, True, **kwargs)
    assert np.array_equal(r_dense, r_packed)
    for name in w_dense:
        assert np.array_equal(w_dense[name], w_packed[name]), name


@pytest.mark.gpu
def test_packed_batches_through_the_prefetcher_and_the_call_surface():
    from flex_dm_b200.data import DevicePrefetcher

    cols = make_input_columns("crello")
    batches = [make_synthetic_batch(cols, 5, 17, seed=40 + i, lengths="ragged") for i in range(3)]  # different row counts per batch
    r_dense, w_dense, m = _run(cols, "random", batches, 3, False)
    from flex_dm_b200.mfp import MFP, Adam

    other = MFP(cols, num_blocks=2, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=5)
    other.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    other.set_deterministic(True)
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in pack_batch(b, cols).items()} for b in batches]
    feeder = DevicePrefetcher(other, iter(pinned))
    rows = [other.train_step(next(feeder), staged=True).clone() for _ in range(3)]
    torch.cuda.synchronize()
    assert np.array_equal(torch.stack(rows).cpu().numpy(), r_dense[:3])
    # model(...) on a packed batch (demo / eval path expands it: merge_inputs_and_prediction copies ground-truth rows)
    masks = {k: np.zeros((5, 17), dtype=bool) for k in other.keys}
    masks["image_embedding"][:, :3] = True
    out_dense = m(batches[0], training=False, demo_args={"masks": masks})
    m2 = MFP(cols, num_blocks=2, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=5)
    m2.set_weights(m.get_weights())
    out_packed = m2(pack_batch(batches[0], cols), training=False, demo_args={"masks": masks})
    for key in other.keys:
        assert torch.equal(out_dense[key], out_packed[key]), key
