"""TEST INFRASTRUCTURE (oracle) -- not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

Philox4x32-10 counter-based generator (Salmon et al., SC'11), vectorised in numpy.

The reference draws its masks and dropout from TensorFlow's stateful RNG (``masking.py:243-256``,
``tf.keras.layers.Dropout``); those streams cannot be reproduced without TensorFlow, so the B200 path defines
its own counter layout (DESIGN.md "RNG contract") and this file restates it bit-for-bit so that masks,
random tokens and dropout keep-masks can be compared exactly.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)

# stream ids shared with flex_dm_b200/csrc/common.cuh
STREAM_RANDOM_U = 0  # x0,x1,x2 = the three uniforms of random_masking for (token, field)
STREAM_RANDOM_CAT = 1  # + c : random replacement token of sub-target c
STREAM_RANDOM_NUM = 16  # + (j >> 2): Box-Muller normals for numerical dims 4*(j>>2) .. +3
FIELD_ELEM = 1000  # elem_masking: per-document uniform
FIELD_TASK = 1001  # task sampler: per-document draw
FIELD_SHUFFLE = 1002  # shuffle_inputs: per-element sort key (x0)
SITE_POS_DROPOUT = 1999  # Dropout of the PositionEmbedding (input_dtype != "set")
SITE_DROPOUT = 2000  # + 2*block + {0: attention branch, 1: FFN branch}


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays; returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & _MASK32 for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
    """uint32 -> float32 uniform in [0,1): top 24 bits, exact in fp32."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0**-24)


def mulhi_range(x, n):
    """uint32 -> integer in [0,n): (x*n) >> 32."""
    return ((np.asarray(x, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def box_muller(xa, xb):
    """Two uint32 arrays -> two float32 standard normals (fp32 arithmetic, like the kernel)."""
    u1 = ((np.asarray(xa, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0**-24)
    u2 = u01(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    t = (np.float32(6.283185307179586) * u2).astype(np.float32)
    return (r * np.cos(t)).astype(np.float32), (r * np.sin(t)).astype(np.float32)


def dropout_keep(n_elements: int, site: int, rate: float, seed: int, step: int, first_element: int = 0) -> np.ndarray:
    """Keep-mask (bool, n_elements) of one dropout site (flex_dm_b200/csrc/gemm.cuh): element e takes the 16-bit half (e & 7) of
    philox(counter = (e >> 3, site, 0, 0), key = (seed, step)) -- words x0..x3 hold halves (0,1), (2,3), (4,5), (6,7), low half first --
    and is kept iff that half >= round(rate * 65536).  ``first_element`` (a multiple of 8) is the global index of element 0 when the rows
    are a shard of a larger batch (mfp_set_doc_offset)."""
    assert first_element % 8 == 0
    n8 = (n_elements + 7) // 8
    x = philox4x32_10(np.arange(n8) + first_element // 8, site, 0, 0, seed, step)
    halves = np.stack([h for w in x for h in (w & np.uint32(0xFFFF), w >> np.uint32(16))], axis=1).reshape(-1)[:n_elements]
    threshold = np.uint32(int(np.float32(rate) * np.float32(65536.0) + np.float32(0.5)))
    return halves >= threshold
