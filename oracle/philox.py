"""Philox4x32-10 counter RNG shared by the oracle and the CUDA kernels (TEST INFRASTRUCTURE).

The reference draws its randomness with unseeded ``tf.random.uniform`` (src/UtilsCV.py:514 for the
importance samples, src/UtilsCV.py:580 for the stratified jitter), so no fixed stream exists in the
reference itself.  To make "bit-exact sample indices for a fixed RNG stream" a testable statement the
new framework defines ONE stream, used by both sides:

    key      = (seed & 0xffffffff, seed >> 32)
    counter  = (ray_index, draw_index // 4, stream_id, step)
    word     = Philox4x32-10(key, counter)[draw_index % 4]
    uniform  = bits_to_float((word & 0x7fffff) | 0x3f800000) - 1.0          in [0, 1)

``stream_id`` 0 = stratified jitter (get_z_values), 1 = inverse-CDF uniforms
(get_z_vals_from_prob_dist_func).  ``ray_index`` is the GLOBAL ray index so the result does not
depend on how rays are sharded over GPUs.  The mantissa trick is the one TensorFlow's
``random::Uint32ToFloat`` uses (TensorFlow 2.7 is a third-party dependency not vendored under
/root/reference; restated from its published algorithm).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)

STREAM_JITTER = 0
STREAM_IMPORTANCE = 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds. All inputs broadcastable uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32)
    c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32)
    c3 = np.asarray(c3, dtype=np.uint32)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = PHILOX_M0 * c0.astype(np.uint64)
            p1 = PHILOX_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def bits_to_uniform(words):
    """uint32 -> float32 in [0,1) using the low 23 bits as the mantissa of a float in [1,2)."""
    words = np.asarray(words, dtype=np.uint32)
    bits = (words & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)
    return bits.view(np.float32) - np.float32(1.0)


def uniform(seed, stream_id, step, n_rays, n_draws, ray_offset=0):
    """(n_rays, n_draws) float32 uniforms of the shared stream."""
    seed = int(seed)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    rays = (np.arange(n_rays, dtype=np.uint64) + np.uint64(ray_offset)).astype(np.uint32)[:, None]
    n_blocks = (n_draws + 3) // 4
    blocks = np.arange(n_blocks, dtype=np.uint32)[None, :]
    w = philox4x32_10(rays, blocks, np.uint32(stream_id), np.uint32(step & 0xFFFFFFFF), k0, k1)
    words = np.stack(w, axis=-1).reshape(n_rays, n_blocks * 4)[:, :n_draws]
    return bits_to_uniform(words)
