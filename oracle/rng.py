"""Oracle (test infrastructure): NumPy restatement of the device Gaussian field generator of
jaxpm_b200/csrc/misc.cu (jpm_normal_field_f32), the product's replacement for the reference's
`normal_field` (/root/reference/jaxpm/distributed.py:193-223; JAX's threefry stream is not reproducible
here, SURVEY.md section 2.2, so the generator is the product's own and this file pins it).

Philox4x32-10 (Salmon, Moraes, Dror & Shaw 2011; published round constants) keyed by the 64-bit seed, counter =
(index of the group of four consecutive cells of the GLOBAL flattened mesh, stream id, 0); Box-Muller on the top
24 bits of each word.  The 32-bit integer stream is compared bit for bit; the normals to float32 rounding of
log / sincospi."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def normal_field_words(seed, shape, stream_id=0):
    """uint32 [ncell/4 rounded up, 4]: the raw Philox output of every 4-cell group of the flattened mesh."""
    n = int(np.prod(shape))
    g = np.arange((n + 3) // 4, dtype=np.uint64)
    lo, hi = (g & MASK).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32)
    sid = np.full_like(lo, stream_id)
    x = philox4x32_10(lo, hi, sid, np.zeros_like(lo), int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    return np.stack(x, axis=-1)


def normal_field(seed, shape, stream_id=0, dtype=np.float64):
    """N(0,1) field of `shape` (C order), Box-Muller in float64 on the same 24-bit uniforms as the device."""
    x = normal_field_words(seed, shape, stream_id)
    u = ((x >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0
    rad0, rad1 = np.sqrt(-2.0 * np.log(u[:, 0])), np.sqrt(-2.0 * np.log(u[:, 2]))
    z = np.stack([rad0 * np.cos(2 * np.pi * u[:, 1]), rad0 * np.sin(2 * np.pi * u[:, 1]),
                  rad1 * np.cos(2 * np.pi * u[:, 3]), rad1 * np.sin(2 * np.pi * u[:, 3])], axis=-1)
    n = int(np.prod(shape))
    return z.reshape(-1)[:n].reshape(shape).astype(dtype)
