"""Host-side Philox4x32-10 for the handful of scalars the host draws (the one
uniform per resample, resampling.py:157's ``random()``): same counter contract as
the device RNG (csrc/auvpf_math.cuh ``philox_for``), so a sharded filter draws the
same ``u`` on every rank without communicating."""

_M0, _M1, _W0, _W1, _MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
STREAM_RESAMPLE = 4


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def resample_uniform(seed, filter_id, step):
    """``uniform53(philox_for(seed, filter_id, step, kStreamResample, 0))``."""
    r = philox4x32_10((filter_id & _MASK, (filter_id >> 32) & _MASK, step & _MASK, STREAM_RESAMPLE * 16),
                      (seed & _MASK, (seed >> 32) & _MASK))
    return float(((r[0] << 32) | r[1]) >> 11) * 2.0 ** -53
