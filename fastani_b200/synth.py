"""Deterministic synthetic genomes: numpy twin of csrc/synth.cu (same bytes for the same arguments)."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def synth_genome(seed, ancestor, strain, ppm, length):
    """Upper-case ACGT bytes (numpy uint8) of one synthetic genome."""
    with np.errstate(over="ignore"):
        seed = np.uint64(seed)
        ka = _splitmix64(np.array([seed ^ np.uint64(0xA5A5A5A5 + (int(ancestor) << 32))], np.uint64))[0]
        ks = _splitmix64(np.array([seed ^ np.uint64((0x5A5A5A5A + (int(ancestor) << 32) + (int(strain) << 8) + 1) & 0xFFFFFFFFFFFFFFFF)], np.uint64))[0]
        i = np.arange(length, dtype=np.uint64)
        base = (_splitmix64(ka + i) >> np.uint64(62)).astype(np.uint32)
        u = _splitmix64(ks + i)
        r = ((u >> np.uint64(40)) % np.uint64(1000000)).astype(np.uint32)
        if strain != 0:
            sub = r < np.uint32(ppm)
            alt = (base + np.uint32(1) + ((u >> np.uint64(8)) % np.uint64(3)).astype(np.uint32)) & np.uint32(3)
            base = np.where(sub, alt, base)
    return np.frombuffer(b"ACGT", np.uint8)[base]
