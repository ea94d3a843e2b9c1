"""Synthetic vortex fields, reproducible from a counter-based generator so that
C++, CUDA and Python agree bit-for-bit.  Neutral helper: used by tests/, bench.py and tools/."""
import numpy as np


def splitmix_uniform(n, k, seed=20241019, i0=0):
    i = np.arange(i0, i0 + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * np.uint64(1 << 32) + np.uint64(4) * i + np.uint64(k)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def s_wake(n, uniform_vc=True, seed=20241019):
    """S-wake(N): x = 0.015*N*r0, z = r1 - 0.5, G = 0.02*(r2 - 0.5), vc = 0.02"""
    x = 0.015 * n * splitmix_uniform(n, 0, seed)
    z = splitmix_uniform(n, 1, seed) - 0.5
    g = 0.02 * (splitmix_uniform(n, 2, seed) - 0.5)
    vc = np.full(n, 0.02) if uniform_vc else 0.02 * (0.5 + splitmix_uniform(n, 3, seed))
    return x, z, g, vc
