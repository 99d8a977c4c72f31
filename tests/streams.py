"""Deterministic, dependency-free random streams shared by the golden generator and the tests.

A counter-based hash (splitmix64 finaliser) turns (seed, index) into 32 random bits, so the same
"pre-drawn prey-move and packet-loss random streams" (BASELINE.json north_star) can be rebuilt
on any machine from a seed stored in the fixture instead of shipping megabytes of uniforms.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_u32(seed, idx):
    """splitmix64 finaliser over (seed * 2^32 + idx); returns the high 32 bits as uint32."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) << np.uint64(32)) + np.asarray(idx, dtype=np.uint64)
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(32)).astype(np.uint32)


def uniforms_f32(seed, shape):
    """U[0,1) float32 on the 2^-24 lattice — the lattice torch.rand(float32) produces
    (custom_implement/env_communication.py:212 draws with torch.rand)."""
    n = int(np.prod(shape))
    h = hash_u32(seed, np.arange(n, dtype=np.uint64))
    return ((h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).reshape(shape)


# cumulative prey-move distribution (predator_prey.py:54 — (0.175, 0.175, 0.175, 0.175, 0.3)),
# normalised the way numpy.random.choice does (cumsum / cumsum[-1]).
PREY_CDF = np.cumsum(np.array((0.175, 0.175, 0.175, 0.175, 0.3), dtype=np.float64))
PREY_CDF /= PREY_CDF[-1]


def prey_candidates(seed, shape):
    """Categorical(.175,.175,.175,.175,.3) prey-move candidates, int8, via inverse CDF of a
    53-bit float64 uniform (what numpy.random.choice(5, 1, p=probs) does, predator_prey.py:401)."""
    n = int(np.prod(shape))
    hi = hash_u32(seed, np.arange(n, dtype=np.uint64) * 2).astype(np.uint64)
    lo = hash_u32(seed, np.arange(n, dtype=np.uint64) * 2 + 1).astype(np.uint64)
    u = ((hi >> np.uint64(5)) * np.float64(67108864.0) + (lo >> np.uint64(6)).astype(np.float64)) / np.float64(9007199254740992.0)
    return PREY_CDF.searchsorted(u, side="right").astype(np.int8).reshape(shape)


def actions(seed, shape, bias_move=False):
    """Uniform{0..4} actions, or a move-biased mix (NOOP only 1/16 of the time)."""
    n = int(np.prod(shape))
    h = hash_u32(seed, np.arange(n, dtype=np.uint64))
    if not bias_move:
        a = (h.astype(np.uint64) * np.uint64(5)) >> np.uint64(32)
    else:
        r = h >> np.uint32(28)  # 0..15
        a = np.where(r == 15, 4, r % 4)
    return a.astype(np.int8).reshape(shape)
