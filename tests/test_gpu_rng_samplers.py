"""Pins for the two pieces of the device path that no image comparison can pin:

* the counter-based generator: Philox4x32-10 evaluated ON THE DEVICE (b200rt_debug_philox runs csrc/rng.cuh's
  philox4x32_10, the function the path kernels call) against the Random123 known-answer vectors
  (Salmon et al., SC'11; Random123 kat_vectors: philox4x32 10 rounds);
* the direct samplers (uniform on the unit sphere from two uniforms; uniform in the unit disk) that replace the
  reference's rejection loops random_unit_vector() / random_vector_in_unit_disk() (reference
  include/math/vec3d.h:64-85): same DISTRIBUTIONS, checked by two-sample Kolmogorov-Smirnov tests of rotation-
  invariant statistics against 1e6 draws of the oracle's restatement of the reference samplers, and by chi-square
  tests of equal-area cells.
"""
import os
import sys

import numpy as np
import pytest
from scipy import stats

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# Random123 known answers, philox4x32 with 10 rounds: (counter[4], key[2]) -> output[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox_py(c, k):
    """Straight transcription of the published algorithm (independent of csrc/rng.cuh)."""
    c = [int(x) for x in c]
    k = [int(x) for x in k]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def test_philox4x32_10_known_answers_on_the_device():
    from cpp_raytracer_b200 import capi
    inp = np.array([list(c) + list(k) for c, k, _ in KAT], dtype=np.uint32)
    got = capi.debug_philox(inp)
    for (c, k, want), row in zip(KAT, got):
        assert tuple(int(x) for x in row) == want, (c, k, [hex(int(x)) for x in row])
        assert tuple(philox_py(c, k)) == want
    # and 4096 random (counter, key) pairs against the independent transcription
    rng = np.random.default_rng(2)
    inp = rng.integers(0, 2 ** 32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    got = capi.debug_philox(inp)
    want = np.array([philox_py(r[:4], r[4:]) for r in inp], dtype=np.uint32)
    assert np.array_equal(got, want)


def test_path_key_layout_gives_distinct_streams():
    """The kernels key Philox by counter = (pixel, sample, bounce, 0) and key = seed: neighbouring pixels, samples and
    bounces must give unrelated words (a counter-based generator guarantees it; this guards the argument ORDER)."""
    from cpp_raytracer_b200 import capi
    rows = [(p, s, b, 0, 0xB200, 0) for p in range(8) for s in range(8) for b in range(8)]
    out = capi.debug_philox(np.array(rows, dtype=np.uint32))
    assert len({tuple(int(x) for x in r) for r in out}) == len(rows)
    u = (out >> 8).astype(np.float64) / 2 ** 24
    assert abs(u.mean() - 0.5) < 0.03


N = 1_000_000


@pytest.fixture(scope="module")
def draws():
    from cpp_raytracer_b200 import capi
    import pt_oracle
    words = np.random.default_rng(7).integers(0, 2 ** 32, size=(N, 2), dtype=np.uint64).astype(np.uint32)
    sph, disk = capi.debug_samplers(words)
    return sph, disk, pt_oracle.random_unit_vectors(N, 4242), pt_oracle.random_vectors_in_unit_disk(N, 777)


def test_sphere_sampler_matches_random_unit_vector_distribution(draws):
    sph, _, ref, _ = draws
    assert np.abs(np.linalg.norm(sph, axis=1) - 1).max() < 2e-6            # FP32 sincos: unit length to a few ulp
    assert np.abs(np.linalg.norm(ref, axis=1) - 1).max() < 1e-12
    # A distribution on the sphere is uniform iff its projection on EVERY axis is U(-1, 1) (Archimedes); test the three
    # coordinate axes and three oblique ones, each against the reference's draws (two-sample KS) and against U(-1,1).
    axes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [1, -2, 0.5], [-0.3, 0.4, 0.87]], dtype=np.float64)
    axes /= np.linalg.norm(axes, axis=1, keepdims=True)
    for a in axes:
        pa, pr = sph @ a, ref @ a
        assert stats.ks_2samp(pa, pr).pvalue > 1e-3, a
        assert stats.kstest(pa, stats.uniform(-1, 2).cdf).pvalue > 1e-3, a
    # chi-square over 16 x 32 equal-area cells (z bands x azimuth sectors)
    zi = np.minimum(((sph[:, 2] + 1) * 8).astype(int), 15)
    ai = np.minimum(((np.arctan2(sph[:, 1], sph[:, 0]) + np.pi) / (2 * np.pi) * 32).astype(int), 31)
    counts = np.bincount(zi * 32 + ai, minlength=512)
    assert stats.chisquare(counts).pvalue > 1e-3
    assert np.abs(sph.mean(axis=0)).max() < 4 / np.sqrt(3 * N) * 2           # mean of each coordinate ~ N(0, 1/(3N))


def test_disk_sampler_matches_random_vector_in_unit_disk_distribution(draws):
    _, disk, _, ref = draws
    r2, rr2 = (disk ** 2).sum(axis=1), (ref ** 2).sum(axis=1)
    assert r2.max() < 1.0 + 1e-6 and rr2.max() < 1.0
    assert stats.ks_2samp(r2, rr2).pvalue > 1e-3                             # |p|^2 ~ U(0, 1) for a uniform disk
    assert stats.kstest(r2, stats.uniform(0, 1).cdf).pvalue > 1e-3
    for a in ([1, 0], [0, 1], [0.6, 0.8], [-0.8, 0.6]):                      # projections: the semicircle law, vs the reference
        assert stats.ks_2samp(disk @ np.array(a), ref @ np.array(a)).pvalue > 1e-3, a
    ri = np.minimum((r2 * 8).astype(int), 7)
    ai = np.minimum(((np.arctan2(disk[:, 1], disk[:, 0]) + np.pi) / (2 * np.pi) * 32).astype(int), 31)
    assert stats.chisquare(np.bincount(ri * 32 + ai, minlength=256)).pvalue > 1e-3
