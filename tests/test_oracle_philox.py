"""CPU: the Philox4x32-10 restatement (oracle/philox.py) against Random123's published known-answer vectors, and the draw mapping of
gpode_philox_fill (uniform / Box-Muller normal) as a distribution."""
import numpy as np
from scipy import stats

from oracle import philox as P


def test_known_answer_vectors():
    for ctr, key, want in P.KAT:
        got = P.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_draw_mapping_is_standard_normal_and_uniform():
    n = 200001
    u = P.fill(n, 1, seed=7, offset=3, segment=2)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    z = P.fill(n, 0, seed=7, offset=3, segment=0)
    assert np.isfinite(z).all()
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1) < 4 * np.sqrt(2.0 / n)
    assert stats.kstest(z, "norm").pvalue > 1e-3
    # segments, offsets and seeds give different, uncorrelated streams; the same arguments replay exactly
    z2 = P.fill(n, 0, seed=7, offset=3, segment=1)
    assert abs(np.corrcoef(z, z2)[0, 1]) < 4 / np.sqrt(n)
    assert np.array_equal(z, P.fill(n, 0, seed=7, offset=3, segment=0))
    assert np.array_equal(P.fill(40, 1, 9, 5, 0)[8:], P.fill(32, 1, 9, 7, 0))      # offset counts quads
