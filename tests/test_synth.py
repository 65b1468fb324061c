"""The numpy generator (product side) and the C generator (oracle side) must be bit-identical."""
import numpy as np
import pytest

from gnn_cpp_b200 import synth


def test_hash_and_uniform_match_c(oracle):
    idx = np.array([0, 1, 2, 12345, 2 ** 33 + 7], dtype=np.uint64)
    h = synth.hash3(1239, 3, idx)
    for i, v in zip(idx, h):
        assert int(v) == oracle.lib().orc_hash3(1239, 3, int(i))
    a = synth.uniform(77, 3, 10007, -1.0, 1.0)
    assert np.array_equal(a, oracle.synth_uniform(77, 3, 10007, -1.0, 1.0))
    bound = float(np.float32(1.0) / np.sqrt(np.float32(1433)))
    assert np.array_equal(synth.uniform(5, 18, 999, -bound, bound), oracle.synth_uniform(5, 18, 999, -bound, bound))
    assert a.min() >= -1.0 and a.max() < 1.0
    assert np.array_equal(synth.labels(9, 4, 5000, 47), oracle.synth_labels(9, 4, 5000, 47))


@pytest.mark.parametrize("powerlaw", [False, True])
@pytest.mark.parametrize("E", [0, 1, 10, 20001])
def test_edges_match_c(oracle, powerlaw, E):
    N = 3001
    s, d = synth.edges(42, E, N, powerlaw)
    cs, cd = oracle.synth_edges(42, E, N, powerlaw)
    assert np.array_equal(s, cs) and np.array_equal(d, cd)
    if E:
        assert s.min() >= 0 and s.max() < N and d.min() >= 0 and d.max() < N
    half = E // 2
    assert np.array_equal(s[:half], d[half:2 * half]) and np.array_equal(d[:half], s[half:2 * half])  # symmetrised


def test_powerlaw_is_skewed():
    s, d = synth.edges(1, 200000, 5000, True)
    deg = np.bincount(s, minlength=5000)
    assert deg.max() > 10 * deg.mean()
