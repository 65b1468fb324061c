"""Property tests (hypothesis) of the oracle's structure code against a literal dense restatement of the reference's
loops (src/graph.cpp:21-75: dense assignment, fill_diagonal_, row-major scan), on random small edge lists incl.
duplicates, self loops, isolated nodes and empty inputs.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st


def _dense(src, dst, N, w=None):
    A = np.zeros((N, N), np.float32)
    for i, (r, c) in enumerate(zip(src, dst)):          # assignment in edge order: the last write wins
        A[r, c] = 1.0 if w is None else w[i]
    return A


edges = st.integers(1, 12).flatmap(
    lambda N: st.tuples(st.just(N), st.lists(st.tuples(st.integers(0, N - 1), st.integers(0, N - 1)), min_size=0, max_size=40)))


@settings(max_examples=150, deadline=None)
@given(edges)
def test_csr_build_equals_dense_round_trip(oracle, e):
    N, el = e
    src = np.array([a for a, _ in el], np.int32); dst = np.array([b for _, b in el], np.int32)
    for fill in (0, 1, 2):
        A = _dense(src, dst, N)
        if fill != 2:
            np.fill_diagonal(A, float(fill))            # tensor::fill_diagonal_(fillValue), include/tensor.h:806-817
        rows, cols = np.nonzero(A)                      # row-major scan of adj_to_edge_list, graph.cpp:52-59
        rowptr, colidx = oracle.csr_build(src, dst, N, fill)
        assert np.array_equal(np.repeat(np.arange(N), np.diff(rowptr)), rows) and np.array_equal(colidx, cols)
    # CSC = CSR of the transpose, perm maps CSC positions to CSR positions
    rowptr, colidx = oracle.csr_build(src, dst, N, 1)
    colptr, rowidx, perm = oracle.csc_from_csr(N, rowptr, colidx)
    At = _dense(src, dst, N); np.fill_diagonal(At, 1.0)
    r2, c2 = np.nonzero(At.T)
    assert np.array_equal(np.repeat(np.arange(N), np.diff(colptr)), r2) and np.array_equal(rowidx, c2)
    rows = np.repeat(np.arange(N), np.diff(rowptr))
    assert np.array_equal(rows[perm], rowidx) and np.array_equal(colidx[perm], np.repeat(np.arange(N), np.diff(colptr)))


@settings(max_examples=100, deadline=None)
@given(edges)
def test_weighted_build_last_write_wins(oracle, e):
    N, el = e
    if not el:
        return
    src = np.array([a for a, _ in el], np.int32); dst = np.array([b for _, b in el], np.int32)
    w = oracle.edge_weights(len(el))
    for fill in (1, 2):
        A = _dense(src, dst, N, w)
        if fill == 1:
            np.fill_diagonal(A, 1.0)
        rowptr, colidx, val0 = oracle.csr_build_weighted(src, dst, w, N, fill)
        rows = np.repeat(np.arange(N), np.diff(rowptr))
        ref_rows, ref_cols = np.nonzero(A)              # the test weights are positive: nonzero pattern == stored pattern
        assert np.array_equal(rows, ref_rows) and np.array_equal(colidx, ref_cols)
        assert np.array_equal(val0, A[ref_rows, ref_cols])
        if fill == 1:
            degf, dinv, val = oracle.degree_norm_weighted(N, rowptr, colidx, val0)
            assert np.allclose(degf, A.sum(1), rtol=1e-6)
            Ahat = (A * dinv[:, None]) * dinv[None, :]
            assert np.allclose(val, Ahat[ref_rows, ref_cols], rtol=1e-6)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 2000), st.integers(1, 9))
def test_partition_ptr_properties(oracle, N, P):
    ptr = oracle.partition_ptr(N, P)
    chunk = (N + P - 1) // P
    assert ptr[0] == 0 and ptr[-1] == N and bool((np.diff(ptr) >= 0).all()) and bool((np.diff(ptr) <= chunk).all())
    assert all(ptr[p] == min(N, p * chunk) for p in range(P + 1))
