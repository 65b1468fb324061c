"""CPU model of the hi/lo operand split behind the 3xTF32 dense transforms (gnn.cpp_b200/csrc/gemm_tc.cu): numpy restatement
of what the converter warps write and what the tensor core reads (only the upper 19 bits of a TF32 operand), checked against
the error bounds DESIGN.md §4 / profiles/r2b_gemm_pair.md quote.  No GPU, no product code: this pins the ARITHMETIC the
kernels rely on, so a change of the split has to change this file too."""
import numpy as np

MASK = np.uint32(0xFFFFE000)


def tc_read(x):
    """what a tcgen05 kind::tf32 operand contributes: the FP32 bit pattern with the low 13 mantissa bits ignored"""
    return (x.view(np.uint32) & MASK).view(np.float32)


def rna_tf32(x):
    """cvt.rna.tf32.f32 for finite values: round to nearest, ties away, on the magnitude bits"""
    return ((x.view(np.uint32) + np.uint32(0x1000)) & MASK).view(np.float32)


def split_trunc(v):
    """shipped split of the streamed operand (split4_trunc): the raw tile is hi; lo = (v - trunc(v)) with +0x1000 on its
    bit pattern so that the tensor core's truncation rounds it to nearest"""
    hi = v                                                  # stored as is; the tensor core sees tc_read(v)
    d = (v - tc_read(v)).astype(np.float32)                 # exact in FP32
    lo = (d.view(np.uint32) + np.uint32(0x1000)).view(np.float32)
    return hi, lo


def split_rna(v):
    """round-to-nearest split (split4_from, and prep_weights_kernel for the weights)"""
    hi = rna_tf32(v)
    lo = rna_tf32((v - hi).astype(np.float32))
    return hi, lo


def _values(n, seed):
    rng = np.random.default_rng(seed)
    v = (rng.uniform(-1, 1, n) * np.exp2(rng.integers(-20, 20, n))).astype(np.float32)
    return v[v != 0]


def test_truncating_split_representation_error():
    v = _values(1 << 20, 1)
    hi, lo = split_trunc(v)
    eff = tc_read(hi).astype(np.float64) + tc_read(lo).astype(np.float64)
    rel = (eff - v.astype(np.float64)) / np.abs(v.astype(np.float64))
    assert np.abs(rel).max() <= 2.0 ** -21            # |v - (hi + lo)| <= 2^-21 |v|
    assert abs(rel.mean()) <= 2.0 ** -27              # zero-mean: lo is ROUNDED (a truncated lo is biased by ~2^-22)
    # d = v - trunc(v) is exact and has the sign of v; lo = rna_tf32(d) exactly
    d = (v - tc_read(v)).astype(np.float32)
    assert np.array_equal((v.astype(np.float64) - tc_read(v).astype(np.float64)).astype(np.float32), d)
    assert np.all(d * v >= 0)
    assert np.array_equal(tc_read(lo), rna_tf32(d))
    # the remainder of an operand that is already a TF32 number: 0x1000, which the tensor core reads as zero
    z = tc_read(v)
    assert np.all(tc_read(split_trunc(z)[1]) == 0)


def test_round_to_nearest_split_representation_error():
    v = _values(1 << 20, 2)
    hi, lo = split_rna(v)
    eff = hi.astype(np.float64) + lo.astype(np.float64)
    rel = (eff - v.astype(np.float64)) / np.abs(v.astype(np.float64))
    assert np.abs(rel).max() <= 2.0 ** -22
    assert np.array_equal(tc_read(hi), hi) and np.array_equal(tc_read(lo), lo)   # already TF32 numbers


def test_three_term_product_error_is_inside_the_contract():
    """a.b ~ lo_a hi_b + hi_a lo_b + hi_a hi_b (the dropped term is lo_a lo_b), exact accumulation: what is left is the
    split's own error, far inside the 1e-5 contract for both operand conventions the kernels use"""
    rng = np.random.default_rng(3)
    K = 256
    a = rng.uniform(-1, 1, (2000, K)).astype(np.float32)
    b = rng.uniform(-1, 1, (K, 64)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)

    def three(sa, sb):
        ah, al = (tc_read(x).astype(np.float64) for x in sa)
        bh, bl = (tc_read(x).astype(np.float64) for x in sb)
        return al @ bh + ah @ bl + ah @ bh

    scale = np.abs(ref).max()
    rows = np.abs(three(split_trunc(a), split_rna(b)) - ref).max() / scale       # NT / NN: streamed x pre-split weights
    tn = np.abs(three(split_trunc(a), split_trunc(b)) - ref).max() / scale       # TN: both operands streamed
    rna = np.abs(three(split_rna(a), split_rna(b)) - ref).max() / scale
    assert rna <= rows <= 5e-7 and tn <= 1e-6, (rna, rows, tn)
