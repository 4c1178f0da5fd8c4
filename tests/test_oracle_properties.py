"""Property tests of the CPU oracle (hypothesis, CPU-only): invariants that follow from the reference's
definitions and that the GPU parity tests then inherit through the oracle."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle

D6 = [1, 2, 4, 8, 12, 24]
dims = st.tuples(st.integers(1, 2), st.integers(2, 5), st.integers(1, 30), st.integers(1, 30))


def _mask(rng, B, C, H, W):
    e = np.exp(rng.randn(B, C, H, W)).astype(np.float32)
    return e / e.sum(1, keepdims=True)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(dims, st.integers(0, 10_000))
def test_pamr_is_a_convex_combination(d, seed):
    B, C, H, W = d
    rng = np.random.RandomState(seed)
    image, mask = rng.rand(B, 3, H, W).astype(np.float32), _mask(rng, B, C, H, W)
    out = oracle.pamr_forward(image, mask, 3, D6)
    lo, hi = mask.reshape(B, C, -1).min(-1), mask.reshape(B, C, -1).max(-1)
    assert (out.reshape(B, C, -1).min(-1) >= lo - 1e-6).all() and (out.reshape(B, C, -1).max(-1) <= hi + 1e-6).all()
    assert np.abs(out.sum(1) - 1).max() <= 1e-5                       # affinity rows sum to one
    aff = oracle.affinity(image, D6)
    assert np.abs(aff.sum(1) - 1).max() <= 1e-5 and (aff >= 0).all()


@settings(max_examples=25, deadline=None, derandomize=True)
@given(dims, st.integers(0, 10_000))
def test_pamr_is_linear_in_the_mask_and_shift_invariant_in_the_image(d, seed):
    B, C, H, W = d
    rng = np.random.RandomState(seed)
    image, a, b = rng.rand(B, 3, H, W).astype(np.float32), _mask(rng, B, C, H, W), _mask(rng, B, C, H, W)
    mix = oracle.pamr_forward(image, (0.25 * a + 0.75 * b).astype(np.float32), 2, D6)
    assert np.abs(mix - (0.25 * oracle.pamr_forward(image, a, 2, D6) + 0.75 * oracle.pamr_forward(image, b, 2, D6))).max() <= 2e-6
    # the affinity depends on intensity differences only (power-of-two shift: exact in fp32 for these ranges)
    assert np.abs(oracle.affinity(image, D6) - oracle.affinity(image + np.float32(0.5), D6)).max() <= 5e-6


@settings(max_examples=25, deadline=None, derandomize=True)
@given(dims, st.integers(0, 10_000))
def test_pseudo_labels_are_consistent_with_pseudo_gt(d, seed):
    B, C, H, W = d
    m = _mask(np.random.RandomState(seed), B, C, H, W)
    pg = oracle.pseudo_gtmask(m)
    lab = oracle.pseudo_labels(m)
    assert set(np.unique(pg)) <= {0.0, 1.0} and (pg.sum(1) <= 1).all()
    assert ((lab == 255) == (pg.sum(1) == 0)).all()
    has = lab != 255
    assert (np.argmax(pg, 1)[has] == lab[has]).all()


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.tuples(st.integers(1, 3), st.integers(2, 6), st.integers(1, 9), st.integers(1, 9), st.integers(1, 20), st.integers(1, 20)),
       st.integers(0, 10_000))
def test_mask_ce_gradient_sums_to_zero_over_classes_and_vanishes_where_ignored(d, seed):
    B, C, h, w, H, W = d
    rng = np.random.RandomState(seed)
    logits = rng.randn(B, C, h, w).astype(np.float32)
    lab = rng.randint(0, C, size=(B, H, W)); lab[rng.rand(B, H, W) < 0.4] = 255
    pg = np.zeros((B, C, H, W), np.float32)
    for c in range(C):
        pg[:, c][lab == c] = 1
    gl = (pg.reshape(B, C, -1).sum(-1)[:, 1:] > 0).astype(np.float32)
    loss, grad = oracle.balanced_mask_loss_ce(logits, pg, gl, np.ones(B, np.float32))
    assert (loss >= 0).all() and np.isfinite(grad).all()
    # each fp32 gradient carries its own rounding (half an ulp of its magnitude): the class sum is zero to a few ulps of the largest entry
    assert np.abs(grad.sum(1)).max() <= 1e-6 * max(float(np.abs(grad).max()), 1e-30)
    empty = pg.reshape(B, -1).sum(-1) == 0
    assert not grad[empty].any() and not loss[empty].any()


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.integers(1, 6), st.integers(2, 5), st.integers(2, 16), st.integers(2, 16), st.integers(0, 10_000))
def test_merge_of_identical_unpadded_scales_is_the_identity(S, C, H, W, seed):
    m = _mask(np.random.RandomState(seed), 1, C, H, W)[0]
    masks = np.repeat(m[None], S, 0)
    pads = np.tile(np.int32([0, 0, H, W]), (S, 1))
    merged, pred = oracle.merge_multiscale(masks, pads, None, (H, W), False, 1.0, 0.0)
    assert np.abs(merged - m).max() <= 1e-6                          # same-size resize + mean of equal copies
    assert (pred == np.argmax(merged, 0)).all()
