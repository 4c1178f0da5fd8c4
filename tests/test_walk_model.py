"""CPU model of the band / ownership logic of the y-first loss backward (csrc/pamr_loss.cu, ce_backward_ywalk_kernel):
a thread owns the logit rows [ia, ib) of one label column, walks the label rows whose source row pair (i0, i1) touches
them, keeps the sums of rows i0 / i1 in two accumulators and stores a row when the walk leaves it.  The model restates
that control flow with torch's align_corners=True source index in float32 (make_lerp, pamr_common.cuh) and checks the
two properties the kernel relies on: every logit row is stored exactly once (no atomics, deterministic), and the stored
value is the transposed interpolation  U[i] = sum_y G[y] * wy(y -> i)."""
import numpy as np
import pytest

f32 = np.float32


def make_lerp(dst, scale, n):
    f = f32(scale) * f32(dst)
    i0 = min(int(f), n - 1)
    i1 = i0 + (1 if i0 < n - 1 else 0)
    l1 = min(max(f32(f - f32(i0)), f32(0)), f32(1))
    return i0, i1, f32(1) - l1, l1


def scale_of(n_in, n_out):
    return f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0)


def ywalk(G, h, H, srows):
    sh = scale_of(h, H)
    U, writes = np.zeros(h), np.zeros(h, int)
    for band in range((h + srows - 1) // srows):
        ia, ib = band * srows, min(h, (band + 1) * srows)

        def store(r, v):
            if ia <= r < ib:
                U[r] = v
                writes[r] += 1

        y = max(0, int(f32(ia - 1) / sh) - 1) if (sh > 0 and ia > 0) else 0
        while y < H and make_lerp(y, sh, h)[1] < ia:
            y += 1
        a0 = a1 = 0.0
        r0 = r1 = -1
        while y < H:
            i0, i1, l0, l1 = make_lerp(y, sh, h)
            if i0 >= ib:
                break
            if (i0, i1) != (r0, r1):
                if r0 >= 0:
                    if i0 == r1 and r1 != r0:   # the pair moved down by one row
                        store(r0, a0)
                        a0, a1 = a1, 0.0
                    else:
                        store(r0, a0 + a1 if r1 == r0 else a0)
                        if r1 != r0:
                            store(r1, a1)
                        a0 = a1 = 0.0
                r0, r1 = i0, i1
            a0 += float(l0) * G[y]
            a1 += float(l1) * G[y]
            y += 1
        if r0 >= 0:
            store(r0, a0 + a1 if r1 == r0 else a0)
            if r1 != r0:
                store(r1, a1)
    return U, writes


def dense(G, h, H):
    sh, U = scale_of(h, H), np.zeros(h)
    for y in range(H):
        i0, i1, l0, l1 = make_lerp(y, sh, h)
        U[i0] += float(l0) * G[y]
        U[i1] += float(l1) * G[y]
    return U


@pytest.mark.parametrize("srows", [4, 5, 8, 12, 16])
def test_every_logit_row_has_one_owner_and_the_transposed_sum(srows):
    rng = np.random.RandomState(srows)
    for h, H in [(81, 321), (41, 321), (1, 9), (1, 2), (2, 4), (2, 5), (3, 6), (8, 17), (9, 19), (16, 40), (20, 97),
                 (33, 66), (17, 100), (5, 1000), (64, 129), (100, 200)]:
        assert H >= 2 * h  # the predicate under which the kernel is chosen (backward_y_first)
        G = rng.randn(H)
        U, writes = ywalk(G, h, H, srows)
        assert (writes == 1).all(), (h, H, srows)
        np.testing.assert_allclose(U, dense(G, h, H), rtol=0, atol=1e-12)
