"""GPU parity tests of the small-map RESIDENT kernel (csrc/pamr_resident.cu): affinity + all iterations in one
cooperative launch, chosen automatically for the shapes stage_net calls PAMR with (reference
models/SoftMaxAE.py:176-179, 251: 41x41 / 81x81 masks for a 321x321 crop).  Same bars as test_gpu_parity.py:
refined masks max-abs <= 1e-5 against the CPU oracle, fused class max bit-equal to the max of the result."""
import numpy as np
import pytest
import torch

import synth
import wseg_b200
from oracle import oracle
from wseg_b200 import _lib

pytestmark = pytest.mark.gpu

D6 = [1, 2, 4, 8, 12, 24]
TOL = 1e-5
DEV = "cuda:0"


def G(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def forward_counting(pamr, image, mask, **kw):
    n0 = _lib.launch_count()
    out = pamr(image, mask, **kw)
    torch.cuda.synchronize()
    return out, _lib.launch_count() - n0


def class_max_as_float(cmax):
    L = _lib.lib()
    return np.array([[L.pamr_float_from_ordered(int(v) & 0xffffffff) for v in row] for row in N(cmax)], dtype=np.float32)


# (B, K, C, H, W, iters): the training shapes, every shared-memory geometry (S / M / L windows), one and three thread
# groups, blocks of one row, several launches per call (B * row blocks > SM count), K != 3, C not a multiple of 3
SHAPES = [(2, 3, 21, 41, 41, 10), (16, 3, 21, 41, 41, 10), (1, 3, 21, 81, 81, 10), (16, 3, 21, 81, 81, 10),
          (3, 3, 5, 33, 57, 3), (2, 1, 4, 20, 100, 1), (2, 5, 7, 64, 64, 2), (1, 3, 21, 7, 9, 10),
          (150, 3, 2, 12, 12, 2), (40, 3, 21, 41, 41, 10), (2, 3, 21, 96, 96, 10), (3, 3, 21, 100, 100, 10),
          (1, 3, 22, 127, 127, 4), (2, 4, 1, 50, 120, 5), (5, 3, 21, 24, 200, 2), (1, 3, 3, 1, 1, 3), (2, 8, 20, 3, 300, 2)]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_resident_vs_oracle(shape):
    B, K, C, H, W, it = shape
    image, mask = synth.image_structured(B, K, H, W, 5), synth.mask_softmax(B, max(C, 2), H, W, 6)[:, :C]
    ref = oracle.pamr_forward(image, mask, it, D6)
    pamr = wseg_b200.PAMR(it, D6).to(DEV)
    (out, cmax), launches = forward_counting(pamr, G(image), G(mask), return_class_max=True)
    o = N(out)
    err = float(np.abs(o - ref).max())
    print("%s: max-abs %.3g, %d launches" % (shape, err, launches))
    assert err <= TOL
    np.testing.assert_array_equal(class_max_as_float(cmax), o.reshape(B, C, -1).max(-1))
    # one small launch that zeroes counters / maxima + one resident launch per group of co-resident images:
    # 1 + iters launches would mean the per-iteration path took the call
    assert launches <= 1 + max(1, (B * min(H, 148) + 147) // 148) and launches < 1 + it + 1


def test_resident_training_shapes_every_sample_and_labels():
    """B = 16 at both stage_net mask sizes: every sample against the oracle, masks and labels (sequence A of
    SoftMaxAE.py:250-259 with the image at 321 x 321)."""
    B, C = 16, 21
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    for (h, seed) in [(41, 71), (81, 72)]:
        image = synth.image_structured(B, 3, 321, 321, seed)
        masks = synth.mask_blobs(B, C, h, h, seed + 10)
        labels = synth.labels_bernoulli(B, C, seed + 20, p=0.3)
        lab, dec = wseg_b200.refine_and_label(pamr, G(image), G(masks), G(labels), return_masks=True)
        dec_ref = oracle.run_pamr(image, masks)
        err = float(np.abs(N(dec) - dec_ref).max())
        cleaned_ref = oracle.rescale_and_clean(dec_ref, (321, 321), labels)
        lab_ref = oracle.pseudo_labels(cleaned_ref)
        near = oracle.near_threshold_set(cleaned_ref)
        print("mask %dx%d: max-abs %.3g, near-threshold pixels %d of %d" % (h, h, err, int(near.sum()), near.size))
        assert err <= TOL
        assert np.array_equal(N(lab)[~near], lab_ref[~near]) and near.mean() < 0.01


def test_resident_fuzz_small_maps():
    """Shape fuzz for H, W < 128: random B, K, C, iterations."""
    rng = np.random.RandomState(321)
    pamr_cache = {}
    for _ in range(60):
        B, K, C = int(rng.randint(1, 20)), int(rng.choice([1, 3, 3, 3, 4])), int(rng.randint(1, 25))
        H, W = int(rng.randint(1, 128)), int(rng.randint(1, 128))
        it = int(rng.choice([1, 2, 3, 10]))
        image = rng.rand(B, K, H, W).astype(np.float32)
        e = np.exp(rng.randn(B, C, H, W)).astype(np.float32)
        mask = e / e.sum(1, keepdims=True)
        pamr = pamr_cache.setdefault(it, wseg_b200.PAMR(it, D6).to(DEV))
        out = N(pamr(G(image), G(mask)))
        assert np.abs(out - oracle.pamr_forward(image, mask, it, D6)).max() <= TOL, (B, K, C, H, W, it)


def test_resident_matches_per_iteration_path_and_shards():
    """The resident kernel adds a pixel's 48 products in the same order as the tile / generic kernels, so the result
    equals affinity + propagate (per-iteration path through the public two-call API) to the last bit of the
    propagation, and sharding the batch changes nothing."""
    B, C, H, W = 6, 21, 81, 81
    image, mask = G(synth.image_structured(B, 3, H, W, 91)), G(synth.mask_softmax(B, C, H, W, 92))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    out = pamr(image, mask)
    two_call = wseg_b200.propagate(wseg_b200.local_affinity(image, D6), mask, D6, 10)
    assert float((out - two_call).abs().max()) <= 2e-6  # (the two affinity kernels may differ by an ulp or two)
    parts = torch.cat([pamr(image[:1], mask[:1]), pamr(image[1:4], mask[1:4]), pamr(image[4:], mask[4:])], 0)
    assert torch.equal(parts, out)


def test_resident_repeatability_under_allocator_churn():
    """Race detector for the stage ring / per-image barriers: bit-identical results over repeated calls."""
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    for (B, H, W) in [(16, 81, 81), (16, 41, 41), (2, 96, 96)]:
        image = G(synth.image_uniform(B, 3, H, W, 51))
        mask = G(synth.mask_softmax(B, 21, H, W, 52))
        ref = pamr(image, mask).clone()
        bad = 0
        for _ in range(300):
            junk = torch.rand((B, 21, H, W), device=DEV)  # noqa: F841  (L2 / allocator churn)
            bad += int(not torch.equal(pamr(image, mask), ref))
        assert bad == 0, (B, H, W)


def test_resident_in_cuda_graph():
    """A cooperative launch inside a captured graph: replay on new inputs equals the eager call."""
    B, C, H, W = 16, 21, 41, 41
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    s_img, s_msk = G(synth.image_structured(B, 3, H, W, 61)), G(synth.mask_softmax(B, C, H, W, 62))
    pamr(s_img, s_msk)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s_out = pamr(s_img, s_msk)
    for seed in (71, 72):
        img, msk = G(synth.image_structured(B, 3, H, W, seed)), G(synth.mask_softmax(B, C, H, W, seed + 10))
        s_img.copy_(img); s_msk.copy_(msk)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(s_out, pamr(img, msk))
