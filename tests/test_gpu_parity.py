"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
libpamr_b200.so, against (1) golden vectors produced by the reference's own modules,
(2) the CPU oracle on seeded inputs, (3) size-independent properties at BASELINE.json's sizes.

Tolerances (north_star): refined masks max-abs <= 1e-5 fp32; labels bit-exact outside the
documented near-threshold set {pixels: min_c |m_c - thr_c| <= 2e-5}."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

import synth
import wseg_b200
from oracle import oracle
from wseg_b200 import _lib

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAMR_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "pamr_*.npz")))
RESIZE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "resize_*.npz")))
STAGE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "stage_*.npz")))
D6 = [1, 2, 4, 8, 12, 24]
TOL = 1e-5
DEV = "cuda:0"


def G(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def assert_labels_match(lab, ref_lab, ref_masks, min_agree=0.99):
    near = oracle.near_threshold_set(ref_masks)
    assert np.array_equal(lab[~near], ref_lab[~near]), "labels differ outside the near-threshold set"
    assert near.mean() <= 1 - min_agree
    return int(near.sum())


# ---------------------------------------------------------------- golden vectors (reference outputs)

@pytest.mark.parametrize("name", PAMR_CASES)
def test_pamr_forward_vs_reference_golden(name):
    g = load(name)
    dil = [int(d) for d in g["dilations"]]
    pamr = wseg_b200.PAMR(int(g["num_iter"]), dil).to(DEV)
    out = N(pamr(G(g["image"]), G(g["mask"])))
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= TOL
    aff = N(wseg_b200.local_affinity(G(g["image"]), dil))
    assert np.abs(aff - g["aff"]).max() <= 2e-6
    if "out_iter1" in g:
        m0 = G(oracle.resize_bilinear(g["mask"], g["image"].shape[-2:]))
        one = N(wseg_b200.propagate(G(g["aff"]), m0, dil, 1))
        assert np.abs(one - g["out_iter1"]).max() <= 1e-6


@pytest.mark.parametrize("name", RESIZE_CASES)
def test_resize_vs_reference_golden(name):
    g = load(name)
    y = N(wseg_b200.resize_bilinear(G(g["x"]), g["y"].shape[-2:]))
    assert np.abs(y - g["y"]).max() <= 2.5e-7
    np.testing.assert_array_equal(y, oracle.resize_bilinear(g["x"], g["y"].shape[-2:]))  # bit-exact vs oracle


@pytest.mark.parametrize("name", STAGE_CASES)
def test_stage_sequence_vs_reference_golden(name):
    g = load(name)
    image, masks, labels = G(g["image"]), G(g["masks"]), G(g["labels"])
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    if "pre" in g:  # sequence B (CAM_CASA_WGAP_tf.py:335-345): clean at mask resolution first
        pre = wseg_b200.rescale_and_clean(masks, masks, labels)
        np.testing.assert_array_equal(N(pre), g["pre"])
        assert pre.data_ptr() != masks.data_ptr()
        np.testing.assert_array_equal(N(masks), g["masks"])  # input not mutated
        masks = pre
    dec = wseg_b200.run_pamr(pamr, image, masks)
    assert np.abs(N(dec) - g["masks_dec"]).max() <= TOL
    cleaned = wseg_b200.rescale_and_clean(dec, image, labels)
    assert np.abs(N(cleaned) - g["cleaned"]).max() <= TOL
    # materialised API-parity path
    pg = wseg_b200.pseudo_gtmask(cleaned)
    lab_api = N(wseg_b200.labels_from_pseudo_gt(pg)).astype(np.uint8)
    n_near = assert_labels_match(lab_api, g["label"], g["cleaned"])
    # fused path (no materialised up-sampled tensors)
    lab, onehot, counts = wseg_b200.pseudo_labels(dec, labels, image.shape[-2:], return_onehot=True, return_counts=True)
    lab = N(lab)
    assert_labels_match(lab, g["label"], g["cleaned"])
    np.testing.assert_array_equal(lab, lab_api)
    np.testing.assert_array_equal(N(onehot), N(pg))
    C = masks.shape[1]
    expect = np.stack([(lab == c).reshape(lab.shape[0], -1).sum(1) for c in range(C)], 1)
    np.testing.assert_array_equal(N(counts), expect)
    if "pre" not in g:
        lab2 = N(wseg_b200.refine_and_label(pamr, image, masks, labels))
        np.testing.assert_array_equal(lab2, lab)
    print("%s: near-threshold pixels %d of %d" % (name, n_near, lab.size))


@pytest.mark.parametrize("name", STAGE_CASES)
def test_epilogue_bit_exact_on_reference_masks(name):
    """Given the reference's own cleaned masks, thresholds/ambiguity/argmax are integer work: bit-exact."""
    g = load(name)
    lab, onehot = wseg_b200.pseudo_labels(G(g["cleaned"]), return_onehot=True)
    np.testing.assert_array_equal(N(lab), g["label"])
    np.testing.assert_array_equal(N(onehot).astype(np.uint8), g["pseudo_gt"])


# ---------------------------------------------------------------- CPU oracle on seeded inputs

@pytest.mark.parametrize("family", ["uniform", "structured", "quantised", "constant"])
@pytest.mark.parametrize("shape", [(2, 21, 64, 80), (1, 21, 161, 161)])
def test_pamr_vs_oracle_families(family, shape):
    B, C, H, W = shape
    image = {"uniform": lambda: synth.image_uniform(B, 3, H, W, 0),
             "structured": lambda: synth.image_structured(B, 3, H, W, 1),
             "quantised": lambda: synth.image_structured(B, 3, H, W, 2, quantise=True),
             "constant": lambda: synth.image_constant(B, 3, H, W)}[family]()
    mask = synth.mask_softmax(B, C, H, W, 10)
    ref = oracle.pamr_forward(image, mask, 10, D6)
    out = N(wseg_b200.PAMR(10, D6).to(DEV)(G(image), G(mask)))
    err = np.abs(out - ref).max()
    print("%s %s max-abs %.3g" % (family, shape, err))
    assert err <= TOL


@pytest.mark.parametrize("hw", [(1, 1), (2, 3), (8, 8), (24, 24), (25, 25), (31, 33), (47, 49), (3, 200), (200, 3),
                                (96, 96), (97, 129)])
def test_pamr_vs_oracle_edge_shapes(hw):
    H, W = hw
    image, mask = synth.image_structured(2, 3, H, W, 5), synth.mask_softmax(2, 21, H, W, 6)
    ref = oracle.pamr_forward(image, mask, 10, D6)
    out = N(wseg_b200.PAMR(10, D6).to(DEV)(G(image), G(mask)))
    assert np.abs(out - ref).max() <= TOL


@pytest.mark.parametrize("hw", [(41, 64), (72, 33), (45, 71), (88, 40), (83, 97), (321, 33)])
def test_pamr_vs_oracle_remainder_strips(hw):
    """Shapes whose remainders (<= 8 rows / columns past the last full tile) go to the strip kernels."""
    H, W = hw
    image, mask = synth.image_structured(2, 3, H, W, 15), synth.mask_softmax(2, 21, H, W, 16)
    ref = oracle.pamr_forward(image, mask, 10, D6)
    out, cmax = wseg_b200.PAMR(10, D6).to(DEV)(G(image), G(mask), return_class_max=True)
    assert np.abs(N(out) - ref).max() <= TOL
    L = _lib.lib()
    got = np.array([[L.pamr_float_from_ordered(int(v) & 0xffffffff) for v in row] for row in N(cmax)], dtype=np.float32)
    np.testing.assert_array_equal(got, N(out.flatten(2).max(-1).values))


@pytest.mark.parametrize("C", [1, 2, 7, 8, 20, 22, 40])
def test_pamr_vs_oracle_class_counts(C):
    image, mask = synth.image_uniform(1, 3, 50, 70, 7), synth.mask_softmax(1, C, 50, 70, 8)
    ref = oracle.pamr_forward(image, mask, 10, D6)
    out = N(wseg_b200.PAMR(10, D6).to(DEV)(G(image), G(mask)))
    assert np.abs(out - ref).max() <= TOL


@pytest.mark.parametrize("cfg", [(1, [1], 3), (3, [1, 3], 1), (2, [2, 5, 9], 4), (0, [1, 2], 3), (10, [1, 2, 4, 8, 12, 24, 32], 3)])
def test_pamr_vs_oracle_constructor_args(cfg):
    iters, dil, K = cfg
    image, mask = synth.image_uniform(2, K, 40, 44, 9), synth.mask_softmax(2, 5, 40, 44, 10)
    ref = oracle.pamr_forward(image, mask, iters, dil)
    out = N(wseg_b200.PAMR(iters, dil).to(DEV)(G(image), G(mask)))
    assert np.abs(out - ref).max() <= TOL


def test_each_iteration_vs_oracle():
    image, mask = synth.image_structured(1, 3, 90, 110, 3), synth.mask_softmax(1, 21, 90, 110, 4)
    aff_ref = oracle.affinity(image, D6)
    aff = wseg_b200.local_affinity(G(image), D6)
    assert np.abs(N(aff) - aff_ref).max() <= 1e-6
    np.testing.assert_allclose(N(aff).sum(1), 1.0, atol=2e-6)
    m_ref, m = mask, G(mask)
    for it in range(10):
        m_ref = oracle.propagate(aff_ref, m_ref, D6, 1)
        m = wseg_b200.propagate(aff, m, D6, 1)
        assert np.abs(N(m) - m_ref).max() <= TOL, "iteration %d" % it


def test_stage_vs_oracle_lowres_mask():
    """Real stage_net shape family: mask 41x41, image 161x161 (stride-4 analogue), labels gate."""
    B, C = 2, 21
    image, masks = synth.image_structured(B, 3, 161, 161, 11), synth.mask_blobs(B, C, 41, 41, 12)
    labels = synth.labels_bernoulli(B, C, 13, p=0.3)
    dec_ref = oracle.run_pamr(image, masks)
    cleaned_ref = oracle.rescale_and_clean(dec_ref, (161, 161), labels)
    lab_ref = oracle.pseudo_labels(cleaned_ref)
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    lab, dec = wseg_b200.refine_and_label(pamr, G(image), G(masks), G(labels), return_masks=True)
    assert np.abs(N(dec) - dec_ref).max() <= TOL
    assert_labels_match(N(lab), lab_ref, cleaned_ref)
    assert (lab_ref != 255).mean() > 0.05


def test_host_buffer_entry_point():
    """pamr_pseudo_labels_host_f32: plain host pointers in, uint8 labels out (the non-PyTorch binding)."""
    B, K, C, H, W, h, w = 2, 3, 21, 80, 96, 20, 24
    image, masks = synth.image_structured(B, K, H, W, 21), synth.mask_blobs(B, C, h, w, 22)
    labels = synth.labels_bernoulli(B, C, 23, p=0.3)
    out = np.zeros((B, H, W), dtype=np.uint8)
    dil = (ctypes.c_int * 6)(*D6)
    rc = _lib.lib().pamr_pseudo_labels_host_f32(image.ctypes.data, masks.ctypes.data, labels.ctypes.data, out.ctypes.data,
                                                B, K, C, H, W, h, w, dil, 6, 10, 0.7, 0.6, 0.2, 0)
    _lib.check(rc)
    cleaned_ref = oracle.rescale_and_clean(oracle.run_pamr(image, masks), (H, W), labels)
    assert_labels_match(out, oracle.pseudo_labels(cleaned_ref), cleaned_ref)
    # full-resolution variant (class max fused into the last propagation step)
    masks_f = synth.mask_blobs(B, C, H, W, 24)
    rc = _lib.lib().pamr_pseudo_labels_host_f32(image.ctypes.data, masks_f.ctypes.data, labels.ctypes.data,
                                                out.ctypes.data, B, K, C, H, W, H, W, dil, 6, 10, 0.7, 0.6, 0.2, 0)
    _lib.check(rc)
    cleaned_ref = oracle.rescale_and_clean(oracle.run_pamr(image, masks_f), (H, W), labels)
    assert_labels_match(out, oracle.pseudo_labels(cleaned_ref), cleaned_ref)


# ---------------------------------------------------------------- boundary behaviour

def test_inputs_not_mutated_and_fresh_output():
    image, mask = G(synth.image_uniform(1, 3, 40, 40, 1)), G(synth.mask_softmax(1, 4, 40, 40, 2))
    i0, m0 = image.clone(), mask.clone()
    out = wseg_b200.PAMR(10, D6).to(DEV)(image, mask)
    assert torch.equal(image, i0) and torch.equal(mask, m0)
    assert out.data_ptr() not in (image.data_ptr(), mask.data_ptr())
    assert out.dtype == torch.float32 and out.device == image.device and not out.requires_grad


def test_non_contiguous_and_requires_grad_inputs():
    image = G(synth.image_uniform(1, 3, 40, 48, 1))
    mask = G(synth.mask_softmax(1, 4, 48, 40, 2)).transpose(2, 3).requires_grad_(True)
    assert not mask.is_contiguous()
    out = wseg_b200.PAMR(2, D6).to(DEV)(image, mask)
    ref = oracle.pamr_forward(N(image), N(mask.detach()).copy(), 2, D6)
    assert np.abs(N(out) - ref).max() <= TOL


def test_bad_arguments_raise():
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    image = G(synth.image_uniform(2, 3, 16, 16, 1))
    with pytest.raises(RuntimeError):
        pamr(image, G(synth.mask_softmax(1, 4, 16, 16, 2)))  # batch mismatch
    with pytest.raises(RuntimeError):
        pamr(image.double(), G(synth.mask_softmax(2, 4, 16, 16, 2)))
    with pytest.raises(RuntimeError):
        pamr(image, torch.from_numpy(synth.mask_softmax(2, 4, 16, 16, 2)))  # CPU mask
    with pytest.raises(RuntimeError):
        wseg_b200.PAMR(1, [0]).to(DEV)(image, G(synth.mask_softmax(2, 4, 16, 16, 2)))  # bad dilation
    with pytest.raises(RuntimeError):
        wseg_b200.pseudo_labels(G(synth.mask_softmax(2, 4, 16, 16, 2)), labels=torch.ones(2, 7, device=DEV))


def test_runs_on_callers_stream_and_threads():
    import threading
    image, mask = G(synth.image_uniform(2, 3, 64, 64, 1)), G(synth.mask_softmax(2, 21, 64, 64, 2))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    ref = pamr(image, mask)
    torch.cuda.synchronize()
    outs = [None] * 4

    def work(i):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            outs[i] = pamr(image, mask)
        s.synchronize()

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for o in outs:
        assert torch.equal(o, ref)


def test_fused_class_max_matches_torch():
    image, mask = G(synth.image_uniform(2, 3, 70, 90, 1)), G(synth.mask_softmax(2, 21, 70, 90, 2) - 0.01)
    out, cmax = wseg_b200.PAMR(3, D6).to(DEV)(image, mask, return_class_max=True)
    L = _lib.lib()
    got = np.array([[L.pamr_float_from_ordered(int(v) & 0xffffffff) for v in row] for row in N(cmax)], dtype=np.float32)
    np.testing.assert_array_equal(got, N(out.flatten(2).max(-1).values))


# ---------------------------------------------------------------- properties at BASELINE.json sizes

@pytest.fixture(scope="module")
def config2():
    B, C, H, W = 16, 21, 321, 321
    image = G(synth.image_structured(B, 3, H, W, 31))
    mask = G(synth.mask_softmax(B, C, H, W, 32))
    labels = G(synth.labels_bernoulli(B, C, 33, p=0.3))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    out = pamr(image, mask)
    return image, mask, labels, pamr, out


def test_config2_convexity_and_mass(config2):
    image, mask, labels, pamr, out = config2
    lo, hi = mask.flatten(2).min(-1).values, mask.flatten(2).max(-1).values
    o_lo, o_hi = out.flatten(2).min(-1).values, out.flatten(2).max(-1).values
    assert bool((o_lo >= lo - 1e-6).all()) and bool((o_hi <= hi + 1e-6).all())  # convex combinations
    assert float((out.sum(1) - 1).abs().max()) <= 2e-5  # softmax input keeps sum_c = 1
    assert bool(torch.isfinite(out).all())


def test_config2_batch_shard_equivalence_and_determinism(config2):
    image, mask, labels, pamr, out = config2
    again = pamr(image, mask)
    assert torch.equal(again, out)
    halves = torch.cat([pamr(image[:8], mask[:8]), pamr(image[8:], mask[8:])], 0)
    assert torch.equal(halves, out)  # per-sample independence: sharding by batch is exact
    lab = wseg_b200.refine_and_label(pamr, image, mask, labels)
    lab_sh = torch.cat([wseg_b200.refine_and_label(pamr, image[i:i + 4], mask[i:i + 4], labels[i:i + 4])
                        for i in range(0, 16, 4)], 0)
    assert torch.equal(lab, lab_sh)
    vals = set(torch.unique(lab).tolist())
    assert vals <= set(range(21)) | {255}


def test_config2_linearity(config2):
    image, mask, labels, pamr, out = config2
    other = G(synth.mask_softmax(16, 21, 321, 321, 34))
    mix = pamr(image, 0.25 * mask + 0.75 * other)
    expect = 0.25 * out + 0.75 * pamr(image, other)
    assert float((mix - expect).abs().max()) <= TOL


def _check_against_oracle(tag, pamr, image, mask, labels, out, samples):
    """Masks (max-abs <= 1e-5) and labels (exact outside the near-threshold set) of the given samples against the
    CPU oracle; prints the size of the near-threshold set."""
    worst, near_total = 0.0, 0
    for i in samples:
        ref = oracle.pamr_forward(N(image[i:i + 1]), N(mask[i:i + 1]), 10, D6)
        worst = max(worst, float(np.abs(N(out[i:i + 1]) - ref).max()))
        assert worst <= TOL, (tag, i)
        H, W = ref.shape[-2:]
        cleaned_ref = oracle.rescale_and_clean(ref, (H, W), N(labels[i:i + 1]))
        lab = wseg_b200.refine_and_label(pamr, image[i:i + 1], mask[i:i + 1], labels[i:i + 1])
        near_total += assert_labels_match(N(lab), oracle.pseudo_labels(cleaned_ref), cleaned_ref)
    print("%s: %d samples vs oracle, masks max-abs %.3g, near-threshold pixels %d" % (tag, len(samples), worst, near_total))


def test_config2_every_sample_vs_oracle(config2):
    """BASELINE.json configs[1] (B=16, 321x321, 21 classes, 10 iterations): all 16 samples against the CPU oracle."""
    image, mask, labels, pamr, out = config2
    _check_against_oracle("config 2", pamr, image, mask, labels, out, range(16))


@pytest.mark.parametrize("side", [256, 512, 768, 1024])
def test_config3_multiscale_full_classes_vs_oracle(side):
    """BASELINE.json configs[2]: B=1, 21 classes at the four inference scales (infer_val.py multi-scale sweep)."""
    image = G(synth.image_structured(1, 3, side, side, side))
    mask = G(synth.mask_softmax(1, 21, side, side, side + 1))
    labels = G(synth.labels_bernoulli(1, 21, side + 2, p=0.3))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    _check_against_oracle("config 3 %dx%d" % (side, side), pamr, image, mask, labels, pamr(image, mask), [0])


def test_config4_highres_full_classes_vs_oracle():
    """BASELINE.json configs[3]: B=8, 1024x2048, 21 classes; every sample against the CPU oracle."""
    B, C, H, W = 8, 21, 1024, 2048
    image = torch.rand((B, 3, H, W), generator=torch.Generator(device=DEV).manual_seed(41), device=DEV)
    mask = torch.softmax(2.0 * torch.randn((B, C, H, W), generator=torch.Generator(device=DEV).manual_seed(42), device=DEV), 1)
    labels = G(synth.labels_bernoulli(B, C, 43, p=0.3))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    out = pamr(image, mask)
    _check_against_oracle("config 4", pamr, image, mask, labels, out, range(B))


def test_config5_b128_shards_and_samples_vs_oracle():
    """BASELINE.json configs[4]: B=128 at the VOC training shape.  The 8-way batch shards of the strong-scaling run
    are bit-identical to the single-GPU result, and 6 samples spread over the batch match the CPU oracle."""
    B, C, H, W = 128, 21, 321, 321
    g = torch.Generator(device=DEV).manual_seed(51)
    image = torch.rand((B, 3, H, W), generator=g, device=DEV)
    mask = torch.softmax(2.0 * torch.randn((B, C, H, W), generator=g, device=DEV), 1)
    labels = G(synth.labels_bernoulli(B, C, 53, p=0.3))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    out = pamr(image, mask)
    lab = wseg_b200.refine_and_label(pamr, image, mask, labels)
    for r in range(8):  # rank r of 8 takes [16 r, 16 r + 16)
        sl = slice(16 * r, 16 * r + 16)
        assert torch.equal(pamr(image[sl], mask[sl]), out[sl]), r
        assert torch.equal(wseg_b200.refine_and_label(pamr, image[sl], mask[sl], labels[sl]), lab[sl]), r
    _check_against_oracle("config 5", pamr, image, mask, labels, out, [0, 17, 63, 64, 100, 127])


def test_local_std_vs_oracle():
    """Row a4 (LocalStDev, pamr.py:77-103) on its own: the GPU's conditioned fp32 std against the oracle's (which is
    bit-equal to the reference's, tests/test_oracle_golden.py)."""
    for (fam, seed) in [("uniform", 3), ("structured", 4), ("quantised", 5), ("constant", 6)]:
        image = {"uniform": lambda: synth.image_uniform(2, 3, 90, 130, seed),
                 "structured": lambda: synth.image_structured(2, 3, 90, 130, seed),
                 "quantised": lambda: synth.image_structured(2, 3, 90, 130, seed, quantise=True),
                 "constant": lambda: synth.image_constant(2, 3, 90, 130)}[fam]()
        ref = oracle.local_std(image, D6)
        got = N(wseg_b200.local_std(G(image), D6))
        err = float(np.abs(got - ref).max())
        rel = float((np.abs(got - ref) / np.maximum(ref, 1e-3)).max())
        print("local std, %s: max-abs %.3g, max-rel %.3g (floor 1e-3)" % (fam, err, rel))
        assert got.shape == ref.shape and err <= 2e-7 and rel <= 2e-6
    # the module form returns the reference's shape [B,K,1,H,W] (pamr.py:103, keepdim over the sample axis)
    mod = wseg_b200.PAMR(10, D6).to(DEV).aff_std(G(image))
    assert tuple(mod.shape) == (2, 3, 1, 90, 130) and np.array_equal(N(mod)[:, :, 0], got)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_data_parallel_two_devices():
    """nn.DataParallel (reference train.py:112): one Python thread per GPU calls the module on its own device, so
    the C ABI runs with dev != the process's current device; results equal the single-device call."""
    B, C, H, W = 8, 21, 97, 129
    image, mask = G(synth.image_structured(B, 3, H, W, 81)), G(synth.mask_softmax(B, C, H, W, 82))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    ref = pamr(image, mask)
    dp = torch.nn.DataParallel(pamr, device_ids=[0, 1])
    for _ in range(3):
        out = dp(image, mask)
        assert out.device == ref.device and torch.equal(out, ref)
    # the resident small-map kernel (cooperative launch, per-device attribute caches) on both devices
    image, mask = G(synth.image_structured(B, 3, 41, 41, 83)), G(synth.mask_softmax(B, C, 41, 41, 84))
    assert torch.equal(dp(image, mask), pamr(image, mask))
    # a call on cuda:1 while cuda:0 is current
    i1, m1 = image.to("cuda:1"), mask.to("cuda:1")
    assert torch.equal(pamr.to("cuda:1")(i1, m1).to(DEV), pamr.to(DEV)(image, mask))


def test_constant_mask_is_fixed_point():
    image = G(synth.image_structured(1, 3, 200, 300, 41))
    mask = torch.full((1, 3, 200, 300), 0.375, device=DEV)
    out = wseg_b200.PAMR(10, D6).to(DEV)(image, mask)
    assert float((out - 0.375).abs().max()) <= 5e-6  # weights sum to 1 only up to fp32 rounding, 10 iterations


def test_host_pipeline_matches_device_path():
    """HostPipeline (pinned host inputs, chunked copies overlapped with compute) == refine_and_label."""
    B, C, H, W = 6, 21, 96, 130
    image, masks = synth.image_structured(B, 3, H, W, 51), synth.mask_blobs(B, C, H, W, 52)
    labels = synth.labels_bernoulli(B, C, 53, p=0.3)
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    ref = N(wseg_b200.refine_and_label(pamr, G(image), G(masks), G(labels)))
    pin = lambda a: torch.from_numpy(a).pin_memory()
    pipe = wseg_b200.HostPipeline(pamr, DEV, chunks=4)
    for _ in range(2):  # second call re-uses the staging buffers
        out = pipe(pin(image), pin(masks), pin(labels))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.numpy(), ref)
    d_out = torch.empty((B, H, W), dtype=torch.uint8, device=DEV)
    pipe(pin(image), pin(masks), pin(labels), d_out=d_out)
    np.testing.assert_array_equal(N(d_out), ref)


def test_repeatability_under_allocator_churn():
    """Race detector for the tile kernel's TMA / mbarrier ring: the same forward, repeated with
    unrelated allocations in between, must be bit-identical every time.  (A ring whose parity waits
    could alias -- 4 barrier pairs for 4 slots -- produced one wrong tile-class in 0.2-2.5 % of the
    calls at this shape; with 8 barrier pairs 0 of 5500.)"""
    B, C, H, W = 16, 21, 320, 320
    image = G(synth.image_uniform(B, 3, H, W, 51))
    mask = G(synth.mask_softmax(B, C, H, W, 52))
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    ref = pamr(image, mask).clone()
    bad = 0
    for _ in range(400):
        junk = torch.rand((B, C, H, W), device=DEV)  # noqa: F841  (L2 / allocator churn)
        bad += int(not torch.equal(pamr(image, mask), ref))
    assert bad == 0


def test_cuda_graph_capture_and_replay():
    """The whole refine_and_label sequence (side-stream fork/join, TMA descriptors passed by value) can be
    captured into a CUDA graph and replayed on new inputs."""
    B, C, H, W = 4, 21, 81, 97
    pamr = wseg_b200.PAMR(10, D6).to(DEV)
    s_img = G(synth.image_structured(B, 3, H, W, 61))
    s_msk = G(synth.mask_softmax(B, C, H, W, 62))
    s_lab = G(synth.labels_bernoulli(B, C, 63, p=0.3))
    wseg_b200.refine_and_label(pamr, s_img, s_msk, s_lab)  # warm-up outside the capture (attributes, side stream)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s_out = wseg_b200.refine_and_label(pamr, s_img, s_msk, s_lab)
    for seed in (71, 72):
        img, msk = G(synth.image_structured(B, 3, H, W, seed)), G(synth.mask_softmax(B, C, H, W, seed + 10))
        s_img.copy_(img); s_msk.copy_(msk)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(s_out, wseg_b200.refine_and_label(pamr, img, msk, s_lab))


def test_plain_c_caller_runs(tmp_path):
    """examples/c_abi_demo.c: a C99 program that only knows include/pamr_b200.h, host buffers in, labels out."""
    import subprocess
    from test_host import _build_c_demo
    r = subprocess.run([_build_c_demo(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kernels launched" in r.stdout and "label" in r.stdout


def test_random_shapes_vs_oracle():
    """Shape fuzz (tail / side lane / strip / padded-tile / generic paths are chosen per shape by tuned_tiling)."""
    rng = np.random.RandomState(123)
    for _ in range(40):
        B, C = int(rng.randint(1, 7)), int(rng.randint(1, 25))
        H, W = int(rng.randint(8, 200)), int(rng.randint(8, 200))
        if rng.rand() < 0.3:
            H, W = 40 * int(rng.randint(1, 5)) + int(rng.randint(0, 9)), 32 * int(rng.randint(1, 6)) + int(rng.randint(0, 9))
        it = int(rng.choice([1, 2, 5]))
        image = rng.rand(B, 3, H, W).astype(np.float32)
        e = np.exp(rng.randn(B, C, H, W)).astype(np.float32)
        mask = e / e.sum(1, keepdims=True)
        out = N(wseg_b200.PAMR(it, D6).to(DEV)(G(image), G(mask)))
        assert np.abs(out - oracle.pamr_forward(image, mask, it, D6)).max() <= TOL, (B, C, H, W, it)


# ---------------------------------------------------------------- resized epilogue (column-walk kernels), bit-exact vs oracle

@pytest.mark.parametrize("shape", [
    (2, 21, 81, 81, 321, 321),   # stage_net, stride-4 masks
    (3, 21, 41, 41, 321, 321),   # stride-8 masks
    (2, 5, 20, 30, 97, 65),      # non-integer enlargement
    (1, 21, 33, 47, 40, 50),     # barely enlarged: the source row pair changes on almost every row
    (2, 21, 64, 64, 32, 32),     # reduction: every output row has a new source row pair
    (1, 22, 16, 16, 64, 64),     # more than 21 classes: the one-thread-per-pixel label kernel
    (2, 1, 9, 9, 40, 40),        # a single class (no gated class at all)
    (1, 21, 1, 1, 17, 13),       # 1x1 source
    (2, 4, 10, 3, 300, 5),       # narrow map: a block of the store pass spans many bands (rows computed in place)
    (1, 21, 8, 8, 8, 64),        # enlarged in x only
    (1, 3, 5, 7, 1, 1),          # 1x1 target (scale 0)
])
def test_resized_epilogue_bit_exact_vs_oracle(shape):
    """_rescale_and_clean -> pseudo_gtmask -> argmax at a target size different from the mask size (SoftMaxAE.py:263-268,
    29-50, 61-67): cleaned masks, class maxima, labels, one-hot and counts are integer / single-expression work and
    must be bit-identical to the oracle, with and without label gating."""
    B, C, h, w, H, W = shape
    masks = synth.mask_blobs(B, C, h, w, 31) if min(h, w) >= 8 else synth.mask_softmax(B, C, h, w, 31)
    L = _lib.lib()
    for labels in ([synth.labels_bernoulli(B, C, 32, p=0.4), None] if C > 1 else [None]):
        lab_np = labels if labels is not None else np.ones((B, max(C - 1, 0)), np.float32)
        ref_clean = oracle.rescale_and_clean(masks, (H, W), lab_np) if C > 1 else oracle.resize_bilinear(masks, (H, W))
        ref_lab, ref_pg = oracle.pseudo_labels(ref_clean), oracle.pseudo_gtmask(ref_clean)
        lab_t = G(labels) if labels is not None else None
        cleaned, cmax = wseg_b200.rescale_and_clean(G(masks), (H, W), lab_t, return_class_max=True)
        np.testing.assert_array_equal(N(cleaned), ref_clean)
        got_max = np.array([[L.pamr_float_from_ordered(int(v) & 0xffffffff) for v in row] for row in N(cmax)], dtype=np.float32)
        np.testing.assert_array_equal(got_max, ref_clean.reshape(B, C, -1).max(-1))
        lab, onehot, counts = wseg_b200.pseudo_labels(G(masks), lab_t, (H, W), return_onehot=True, return_counts=True)
        np.testing.assert_array_equal(N(lab), ref_lab)
        np.testing.assert_array_equal(N(onehot), ref_pg)
        np.testing.assert_array_equal(N(counts), np.stack([(ref_lab == c).reshape(B, -1).sum(1) for c in range(C)], 1))
        np.testing.assert_array_equal(N(wseg_b200.pseudo_labels(G(masks), lab_t, (H, W))), ref_lab)  # labels only
        np.testing.assert_array_equal(N(wseg_b200.resize_bilinear(G(masks), (H, W))), oracle.resize_bilinear(masks, (H, W)))
