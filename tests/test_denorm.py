"""the data set's denorm (datasets/pascal_voc.py:85-101, train.py:120) folded into the image
resize of run_pamr (SoftMaxAE.py:177).  Oracle vs goldens from the reference (CPU); CUDA path vs both (GPU).
The denorm itself is bit-exact; the resize carries the 2.5e-7 tolerance of the other resize tests (torch's
vectorised CPU kernel rounds one intermediate differently)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "denorm_*.npz")))
TOL = 2.5e-7 * 4  # values reach ~|3| after denorm of unit-normal inputs: scale the resize tolerance


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_denorm_matches_reference(name):
    g = load(name)
    assert np.array_equal(oracle.denorm_resize(g["image"], g["mean"], g["std"]), g["raw"])
    out = oracle.denorm_resize(g["image"], g["mean"], g["std"], g["out"].shape[-2:])
    assert out.shape == g["out"].shape and np.abs(out - g["out"]).max() <= TOL


torch = pytest.importorskip("torch")
DEV = "cuda:0"


def G(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_denorm_vs_reference_and_oracle(name):
    import wseg_b200
    g = load(name)
    x = G(g["image"])
    keep = x.clone()
    raw = wseg_b200.denorm_resize(x, g["mean"].tolist(), g["std"].tolist())
    assert np.array_equal(raw.cpu().numpy(), g["raw"]) and torch.equal(x, keep)  # bit-exact, input untouched
    size = g["out"].shape[-2:]
    out = wseg_b200.denorm_resize(x, g["mean"].tolist(), g["std"].tolist(), size).cpu().numpy()
    assert np.abs(out - g["out"]).max() <= TOL
    assert np.array_equal(out, oracle.denorm_resize(g["image"], g["mean"], g["std"], size))  # same expression as the oracle


@pytest.mark.gpu
def test_gpu_denorm_folded_into_refine_and_label():
    import synth
    import wseg_b200
    B, C, H, W, h, w = 2, 21, 65, 77, 17, 20
    norm = np.random.RandomState(3).randn(B, 3, H, W).astype(np.float32)
    masks = synth.mask_blobs(B, C, h, w, 4)
    labels = synth.labels_bernoulli(B, C, 5, p=0.3)
    pamr = wseg_b200.PAMR(10, [1, 2, 4, 8, 12, 24]).to(DEV)
    raw = wseg_b200.denorm_resize(G(norm))                                   # what train.py:120 hands to the model
    a = wseg_b200.refine_and_label(pamr, raw, G(masks), G(labels))
    b = wseg_b200.refine_and_label(pamr, G(norm), G(masks), G(labels), denorm=(wseg_b200.VOC_MEAN, wseg_b200.VOC_STD))
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        wseg_b200.denorm_resize(G(norm), (0.5,), (0.5,))                    # one entry per channel
