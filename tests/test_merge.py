"""MergeMultiScale._merge_masks (utils/inference_tools.py:134-161) + the no-CRF
prediction of ResultWriter.save (:85-88).

CPU part: the oracle against golden vectors produced by the reference's own class (oracle/gen_golden_merge.py).
GPU part: the CUDA path (through the C ABI) against goldens and oracle.
Tolerance: merged scores 3e-7 absolute (powf implementations differ by an ulp; with BG_POW = 1 the path is
bit-exact); predictions identical except where the top two scores are within that tolerance of each other
or of the threshold."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MERGE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "merge_*.npz")))
ATOL = 3e-7


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def decisive(merged, thresh, tol=1e-6):
    """Pixels whose prediction cannot flip under a perturbation of `tol`."""
    m = merged.astype(np.float64).copy()
    near_thr = (np.abs(m[1:] - thresh) < tol).any(0) if thresh > 0 else np.zeros(m.shape[1:], dtype=bool)
    m[1:][m[1:] < thresh] = 0
    srt = np.sort(m, 0)
    return (srt[-1] - srt[-2] > 2 * tol) & ~near_thr


def test_merge_golden_inventory():
    assert len(MERGE_CASES) == 4


@pytest.mark.parametrize("name", MERGE_CASES)
def test_oracle_merge_matches_reference(name):
    g = load(name)
    H, W = g["merged"].shape[-2:]
    merged, pred = oracle.merge_multiscale(g["masks"], g["pads"], g["labels"], (H, W), int(g["flip"]), float(g["bg_pow"]),
                                           float(g["thresh"]))
    assert np.abs(merged - g["merged"]).max() <= ATOL
    if float(g["bg_pow"]) == 1.0:
        assert np.array_equal(merged, g["merged"])  # interpolation, flip, gate and mean are bit-exact
    ok = decisive(g["merged"], float(g["thresh"]))
    assert ok.mean() > 0.9 and np.array_equal(pred[ok], g["pred"][ok])


def _random_case(seed, S, C, Hp, Wp, H, W, flip):
    rng = np.random.RandomState(seed)
    masks = np.zeros((S, C, Hp, Wp), dtype=np.float32)
    pads = np.zeros((S, 4), dtype=np.int32)
    for s in range(S):
        h, w = rng.randint(max(1, H // 2), Hp + 1), rng.randint(max(1, W // 2), Wp + 1)
        pt, pl = rng.randint(0, Hp - h + 1), rng.randint(0, Wp - w + 1)
        e = np.exp(2.0 * rng.randn(C, h, w)).astype(np.float32)
        masks[s, :, pt:pt + h, pl:pl + w] = e / e.sum(0, keepdims=True)
        pads[s] = (pt, pl, h, w)
    labels = (rng.rand(C - 1) < 0.5).astype(np.float32)
    return masks, pads, labels


# ------------------------------------------------------------------------------------------ GPU
torch = pytest.importorskip("torch")
DEV = "cuda:0"


def G(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MERGE_CASES)
def test_gpu_merge_vs_reference_golden(name):
    import wseg_b200
    g = load(name)
    H, W = g["merged"].shape[-2:]
    flip, bg_pow, thr = int(g["flip"]), float(g["bg_pow"]), float(g["thresh"])
    pred, merged = wseg_b200.merge_and_predict(G(g["masks"]), G(g["labels"]), g["pads"].tolist(), (H, W), thr, flip=flip,
                                               bg_pow=bg_pow, return_merged=True)
    merged, pred = merged.cpu().numpy(), pred.cpu().numpy()
    assert np.abs(merged - g["merged"]).max() <= ATOL
    if bg_pow == 1.0:
        assert np.array_equal(merged, g["merged"])
    ok = decisive(g["merged"], thr)
    assert np.array_equal(pred[ok], g["pred"][ok])
    only = wseg_b200.merge_masks(G(g["masks"]), G(g["labels"]), torch.from_numpy(g["pads"]), (H, W), flip=flip, bg_pow=bg_pow)
    assert np.array_equal(only.cpu().numpy(), merged)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(8, 21, 96, 128, 75, 100, True), (4, 21, 64, 64, 64, 64, False), (1, 2, 5, 7, 11, 3, False),
                                 (16, 3, 40, 40, 33, 29, True), (8, 21, 256, 256, 200, 250, True)])
def test_gpu_merge_vs_oracle(cfg):
    import wseg_b200
    S, C, Hp, Wp, H, W, flip = cfg
    masks, pads, labels = _random_case(7 + S, S, C, Hp, Wp, H, W, flip)
    o_merged, o_pred = oracle.merge_multiscale(masks, pads, labels, (H, W), flip, 3.0, 0.3)
    pred, merged = wseg_b200.merge_and_predict(G(masks), G(labels), pads.tolist(), (H, W), 0.3, flip=flip, bg_pow=3,
                                               return_merged=True)
    merged, pred = merged.cpu().numpy(), pred.cpu().numpy()
    assert np.abs(merged - o_merged).max() <= ATOL
    ok = decisive(o_merged, 0.3)
    assert ok.mean() > 0.9 and np.array_equal(pred[ok], o_pred[ok])
    # no labels = no gate; exponent 1 = plain mean: bit-exact against the oracle
    o2, _ = oracle.merge_multiscale(masks, pads, None, (H, W), flip, 1.0, 0.0)
    m2 = wseg_b200.merge_masks(G(masks), None, pads.tolist(), (H, W), flip=flip, bg_pow=1)
    assert np.array_equal(m2.cpu().numpy(), o2)


@pytest.mark.gpu
def test_gpu_merge_bad_arguments_raise():
    import wseg_b200
    masks, pads, labels = _random_case(3, 2, 4, 16, 16, 10, 12, False)
    with pytest.raises(RuntimeError):
        wseg_b200.merge_masks(torch.from_numpy(masks), G(labels), pads.tolist(), (10, 12))       # CPU scores
    bad = pads.copy(); bad[1] = (10, 10, 10, 10)                                                  # outside the padded mask
    with pytest.raises(RuntimeError):
        wseg_b200.merge_masks(G(masks), G(labels), bad.tolist(), (10, 12))
    with pytest.raises(RuntimeError):
        wseg_b200.merge_masks(G(masks), G(labels[:2]), pads.tolist(), (10, 12))                  # wrong labels shape
    with pytest.raises(RuntimeError):
        wseg_b200.merge_masks(G(masks), G(labels), pads[:1].tolist(), (10, 12))                  # pads for one scale only
