"""balanced_mask_loss_ce (models/SoftMaxAE.py:52-88) forward + gradient.

CPU part: the oracle (oracle/pamr_oracle.c:pamr_oracle_mask_ce) against golden vectors produced by the
reference's own function under torch autograd (oracle/gen_golden_loss.py).
GPU part: the CUDA path (through the C ABI) against those goldens and against the oracle on seeded
inputs, plus the properties the reference's definition implies.
Tolerances: loss 2e-6 relative, gradient 2e-6 of the gradient's max magnitude (fp32 exp/log and the
summation order differ; nothing else does)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz")))
RTOL = 2e-6


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def close(a, b, scale=None):
    scale = float(np.abs(b).max()) if scale is None else scale
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max()) <= RTOL * max(scale, 1e-30)


def test_loss_golden_inventory():
    assert len(LOSS_CASES) == 4


@pytest.mark.parametrize("name", LOSS_CASES)
def test_oracle_loss_matches_reference(name):
    g = load(name)
    loss, grad = oracle.balanced_mask_loss_ce(g["logits"], g["pseudo_gt"].astype(np.float32), g["gt_labels"], g["gout"])
    assert close(loss, g["loss"], scale=max(1.0, float(np.abs(g["loss"]).max())) if not np.any(g["loss"]) else None)
    assert close(grad, g["grad"])
    only = oracle.balanced_mask_loss_ce(g["logits"], g["pseudo_gt"].astype(np.float32), g["gt_labels"])
    assert np.array_equal(only, loss)


def _random_case(seed, B, C, h, w, H, W, p_ignore=0.3):
    rng = np.random.RandomState(seed)
    logits = (2.5 * rng.randn(B, C, h, w)).astype(np.float32)
    lab = rng.randint(0, C, size=(B, H, W))
    lab[rng.rand(B, H, W) < p_ignore] = 255
    pg = np.zeros((B, C, H, W), dtype=np.float32)
    for c in range(C):
        pg[:, c][lab == c] = 1.0
    n = pg.reshape(B, C, -1).sum(-1)
    gl = (n[:, 1:] > 0).astype(np.float32)
    if B > 1:
        gl[-1] = 1.0 - gl[-1]  # last sample: label set differs from the pseudo mask -> batch weight 0
    gout = (0.5 + rng.rand(B)).astype(np.float32)
    return logits, pg, lab.astype(np.uint8), gl, gout


def test_oracle_loss_properties():
    logits, pg, lab, gl, gout = _random_case(3, 3, 6, 9, 11, 20, 23)
    loss, grad = oracle.balanced_mask_loss_ce(logits, pg, gl, gout)
    assert loss[-1] == 0.0 and not grad[-1].any()           # batch weight 0: no loss, no gradient
    assert (loss[:-1] > 0).all()
    # softmax - onehot sums to zero over classes at every label pixel, and so does its interpolation transpose
    assert np.abs(grad.sum(1)).max() <= 1e-9
    # shifting all logits of a pixel column by a constant leaves the loss unchanged
    loss2 = oracle.balanced_mask_loss_ce(logits + 7.0, pg, gl)
    assert close(loss2, loss)
    # finite-difference check of one entry (double-accumulated oracle: tight)
    e = np.zeros_like(logits); e[0, 2, 4, 5] = 1e-2
    lp = oracle.balanced_mask_loss_ce(logits + e, pg, gl); lm = oracle.balanced_mask_loss_ce(logits - e, pg, gl)
    fd = float(((lp - lm) * gout).sum() / 2e-2)
    assert abs(fd - grad[0, 2, 4, 5]) <= 2e-3 * abs(grad[0, 2, 4, 5]) + 1e-7


# ------------------------------------------------------------------------------------------ GPU
torch = pytest.importorskip("torch")
DEV = "cuda:0"


def G(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _gpu_loss_and_grad(logits, pg, gl, gout, from_labels=None):
    import wseg_b200
    x = G(logits).requires_grad_(True)
    if from_labels is None:
        loss = wseg_b200.balanced_mask_loss_ce(x, G(pg), G(gl))
    else:
        lab, cnt = from_labels
        loss = wseg_b200.balanced_mask_loss_ce_from_labels(x, lab, cnt, G(gl))
    (loss * G(gout)).sum().backward()
    return loss.detach().cpu().numpy(), x.grad.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOSS_CASES)
def test_gpu_loss_vs_reference_golden(name):
    g = load(name)
    loss, grad = _gpu_loss_and_grad(g["logits"], g["pseudo_gt"].astype(np.float32), g["gt_labels"], g["gout"])
    assert close(loss, g["loss"], scale=max(1.0, float(np.abs(g["loss"]).max())))
    assert close(grad, g["grad"])


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(2, 21, 41, 41, 161, 161), (3, 21, 33, 37, 33, 37), (2, 5, 50, 40, 17, 13),
                                  (1, 2, 1, 1, 9, 7), (2, 40, 11, 13, 40, 45), (4, 21, 81, 81, 321, 321),
                                  (2, 7, 20, 30, 97, 65), (1, 21, 8, 40, 33, 40), (2, 21, 16, 9, 40, 8)])
def test_gpu_loss_vs_oracle(dims):
    B, C, h, w, H, W = dims
    logits, pg, lab, gl, gout = _random_case(11 + h, B, C, h, w, H, W)
    o_loss, o_grad = oracle.balanced_mask_loss_ce(logits, pg, gl, gout)
    loss, grad = _gpu_loss_and_grad(logits, pg, gl, gout)
    assert close(loss, o_loss, scale=max(1.0, float(np.abs(o_loss).max())))
    assert close(grad, o_grad)


@pytest.mark.gpu
def test_gpu_loss_fused_label_path_and_helpers():
    import wseg_b200
    B, C, h, w, H, W = 2, 21, 21, 25, 81, 97
    logits, pg, lab, gl, gout = _random_case(5, B, C, h, w, H, W)
    d_lab, d_cnt = wseg_b200.labels_from_onehot(G(pg))
    assert np.array_equal(d_lab.cpu().numpy(), lab)
    assert np.array_equal(d_cnt.cpu().numpy(), pg.reshape(B, C, -1).sum(-1).astype(np.int32))
    a = _gpu_loss_and_grad(logits, pg, gl, gout)
    b = _gpu_loss_and_grad(logits, pg, gl, gout, from_labels=(d_lab, d_cnt))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    again = _gpu_loss_and_grad(logits, pg, gl, gout)
    assert np.array_equal(a[1], again[1])  # the backward pass is a deterministic gather
    # labels + counts straight from the pseudo-label kernel feed the loss without a one-hot tensor
    masks = torch.softmax(G(2.0 * np.random.RandomState(1).randn(B, C, H, W).astype(np.float32)), 1)
    lab2, cnt2 = wseg_b200.pseudo_labels(masks, None, None, None, return_counts=True)
    loss = wseg_b200.balanced_mask_loss_ce_from_labels(G(logits), lab2, cnt2, G(gl))
    assert loss.shape == (B,) and bool(torch.isfinite(loss).all())


@pytest.mark.gpu
def test_gpu_loss_bad_arguments_raise():
    import wseg_b200
    logits, pg, lab, gl, gout = _random_case(6, 2, 4, 5, 6, 10, 12)
    with pytest.raises(RuntimeError):
        wseg_b200.balanced_mask_loss_ce(torch.from_numpy(logits), G(pg), G(gl))            # CPU logits
    with pytest.raises(RuntimeError):
        wseg_b200.balanced_mask_loss_ce(G(logits), G(pg), G(gl[:, :2]))                    # wrong gt_labels shape
    with pytest.raises(RuntimeError):
        wseg_b200.balanced_mask_loss_ce(G(logits).double(), G(pg), G(gl))                  # wrong dtype
