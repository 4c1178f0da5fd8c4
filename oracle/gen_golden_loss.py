"""Golden vectors for SURVEY 8(f) row 2 -- balanced_mask_loss_ce (models/SoftMaxAE.py:52-88) and its
gradient w.r.t. the mask logits -- FROM THE REFERENCE ITSELF (CPU fp32, torch autograd).

Build container only (needs /root/reference):   python oracle/gen_golden_loss.py
Writes tests/golden/loss_*.npz; oracle/pamr_oracle.c:pamr_oracle_mask_ce and the CUDA path are tested
against them.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PAMR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from models.SoftMaxAE import balanced_mask_loss_ce, pseudo_gtmask  # noqa: E402  (reference)

OUT = os.path.join(ROOT, "tests", "golden")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def case(name, logits, pseudo_gt, gt_labels, seed):
    """pseudo_gt: float one-hot-or-empty [B,C,H,W]; logits [B,C,h,w]; gt_labels [B,C-1]."""
    rng = np.random.RandomState(seed)
    gout = (0.25 + rng.rand(logits.shape[0])).astype(np.float32)
    x = T(logits).clone().requires_grad_(True)
    loss = balanced_mask_loss_ce(x, T(pseudo_gt), T(gt_labels))
    (loss * T(gout)).sum().backward()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, logits=logits, pseudo_gt=pseudo_gt.astype(np.uint8), gt_labels=gt_labels, gout=gout,
                        loss=loss.detach().numpy(), grad=x.grad.numpy())
    print("%-30s %7.1f KB  loss %s" % (name, os.path.getsize(path) / 1024, loss.detach().numpy()))


def present(pg):
    """gt_labels [B,C-1] that make batch_weight 1: exactly the foreground classes present in pseudo_gt."""
    n = pg.reshape(pg.shape[0], pg.shape[1], -1).sum(-1)
    return (n[:, 1:] > 0).astype(np.float32)


def main():
    rng = np.random.RandomState(7)
    # 1. the stage_net call shape family: logits at 21x25, pseudo_gt at 81x97 (from the seqA golden), 21 classes
    g = np.load(os.path.join(OUT, "stage_seqA_21x25_to_81x97.npz"))
    pg = g["pseudo_gt"].astype(np.float32)
    logits = (3.0 * rng.randn(2, 21, 21, 25)).astype(np.float32)
    gl = present(pg)
    gl[1] = g["labels"][1]  # second sample: the data set's labels (batch weight 0 unless they happen to match)
    case("loss_seqA_21x25_to_81x97", logits, pg, gl, 1)

    # 2. same resolution, 5 classes, one sample whose pseudo_gt is empty (everything ignored)
    with torch.no_grad():
        pg = pseudo_gtmask(T(synth.mask_blobs(3, 5, 24, 31, 41))).numpy()
    pg[2] = 0.0
    logits = (2.0 * rng.randn(3, 5, 24, 31)).astype(np.float32)
    case("loss_sameres_24x31", logits, pg, present(pg), 2)

    # 3. non-integer ratio, 4 classes, large logits (tests the log-sum-exp shift)
    with torch.no_grad():
        pg = pseudo_gtmask(T(synth.mask_blobs(2, 4, 40, 33, 42))).numpy()
    logits = (30.0 * rng.randn(2, 4, 13, 17)).astype(np.float32)
    case("loss_13x17_to_40x33", logits, pg, present(pg), 3)

    # 4. down-sampling direction (logits finer than the labels) and B=1
    with torch.no_grad():
        pg = pseudo_gtmask(T(synth.mask_blobs(1, 6, 13, 17, 43))).numpy()
    logits = rng.randn(1, 6, 40, 33).astype(np.float32)
    case("loss_40x33_to_13x17", logits, pg, present(pg), 4)


if __name__ == "__main__":
    main()
