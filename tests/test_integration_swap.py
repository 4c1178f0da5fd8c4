"""INTEGRATION.md section 2 executed: the two-line swap applied to the imported, otherwise unmodified reference.

Runs only where /root/reference exists (the build container; never on the GPU box).  In a child process:
  1. build the reference's SoftMaxAE (models/SoftMaxAE.py:118-174, ResNet38 backbone) with the reference's own PAMR and
     keep its full state dict -- what utils/checkpoints.py:99 would torch.load from a snapshot;
  2. purge the reference package from sys.modules, put the shim of INTEGRATION.md section 2 in place of
     models/mods/pamr.py, import the reference again and build the same model;
  3. load_state_dict(strict=True) of (1) into (2), as utils/checkpoints.py:99 does; key sets, shapes and the PAMR
     buffers must be identical, the model's `_aff` must be the B200 module, and the swapped helper trio must bind.
No forward pass here: the product has no CPU path (the forward of the swapped module is covered by the GPU tests)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"

CHILD = textwrap.dedent(r"""
    import sys, types
    sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
    import torch

    class Cfg:  # the fields the model constructors read (core/config.py:79-93)
        MODEL = "ae"; BACKBONE = "resnet38"; PAMR_ITER = 10; PAMR_KERNEL = [1, 2, 4, 8, 12, 24]
        PRE_WEIGHTS_PATH = None; MASK_LOSS_BCE = 1.0; FOCAL_P = 3; FOCAL_LAMBDA = 0.01; SG_PSI = 0.3

    def build():
        import models.backbones.base_net as base_net
        base_net.BaseNet._init_weights = lambda self, path: None  # no ImageNet snapshot here (and none is needed)
        from models.SoftMaxAE import network_SoftMaxAE
        torch.manual_seed(0)
        return network_SoftMaxAE(Cfg)(Cfg, None, 21)

    ref_model = build()
    import models.mods.pamr as ref_pamr
    assert type(ref_model._aff) is ref_pamr.PAMR and ref_pamr.__file__.startswith(%(ref)r)
    snapshot = {k: v.clone() for k, v in ref_model.state_dict().items()}

    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        del sys.modules[name]
    # INTEGRATION.md section 2: the body of models/mods/pamr.py becomes this one import
    shim = types.ModuleType("models.mods.pamr")
    exec("from wseg_b200.pamr import PAMR, LocalAffinity, LocalAffinityCopy, LocalStDev, LocalAffinityAbs  # noqa: F401",
         shim.__dict__)
    sys.modules["models.mods.pamr"] = shim

    model = build()
    import wseg_b200
    assert type(model._aff) is wseg_b200.PAMR, type(model._aff)
    assert model._aff.num_iter == 10 and list(model._aff.dilations) == Cfg.PAMR_KERNEL
    own = model.state_dict()
    assert list(own.keys()) == list(snapshot.keys())
    assert all(own[k].shape == snapshot[k].shape and own[k].dtype == snapshot[k].dtype for k in own)
    res = model.load_state_dict(snapshot, strict=True)  # utils/checkpoints.py:99
    assert not res.missing_keys and not res.unexpected_keys
    aff_keys = [k for k in own if k.startswith("_aff.")]
    assert sorted(aff_keys) == ["_aff.aff_m.kernel", "_aff.aff_std.kernel", "_aff.aff_x.kernel"], aff_keys
    for k in aff_keys:
        assert torch.equal(model.state_dict()[k], snapshot[k]), k
    # a tampered PAMR buffer in a snapshot is refused (the reference asserts this on every forward, pamr.py:42-43)
    bad = dict(snapshot); bad["_aff.aff_m.kernel"] = snapshot["_aff.aff_m.kernel"] + 1
    try:
        model.load_state_dict(bad, strict=True)
    except (RuntimeError, AssertionError):
        pass
    else:
        raise AssertionError("tampered PAMR buffer accepted")
    # nn.DataParallel replication (train.py:112) keeps the module type and its buffers
    rep = torch.nn.DataParallel(model)
    assert type(rep.module._aff) is wseg_b200.PAMR
    # the optional helper swap of section 2 binds against the model's own attributes
    from wseg_b200.stage import run_pamr, rescale_and_clean, pseudo_gtmask, refine_and_label  # noqa: F401
    import inspect
    assert list(inspect.signature(run_pamr).parameters)[:3] == ["pamr", "im", "mask"]
    print("swap ok: %%d tensors strict-loaded, %%d PAMR buffers" %% (len(own), len(aff_keys)))
""")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "models", "mods")),
                    reason="the reference tree exists only in the build container")
def test_reference_model_with_swapped_pamr_strict_loads_reference_snapshot():
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "ref": REFERENCE}], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "swap ok" in r.stdout
