"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, the
nn.Module mirrors the reference's state dict, the product never touches the oracle, and the
batch-shard / gather logic works with world_size 2 on gloo.  No GPU compute is attempted."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest
import torch

import wseg_b200
from wseg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "1-stage-wseg_b200")
D6 = [1, 2, 4, 8, 12, 24]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def header_functions():
    src = open(os.path.join(ROOT, "include", "pamr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pamr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = header_functions()
    assert len(names) >= 12
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libpamr_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding lacks %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.pamr_b200_abi_version() == _lib.ABI_VERSION


def test_library_is_sm100a_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % _lib.LIB_PATH).read()
    if not out.strip():
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_ordered_encoding_monotone(lib):
    vals = np.array([-np.inf, -3.5, -1e-30, -0.0, 0.0, 1e-30, 0.2, 0.7, 1.0, 3e38, np.inf], dtype=np.float32)
    enc = [lib.pamr_ordered_from_float(float(v)) for v in vals]
    assert all(a <= b for a, b in zip(enc, enc[1:]))
    assert all(e > 0 for e in enc)
    for v, e in zip(vals, enc):
        assert lib.pamr_float_from_ordered(e) == v


def test_no_gpu_means_error_not_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x = np.zeros(16, dtype=np.float32)
    rc = lib.pamr_resize_bilinear_f32(x.ctypes.data, x.ctypes.data, 1, 2, 2, 4, 4, 0, None)
    assert rc != 0
    assert len(lib.pamr_last_error()) > 0
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_cpu_tensors_raise():
    m = wseg_b200.PAMR(10, D6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError):
        wseg_b200.pseudo_labels(torch.zeros(1, 3, 8, 8))


def test_state_dict_matches_reference(golden_dir):
    ref = np.load(os.path.join(golden_dir, "state_dict.npz"))
    m = wseg_b200.PAMR(10, D6)
    sd = m.state_dict()
    assert list(sd.keys()) == ["aff_x.kernel", "aff_m.kernel", "aff_std.kernel"]
    assert sorted(sd.keys()) == sorted(ref.files)
    for k in ref.files:
        assert tuple(sd[k].shape) == ref[k].shape
        np.testing.assert_array_equal(sd[k].numpy(), ref[k])
    assert len(list(m.parameters())) == 0
    # reference attributes (pamr.py:119-122, :14)
    assert m.num_iter == 10 and m.aff_x.dilations == D6 and m.aff_m.dilations == D6 and m.aff_std.dilations == D6
    d = wseg_b200.PAMR()
    assert d.num_iter == 1 and d.dilations == [1]


def test_strict_load_and_tamper_check(golden_dir):
    ref = np.load(os.path.join(golden_dir, "state_dict.npz"))
    sd = {k: torch.from_numpy(ref[k]) for k in ref.files}
    m = wseg_b200.PAMR(10, D6)
    m.load_state_dict(sd, strict=True)
    bad = {k: v.clone() for k, v in sd.items()}
    bad["aff_m.kernel"][0, 0, 0, 0] = 0.5
    with pytest.raises(AssertionError):
        wseg_b200.PAMR(10, D6).load_state_dict(bad, strict=True)
    # embedded in a parent module, keys get the parent's prefix like `_aff.aff_x.kernel`
    parent = torch.nn.Module()
    parent._aff = wseg_b200.PAMR(10, D6)
    assert sorted(parent.state_dict()) == ["_aff.aff_m.kernel", "_aff.aff_std.kernel", "_aff.aff_x.kernel"]


def test_product_never_touches_oracle():
    files = glob.glob(os.path.join(PKG, "**", "*.py"), recursive=True) + \
        glob.glob(os.path.join(PKG, "csrc", "*.cu*")) + [os.path.join(ROOT, "include", "pamr_b200.h")]
    assert len(files) > 8
    for f in files:
        assert "oracle" not in open(f).read().lower(), f


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_shard_range_partitions():
    for B in (1, 7, 16, 128):
        for world in (1, 2, 3, 4, 8):
            spans = [wseg_b200.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        wseg_b200.shard_range(4, 2, 2)


def _gather_worker(rank, world, port, batch, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = (torch.arange(batch * 3 * 5, dtype=torch.int64) % 251).to(torch.uint8).view(batch, 3, 5)
        local = wseg_b200.shard_batch(full, rank, world).clone()
        out = wseg_b200.gather_labels(local, batch)
        q.put((rank, bool(torch.equal(out, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_gather_labels_gloo_world2(batch):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + batch
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _overlapped_gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = wseg_b200.OverlappedLabelGather("cpu", depth=2)
        ok = True
        outs = []
        for step in range(5):  # the two output buffers rotate: results of steps i and i+1 are alive together
            full = ((torch.arange(world * 2 * 3 * 4, dtype=torch.int64) + 7 * step) % 251).to(torch.uint8).view(world * 2, 3, 4)
            host = torch.empty_like(full)
            out = g.submit(full[2 * rank: 2 * rank + 2].clone(), host_out=host)
            g.wait()
            ok = ok and bool(torch.equal(out, full)) and bool(torch.equal(host, full))
            outs.append((out, full))
            if len(outs) >= 2:
                ok = ok and bool(torch.equal(outs[-2][0], outs[-2][1]))  # the previous step's buffer is still intact
        ok = ok and len({o.data_ptr() for o, _ in outs}) == 2  # two rotating output buffers
        own = torch.empty((2, 3, 4), dtype=torch.uint8)
        g.submit(full[2 * rank: 2 * rank + 2].clone(), host_out=own)  # [b,...]: this rank's slice of the gathered buffer
        g.wait()
        ok = ok and bool(torch.equal(own, full[2 * rank: 2 * rank + 2]))
        # a rank with an empty shard (B < world) still takes part in the gather of ShardedPseudoLabeler
        empty = wseg_b200.gather_labels(torch.full((1 if rank == 0 else 0, 2, 2), 9, dtype=torch.uint8), 1)
        ok = ok and tuple(empty.shape) == (1, 2, 2) and int(empty.sum()) == 36
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_overlapped_gather_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_overlapped_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _build_c_demo(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "1-stage-wseg_b200")
    exe = os.path.join(str(tmp_path), "c_abi_demo")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                        os.path.join(root, "examples", "c_abi_demo.c"), "-L" + pkg, "-lpamr_b200", "-Wl,-rpath," + pkg,
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_links(tmp_path):
    """include/pamr_b200.h compiles as pedantic C99 and a C program links against the library with no C++ /
    torch / Python involved (examples/c_abi_demo.c); running it needs a GPU (tests/test_gpu_parity.py)."""
    _lib.build()
    assert os.path.exists(_build_c_demo(tmp_path))


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU leg the driver runs beside the GPU arm) prints one JSON line with the
    contract's keys, honours --steps / --warmup as given and scales its sample with --gpus."""
    import json
    import subprocess
    import sys
    for gpus in (1, 2):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus),
                            "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["metric"].startswith("PAMR Mpix/s") and d["unit"] == "Mpix/s"
        assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == gpus and d["higher_is_better"] is True
        assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
        assert d["config"]["sample_batch"] == gpus and "B=%d per step" % gpus in d["config"]["workload"]
        assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert d["gpu_launches"] == 0


def test_mask_ce_workspace_covers_both_backward_orders():
    """The loss workspace holds lse + coef [B,H,W] and the backward pass's intermediate, which is [B,C,H,w] (x first)
    or [B,C,h,W] (y first, logits enlarged >= 2x): a host-only query, no device needed."""
    from wseg_b200 import _lib
    L = _lib.lib()
    for (B, C, h, w, H, W) in [(16, 21, 81, 81, 321, 321), (2, 21, 8, 40, 33, 40), (2, 21, 16, 9, 40, 8), (2, 5, 50, 40, 17, 13),
                               (3, 7, 10, 100, 20, 10)]:
        need = L.pamr_mask_ce_workspace_bytes(B, C, h, w, H, W)
        inter = 4 * B * C * (h * W if H >= 2 * h else H * w)
        assert need >= 2 * 4 * B * H * W + inter
        assert need <= 2 * 4 * B * H * W + inter + 8 * 256 + 16 * B
    assert L.pamr_mask_ce_workspace_bytes(4, 21, 33, 37, 33, 37) < 2 * 4 * 4 * 33 * 37 + 8 * 256 + 16 * 4 + 1  # same size: no intermediate
