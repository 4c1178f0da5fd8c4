"""ctypes binding of libpamr_b200.so (C ABI declared in include/pamr_b200.h).

The product path has no CPU fallback: if the library is missing, cannot be loaded, or the
device is not an sm_100 GPU, calls raise RuntimeError.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAMR_B200_LIB selects another build of the same ABI (experiment variants, tools/build_variant.sh)
LIB_PATH = os.environ.get("PAMR_B200_LIB") or os.path.join(_HERE, "libpamr_b200.so")
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 2

_lock = threading.Lock()
_lib = None

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float

# name -> (restype, argtypes); must list every symbol include/pamr_b200.h declares
SIGNATURES = {
    "pamr_b200_abi_version": (_i, []),
    "pamr_last_error": (ctypes.c_char_p, []),
    "pamr_device_info": (_i, [_i, _vp, _vp, _vp, _vp]),
    "pamr_launch_count": (ctypes.c_ulonglong, []),
    "pamr_resize_bilinear_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pamr_affinity_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "pamr_local_std_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "pamr_propagate_scratch_bytes": (ctypes.c_size_t, [_i] * 4 + [_vp, _i, _i]),
    "pamr_propagate_f32": (_i, [_vp, _vp, _vp, _vp, ctypes.c_size_t, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp]),
    "pamr_forward_workspace_bytes": (ctypes.c_size_t, [_i] * 7 + [_vp, _i, _i]),
    "pamr_forward_f32": (_i, [_vp, _vp, _vp, _vp, ctypes.c_size_t] + [_i] * 7 + [_vp, _i, _i, _vp, _i, _vp]),
    "pamr_clean_f32": (_i, [_vp, _vp, _vp, _vp] + [_i] * 6 + [_i, _vp]),
    "pamr_pseudo_labels_f32": (_i, [_vp] * 6 + [_i] * 6 + [_f, _f, _f, _i, _i, _vp]),
    "pamr_pseudo_labels_host_f32": (_i, [_vp] * 4 + [_i] * 7 + [_vp, _i, _i, _f, _f, _f, _i]),
    "pamr_denorm_resize_f32": (_i, [_vp] * 4 + [_i] * 6 + [_i, _vp]),
    "pamr_merge_multiscale_f32": (_i, [_vp] * 5 + [_i] * 7 + [_f, _f, _i, _vp]),
    "pamr_mask_ce_workspace_bytes": (ctypes.c_size_t, [_i] * 6),
    "pamr_labels_from_onehot_f32": (_i, [_vp, _vp, _vp] + [_i] * 4 + [_i, _vp]),
    "pamr_mask_ce_forward_f32": (_i, [_vp] * 6 + [ctypes.c_size_t] + [_i] * 6 + [_i, _vp]),
    "pamr_mask_ce_backward_f32": (_i, [_vp] * 5 + [ctypes.c_size_t] + [_i] * 6 + [_i, _vp]),
    "pamr_ordered_from_float": (ctypes.c_uint, [_f]),
    "pamr_float_from_ordered": (_f, [ctypes.c_uint]),
}


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libpamr_b200.so (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    out = subprocess.run(["make", "-C", CSRC, "-j8"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libpamr_b200.so failed")
    return LIB_PATH


def lib():
    """Load the shared library once per process; raises if it is missing (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "libpamr_b200.so not found at %s: build it with `make -C %s` "
                        "(there is no CPU or PyTorch fallback)" % (LIB_PATH, CSRC))
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    if not hasattr(L, name) and os.environ.get("PAMR_B200_OLD_LIBRARY") == "1":
                        continue  # A/B experiments against a library built from an older commit (tools/build_variant.sh)
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                if L.pamr_b200_abi_version() != ABI_VERSION:
                    raise RuntimeError("libpamr_b200.so ABI %d != expected %d" % (L.pamr_b200_abi_version(), ABI_VERSION))
                _lib = L
    return _lib


def check(rc):
    """Turn a PAMR_* status into RuntimeError carrying pamr_last_error()."""
    if rc != 0:
        msg = lib().pamr_last_error()
        raise RuntimeError("libpamr_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def launch_count():
    return int(lib().pamr_launch_count())
