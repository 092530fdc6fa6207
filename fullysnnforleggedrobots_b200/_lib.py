"""ctypes binding of the C-ABI in include/snn_b200.h (lib/libsnn_b200.so).

There is no CPU / PyTorch fallback: if the shared library is missing, or a tensor is not on a
CUDA device, the calls raise.  `build()` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsnn_b200.so")
# tools-only build with -DSNN_ABLATE (work-skipping SNN_DBG switches + barrier-wait cycle counters), selected with
# SNN_LIB=ablate; tests and bench.py check snn_build_flags() == 0, i.e. that they run the release library
ABLATE_LIB_PATH = os.path.join(LIB_DIR, "libsnn_b200_ablate.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SNN_MAX_LAYERS = 8
ERRORS = {-1: "SNN_EINVAL", -2: "SNN_EARCH", -3: "SNN_ECUDA", -4: "SNN_ENOMEM"}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


class SnnDesc(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("en_pop", C.c_int32), ("de_pop", C.c_int32),
                ("num_hidden", C.c_int32), ("hidden", C.c_int32 * SNN_MAX_LAYERS), ("spike_ts", C.c_int32),
                ("dec_tanh", C.c_int32), ("enc_eps", C.c_float), ("fwd_planes", C.c_int32)]


class SnnParams(C.Structure):
    _fields_ = [("enc_mean", C.c_void_p), ("enc_std", C.c_void_p),
                ("weight", C.c_void_p * SNN_MAX_LAYERS), ("bias", C.c_void_p * SNN_MAX_LAYERS),
                ("dec_w", C.c_void_p), ("dec_b", C.c_void_p)]


class SnnGrads(C.Structure):
    _fields_ = [("enc_mean", C.c_void_p), ("enc_std", C.c_void_p),
                ("weight", C.c_void_p * SNN_MAX_LAYERS), ("bias", C.c_void_p * SNN_MAX_LAYERS),
                ("dec_w", C.c_void_p), ("dec_b", C.c_void_p), ("obs", C.c_void_p)]


class SnnPpoActorHead(C.Structure):
    _fields_ = [("log_std", C.c_void_p), ("actions", C.c_void_p), ("old_logp", C.c_void_p), ("adv", C.c_void_p),
                ("old_mu", C.c_void_p), ("old_sigma", C.c_void_p), ("clip", C.c_float), ("mu_out", C.c_void_p),
                ("g_mu_out", C.c_void_p), ("part", C.c_void_p), ("part_bytes", C.c_size_t)]


class SnnMlpDesc(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("num_hidden", C.c_int32), ("hidden", C.c_int32 * SNN_MAX_LAYERS),
                ("out_dim", C.c_int32)]


class SnnMlpParams(C.Structure):
    _fields_ = [("weight", C.c_void_p * (SNN_MAX_LAYERS + 1)), ("bias", C.c_void_p * (SNN_MAX_LAYERS + 1))]


class SnnMlpGrads(C.Structure):
    _fields_ = [("weight", C.c_void_p * (SNN_MAX_LAYERS + 1)), ("bias", C.c_void_p * (SNN_MAX_LAYERS + 1))]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(INCLUDE, "snn_b200.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, ablate: bool = False) -> str:
    """nvcc -> lib/libsnn_b200.so (sm_100a only; cross-compiles without a GPU).
    ablate=True builds lib/libsnn_b200_ablate.so with -DSNN_ABLATE instead (tools only)."""
    out = ABLATE_LIB_PATH if ablate else LIB_PATH
    if not force and not ablate and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, "-I", INCLUDE, "-o", out, os.path.join(CSRC, "api.cu")]
    if ablate:
        cmd.insert(1, "-DSNN_ABLATE")
        # tools only: extra -D switches for A/B experiments of kernel geometry (e.g. -DSNN_FWD_NW=9 -DSNN_FWD_NS=2)
        cmd[2:2] = os.environ.get("SNN_EXTRA_NVCC", "").split()
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


_lib = None
_lock = threading.Lock()

_SIGS = {
    "snn_last_error": (C.c_char_p, []),
    "snn_abi_version": (C.c_int, []),
    "snn_build_flags": (C.c_int, []),
    "snn_packed_weights_bytes": (C.c_size_t, [C.POINTER(SnnDesc)]),
    "snn_trace_bytes": (C.c_size_t, [C.POINTER(SnnDesc), C.c_int64]),
    "snn_workspace_bytes": (C.c_size_t, [C.POINTER(SnnDesc), C.c_int64, C.c_int]),
    "snn_pack_weights": (C.c_int, [C.POINTER(SnnDesc), C.POINTER(SnnParams), C.c_void_p, C.c_void_p]),
    "snn_fwd_infer": (C.c_int, [C.POINTER(SnnDesc), C.POINTER(SnnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_fwd_act_supported": (C.c_int, [C.POINTER(SnnDesc), C.c_int64]),
    "snn_fwd_act_preferred": (C.c_int, [C.POINTER(SnnDesc), C.c_int64]),
    "snn_fwd_act": (C.c_int, [C.POINTER(SnnDesc), C.POINTER(SnnParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snn_fwd_train": (C.c_int, [C.POINTER(SnnDesc), C.POINTER(SnnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_bwd": (C.c_int, [C.POINTER(SnnDesc), C.POINTER(SnnParams), C.c_void_p, C.c_void_p, C.c_int64,
                          C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SnnGrads), C.c_void_p, C.c_size_t,
                          C.c_void_p]),
    "snn_ppo_head_part_bytes": (C.c_size_t, [C.POINTER(SnnDesc), C.c_int64]),
    "snn_bwd_ppo_top": (C.c_int, [C.POINTER(SnnDesc), C.POINTER(SnnParams), C.c_void_p, C.c_void_p, C.c_int64,
                                  C.POINTER(SnnPpoActorHead), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_bwd_rest": (C.c_int, [C.POINTER(SnnDesc), C.POINTER(SnnParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                               C.POINTER(SnnGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_ppo_value_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_ppo_finalize": (C.c_int, [C.POINTER(SnnDesc), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snn_gae": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float,
                          C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "snn_gather_rows": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "snn_mlp_packed_bytes": (C.c_size_t, [C.POINTER(SnnMlpDesc)]),
    "snn_mlp_trace_bytes": (C.c_size_t, [C.POINTER(SnnMlpDesc), C.c_int64]),
    "snn_mlp_workspace_bytes": (C.c_size_t, [C.POINTER(SnnMlpDesc), C.c_int64]),
    "snn_mlp_pack": (C.c_int, [C.POINTER(SnnMlpDesc), C.POINTER(SnnMlpParams), C.c_void_p, C.c_void_p]),
    "snn_mlp_fwd": (C.c_int, [C.POINTER(SnnMlpDesc), C.POINTER(SnnMlpParams), C.c_void_p, C.c_void_p, C.c_int64,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_mlp_bwd": (C.c_int, [C.POINTER(SnnMlpDesc), C.POINTER(SnnMlpParams), C.c_void_p, C.c_int64, C.c_void_p,
                              C.c_void_p, C.POINTER(SnnMlpGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_mlp_fwd_ex": (C.c_int, [C.POINTER(SnnMlpDesc), C.POINTER(SnnMlpParams), C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_mlp_bwd_ex": (C.c_int, [C.POINTER(SnnMlpDesc), C.POINTER(SnnMlpParams), C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(SnnMlpGrads), C.c_void_p,
                                 C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snn_ppo_head": (C.c_int, [C.c_void_p] * 10 + [C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
                     + [C.c_void_p] * 5 + [C.c_size_t, C.c_void_p]),
    "snn_sample_logprob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "snn_lr_adapt": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "snn_clip_adam": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_double] * 2 + [C.c_float] * 4
                      + [C.c_void_p, C.c_void_p, C.c_void_p]),
    "snn_clip_adam_kl": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_double] * 2 + [C.c_float] * 4
                         + [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snn_p2p_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "snn_p2p_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "snn_p2p_close": (C.c_int, [C.c_void_p]),
    "snn_p2p_free": (C.c_int, [C.c_void_p]),
    "snn_clip_adam_kl_p2p": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p] + [C.c_double] * 2
                             + [C.c_float] * 4 + [C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64,
                                                  C.c_void_p, C.c_void_p]),
    "snn_mse_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snn_profile_enable": (C.c_int, [C.c_int]),
    "snn_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "snn_launch_count": (C.c_int64, []),
    "snn_debug_set_cycle_buffer": (C.c_int, [C.c_void_p]),
    "snn_debug_trace_layout": (C.c_int, [C.POINTER(SnnDesc), C.c_int64, C.POINTER(C.c_int64), C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Load the shared library (once).  Raises if it has not been built: no fallback path exists."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = ABLATE_LIB_PATH if os.environ.get("SNN_LIB") == "ablate" else LIB_PATH
                if not os.path.exists(path):
                    raise RuntimeError(
                        f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  The spiking actor has no CPU or PyTorch fallback.")
                l = C.CDLL(path)
                for name, (res, args) in _SIGS.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(rc: int, what: str):
    if rc < 0:
        msg = lib().snn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {ERRORS.get(rc, rc)}: {msg}")
    return rc
