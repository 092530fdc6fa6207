"""Import the UNMODIFIED reference packages staged by oracle/make_ref.sh (TEST INFRASTRUCTURE: only tests/, smoke() and
bench.py's CPU arm may use this).

    mods = load("cassie", "modules")        # rsl_rl.modules of CASSIE/rsl_rl-master
    algs = load("cassie", "algorithms")     # rsl_rl.algorithms (PPO)

Search order: oracle/_ref/<fork> (travels to the GPU box), then the read-only tree under /root/reference (build
container only).  The four forks all call their package `rsl_rl`, so every load() first drops any `rsl_rl*` module
already imported; keep references to what you need before loading another fork.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
FORK_ROOT = {
    "amp": "AMP/AMP_for_hardware-main/rsl_rl",
    "cassie": "CASSIE/rsl_rl-master",
    "humanoid": "HUMANOID/pbrs-humanoid-dev/gpu_rl",
    "rma": "RMA/quadruped-robot-master/rl-robotics/rsl_rl",
}


def root_of(fork: str):
    """Directory that contains the fork's `rsl_rl` package, or None when the reference is not available."""
    staged = os.path.join(STAGED, fork)
    if os.path.isdir(os.path.join(staged, "rsl_rl")):
        return staged
    src = os.path.join(REFERENCE, FORK_ROOT[fork])
    if os.path.isdir(os.path.join(src, "rsl_rl")):
        return src
    return None


def available(fork: str = "cassie") -> bool:
    return root_of(fork) is not None


def kind(fork: str = "cassie") -> str:
    r = root_of(fork)
    return "none" if r is None else ("staged" if r.startswith(STAGED) else "tree")


def load(fork: str, sub: str):
    """Import rsl_rl.<sub> of `fork` (never rsl_rl.runners: it calls wandb.init at import)."""
    root = root_of(fork)
    if root is None:
        raise RuntimeError(f"reference fork {fork!r} is not staged: run `bash oracle/make_ref.sh` in the build container")
    for k in [k for k in sys.modules if k == "rsl_rl" or k.startswith("rsl_rl.")]:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return importlib.import_module("rsl_rl." + sub)
    finally:
        sys.path.remove(root)


def load_all(fork: str):
    """(rsl_rl.modules, rsl_rl.algorithms, rsl_rl.storage) of `fork`, imported together (one consistent package)."""
    mods = load(fork, "modules")
    root = root_of(fork)
    sys.path.insert(0, root)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            algs = importlib.import_module("rsl_rl.algorithms")
            sto = importlib.import_module("rsl_rl.storage")
    finally:
        sys.path.remove(root)
    return mods, algs, sto


@contextlib.contextmanager
def quiet():
    """The AMP / CASSIE / HUMANOID forwards print on every call (a host sync on the GPU): silence them."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
