"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, size queries work without a GPU, the modules keep the reference's state-dict keys and
constructor surface, and CUDA-less use fails loudly instead of falling back."""
import ctypes as C
import math
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "snn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fullysnnforleggedrobots_b200 import _lib
    lib = _lib.lib()
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in snn_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    assert lib.snn_abi_version() == 1


def test_size_queries_without_gpu():
    from fullysnnforleggedrobots_b200 import SnnEngine, _lib
    eng = SnnEngine(48, 12, 10, 10, [256, 256], spike_ts=5)
    L = _lib.lib()
    assert L.snn_packed_weights_bytes(C.byref(eng.desc)) > 3 * 2 * (512 * 256 + 256 * 256 + 256 * 128)
    # PPO-minibatch-sized batches are cut into tiles of 48 samples for T = 5 (csrc/plan.cuh): 24576 = 512 tiles,
    # 24577..24624 = 513 tiles; rollout-sized and tiny batches get narrower tiles (16 samples for B <= 16)
    tb = lambda B: L.snn_trace_bytes(C.byref(eng.desc), B)
    assert 0 < tb(24576) < tb(24577) and tb(24577) == tb(24624)
    assert 0 < tb(1) == tb(16) < tb(17)
    assert L.snn_workspace_bytes(C.byref(eng.desc), 4096, 1) > L.snn_workspace_bytes(C.byref(eng.desc), 4096, 0)
    bad = SnnEngine(48, 12, 10, 10, [256, 256], spike_ts=40)                      # unsupported T (> 32)
    assert L.snn_trace_bytes(C.byref(bad.desc), 128) == 0


def test_state_dict_keys_and_init_match_reference_layout():
    from fullysnnforleggedrobots_b200 import ActorCritic, ActorCriticRMA
    ac = ActorCritic(48, 48, 12, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256])
    keys = set(ac.state_dict().keys())
    expect = {"std", "actor.log_std", "actor.encoder.mean", "actor.encoder.std", "actor.snn.hidden_layers.0.weight",
              "actor.snn.hidden_layers.0.bias", "actor.snn.hidden_layers.1.weight", "actor.snn.hidden_layers.1.bias",
              "actor.snn.out_pop_layer.weight", "actor.snn.out_pop_layer.bias", "actor.decoder.decoder.weight",
              "actor.decoder.decoder.bias", "critic.0.weight", "critic.0.bias", "critic.2.weight", "critic.2.bias",
              "critic.4.weight", "critic.4.bias"}
    assert keys == expect
    sd = ac.state_dict()
    assert sd["actor.encoder.mean"].shape == (1, 48, 10) and sd["actor.decoder.decoder.weight"].shape == (12, 1, 10)
    assert torch.allclose(sd["actor.encoder.mean"][0, 0], torch.linspace(-5, 5, 10))
    assert torch.allclose(sd["actor.encoder.std"], torch.full((1, 48, 10), math.sqrt(0.15)))
    assert torch.allclose(sd["actor.log_std"], torch.full((12,), -0.001))
    assert ac.is_recurrent is False and ac.reset() is None
    rma = ActorCriticRMA(30, 20, 60, 12, actor_hidden_dims=[64], critic_hidden_dims=[64])
    rk = set(rma.state_dict().keys())
    assert {"env_factor_encoder.0.weight", "encoder.0.weight", "adaptation_module.0.weight"} <= rk
    assert rma.actor.obs_dim == 30 + 18


def test_unknown_kwargs_are_reported_and_ignored(capsys):
    from fullysnnforleggedrobots_b200 import ActorCritic
    ActorCritic(8, 8, 2, actor_hidden_dims=[16], critic_hidden_dims=[16], some_unknown_flag=3)
    assert "unexpected arguments" in capsys.readouterr().out


def test_no_cpu_fallback():
    from fullysnnforleggedrobots_b200 import ActorCritic, PPO
    ac = ActorCritic(8, 8, 2, actor_hidden_dims=[16], critic_hidden_dims=[16])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ac.act(torch.randn(4, 8))
    with pytest.raises(RuntimeError):
        PPO(ac, device="cpu")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fullysnnforleggedrobots_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_voltage_code_format_and_trace_decoder_on_the_host():
    """The 16-bit voltage trace (csrc/umma.cuh: volt_code2) restated in numpy: k = ceil(65534 v) saturated to
    [0, 65535] — spike flag exact around 0.5, value inside the surrogate window to 2^-17 — and `engine.unpack_trace`
    decoding a trace buffer laid out by the plan (no GPU: the layout query is host code)."""
    import ctypes as C
    import numpy as np
    from fullysnnforleggedrobots_b200 import _lib
    from fullysnnforleggedrobots_b200.engine import SnnEngine

    def code(v):                      # exact product in float64 (an fp32 x 17-bit constant fits), rounded up, saturated
        v = np.asarray(v, dtype=np.float32)
        hi = np.float32(65535.0 / 65534.0)
        return np.ceil(np.clip(v, 0.0, hi).astype(np.float64) * 65534.0).astype(np.int64)

    half = np.float32(0.5)
    near = np.array([np.nextafter(half, np.float32(0)), half, np.nextafter(half, np.float32(1))], dtype=np.float32)
    assert list(code(near) > 32767) == [False, False, True]                      # spike flag <=> v > 0.5, to the last ulp
    assert list(code(np.float32([-3.0, 0.0, 1e-30, 1.0, 1.5, 77.0]))) == [0, 0, 1, 65534, 65535, 65535]
    v = np.random.default_rng(0).uniform(1e-6, 1 - 1e-6, 200000).astype(np.float32)
    k = code(v)
    assert k.min() >= 1 and k.max() <= 65534
    assert np.abs((k - 0.5) / 65534.0 - v.astype(np.float64)).max() <= 2.0 ** -17 + 1e-9

    eng = SnnEngine(3, 2, 4, 4, [128], spike_ts=5)
    B = 20
    arr = (C.c_int64 * (8 + 4 * _lib.SNN_MAX_LAYERS))()
    _lib.check(_lib.lib().snn_debug_trace_layout(C.byref(eng.desc), B, arr, len(arr)), "trace layout")
    Lc, T, Bt, CT, ntile, Rt = (int(arr[i]) for i in range(6))
    nbytes = int(_lib.lib().snn_trace_bytes(C.byref(eng.desc), B))
    raw = torch.zeros(nbytes, dtype=torch.uint8)
    rng = np.random.default_rng(1)
    want = []
    for l in range(Lc):
        NP, voff = int(arr[8 + 4 * l]), int(arr[11 + 4 * l])
        volt = rng.uniform(-0.5, 1.5, size=(T, B, NP)).astype(np.float32)
        codes = np.zeros((NP // 32, Rt // 8, 32, 8), dtype=np.uint16)
        for t in range(T):
            for b in range(B):
                r = (b // Bt) * CT + t * Bt + b % Bt
                codes[:, r // 8, :, r % 8] = code(volt[t, b]).reshape(NP // 32, 32).astype(np.uint16)
        raw[voff: voff + Rt * NP * 2] = torch.from_numpy(codes.reshape(-1).view(np.uint8).copy())
        want.append(volt)
    tr = eng.unpack_trace(raw, B)
    for l in range(Lc):
        n = eng.dims[l + 1]
        vw = torch.from_numpy(want[l][:, :, :n])
        win = (vw > 0) & (vw < 1)
        assert torch.equal(tr["volt_code"][l] > 32767, vw > 0.5)
        assert (tr["volt"][l][win] - vw[win]).abs().max() <= 2.0 ** -17 + 1e-7
        assert torch.equal(tr["volt"][l][vw <= 0], torch.zeros_like(vw[vw <= 0]))
        assert torch.equal(tr["volt"][l][vw > 1], torch.ones_like(vw[vw > 1]))
