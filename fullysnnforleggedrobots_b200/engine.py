"""SnnEngine — owns the device buffers one PopSpikeActor needs and drives the C-ABI kernels.

Buffers (all torch-allocated, caller-owned from the library's point of view):
  packed   bf16 hi/mid/lo plane images of every layer's W and W^T (re-packed when parameters change)
  ws_fwd   bit-packed spike trains of one inference call
  trace    spike bits + 16-bit voltage-code traces of one training forward (lives until its backward)
  ws_bwd   g_c plane images, split-K partials, small reduction scratch
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_f32_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the spiking actor has no CPU fallback")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous float32")


class SnnEngine:
    def __init__(self, obs_dim: int, act_dim: int, en_pop: int, de_pop: int, hidden: Sequence[int],
                 spike_ts: int = 5, dec_tanh: bool = False, enc_eps: float = 0.0, fwd_planes: int = 3):
        if len(hidden) + 1 > _lib.SNN_MAX_LAYERS:
            raise ValueError("too many hidden layers")
        d = _lib.SnnDesc()
        d.obs_dim, d.act_dim, d.en_pop, d.de_pop = obs_dim, act_dim, en_pop, de_pop
        d.num_hidden = len(hidden)
        for i, h in enumerate(hidden):
            d.hidden[i] = int(h)
        d.spike_ts = spike_ts
        d.dec_tanh = 1 if dec_tanh else 0
        d.enc_eps = float(enc_eps)
        d.fwd_planes = fwd_planes
        self.desc = d
        self.L = len(hidden) + 1
        self.dims = [obs_dim * en_pop] + [int(h) for h in hidden] + [act_dim * de_pop]
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._ws: Dict[tuple, torch.Tensor] = {}
        # inference forward: the one-launch fused kernel (snn_fwd_act) whenever it takes the geometry; False keeps the
        # layered path (encode -> one scan-GEMM per layer -> decode) — same mu bit for bit, used by tests and A/B timing
        self.fused_infer = True

    # ------------------------------------------------------------------ helpers
    def _params_struct(self, enc_mean, enc_std, weights, biases, dec_w, dec_b) -> _lib.SnnParams:
        p = _lib.SnnParams()
        p.enc_mean, p.enc_std = enc_mean.data_ptr(), enc_std.data_ptr()
        for i in range(self.L):
            p.weight[i] = weights[i].data_ptr()
            p.bias[i] = biases[i].data_ptr()
        p.dec_w, p.dec_b = dec_w.data_ptr(), dec_b.data_ptr()
        return p

    def _buffer(self, kind: str, nbytes: int, device) -> torch.Tensor:
        key = (kind, str(device))
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1024), dtype=torch.uint8, device=device)
            self._ws[key] = buf
        return buf

    def packed_weights(self, weights: List[torch.Tensor], biases: List[torch.Tensor], params_struct, enc=()) -> torch.Tensor:
        """bf16 plane images (+ the encoder table of the fused rollout kernel); re-packed only when a weight / bias /
        encoder tensor changed (version counter)."""
        key = tuple((t.data_ptr(), t._version) for t in (*weights, *biases, *enc))
        dev = weights[0].device
        if self._packed is None or self._packed.device != dev:
            n = _lib.lib().snn_packed_weights_bytes(C.byref(self.desc))
            if n == 0:
                raise RuntimeError("snn_packed_weights_bytes: unsupported network description")
            self._packed = torch.empty(n, dtype=torch.uint8, device=dev)
            self._packed_key = None
        if key != self._packed_key:
            _lib.check(_lib.lib().snn_pack_weights(C.byref(self.desc), C.byref(params_struct), _ptr(self._packed),
                                                   _stream_ptr(dev)), "snn_pack_weights")
            self._packed_key = key
        return self._packed

    def invalidate(self):
        self._packed_key = None

    def repack(self, enc_mean, enc_std, weights, biases, dec_w, dec_b):
        """Unconditional re-pack (the fused optimizer updates parameters behind torch's version counters).  Pass the
        Parameters themselves, not `.data` views: the "changed since the last pack" key is (data_ptr, _version) of what
        the caller passes, and a `.data` view always reports version 0 — a later ensure_packed() on the Parameters
        would then re-pack a second time.  A write through `param.data.<op>_()` bumps no counter at all: call
        PPO.refresh_packed() (or invalidate()) after such a write."""
        self._packed_key = None
        ps = self._params_struct(enc_mean, enc_std, weights, biases, dec_w, dec_b)
        with torch.cuda.device(weights[0].device):
            self.packed_weights(weights, biases, ps, enc=(enc_mean, enc_std))

    def ensure_packed(self, enc_mean, enc_std, weights, biases, dec_w, dec_b):
        """Re-pack only if a weight / bias tensor changed since the last pack (torch version counters).  Callers that
        replay CUDA graphs over the packed planes run this before the replay: the check is host-side."""
        key = tuple((t.data_ptr(), t._version) for t in (*weights, *biases, enc_mean, enc_std))
        if key == self._packed_key and self._packed is not None:
            return
        ps = self._params_struct(enc_mean.data, enc_std.data, [w.data for w in weights], [b.data for b in biases],
                                 dec_w.data, dec_b.data)
        with torch.cuda.device(weights[0].device):
            self.packed_weights(weights, biases, ps, enc=(enc_mean, enc_std))

    def trace_bytes(self, B: int) -> int:
        return int(_lib.lib().snn_trace_bytes(C.byref(self.desc), B))

    def workspace_bytes(self, B: int, backward: bool) -> int:
        return int(_lib.lib().snn_workspace_bytes(C.byref(self.desc), B, 1 if backward else 0))

    # ------------------------------------------------------------------ calls
    def forward(self, obs, enc_mean, enc_std, weights, biases, dec_w, dec_b, train: bool, out_mu=None,
                out_trace=None, assume_packed: bool = False, ws=None):
        """Returns (mu, trace or None).  out_mu / out_trace: optional caller-owned (static) output buffers.
        assume_packed: the caller re-packs explicitly after every optimizer step (fused PPO path).
        ws: caller-owned inference workspace (>= workspace_bytes(B, backward=False)); a CUDA graph that captures this
        call must pass its own, the engine's internal buffer may be re-allocated by a later, larger eager call."""
        _check_f32_cuda(obs, "obs")
        for t, n in ((enc_mean, "encoder.mean"), (enc_std, "encoder.std"), (dec_w, "decoder.weight"),
                     (dec_b, "decoder.bias"), *[(w, "weight") for w in weights], *[(b, "bias") for b in biases]):
            _check_f32_cuda(t, n)
        if obs.dim() != 2 or obs.shape[1] != self.desc.obs_dim:
            raise RuntimeError(f"obs must be [B, {self.desc.obs_dim}], got {tuple(obs.shape)}")
        dev = obs.device
        B = obs.shape[0]
        L = _lib.lib()
        with torch.cuda.device(dev):
            ps = self._params_struct(enc_mean, enc_std, weights, biases, dec_w, dec_b)
            packed = self._packed if (assume_packed and self._packed is not None) else \
                self.packed_weights(weights, biases, ps, enc=(enc_mean, enc_std))
            if out_mu is False:      # training forward of the fused PPO step: the decoder runs in backward_ppo_top
                if not train:
                    raise RuntimeError("out_mu=False (no decoder pass) is a training-forward option")
                mu = None
            else:
                mu = out_mu if out_mu is not None else torch.empty(B, self.desc.act_dim, dtype=torch.float32, device=dev)
            if B == 0:
                return mu, None
            if train:
                nbytes = L.snn_trace_bytes(C.byref(self.desc), B)
                trace = out_trace if out_trace is not None else torch.empty(nbytes, dtype=torch.uint8, device=dev)
                if trace.numel() < nbytes:
                    raise RuntimeError("trace buffer too small")
                _lib.check(L.snn_fwd_train(C.byref(self.desc), C.byref(ps), _ptr(packed), _ptr(obs), B, _ptr(mu),
                                           _ptr(trace), nbytes, None, 0, _stream_ptr(dev)), "snn_fwd_train")
                return mu, trace
            if self.fused_infer and L.snn_fwd_act_preferred(C.byref(self.desc), B):
                _lib.check(L.snn_fwd_act(C.byref(self.desc), C.byref(ps), _ptr(packed), _ptr(obs), B, _ptr(mu), None, None,
                                         None, None, None, _stream_ptr(dev)), "snn_fwd_act")
                return mu, None
            nbytes = L.snn_workspace_bytes(C.byref(self.desc), B, 0)
            if ws is None:
                ws = self._buffer("fwd", nbytes, dev)
            elif ws.numel() < nbytes:
                raise RuntimeError("inference workspace too small")
            _lib.check(L.snn_fwd_infer(C.byref(self.desc), C.byref(ps), _ptr(packed), _ptr(obs), B, _ptr(mu),
                                       _ptr(ws), ws.numel(), _stream_ptr(dev)), "snn_fwd_infer")
            return mu, None

    def act_supported(self, B: int) -> bool:
        """The one-launch kernel takes this geometry at all."""
        return bool(B > 0 and _lib.lib().snn_fwd_act_supported(C.byref(self.desc), B))

    def act_preferred(self, B: int) -> bool:
        """... and is the faster path for this batch (rollout-sized: at most two waves of tiles): what forward() uses."""
        return bool(self.fused_infer and B > 0 and _lib.lib().snn_fwd_act_preferred(C.byref(self.desc), B))

    def act(self, obs, enc_mean, enc_std, weights, biases, dec_w, dec_b, noise, log_std, out_mu, out_actions, out_logp,
            out_sigma, assume_packed: bool = False):
        """Rollout step in one launch (snn_fwd_act): mu, actions = mu + exp(log_std) * noise, summed Normal log-prob and
        sigma, all written into caller-owned buffers.  noise = None: mu only (the other outputs are ignored).  Only when
        act_supported(B); forward(train=False) picks this kernel by itself for rollout-sized batches."""
        _check_f32_cuda(obs, "obs")
        _check_f32_cuda(out_mu, "mu")
        if noise is not None:
            for t, n in ((noise, "noise"), (log_std, "log_std"), (out_actions, "actions"), (out_logp, "logp"),
                         (out_sigma, "sigma")):
                _check_f32_cuda(t, n)
        else:
            log_std = out_actions = out_logp = out_sigma = None
        dev = obs.device
        B = obs.shape[0]
        L = _lib.lib()
        with torch.cuda.device(dev):
            ps = self._params_struct(enc_mean, enc_std, weights, biases, dec_w, dec_b)
            packed = self._packed if (assume_packed and self._packed is not None) else \
                self.packed_weights(weights, biases, ps, enc=(enc_mean, enc_std))
            _lib.check(L.snn_fwd_act(C.byref(self.desc), C.byref(ps), _ptr(packed), _ptr(obs), B, _ptr(out_mu), _ptr(noise),
                                     _ptr(log_std), _ptr(out_actions), _ptr(out_logp), _ptr(out_sigma), _stream_ptr(dev)),
                       "snn_fwd_act")

    def backward(self, obs, mu, g_mu, trace, enc_mean, enc_std, weights, biases, dec_w, dec_b, need_dobs: bool,
                 out=None, assume_packed: bool = False, ws=None):
        """out: optional dict of caller-owned gradient tensors (same keys as the returned dict); overwritten.
        ws: caller-owned backward workspace (>= workspace_bytes(B, backward=True)), see forward()."""
        dev = obs.device
        B = obs.shape[0]
        L = _lib.lib()
        g_mu = g_mu.contiguous().float()
        with torch.cuda.device(dev):
            ps = self._params_struct(enc_mean, enc_std, weights, biases, dec_w, dec_b)
            packed = self._packed if (assume_packed and self._packed is not None) else \
                self.packed_weights(weights, biases, ps, enc=(enc_mean, enc_std))
            g = _lib.SnnGrads()
            if out is None:
                out = {
                    "enc_mean": torch.empty_like(enc_mean), "enc_std": torch.empty_like(enc_std),
                    "weights": [torch.empty_like(w) for w in weights], "biases": [torch.empty_like(b) for b in biases],
                    "dec_w": torch.empty_like(dec_w), "dec_b": torch.empty_like(dec_b),
                    "obs": torch.empty_like(obs) if need_dobs else None,
                }
            g.enc_mean, g.enc_std = out["enc_mean"].data_ptr(), out["enc_std"].data_ptr()
            for i in range(self.L):
                g.weight[i] = out["weights"][i].data_ptr()
                g.bias[i] = out["biases"][i].data_ptr()
            g.dec_w, g.dec_b = out["dec_w"].data_ptr(), out["dec_b"].data_ptr()
            g.obs = out["obs"].data_ptr() if need_dobs else None
            nbytes = L.snn_workspace_bytes(C.byref(self.desc), B, 1)
            if ws is None:
                ws = self._buffer("bwd", nbytes, dev)
            elif ws.numel() < nbytes:
                raise RuntimeError("backward workspace too small")
            _lib.check(L.snn_bwd(C.byref(self.desc), C.byref(ps), _ptr(packed), _ptr(obs), B, _ptr(mu), _ptr(g_mu),
                                 _ptr(trace), C.byref(g), _ptr(ws), ws.numel(), _stream_ptr(dev)), "snn_bwd")
        return out

    def _grads_struct(self, out, need_dobs):
        g = _lib.SnnGrads()
        g.enc_mean, g.enc_std = out["enc_mean"].data_ptr(), out["enc_std"].data_ptr()
        for i in range(self.L):
            g.weight[i] = out["weights"][i].data_ptr()
            g.bias[i] = out["biases"][i].data_ptr()
        g.dec_w, g.dec_b = out["dec_w"].data_ptr(), out["dec_b"].data_ptr()
        g.obs = out["obs"].data_ptr() if need_dobs else None
        return g

    def ppo_head_part_bytes(self, B: int) -> int:
        return int(_lib.lib().snn_ppo_head_part_bytes(C.byref(self.desc), B))

    def backward_ppo_top(self, obs, trace, params, log_std, actions, old_logp, adv, old_mu, old_sigma, clip: float, part,
                         ws, out_mu, out_g_mu):
        """Stage 1 of the fused PPO backward (snn_bwd_ppo_top): decoder + the policy terms of the PPO head in one kernel
        (-> out_mu, out_g_mu), then the output layer's BPTT scan.  params = (enc_mean, enc_std, weights, biases, dec_w,
        dec_b); the planes must be current (the fused step re-packs after every optimizer step).  part: uint8 buffer of
        ppo_head_part_bytes(B)."""
        dev = obs.device
        B = obs.shape[0]
        for t, n in ((log_std, "log_std"), (actions, "actions"), (old_logp, "old_logp"), (adv, "adv"), (old_mu, "old_mu"),
                     (old_sigma, "old_sigma")):
            _check_f32_cuda(t, n)
        with torch.cuda.device(dev):
            ps = self._params_struct(*params)
            h = _lib.SnnPpoActorHead()
            h.log_std, h.actions, h.old_logp, h.adv = log_std.data_ptr(), actions.data_ptr(), old_logp.data_ptr(), adv.data_ptr()
            h.old_mu, h.old_sigma, h.clip = old_mu.data_ptr(), old_sigma.data_ptr(), float(clip)
            h.mu_out, h.g_mu_out = out_mu.data_ptr(), out_g_mu.data_ptr()
            h.part, h.part_bytes = part.data_ptr(), part.numel()
            _lib.check(_lib.lib().snn_bwd_ppo_top(C.byref(self.desc), C.byref(ps), _ptr(self._packed), _ptr(obs), B, C.byref(h),
                                                  _ptr(trace), _ptr(ws), ws.numel(), _stream_ptr(dev)), "snn_bwd_ppo_top")

    def backward_rest(self, obs, trace, params, out, ws, need_dobs: bool = False):
        """Stage 2 (snn_bwd_rest): the dX chain below the output layer, dW, the deferred reductions -> `out`."""
        dev = obs.device
        with torch.cuda.device(dev):
            ps = self._params_struct(*params)
            g = self._grads_struct(out, need_dobs)
            _lib.check(_lib.lib().snn_bwd_rest(C.byref(self.desc), C.byref(ps), _ptr(self._packed), _ptr(obs), obs.shape[0],
                                               _ptr(trace), C.byref(g), _ptr(ws), ws.numel(), _stream_ptr(dev)), "snn_bwd_rest")
        return out

    # ------------------------------------------------------------------ test hooks
    def unpack_trace(self, trace: torch.Tensor, B: int, rows: Optional[torch.Tensor] = None):
        """Decode a training trace into dense tensors (tests / debugging only; not on the product path).
        rows: optional 1-D index tensor — decode only these samples (full-size traces are hundreds of MB)."""
        arr = (C.c_int64 * (8 + 4 * _lib.SNN_MAX_LAYERS))()
        n = _lib.check(_lib.lib().snn_debug_trace_layout(C.byref(self.desc), B, arr, len(arr)), "trace layout")
        Lc, T, Bt, CT, ntile, Rt, K0P, enc_off = (int(arr[i]) for i in range(8))
        raw = trace.cpu()
        # sample-step (t, b) lives at column tile * CT + t * Bt + b % Bt (csrc/plan.cuh)
        bb = torch.arange(B) if rows is None else rows.detach().cpu().to(torch.int64)
        col = ((bb // Bt) * CT + bb % Bt).unsqueeze(0) + (torch.arange(T) * Bt).unsqueeze(1)      # [T, rows]

        def bits(off, ncols_p, ncols):
            words = raw[off: off + (ncols_p // 32) * Rt * 4].view(torch.int32).reshape(ncols_p // 32, Rt)
            w = words[:, col]                                                         # [ncols_p / 32, T, rows]
            sh = torch.arange(32, dtype=torch.int32).reshape(1, 32, 1, 1)
            b = ((w.unsqueeze(1) >> sh) & 1).reshape(ncols_p, T, -1)                 # [neuron, T, rows]
            return b.permute(1, 2, 0)[:, :, :ncols].to(torch.float32).contiguous()

        res = {"enc_spikes": bits(enc_off, K0P, self.dims[0]), "spikes": [], "volt": []}
        for l in range(Lc):
            NP, KP, boff, voff = (int(arr[8 + 4 * l + i]) for i in range(4))
            res["spikes"].append(bits(boff, NP, self.dims[l + 1]))
            # voltage trace: 16-bit codes [NP / 32][Rt / 8][32 neurons][8 sample-steps] (csrc/umma.cuh: volt_code2).
            # 0 = below the surrogate window (decoded 0.0), 0xFFFF = above it (decoded 1.0), k in 1..65534 = inside,
            # v = (k - 0.5) / 65534 to 2^-17 — the value the backward kernels use.
            c = raw[voff: voff + Rt * NP * 2].view(torch.int16).reshape(NP // 32, Rt // 8, 32, 8)
            c = c.permute(0, 1, 3, 2).reshape(NP // 32, Rt, 32)[:, col]               # [NP / 32, T, rows, 32]
            c = c.permute(1, 2, 0, 3).reshape(T, -1, NP)[:, :, :self.dims[l + 1]].to(torch.int32) & 0xFFFF
            v = (c.to(torch.float32) - 0.5) * (1.0 / 65534.0)
            v = torch.where(c == 0, torch.zeros_like(v), torch.where(c == 0xFFFF, torch.ones_like(v), v))
            res["volt"].append(v.contiguous())
            res.setdefault("volt_code", []).append(c.contiguous())
        return res


def gae(rewards, dones, values, last_values, gamma: float, lam: float, normalize: bool = True):
    """RolloutStorage.compute_returns on [S, N] tensors -> (returns, advantages)."""
    for t, n in ((rewards, "rewards"), (values, "values"), (last_values, "last_values")):
        _check_f32_cuda(t, n)
    if dones.dtype != torch.uint8 or not dones.is_cuda or not dones.is_contiguous():
        raise RuntimeError("dones must be a contiguous CUDA uint8 tensor")
    S, N = rewards.shape[0], rewards.shape[1]
    dev = rewards.device
    returns = torch.empty_like(rewards)
    adv = torch.empty_like(rewards)
    scratch = torch.empty(16384, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().snn_gae(_ptr(rewards), _ptr(dones), _ptr(values), _ptr(last_values), S, N,
                                      float(gamma), float(lam), _ptr(returns), _ptr(adv), 1 if normalize else 0,
                                      _ptr(scratch), _stream_ptr(dev)), "snn_gae")
    return returns, adv


def _check_rows_f32_cuda(t: torch.Tensor, name: str):
    """2-D (or 1-D) float32 CUDA rows whose elements are contiguous within a row; the row stride is free (a column slice of
    a wider buffer, e.g. the latent part of cat(obs, latent))."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: no CPU fallback")
    if t.dtype != torch.float32 or t.dim() not in (1, 2) or (t.dim() == 2 and t.shape[1] > 1 and t.stride(1) != 1):
        raise RuntimeError(f"{name} must be float32 rows with unit element stride")


def _row_stride(t: torch.Tensor) -> int:
    if t.dim() == 1:
        return 1
    return int(t.stride(0)) if t.shape[0] > 1 else max(int(t.stride(0)), int(t.shape[1]))


class MlpEngine:
    """Dense ELU MLP (Linear -> ELU -> ... -> Linear(., out_dim)) on the tensor-core kernels (snn_mlp_*).

    out_dim == 1: the reference's critic (AMP/.../modules/actor_critic.py:359-368), 1-wide SIMT head.
    out_dim >= 2: the RMA fork's env_factor_encoder / adaptation_module (RMA/.../modules/actor_critic.py:394-418):
    the linear output layer is one more tensor-core GEMM.  Rows may be strided (a column slice of a wider buffer),
    `backward` can sum two gradient sources and return the gradient into the input rows.

    `linears` is the list of nn.Linear modules of an nn.Sequential of that form."""

    def __init__(self, linears):
        if len(linears) < 2:
            raise ValueError("MLP needs at least one hidden layer")
        if len(linears) - 1 > _lib.SNN_MAX_LAYERS:
            raise ValueError("too many MLP layers")
        d = _lib.SnnMlpDesc()
        d.in_dim = linears[0].in_features
        d.num_hidden = len(linears) - 1
        for i, l in enumerate(linears[:-1]):
            d.hidden[i] = l.out_features
        d.out_dim = linears[-1].out_features
        if d.out_dim == 1 and linears[-2].out_features > 256:
            raise ValueError("last hidden layer wider than 256 is not supported by the value-head kernels")
        if d.out_dim > 128:
            raise ValueError("output layers wider than 128 are not supported")
        self.desc = d
        self.out_dim = int(d.out_dim)
        self.linears = list(linears)
        self._packed = None
        self._ws = None
        if _lib.lib().snn_mlp_packed_bytes(C.byref(d)) == 0:
            raise ValueError("unsupported MLP dimensions")

    @classmethod
    def from_sequential(cls, seq):
        """Return an engine if `seq` is Linear/ELU/.../Linear (and, for CriticEngine, ends in Linear(., 1)), else None."""
        mods = list(seq)
        lin = [m for m in mods if isinstance(m, torch.nn.Linear)]
        act = [m for m in mods if not isinstance(m, torch.nn.Linear)]
        ok = (len(mods) == 2 * len(lin) - 1 and all(isinstance(m, torch.nn.ELU) and m.alpha == 1.0 for m in act)
              and all(isinstance(mods[2 * i], torch.nn.Linear) for i in range(len(lin))))
        if not ok:
            return None
        try:
            return cls(lin)
        except ValueError:
            return None

    def _params(self):
        p = _lib.SnnMlpParams()
        for i, l in enumerate(self.linears):
            p.weight[i] = l.weight.data_ptr()
            p.bias[i] = l.bias.data_ptr()
        return p

    def trace_bytes(self, B):
        return int(_lib.lib().snn_mlp_trace_bytes(C.byref(self.desc), B))

    def repack(self):
        dev = self.linears[0].weight.device
        L = _lib.lib()
        if self._packed is None or self._packed.device != dev:
            self._packed = torch.empty(L.snn_mlp_packed_bytes(C.byref(self.desc)), dtype=torch.uint8, device=dev)
        ps = self._params()
        with torch.cuda.device(dev):
            _lib.check(L.snn_mlp_pack(C.byref(self.desc), C.byref(ps), _ptr(self._packed), _stream_ptr(dev)), "snn_mlp_pack")

    def forward(self, x, out_value=None, out_trace=None):
        """y = mlp(x).  out_value: [B] (out_dim 1) or a [B, out_dim] view with unit element stride (any row stride)."""
        _check_rows_f32_cuda(x, "mlp input")
        dev = x.device
        B = x.shape[0]
        L = _lib.lib()
        if self._packed is None:
            self.repack()
        nbytes = L.snn_mlp_trace_bytes(C.byref(self.desc), B)
        trace = out_trace if out_trace is not None else torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if out_value is None:
            out_value = torch.empty((B,) if self.out_dim == 1 else (B, self.out_dim), dtype=torch.float32, device=dev)
        _check_rows_f32_cuda(out_value, "mlp output")
        ps = self._params()
        with torch.cuda.device(dev):
            _lib.check(L.snn_mlp_fwd_ex(C.byref(self.desc), C.byref(ps), _ptr(self._packed), _ptr(x), _row_stride(x), B,
                                        _ptr(out_value), _row_stride(out_value), _ptr(trace), trace.numel(),
                                        _stream_ptr(dev)), "snn_mlp_fwd_ex")
        return out_value, trace

    def workspace_bytes(self, B):
        return int(_lib.lib().snn_mlp_workspace_bytes(C.byref(self.desc), B))

    def backward(self, g_value, trace, B, out=None, ws=None, g2=None, d_x=None):
        """Gradients of sum_b <g_value[b] (+ g2[b]), y[b]>; `out` = (list of weight grads, list of bias grads) to overwrite.
        d_x: optional [B, in_dim] rows (any row stride) receiving the gradient into the input.
        ws: caller-owned workspace (a CUDA graph capturing this call passes its own)."""
        dev = g_value.device
        L = _lib.lib()
        _check_rows_f32_cuda(g_value, "mlp output gradient")
        if g2 is not None:
            _check_rows_f32_cuda(g2, "mlp output gradient (second source)")
        if d_x is not None:
            _check_rows_f32_cuda(d_x, "mlp input gradient")
        if out is None:
            out = ([torch.empty_like(l.weight) for l in self.linears], [torch.empty_like(l.bias) for l in self.linears])
        g = _lib.SnnMlpGrads()
        for i in range(len(self.linears)):
            g.weight[i] = out[0][i].data_ptr()
            g.bias[i] = out[1][i].data_ptr()
        nbytes = L.snn_mlp_workspace_bytes(C.byref(self.desc), B)
        if ws is None:
            if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
                self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            ws = self._ws
        elif ws.numel() < nbytes:
            raise RuntimeError("mlp workspace too small")
        ps = self._params()
        with torch.cuda.device(dev):
            _lib.check(L.snn_mlp_bwd_ex(C.byref(self.desc), C.byref(ps), _ptr(self._packed), B, _ptr(g_value),
                                        _row_stride(g_value), _ptr(g2), 0 if g2 is None else _row_stride(g2), _ptr(trace),
                                        C.byref(g), _ptr(d_x), 0 if d_x is None else _row_stride(d_x), _ptr(ws), ws.numel(),
                                        _stream_ptr(dev)), "snn_mlp_bwd_ex")
        return out


class CriticEngine(MlpEngine):
    """The dense critic: an MlpEngine that must end in Linear(., 1).  Used by the fused PPO update and by the critic
    branch of the graphed rollout step (PPO._act_fused); a direct `ActorCritic.evaluate` call keeps the torch path."""

    def __init__(self, linears):
        if len(linears) < 2 or linears[-1].out_features != 1:
            raise ValueError("critic must end in Linear(., 1) and have at least one hidden layer")
        super().__init__(linears)
