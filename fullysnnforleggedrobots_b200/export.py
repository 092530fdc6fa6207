"""Checkpoint compatibility and the deployment artefact.

Reference behaviour replaced here:
  * OnPolicyRunner.save / load (CASSIE/rsl_rl-master/rsl_rl/runners/on_policy_runner.py:224-244): a dict with
    'model_state_dict', 'optimizer_state_dict', 'iter', 'infos'.  `save_checkpoint` / `load_checkpoint` read and write
    exactly that layout, so `model_*.pt` files move between the reference and this package in both directions (the
    state-dict keys are the reference's, FlatAdam's state dict has torch.optim.Adam's layout).
  * export_policy_as_jit (legged_gym/utils/helpers.py:180-191) deep-copies `.actor` and pickles the module with
    torch.jit — which cannot carry a custom-kernel actor.  `export_policy` writes a self-contained artefact instead:
    the network description (dimensions + per-fork switches) and the actor's tensors, nothing pickled
    (`torch.load(..., weights_only=True)` reads it).  `load_policy` rebuilds a PopSpikeActor from it and returns
    the `obs -> mu` callable that `get_inference_policy` (on_policy_runner.py:240-244) hands to play.py.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from .actor import PopSpikeActor

ARTEFACT_FORMAT = "snn_b200_policy_v1"


def save_checkpoint(path, alg, it: int = 0, infos=None):
    torch.save({'model_state_dict': alg.actor_critic.state_dict(),
                'optimizer_state_dict': alg.optimizer.state_dict(),
                'iter': it, 'infos': infos}, path)


def load_checkpoint(path_or_dict, alg, load_optimizer: bool = True):
    """Returns `infos` like the reference runner's load()."""
    d = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu",
                                                                       weights_only=False)
    # in-place copy: the parameters are views of FlatAdam's flat buffer and must stay so
    alg.actor_critic.load_state_dict(d['model_state_dict'])
    if load_optimizer and d.get('optimizer_state_dict') is not None:
        alg.optimizer.load_state_dict(d['optimizer_state_dict'])
        if hasattr(alg, "learning_rate"):
            alg.learning_rate = alg.optimizer.param_groups[0]["lr"]
    alg.actor_critic.actor.engine.invalidate()          # packed bf16 planes are stale now
    if hasattr(alg, "refresh_packed") and next(alg.actor_critic.parameters()).is_cuda:
        alg.refresh_packed()
    return d.get('infos')


def export_policy(actor_critic_or_actor, path):
    actor = getattr(actor_critic_or_actor, "actor", actor_critic_or_actor)
    if not isinstance(actor, PopSpikeActor):
        raise TypeError("export_policy expects an ActorCritic of this package or its PopSpikeActor")
    desc = {"obs_dim": actor.obs_dim, "act_dim": actor.act_dim, "en_pop": actor.encoder.pop_dim,
            "de_pop": actor.decoder.pop_dim, "hidden": [int(h) for h in actor.snn.hidden_sizes],
            "spike_ts": actor.spike_ts, "enc_eps": float(actor.enc_eps), "dec_tanh": bool(actor.dec_tanh),
            "fwd_planes": int(actor.fwd_planes)}
    torch.save({"format": ARTEFACT_FORMAT, "desc": desc,
                "state_dict": {k: v.detach().cpu().clone() for k, v in actor.state_dict().items()}}, path)
    return desc


class InferencePolicy:
    """obs [B, obs_dim] (CUDA, fp32) -> action means [B, act_dim]; no autograd, no sampling."""

    def __init__(self, actor: PopSpikeActor):
        self.actor = actor.eval()

    @torch.inference_mode()
    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        return self.actor.mean_only(obs)


def load_policy(path, device: Optional[str] = "cuda") -> InferencePolicy:
    d = torch.load(path, map_location="cpu", weights_only=True)
    if d.get("format") != ARTEFACT_FORMAT:
        raise ValueError(f"not a {ARTEFACT_FORMAT} artefact: {d.get('format')!r}")
    c = d["desc"]
    actor = PopSpikeActor(c["obs_dim"], c["act_dim"], c["en_pop"], c["de_pop"], list(c["hidden"]), (-5, 5),
                          math.sqrt(0.15), spike_ts=c["spike_ts"], enc_eps=c["enc_eps"], dec_tanh=c["dec_tanh"],
                          fwd_planes=c["fwd_planes"])
    actor.load_state_dict(d["state_dict"])
    return InferencePolicy(actor.to(device))
