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


# ---------------------------------------------------------------- optimizer state: reference layout <-> ours
# The reference keeps ONE torch.optim.Adam over `actor_critic.parameters()` — which starts with the dead `std`
# parameter (index 0, never gets a state entry) — plus, in the AMP fork, two discriminator groups (amp_ppo.py:87-92).
# Here the actor-critic parameters live in FlatAdam without `std`, the discriminator's in a second Adam.  Checkpoints
# are written and read in the REFERENCE's indexing so that `model_*.pt` files move both ways.

def _index_maps(alg):
    full = list(alg.actor_critic.parameters())
    pos = {id(p): i for i, p in enumerate(full)}
    ours_to_ref = [pos[id(p)] for p in alg.optimizer.params]
    return len(full), ours_to_ref


def optimizer_state_to_reference(alg) -> dict:
    n_full, ours_to_ref = _index_maps(alg)
    sd = alg.optimizer.state_dict()
    g0 = {k: v for k, v in sd["param_groups"][0].items() if k != "params"}
    g0["lr"] = float(g0["lr"])
    def plain(st):      # torch.optim.Adam's per-parameter entry with a host step count (any torch version loads it)
        step = st["step"]
        return {"step": float(step.item()) if torch.is_tensor(step) else float(step),
                "exp_avg": st["exp_avg"].detach().clone(), "exp_avg_sq": st["exp_avg_sq"].detach().clone()}

    state = {ours_to_ref[j]: plain(st) for j, st in sd["state"].items()}
    groups = [dict(g0, params=list(range(n_full)), name="actor_critic")]
    disc_opt = getattr(alg, "disc_optimizer", None)
    if disc_opt is not None:
        dsd = disc_opt.state_dict()
        off = n_full
        for g in dsd["param_groups"]:
            ids = [off + i for i in g["params"]]
            groups.append(dict(g0, params=ids, weight_decay=g["weight_decay"], name=g.get("name", "amp")))
            for i in g["params"]:
                if i in dsd["state"]:
                    state[off + i] = plain(dsd["state"][i])
    return {"state": state, "param_groups": groups}


def split_reference_optimizer_state(sd: dict, n_full: int, ours_to_ref, n_ours: int):
    """Pure dict surgery (unit-tested on the CPU): reference-indexed state dict -> (actor-critic part in FlatAdam's
    indexing, discriminator part in the second Adam's indexing or None).  Also accepts the std-less layout earlier
    versions of this package wrote."""
    groups = sd["param_groups"]
    ids0 = list(groups[0]["params"])
    if len(ids0) == n_ours:                                   # our own old layout: indices already FlatAdam's
        ref_to_ours = {i: j for j, i in enumerate(ids0)}
    elif len(ids0) == n_full:
        ref_to_ours = {ids0[r]: j for j, r in enumerate(ours_to_ref)}
    else:
        raise ValueError(f"optimizer state has {len(ids0)} actor-critic parameters, expected {n_full} (reference "
                         f"layout) or {n_ours}")
    ac_state = {ref_to_ours[i]: st for i, st in sd["state"].items() if i in ref_to_ours}
    ac = {"state": ac_state, "param_groups": [dict({k: v for k, v in groups[0].items() if k != "name"},
                                                   params=list(range(n_ours)))]}
    disc = None
    if len(groups) > 1:
        dstate, dgroups, nxt = {}, [], 0
        for g in groups[1:]:
            new_ids = list(range(nxt, nxt + len(g["params"])))
            for old, new in zip(g["params"], new_ids):
                if old in sd["state"]:
                    dstate[new] = sd["state"][old]
            dgroups.append(dict(g, params=new_ids))
            nxt += len(g["params"])
        disc = {"state": dstate, "param_groups": dgroups}
    return ac, disc


def _normalizer_state(norm):
    if norm is None:
        return None
    to_np = lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else x
    return {"mean": to_np(norm.mean), "var": to_np(norm.var), "count": float(norm.count)}


def _load_normalizer(norm, saved):
    """`saved`: the dict written here, or the reference's pickled Normalizer (numpy mean / var / count attributes)."""
    if norm is None or saved is None:
        return
    get = (lambda k: saved[k]) if isinstance(saved, dict) else (lambda k: getattr(saved, k))
    norm.mean.copy_(torch.as_tensor(get("mean"), dtype=norm.mean.dtype))
    norm.var.copy_(torch.as_tensor(get("var"), dtype=norm.var.dtype))
    if torch.is_tensor(norm.count):
        norm.count.fill_(float(get("count")))
    else:
        norm.count = float(get("count"))


def save_checkpoint(path, alg, it: int = 0, infos=None):
    d = {'model_state_dict': alg.actor_critic.state_dict(),
         'optimizer_state_dict': optimizer_state_to_reference(alg),
         'iter': it, 'infos': infos}
    if hasattr(alg, "discriminator"):                          # AMP fork (amp_on_policy_runner.py:268-276)
        d['discriminator_state_dict'] = alg.discriminator.state_dict()
        d['amp_normalizer'] = _normalizer_state(alg.amp_normalizer)
    torch.save(d, path)


def load_checkpoint(path_or_dict, alg, load_optimizer: bool = True):
    """Returns `infos` like the reference runner's load()."""
    d = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu",
                                                                       weights_only=False)
    # in-place copy: the parameters are views of FlatAdam's flat buffer and must stay so
    alg.actor_critic.load_state_dict(d['model_state_dict'])
    if hasattr(alg, "discriminator") and d.get('discriminator_state_dict') is not None:
        alg.discriminator.load_state_dict(d['discriminator_state_dict'])
        _load_normalizer(alg.amp_normalizer, d.get('amp_normalizer'))
    if load_optimizer and d.get('optimizer_state_dict') is not None:
        n_full, ours_to_ref = _index_maps(alg)
        ac_sd, disc_sd = split_reference_optimizer_state(d['optimizer_state_dict'], n_full, ours_to_ref,
                                                         len(alg.optimizer.params))
        alg.optimizer.load_state_dict(ac_sd)
        if hasattr(alg, "learning_rate"):
            alg.learning_rate = float(alg.optimizer.param_groups[0]["lr"])
        disc_opt = getattr(alg, "disc_optimizer", None)
        if disc_opt is not None and disc_sd is not None:
            disc_opt.load_state_dict(disc_sd)
            for g in disc_opt.param_groups:                    # keep the learning rate a view of the device scalar
                g["lr"], g["capturable"], g["foreach"] = alg.optimizer.lr_dev[0], True, False
            dev = alg.optimizer.flat_params.device
            for st in disc_opt.state.values():                 # capturable Adam: step counters live on the device
                if "step" in st:
                    st["step"] = torch.as_tensor(st["step"], dtype=torch.float32).to(dev)
            if getattr(alg, "_disc_state", None) is not None:
                alg._disc_state = None                         # a captured discriminator graph holds the old state
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
