"""Generate the golden fixtures by running the UNMODIFIED reference modules from /root/reference.

Runs only in the build container (the GPU box has no /root/reference); its outputs
(tests/golden/*.npz) are committed.  Usage:  python tests/golden/make_golden.py

For every case in cases.py: load the case's seeded state-dict into the reference
PopSpikeActor of that fork, run forward on the case's observations, record the encoder spikes,
every LIF layer's (voltage, spike) trace (by wrapping SpikeMLP.neuron_model), mu and std, then
backpropagate sum(mu * G) and record all parameter gradients and d/d obs.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import (CASES, PPO_CASE, RMA_CASE, make_gmu, make_obs, make_ppo_rollout, make_ppo_state_dict,  # noqa: E402
                   make_rma_rollout, make_rma_state_dict, make_state_dict)

REF = "/root/reference"
PKG_ROOT = {
    "amp": f"{REF}/AMP/AMP_for_hardware-main/rsl_rl",
    "cassie": f"{REF}/CASSIE/rsl_rl-master",
    "humanoid": f"{REF}/HUMANOID/pbrs-humanoid-dev/gpu_rl",
    "rma": f"{REF}/RMA/quadruped-robot-master/rl-robotics/rsl_rl",
}


def import_ref(variant: str, sub: str = "modules.actor_critic"):
    """Import rsl_rl.<sub> from the given fork (never rsl_rl.runners: wandb.init at import)."""
    for k in [k for k in sys.modules if k == "rsl_rl" or k.startswith("rsl_rl.")]:
        del sys.modules[k]
    sys.path.insert(0, PKG_ROOT[variant])
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mod = importlib.import_module("rsl_rl." + sub)
    finally:
        sys.path.pop(0)
    return mod


def pack_bits(x: torch.Tensor) -> np.ndarray:
    return np.packbits(x.to(torch.uint8).numpy().reshape(-1))


def run_actor_case(c):
    ac = import_ref(c.variant)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        actor = ac.PopSpikeActor(c.obs_dim, c.act_dim, c.pop, c.pop, list(c.hidden), (-5, 5),
                                 math.sqrt(0.15), spike_ts=c.spike_ts, device="cpu")
    sd = {k[len("actor."):]: v for k, v in make_state_dict(c).items()}
    actor.load_state_dict(sd)
    L = len(c.hidden) + 1
    rec = []
    orig = actor.snn.neuron_model

    def wrapped(syn, pre, cur, volt, spk):
        out = orig(syn, pre, cur, volt, spk)
        rec.append((out[1].detach().clone(), out[2].detach().clone()))
        return out

    actor.snn.neuron_model = wrapped
    enc_out = {}
    h = actor.encoder.register_forward_hook(lambda m, i, o: enc_out.__setitem__("s", o.detach().clone()))
    obs = make_obs(c).requires_grad_(True)
    G = make_gmu(c)
    with contextlib.redirect_stdout(io.StringIO()):
        mu, std = actor(obs)
    h.remove()
    (mu * G).sum().backward()
    T = c.spike_ts
    volt = [torch.stack([rec[t * L + l][0] for t in range(T)]) for l in range(L)]
    spk = [torch.stack([rec[t * L + l][1] for t in range(T)]) for l in range(L)]
    enc = enc_out["s"].permute(2, 0, 1).contiguous()       # [B,K0,T] -> [T,B,K0]
    out = {"mu": mu.detach().numpy(), "std": std.detach().numpy()}
    # log_std only feeds the Normal head, so it has no gradient from mu
    grads = {n: p.grad.detach() for n, p in actor.named_parameters() if p.grad is not None}
    assert set(n for n, _ in actor.named_parameters()) - set(grads) == {"log_std"}
    if c.full:
        out["enc_spikes"] = enc.numpy().astype(np.uint8)
        for l in range(L):
            out[f"volt{l}"] = volt[l].numpy()
            out[f"spk{l}"] = spk[l].numpy().astype(np.uint8)
        for n, g in grads.items():
            out["grad." + n] = g.numpy()
        out["grad.obs"] = obs.grad.numpy()
    else:
        out["enc_spikes_bits"] = pack_bits(enc)
        for l in range(L):
            out[f"spk{l}_bits"] = pack_bits(spk[l])
            out[f"volt{l}_absmax"] = np.float32(volt[l].abs().max())
        for n, g in grads.items():
            flat = g.reshape(-1)
            out["gradnorm." + n] = np.float64(flat.double().norm())
            out["gradsamp." + n] = flat[:: max(1, flat.numel() // 4096)].numpy()
        out["grad.obs"] = obs.grad.numpy()
    return out


def run_gae_case():
    """RolloutStorage.compute_returns of the reference on a seeded rollout (CASSIE copy)."""
    st = import_ref("cassie", "storage.rollout_storage")
    g = torch.Generator().manual_seed(77)
    S, N = 24, 37
    storage = st.RolloutStorage(N, S, [5], [None], [3], device="cpu")
    storage.rewards.copy_(torch.randn(S, N, 1, generator=g))
    storage.values.copy_(torch.randn(S, N, 1, generator=g))
    storage.dones.copy_((torch.rand(S, N, 1, generator=g) < 0.1).to(torch.uint8))
    last = torch.randn(N, 1, generator=g)
    storage.compute_returns(last, 0.99, 0.95)
    return {"returns": storage.returns.numpy(), "advantages": storage.advantages.numpy()}


def run_ppo_case(pc):
    """The reference PPO.update (CASSIE copy) on a seeded rollout: 1 minibatch x `epochs` Adam steps.
    Records what the rollout phase produced (so tests can re-feed it) and the parameters after the update."""
    import_ref(pc.variant, "modules.actor_critic")
    mods = importlib.import_module("rsl_rl.modules")
    algs = importlib.import_module("rsl_rl.algorithms")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = mods.ActorCritic(pc.obs_dim, pc.obs_dim, pc.act_dim, actor_hidden_dims=list(pc.hidden),
                              critic_hidden_dims=list(pc.critic_hidden), activation="elu", init_noise_std=1.0)
    ac.actor.encoder.device = ac.actor.snn.device = "cpu"
    ac.load_state_dict(make_ppo_state_dict(pc))
    alg = algs.PPO(ac, num_learning_epochs=pc.epochs, num_mini_batches=1, clip_param=0.2, gamma=pc.gamma,
                   lam=pc.lam, value_loss_coef=1.0, entropy_coef=pc.entropy_coef, learning_rate=pc.lr,
                   max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cpu")
    alg.init_storage(pc.num_envs, pc.num_steps, [pc.obs_dim], [None], [pc.act_dim])
    ro = make_ppo_rollout(pc)
    st = alg.storage
    with contextlib.redirect_stdout(io.StringIO()), torch.inference_mode():
        for s in range(pc.num_steps):
            mu = ac.act_inference(ro["obs"][s])
            sigma = torch.exp(ac.actor.log_std).expand_as(mu)
            actions = mu + sigma * ro["noise"][s]
            ac.update_distribution(ro["obs"][s])
            st.observations[s].copy_(ro["obs"][s])
            st.actions[s].copy_(actions)
            st.rewards[s].copy_(ro["rewards"][s].view(-1, 1))
            st.dones[s].copy_(ro["dones"][s].view(-1, 1))
            st.values[s].copy_(ac.evaluate(ro["obs"][s]))
            st.actions_log_prob[s].copy_(ac.get_actions_log_prob(actions).view(-1, 1))
            st.mu[s].copy_(ac.action_mean)
            st.sigma[s].copy_(ac.action_std)
            st.step += 1
        alg.compute_returns(ro["last_obs"])
    out = {"actions": st.actions.numpy().copy(), "values": st.values.numpy().copy(),
           "actions_log_prob": st.actions_log_prob.numpy().copy(), "mu": st.mu.numpy().copy(),
           "sigma": st.sigma.numpy().copy(), "returns": st.returns.numpy().copy(),
           "advantages": st.advantages.numpy().copy()}
    with contextlib.redirect_stdout(io.StringIO()):
        vl, sl = alg.update()
    out["mean_value_loss"] = np.float64(vl)
    out["mean_surrogate_loss"] = np.float64(sl)
    out["final_lr"] = np.float64(alg.learning_rate)
    for k, v in ac.state_dict().items():
        out["final." + k] = v.detach().numpy().copy()
    return out


def run_rma_ppo_case(rc):
    """The RMA fork's PPO.update (RMA-rl/algorithms/ppo.py:134-221) on a seeded rollout: 1 minibatch x `epochs` PPO
    steps, each followed by `substeps` adaptation-module regression steps (second Adam, fixed lr).  Records what the
    rollout phase produced and every parameter after the update."""
    import_ref("rma", "modules.actor_critic")
    mods = importlib.import_module("rsl_rl.modules")
    algs = importlib.import_module("rsl_rl.algorithms")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = mods.ActorCritic(rc.num_obs, rc.num_priv, rc.num_hist, rc.act_dim, actor_hidden_dims=list(rc.hidden),
                              critic_hidden_dims=list(rc.critic_hidden), encoder_hidden_dims=list(rc.enc_hidden),
                              adaptation_hidden_dims=list(rc.adapt_hidden), encoder_latent_dims=rc.latent,
                              activation="elu", init_noise_std=1.0)
    ac.actor.encoder.device = ac.actor.snn.device = "cpu"
    ac.load_state_dict(make_rma_state_dict(rc))
    alg = algs.PPO(ac, num_learning_epochs=rc.epochs, num_mini_batches=1, clip_param=0.2, gamma=rc.gamma, lam=rc.lam,
                   value_loss_coef=1.0, entropy_coef=rc.entropy_coef, learning_rate=rc.lr, max_grad_norm=1.0,
                   use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01,
                   num_adaptation_module_substeps=rc.substeps, device="cpu")
    alg.init_storage(rc.num_envs, rc.num_steps, [rc.num_obs], [rc.num_priv], [rc.num_hist], [rc.act_dim])
    ro = make_rma_rollout(rc)
    st = alg.storage
    with contextlib.redirect_stdout(io.StringIO()), torch.inference_mode():
        for s in range(rc.num_steps):
            obs, priv = ro["obs"][s], ro["priv"][s]
            ac.update_distribution(obs, priv)
            mu = ac.action_mean
            actions = mu + ac.action_std * ro["noise"][s]
            st.observations[s].copy_(obs)
            st.privileged_observations[s].copy_(priv)
            st.observation_histories[s].copy_(ro["hist"][s])
            st.actions[s].copy_(actions)
            st.rewards[s].copy_(ro["rewards"][s].view(-1, 1))
            st.dones[s].copy_(ro["dones"][s].view(-1, 1))
            st.values[s].copy_(ac.evaluate(obs, priv))
            st.actions_log_prob[s].copy_(ac.get_actions_log_prob(actions).view(-1, 1))
            st.mu[s].copy_(mu)
            st.sigma[s].copy_(ac.action_std)
            st.step += 1
        alg.compute_returns(ro["last_obs"], ro["last_priv"])
    out = {"actions": st.actions.numpy().copy(), "values": st.values.numpy().copy(),
           "actions_log_prob": st.actions_log_prob.numpy().copy(), "mu": st.mu.numpy().copy(),
           "sigma": st.sigma.numpy().copy(), "returns": st.returns.numpy().copy(),
           "advantages": st.advantages.numpy().copy()}
    with contextlib.redirect_stdout(io.StringIO()):
        vl, sl, al = alg.update()
    out["mean_value_loss"] = np.float64(vl)
    out["mean_surrogate_loss"] = np.float64(sl)
    out["mean_adaptation_loss"] = np.float64(al)
    out["final_lr"] = np.float64(alg.learning_rate)
    for k, v in ac.state_dict().items():
        out["final." + k] = v.detach().numpy().copy()
    return out


def _emit(path, out, check):
    """Write the fixture, or (--check) compare a fresh run of the reference against the committed file."""
    if not check:
        np.savez_compressed(path, **out)
        return "written"
    ref = dict(np.load(path))
    assert set(ref) == set(out), f"{path}: keys differ: {sorted(set(ref) ^ set(out))}"
    worst = 0.0
    for k, v in out.items():
        a, b = np.asarray(v), ref[k]
        assert a.shape == b.shape and a.dtype == b.dtype, f"{path}:{k}: {a.dtype}{a.shape} vs {b.dtype}{b.shape}"
        if not np.array_equal(a, b):
            d = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))))
            worst = max(worst, d / (float(np.max(np.abs(b.astype(np.float64)))) + 1e-30))
    return "bit-identical" if worst == 0.0 else f"max rel diff {worst:.2e}"


def main(check=False):
    """`--check`: re-run the unmodified reference here and compare with the committed fixtures instead of writing."""
    for c in CASES:
        out = run_actor_case(c)
        path = os.path.join(HERE, f"actor_{c.name}.npz")
        print(f"{c.name:22s} {_emit(path, out, check)}  mu[0]={out['mu'][0][:3]}")
    print("gae_cassie", _emit(os.path.join(HERE, "gae_cassie.npz"), run_gae_case(), check))
    print(PPO_CASE.name, _emit(os.path.join(HERE, f"{PPO_CASE.name}.npz"), run_ppo_case(PPO_CASE), check))
    print(RMA_CASE.name, _emit(os.path.join(HERE, f"{RMA_CASE.name}.npz"), run_rma_ppo_case(RMA_CASE), check))


if __name__ == "__main__":
    main(check="--check" in sys.argv[1:])
