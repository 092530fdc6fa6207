"""PPORMA.update at RMA-actual dimensions (RMA/.../legged_gym/envs/base/legged_robot_config.py:38-41,283-305 and
lite3_config.py:57-59: 320 observations + 18-dim latent -> [512, 256, 128] spiking actor, 54 privileged observations,
40 x 320 observation history, 4096 envs x 24 steps, 5 epochs x 4 minibatches), fused + CUDA graph vs the autograd path."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fullysnnforleggedrobots_b200 import PPORMA, ActorCriticRMA

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S, no, npriv, nh, na = 24, 320, 54, int(os.environ.get("RMA_HIST", 40 * 320)), 12


def run(fused):
    torch.manual_seed(0)
    ac = ActorCriticRMA(no, npriv, nh, na, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[512, 256, 128])
    alg = PPORMA(ac, num_learning_epochs=5, num_mini_batches=4, entropy_coef=0.01, schedule="adaptive", device="cuda",
                 num_adaptation_module_substeps=1, fused=fused)
    alg.init_storage(N, S, [no], [npriv], [nh], [na])

    def rollout():
        with torch.inference_mode():
            for _ in range(S):
                obs, priv, hist = (torch.randn(N, no, device="cuda"), torch.randn(N, npriv, device="cuda"),
                                   torch.randn(N, nh, device="cuda"))
                alg.act(obs, priv, hist)
                alg.process_env_step(torch.randn(N, device="cuda"), torch.rand(N, device="cuda") < 0.02, {})
            alg.compute_returns(obs, priv)

    times = []
    for it in range(4):
        rollout()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = alg.update()
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    flops = 2 * 5 * (3380 * 512 + 512 * 256 + 256 * 128 + 128 * 120)          # actor forward per sample
    per_update = N * S * 5 * flops * 3.0                                        # fwd + dX (incl. d obs) + dW
    print(f"fused={fused} graphs={len(alg._graphs)} update ms {['%.2f' % t for t in times]}  losses {out}  "
          f"actor algo TFLOP/s (best) {per_update / (min(times) * 1e-3) / 1e12:.1f}")
    return min(times)


a = run(True)
b = run(False)
print(f"fused / generic = {a:.2f} / {b:.2f} ms -> x{b / a:.2f}")
