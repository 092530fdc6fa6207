"""The general dense ELU MLP on the tensor-core kernels (snn_mlp_fwd_ex / snn_mlp_bwd_ex) against plain PyTorch fp32:
the RMA fork's env_factor_encoder / adaptation_module (RMA/.../rsl_rl/modules/actor_critic.py:394-418; out_dim-wide
linear head, rows that are column slices of the cat(obs, latent) buffer, the SUM of two gradient sources) and the
critic with a gradient into its input (RMA: the critic reads cat(obs, latent) too, :633-634)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _mlp(dims):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(torch.nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(torch.nn.ELU())
    return torch.nn.Sequential(*layers).cuda()


@pytest.mark.parametrize("dims,B", [([17, 256, 128, 18], 1000), ([9, 24, 12, 6], 192), ([1530, 256, 32, 18], 3072),
                                    ([50, 64, 2], 130)])
def test_wide_head_mlp_matches_torch(dims, B):
    from fullysnnforleggedrobots_b200 import MlpEngine
    torch.manual_seed(sum(dims) + B)
    seq = _mlp(dims)
    eng = MlpEngine.from_sequential(seq)
    assert eng is not None and eng.out_dim == dims[-1]
    O = 11
    wide_in = torch.randn(B, dims[0] + 5, device="cuda")            # the input rows are a column slice of a wider buffer
    x = wide_in[:, 3:3 + dims[0]]
    cat = torch.zeros(B, O + dims[-1], device="cuda")               # the output lands in the latent columns of cat(obs, latent)
    y, trace = eng.forward(x, out_value=cat[:, O:])
    xr = x.detach().clone().requires_grad_(True)
    ref = seq(xr)
    assert rel_err(y.cpu(), ref.detach().cpu()) < 1e-4
    assert torch.equal(cat[:, :O], torch.zeros(B, O, device="cuda"))          # nothing outside the slice is touched
    ga = torch.randn(B, O + dims[-1], device="cuda")
    gb = torch.randn(B, O + dims[-1], device="cuda")
    ref.backward(ga[:, O:] + gb[:, O:])
    d_x = torch.full((B, dims[0] + 2), 7.0, device="cuda")
    gw, gbias = eng.backward(ga[:, O:], trace, B, g2=gb[:, O:], d_x=d_x[:, :dims[0]])
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    for i, l in enumerate(lin):
        assert rel_err(gw[i].cpu(), l.weight.grad.cpu()) < 1e-4, f"dW{i}"
        assert rel_err(gbias[i].cpu(), l.bias.grad.cpu()) < 1e-4, f"db{i}"
    assert rel_err(d_x[:, :dims[0]].cpu(), xr.grad.cpu()) < 1e-4
    assert torch.equal(d_x[:, dims[0]:], torch.full((B, 2), 7.0, device="cuda"))


@pytest.mark.parametrize("dims,B", [([66, 256, 256, 1], 3072), ([30, 32, 16, 1], 200)])
def test_critic_input_gradient_matches_torch(dims, B):
    from fullysnnforleggedrobots_b200 import CriticEngine
    torch.manual_seed(5 + B)
    seq = _mlp(dims)
    ce = CriticEngine.from_sequential(seq)
    x = torch.randn(B, dims[0], device="cuda")
    gv = torch.randn(B, device="cuda")
    value, trace = ce.forward(x)
    xr = x.clone().requires_grad_(True)
    ref = seq(xr)
    assert rel_err(value.cpu(), ref.detach().squeeze(-1).cpu()) < 1e-4
    ref.backward(gv.view_as(ref))
    d_x = torch.empty(B, dims[0], device="cuda")
    gw, gb = ce.backward(gv, trace, B, d_x=d_x)
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    for i, l in enumerate(lin):
        assert rel_err(gw[i].cpu(), l.weight.grad.cpu()) < 1e-4, f"dW{i}"
        assert rel_err(gb[i].cpu(), l.bias.grad.cpu()) < 1e-4, f"db{i}"
    assert rel_err(d_x.cpu(), xr.grad.cpu()) < 1e-4


def test_mse_grad_matches_torch():
    import ctypes as C

    from fullysnnforleggedrobots_b200 import _lib
    torch.manual_seed(3)
    B, D = 3072, 18
    pred = torch.randn(B, D, device="cuda", requires_grad=True)
    tgt = torch.randn(B, D, device="cuda")
    g = torch.empty(B, D, device="cuda")
    acc = torch.full((1,), 0.25, device="cuda")
    scratch = torch.empty(4096, dtype=torch.uint8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().snn_mse_grad(p(pred), p(tgt), B * D, p(g), p(acc), p(scratch),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "snn_mse_grad")
    loss = torch.nn.functional.mse_loss(pred, tgt)
    loss.backward()
    assert abs(acc.item() - 0.25 - loss.item()) < 1e-5 * loss.item()
    assert rel_err(g.cpu(), pred.grad.cpu()) < 1e-6
