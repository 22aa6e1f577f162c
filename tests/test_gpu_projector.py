"""GPU parity of the projection-head kernel (``cfm_projector_fwd/bwd`` through ``ops.projector_heads``) against the
reference's own expression of the heads — ``F.normalize(nn.Sequential(Linear, ReLU, Linear)(x), dim=1)``
(contrastive.py:41-50, 88-97) — evaluated by torch on the CPU in float64.  fp32 FMA arithmetic: the bar is the
fp32 one, 1e-5 of the tensor's scale for outputs and for every gradient."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import assert_close_scaled

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _head(d_in, d_hid, d_out, seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(d_in, d_hid), nn.ReLU(), nn.Linear(d_hid, d_out))


def _reference(seq, x, g):
    """float64 forward + backward of the reference expression; returns (out, dx, {param grads})."""
    ref = nn.Sequential(nn.Linear(seq[0].in_features, seq[0].out_features), nn.ReLU(),
                        nn.Linear(seq[2].in_features, seq[2].out_features)).double()
    ref.load_state_dict({k: v.double() for k, v in seq.state_dict().items()})
    xr = x.double().clone().requires_grad_(True)
    out = F.normalize(ref(xr), dim=1)
    out.backward(g.double())
    return out.detach(), xr.grad, {k: p.grad for k, p in ref.named_parameters()}


@pytest.mark.parametrize("B", [1, 63, 64, 65, 300, 4097])
def test_two_heads_forward_and_every_gradient(B):
    """The contrastive model's shape (60 -> 60 -> 30 per side), both heads in one launch, ragged last tiles."""
    from ceo_firm_matching import ops
    gen = torch.Generator().manual_seed(B)
    heads = [_head(60, 60, 30, 1), _head(60, 60, 30, 2)]
    xs = [F.normalize(torch.randn(B, 60, generator=gen), dim=1) for _ in heads]
    gs = [torch.randn(B, 30, generator=gen) for _ in heads]
    dheads = [_head(60, 60, 30, 1).to(DEV), _head(60, 60, 30, 2).to(DEV)]
    dxs = [x.to(DEV).requires_grad_(True) for x in xs]
    outs = ops.projector_heads(list(zip(dxs, dheads)), 1e-12)
    torch.autograd.backward(outs, [g.to(DEV) for g in gs])
    for i, (seq, x, g) in enumerate(zip(heads, xs, gs)):
        out_r, dx_r, pg_r = _reference(seq, x, g)
        assert_close_scaled(outs[i], out_r, 1e-5, f"head {i} output")
        assert torch.allclose(outs[i].norm(dim=1), torch.ones(B, device=DEV), atol=1e-5)
        assert_close_scaled(dxs[i].grad, dx_r, 1e-5, f"head {i} dx")
        for k, p in dheads[i].named_parameters():
            assert_close_scaled(p.grad, pg_r[k], 1e-5, f"head {i} grad {k}")


@pytest.mark.parametrize("dims", [(7, 5, 3), (64, 64, 64), (33, 17, 2), (12, 64, 5)])
def test_other_widths_single_head(dims):
    from ceo_firm_matching import ops
    d_in, d_hid, d_out = dims
    B = 130
    gen = torch.Generator().manual_seed(d_in)
    seq = _head(d_in, d_hid, d_out, 3)
    x, g = torch.randn(B, d_in, generator=gen), torch.randn(B, d_out, generator=gen)
    dseq = _head(d_in, d_hid, d_out, 3).to(DEV)
    dx = x.to(DEV).requires_grad_(True)
    (out,) = ops.projector_heads([(dx, dseq)], 1e-12)
    out.backward(g.to(DEV))
    out_r, dx_r, pg_r = _reference(seq, x, g)
    assert_close_scaled(out, out_r, 1e-5, "output")
    assert_close_scaled(dx.grad, dx_r, 1e-5, "dx", floor=1e-6)
    for k, p in dseq.named_parameters():
        assert_close_scaled(p.grad, pg_r[k], 1e-5, "grad " + k, floor=1e-6)


def test_bitwise_reproducible_and_wide_layers_refused():
    """Per-CTA partials are added in CTA order: two runs give identical bits.  Layers wider than 64 raise."""
    from ceo_firm_matching import ops
    from ceo_firm_matching._native import CfmError
    gen = torch.Generator().manual_seed(0)
    B = 20000
    x, g = torch.randn(B, 60, generator=gen).to(DEV), torch.randn(B, 30, generator=gen).to(DEV)
    grads = []
    for _ in range(2):
        seq = _head(60, 60, 30, 4).to(DEV)
        (out,) = ops.projector_heads([(x, seq)], 1e-12)
        out.backward(g)
        grads.append([p.grad.clone() for p in seq.parameters()] + [out.detach().clone()])
    for a, b in zip(*grads):
        assert torch.equal(a, b)
    with pytest.raises(CfmError):
        ops.projector_heads([(torch.randn(8, 65, device=DEV), _head(65, 60, 30, 0).to(DEV))], 1e-12)
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.projector_heads([(torch.randn(8, 60), _head(60, 60, 30, 0))], 1e-12)
