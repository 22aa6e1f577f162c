"""FusedAdam (cfm_adam_step) against torch.optim.Adam: bit-equal parameters and state over several steps, state_dict
interchange both ways, and use inside the CUDA-graph training step."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPES = [(1000, 48), (7,), (64, 205), (1,), (333, 8), (70001,), (3, 5, 7)]


def _params(seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return [torch.randn(s, device=DEV, generator=g).requires_grad_(True) for s in SHAPES], g


def _set_grads(pa, pb, g, it):
    for a, b in zip(pa, pb):
        gr = torch.randn(a.shape, device=DEV, generator=g) * (10.0 ** (it % 5 - 3))
        if it == 3:
            gr.view(-1)[::2] = 0                      # untouched embedding rows have exactly-zero gradients
        a.grad, b.grad = gr.clone(), gr.clone()


@pytest.mark.parametrize("lr,betas,eps", [(1e-3, (0.9, 0.999), 1e-8), (3e-2, (0.8, 0.99), 1e-6)])
def test_bit_equal_to_torch_adam(lr, betas, eps):
    from ceo_firm_matching.optim import FusedAdam
    pa, g = _params(0)
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = torch.optim.Adam(pa, lr=lr, betas=betas, eps=eps, capturable=True, foreach=True)
    ob = FusedAdam(pb, lr=lr, betas=betas, eps=eps)
    for it in range(8):
        _set_grads(pa, pb, g, it)
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)
        for k in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(oa.state[a][k], ob.state[b][k]), k
        assert float(oa.state[a]["step"]) == float(ob.state[b]["step"]) == 8.0


def test_state_dict_round_trips_with_torch_adam():
    from ceo_firm_matching.optim import FusedAdam
    pa, g = _params(1)
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa = torch.optim.Adam(pa, lr=1e-3, capturable=True, foreach=True)
    ob = FusedAdam(pb, lr=1e-3)
    for it in range(3):
        _set_grads(pa, pb, g, it)
        oa.step()
        ob.step()
    # swap the optimiser states: fused continues from torch's checkpoint and vice versa
    sa, sb = oa.state_dict(), ob.state_dict()
    oa.load_state_dict(sb)
    ob.load_state_dict(sa)
    for it in range(3, 6):
        _set_grads(pa, pb, g, it)
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)


def test_skips_parameters_without_gradient_and_rejects_bad_inputs():
    from ceo_firm_matching.optim import FusedAdam
    pa, g = _params(2)
    opt = FusedAdam(pa, lr=1e-3)
    before = [p.detach().clone() for p in pa]
    pa[0].grad = torch.ones_like(pa[0])
    opt.step()
    assert not torch.equal(pa[0], before[0]) and all(torch.equal(p, b) for p, b in zip(pa[1:], before[1:]))
    pa[1].grad = torch.ones(7, 2, device=DEV)[:, 0]            # non-contiguous gradient: no silent fallback
    with pytest.raises(RuntimeError):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam(pa, lr=1e-3, betas=(0.4, 0.999))


def test_inside_the_graphed_training_step_matches_torch_adam():
    """The whole step (fwd, loss, bwd, fused Adam) as one graph replay == the same step with torch's Adam."""
    import oracle
    from helpers import load_into
    from ceo_firm_matching import CEOFirmMatcher, Config
    from ceo_firm_matching.optim import FusedAdam
    from ceo_firm_matching.training import GraphedTwoTowerStep
    f_cards, c_cards, B = [50, 5, 3, 2], [2, 4, 30, 2, 2, 5, 2], 256
    p = oracle.init_two_tower_params(12, f_cards, 2, c_cards, seed=5)
    meta = {"n_firm_numeric": 12, "firm_cat_counts": f_cards, "n_ceo_numeric": 2, "ceo_cat_counts": c_cards}
    gen = torch.Generator().manual_seed(9)
    batches = [[x.to(DEV) for x in (
        torch.randn(B, 12, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in f_cards], 1),
        torch.randn(B, 2, generator=gen), torch.stack([torch.randint(0, n, (B,), generator=gen) for n in c_cards], 1),
        torch.randn(B, 1, generator=gen), torch.rand(B, 1, generator=gen) + 0.5)] for _ in range(4)]
    finals = []
    for fused in (False, True):
        m = load_into(CEOFirmMatcher(meta, Config()), p).to(DEV).train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.use_persistent_table_grads(True)
        opt = FusedAdam(m.parameters(), lr=1e-3) if fused else torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
        runner = GraphedTwoTowerStep(m, batches[0], optimizer=opt, warmup=1)
        for b in batches:
            runner.step(b)
        torch.cuda.synchronize()
        finals.append({k: v.detach().clone() for k, v in m.state_dict().items()})
    for k in finals[0]:
        assert torch.equal(finals[0][k], finals[1][k]), k
