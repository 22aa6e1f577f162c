"""The reference's own model unit tests, restated one for one against the CUDA build:
``/root/reference/tests/test_model.py:23-84`` and ``/root/reference/tests/test_structural_model.py:34-178`` — same
fixtures, same assertions, tensors on ``cuda`` (the product has no CPU path).  The bilinear known-answer test
(``test_structural_model.py:158-178``) lives in ``test_gpu_structural.py::test_reference_known_answer_bilinear``."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ---- test_model.py ---------------------------------------------------------------------------------------------
@pytest.fixture
def sample_metadata():
    return {"n_firm_numeric": 12, "firm_cat_counts": [4, 4, 2, 2], "n_ceo_numeric": 2,
            "ceo_cat_counts": [2, 4, 2, 2, 2, 2, 2]}


def _two_tower_inputs(md, batch_size):
    return (torch.randn(batch_size, md["n_firm_numeric"], device=DEV),
            torch.randint(0, 2, (batch_size, len(md["firm_cat_counts"])), device=DEV),
            torch.randn(batch_size, md["n_ceo_numeric"], device=DEV),
            torch.randint(0, 2, (batch_size, len(md["ceo_cat_counts"])), device=DEV))


def test_model_initialization(sample_metadata):                       # test_model.py:23-31
    from ceo_firm_matching import CEOFirmMatcher, Config
    model = CEOFirmMatcher(sample_metadata, Config())
    assert model is not None
    assert hasattr(model, "firm_tower") and hasattr(model, "ceo_tower") and hasattr(model, "logit_scale")


def test_model_forward_shape(sample_metadata):                        # test_model.py:33-50
    from ceo_firm_matching import CEOFirmMatcher, Config
    model = CEOFirmMatcher(sample_metadata, Config()).to(DEV)
    model.eval()
    with torch.no_grad():
        output = model(*_two_tower_inputs(sample_metadata, 32))
    assert output.shape == (32, 1)


def test_model_gradients_flow(sample_metadata):                       # test_model.py:52-74
    from ceo_firm_matching import CEOFirmMatcher, Config
    model = CEOFirmMatcher(sample_metadata, Config()).to(DEV)
    model.train()
    target = torch.randn(8, 1, device=DEV)
    output = model(*_two_tower_inputs(sample_metadata, 8))
    loss = ((output - target) ** 2).mean()
    loss.backward()
    for name, param in model.named_parameters():
        if param.requires_grad:
            assert param.grad is not None, f"No gradient for {name}"
            assert torch.isfinite(param.grad).all(), f"Non-finite gradient for {name}"


def test_embedding_dimensions(sample_metadata):                       # test_model.py:76-84
    from ceo_firm_matching import CEOFirmMatcher, Config
    model = CEOFirmMatcher(sample_metadata, Config())
    assert len(model.firm_embeddings) == len(sample_metadata["firm_cat_counts"])
    assert len(model.ceo_embeddings) == len(sample_metadata["ceo_cat_counts"])


# ---- test_structural_model.py ------------------------------------------------------------------------------------
@pytest.fixture
def structural_metadata():
    return {"n_firm_num": 12, "n_ceo_num": 2, "firm_cat_cards": [4, 4, 2, 2], "ceo_cat_cards": [2, 4, 2, 2, 2, 2, 2]}


@pytest.fixture
def model(structural_metadata):
    from ceo_firm_matching import StructuralConfig, StructuralDistillationNet
    return StructuralDistillationNet(structural_metadata, StructuralConfig()).to(DEV)


def _structural_inputs(md, batch_size):
    return (torch.randn(batch_size, md["n_firm_num"], device=DEV),
            torch.randint(0, 2, (batch_size, len(md["firm_cat_cards"])), device=DEV),
            torch.randn(batch_size, md["n_ceo_num"], device=DEV),
            torch.randint(0, 2, (batch_size, len(md["ceo_cat_cards"])), device=DEV))


def test_structural_model_initialization(model):                      # test_structural_model.py:34-41
    assert hasattr(model, "firm_tower") and hasattr(model, "ceo_tower") and hasattr(model, "A")


def test_interaction_matrix_frozen(model):                            # :43-53
    assert "A" not in dict(model.named_parameters())
    assert "A" in [name for name, _ in model.named_buffers()]
    assert not model.A.requires_grad


def test_interaction_matrix_shape(model):                             # :55-57
    assert model.A.shape == (5, 5)


def test_forward_output_shapes(model, structural_metadata):           # :59-76
    model.eval()
    with torch.no_grad():
        c_logits, f_logits, expected_match = model(*_structural_inputs(structural_metadata, 32))
    assert c_logits.shape == (32, 5), "CEO logits should be (batch, 5)"
    assert f_logits.shape == (32, 5), "Firm logits should be (batch, 5)"
    assert expected_match.shape == (32, 1), "Expected match should be (batch, 1)"


def test_logits_produce_valid_probabilities(model, structural_metadata):      # :78-101
    model.eval()
    with torch.no_grad():
        c_logits, f_logits, _ = model(*_structural_inputs(structural_metadata, 16))
    c_probs, f_probs = F.softmax(c_logits, dim=1), F.softmax(f_logits, dim=1)
    ones = torch.ones(16, device=DEV)
    assert torch.allclose(c_probs.sum(dim=1), ones, atol=1e-5) and torch.allclose(f_probs.sum(dim=1), ones, atol=1e-5)
    assert (c_probs >= 0).all() and (c_probs <= 1).all()
    assert (f_probs >= 0).all() and (f_probs <= 1).all()


def test_gradients_flow_to_towers(model, structural_metadata):        # :103-130
    model.train()
    target_ceo = F.softmax(torch.randn(8, 5, device=DEV), dim=1)
    target_firm = F.softmax(torch.randn(8, 5, device=DEV), dim=1)
    c_logits, f_logits, _ = model(*_structural_inputs(structural_metadata, 8))
    loss = (F.kl_div(F.log_softmax(c_logits, dim=1), target_ceo, reduction="batchmean") +
            F.kl_div(F.log_softmax(f_logits, dim=1), target_firm, reduction="batchmean"))
    loss.backward()
    trainable_params = [p for p in model.parameters() if p.requires_grad]
    assert len(trainable_params) > 0, "Should have trainable parameters"
    for param in trainable_params:
        assert param.grad is not None, "Gradient should exist"
        assert torch.isfinite(param.grad).all()


def test_embedding_counts(model, structural_metadata):                # :132-135
    assert len(model.firm_embeddings) == len(structural_metadata["firm_cat_cards"])
    assert len(model.ceo_embeddings) == len(structural_metadata["ceo_cat_cards"])


def test_get_type_probabilities(model, structural_metadata):          # :137-156
    model.eval()
    with torch.no_grad():
        ceo_probs, firm_probs = model.get_type_probabilities(*_structural_inputs(structural_metadata, 16))
    assert ceo_probs.shape == (16, 5) and firm_probs.shape == (16, 5)
    ones = torch.ones(16, device=DEV)
    assert torch.allclose(ceo_probs.sum(dim=1), ones, atol=1e-5)
    assert torch.allclose(firm_probs.sum(dim=1), ones, atol=1e-5)
