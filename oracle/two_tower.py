"""Functional CPU restatement of the reference's two-tower hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  All citations are
relative to ``/root/reference/``.

Parameters travel as a flat ``dict[str, Tensor]`` whose keys are exactly the
reference modules' ``state_dict()`` keys (``firm_embeddings.0.weight``,
``firm_tower.0.weight``, ``firm_tower.1.running_mean``, ``logit_scale`` ...),
so weights move reference <-> oracle <-> product without renaming.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class TowerSpec:
    """Shape of one tower's nn.Sequential.

    ``lin``: indices of the three Linear layers inside the Sequential,
    ``bn``: indices of BatchNorm1d layers (after Linear 1 and, optionally, 2),
    ``drop``: dropout probability after each of the two hidden activations
    (0.0 where the reference has no Dropout there).
    """
    prefix: str
    emb_prefix: str
    lin: Tuple[int, int, int]
    bn: Tuple[Optional[int], Optional[int]]
    drop: Tuple[float, float]


def two_tower_spec(side: str) -> TowerSpec:
    # ceo_firm_matching/model.py:37-47 (firm) and :52-62 (ceo):
    # Linear(0) BN(1) ReLU Dropout(.1) Linear(4) BN(5) ReLU Dropout(.1) Linear(8)
    return TowerSpec(f"{side}_tower", f"{side}_embeddings", (0, 4, 8), (1, 5), (0.1, 0.1))


def structural_spec(side: str, dropout: float = 0.2) -> TowerSpec:
    # ceo_firm_matching/structural_model.py:72-80 (firm) and :87-95 (ceo):
    # Linear(0) BN(1) ReLU Dropout(p) Linear(4) ReLU Linear(6)
    return TowerSpec(f"{side}_tower", f"{side}_embeddings", (0, 4, 6), (1, None), (dropout, 0.0))


# --------------------------------------------------------------------------
# parameter initialisation (same distributions as torch's module defaults)
# --------------------------------------------------------------------------
def _linear_init(out_f: int, in_f: int, gen: torch.Generator) -> Tuple[torch.Tensor, torch.Tensor]:
    # nn.Linear.reset_parameters: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(in)), bias U(+-1/sqrt(in))
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def _bn_init(n: int) -> Dict[str, torch.Tensor]:
    return {
        "weight": torch.ones(n), "bias": torch.zeros(n),
        "running_mean": torch.zeros(n), "running_var": torch.ones(n),
        "num_batches_tracked": torch.zeros((), dtype=torch.long),
    }


def _init_tower(p: Params, spec: TowerSpec, dims: Sequence[int], gen: torch.Generator) -> None:
    for li, (i, o) in zip(spec.lin, zip(dims[:-1], dims[1:])):
        w, b = _linear_init(o, i, gen)
        p[f"{spec.prefix}.{li}.weight"], p[f"{spec.prefix}.{li}.bias"] = w, b
    for bi, n in zip(spec.bn, dims[1:3]):
        if bi is not None:
            for k, v in _bn_init(n).items():
                p[f"{spec.prefix}.{bi}.{k}"] = v


def init_two_tower_params(n_firm_numeric: int, firm_cat_counts: Sequence[int], n_ceo_numeric: int,
                          ceo_cat_counts: Sequence[int], emb_large: int = 48, emb_medium: int = 8,
                          latent: int = 60, seed: int = 0) -> Params:
    """Random parameters with the layout of ``CEOFirmMatcher`` (model.py:19-65)."""
    gen = torch.Generator().manual_seed(seed)
    p: Params = {}
    for i, n in enumerate(firm_cat_counts):
        p[f"firm_embeddings.{i}.weight"] = torch.randn(n, emb_large, generator=gen)
    for i, n in enumerate(ceo_cat_counts):
        p[f"ceo_embeddings.{i}.weight"] = torch.randn(n, emb_medium, generator=gen)
    _init_tower(p, two_tower_spec("firm"), [n_firm_numeric + len(firm_cat_counts) * emb_large, 64, 32, latent], gen)
    _init_tower(p, two_tower_spec("ceo"), [n_ceo_numeric + len(ceo_cat_counts) * emb_medium, 64, 32, latent], gen)
    p["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32)  # model.py:65
    return p


BLM_A = [[-0.5, -0.3, 0.0, 0.1, 0.2], [-0.2, -0.1, 0.1, 0.3, 0.4], [0.0, 0.2, 0.4, 0.6, 0.7],
         [0.1, 0.4, 0.7, 0.9, 1.1], [0.3, 0.6, 0.9, 1.2, 1.5]]  # structural_config.py:41-47


def init_structural_params(n_firm_num: int, firm_cat_cards: Sequence[int], n_ceo_num: int,
                           ceo_cat_cards: Sequence[int], emb: int = 8, latent: int = 128,
                           seed: int = 0) -> Params:
    """Random parameters with the layout of ``StructuralDistillationNet`` (structural_model.py:31-95)."""
    gen = torch.Generator().manual_seed(seed)
    p: Params = {"A": torch.tensor(BLM_A, dtype=torch.float32)}
    for i, n in enumerate(firm_cat_cards):
        p[f"firm_embeddings.{i}.weight"] = torch.randn(n, emb, generator=gen)
    for i, n in enumerate(ceo_cat_cards):
        p[f"ceo_embeddings.{i}.weight"] = torch.randn(n, emb, generator=gen)
    _init_tower(p, structural_spec("firm"), [n_firm_num + len(firm_cat_cards) * emb, latent, 64, 5], gen)
    _init_tower(p, structural_spec("ceo"), [n_ceo_num + len(ceo_cat_cards) * emb, latent, 64, 5], gen)
    return p


# --------------------------------------------------------------------------
# towers
# --------------------------------------------------------------------------
def _n_tables(p: Params, emb_prefix: str) -> int:
    n = 0
    while f"{emb_prefix}.{n}.weight" in p:
        n += 1
    return n


def gather_concat(p: Params, emb_prefix: str, x_num: torch.Tensor, x_cat: torch.Tensor) -> torch.Tensor:
    """``[x_num | E_0[x_cat[:,0]] | E_1[x_cat[:,1]] | ...]`` — model.py:69-70,74-75;
    structural_model.py:120-121,125-126."""
    embs = [F.embedding(x_cat[:, i], p[f"{emb_prefix}.{i}.weight"]) for i in range(_n_tables(p, emb_prefix))]
    return torch.cat([x_num] + embs, dim=1)


def tower_forward(p: Params, spec: TowerSpec, x_num: torch.Tensor, x_cat: torch.Tensor, training: bool,
                  masks: Optional[Sequence[Optional[torch.Tensor]]] = None, update_stats: bool = True
                  ) -> torch.Tensor:
    """One tower: gather+concat then Linear-[BN]-ReLU-[Dropout] x2, Linear (model.py:37-47,71,76).

    ``masks``: optional keep-masks (bool/0-1 float, shape [B, hidden]) for the two
    dropout sites; a kept unit is scaled by 1/(1-p) exactly like ``nn.Dropout``.
    With ``masks=None`` dropout is the identity (== eval mode or p = 0).
    BatchNorm1d semantics (eps 1e-5, momentum 0.1): training normalises by the
    biased batch variance and updates running stats with the unbiased one.
    """
    h = gather_concat(p, spec.emb_prefix, x_num, x_cat)
    for s in range(2):
        li, bi = spec.lin[s], spec.bn[s]
        h = F.linear(h, p[f"{spec.prefix}.{li}.weight"], p[f"{spec.prefix}.{li}.bias"])
        if bi is not None:
            pre = f"{spec.prefix}.{bi}"
            if training and h.shape[0] == 1:
                raise ValueError("Expected more than 1 value per channel when training")
            rm, rv = p[pre + ".running_mean"], p[pre + ".running_var"]
            if training and not update_stats:
                rm, rv = rm.clone(), rv.clone()
            h = F.batch_norm(h, rm, rv, p[pre + ".weight"], p[pre + ".bias"], training, 0.1, 1e-5)
            if training and update_stats:
                p[pre + ".num_batches_tracked"] += 1
        h = F.relu(h)
        if training and masks is not None and masks[s] is not None and spec.drop[s] > 0:
            h = h * masks[s].to(h.dtype) / (1.0 - spec.drop[s])
    li = spec.lin[2]
    return F.linear(h, p[f"{spec.prefix}.{li}.weight"], p[f"{spec.prefix}.{li}.bias"])


def min_abs_relu_input(p: Params, specs: Sequence[TowerSpec], inputs: Sequence[Tuple[torch.Tensor, torch.Tensor]]
                       ) -> float:
    """Smallest |ReLU input| over all hidden units, rows and towers in train mode (float64).  ReLU has no derivative
    at 0: when a pre-activation lies within an implementation's rounding error of 0, the sign of the unit - and with
    it a 1/B share of that column's gradients - is decided by rounding, in torch as much as in any other
    implementation.  Parity tests draw their inputs so that this margin is far above the arithmetic's resolution."""
    worst = float("inf")
    for spec, (x_num, x_cat) in zip(specs, inputs):
        pd = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
        h = gather_concat(pd, spec.emb_prefix, x_num.double(), x_cat)
        for s in range(2):
            li, bi = spec.lin[s], spec.bn[s]
            h = F.linear(h, pd[f"{spec.prefix}.{li}.weight"], pd[f"{spec.prefix}.{li}.bias"])
            if bi is not None:
                pre = f"{spec.prefix}.{bi}"
                h = F.batch_norm(h, None, None, pd[pre + ".weight"], pd[pre + ".bias"], True, 0.1, 1e-5)
            worst = min(worst, float(h.abs().min()))
            h = F.relu(h)
    return worst


def two_tower_forward(p: Params, f_num, f_cat, c_num, c_cat, training: bool = False,
                      masks: Optional[Dict[str, Sequence[Optional[torch.Tensor]]]] = None,
                      update_stats: bool = True) -> torch.Tensor:
    """``CEOFirmMatcher.forward`` — model.py:67-89.  L2-normalise WITHOUT eps (model.py:79-80),
    row-wise dot, times ``exp(logit_scale)`` (model.py:86-87).  Returns ``[B,1]``."""
    masks = masks or {}
    u = tower_forward(p, two_tower_spec("firm"), f_num, f_cat, training, masks.get("firm"), update_stats)
    v = tower_forward(p, two_tower_spec("ceo"), c_num, c_cat, training, masks.get("ceo"), update_stats)
    u = u / u.norm(dim=1, keepdim=True)
    v = v / v.norm(dim=1, keepdim=True)
    return (u * v).sum(dim=1, keepdim=True) * p["logit_scale"].exp()


def weighted_mse(preds: torch.Tensor, target: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """training.py:52 — ``(w * (pred - tgt)**2).mean()``."""
    return (weights * (preds - target) ** 2).mean()


# --------------------------------------------------------------------------
# contrastive (contrastive.py)
# --------------------------------------------------------------------------
def contrastive_forward(p: Params, f_num, f_cat, c_num, c_cat, training: bool = False,
                        masks=None, update_stats: bool = True):
    """``ContrastiveCEOFirmMatcher.forward`` — contrastive.py:74-99 (+ get_embeddings :52-72).
    Keys: base model under ``base_model.*``, projectors ``{firm,ceo}_projector.{0,2}.*``.
    Uses ``F.normalize`` (eps 1e-12) unlike the base model."""
    base = {k[len("base_model."):]: v for k, v in p.items() if k.startswith("base_model.")}
    masks = masks or {}
    u = tower_forward(base, two_tower_spec("firm"), f_num, f_cat, training, masks.get("firm"), update_stats)
    v = tower_forward(base, two_tower_spec("ceo"), c_num, c_cat, training, masks.get("ceo"), update_stats)
    u, v = F.normalize(u, dim=1), F.normalize(v, dim=1)
    score = (u * v).sum(dim=1, keepdim=True) * base["logit_scale"].exp()

    def proj(x, pre):
        h = F.relu(F.linear(x, p[pre + ".0.weight"], p[pre + ".0.bias"]))
        return F.normalize(F.linear(h, p[pre + ".2.weight"], p[pre + ".2.bias"]), dim=1)

    return score, proj(u, "firm_projector"), proj(v, "ceo_projector")


def info_nce(firm_proj: torch.Tensor, ceo_proj: torch.Tensor, temperature: float = 0.07) -> torch.Tensor:
    """``info_nce_loss`` — contrastive.py:102-138.  S = F C^T / T; mean CE of rows and of
    columns against the diagonal, averaged; 0 when B <= 1 (contrastive.py:124-126)."""
    B = firm_proj.shape[0]
    if B <= 1:
        return torch.tensor(0.0)
    sim = firm_proj @ ceo_proj.t() / temperature
    labels = torch.arange(B)
    return (F.cross_entropy(sim, labels) + F.cross_entropy(sim.t(), labels)) / 2


# --------------------------------------------------------------------------
# structural distillation (structural_model.py / structural_training.py)
# --------------------------------------------------------------------------
def structural_forward(p: Params, f_num, f_cat, c_num, c_cat, training: bool = False, masks=None,
                       dropout: float = 0.2, update_stats: bool = True):
    """``StructuralDistillationNet.forward`` — structural_model.py:97-143.
    Returns ``(c_logits, f_logits, expected_match)``: CEO logits FIRST (:143)."""
    masks = masks or {}
    f_logits = tower_forward(p, structural_spec("firm", dropout), f_num, f_cat, training, masks.get("firm"), update_stats)
    c_logits = tower_forward(p, structural_spec("ceo", dropout), c_num, c_cat, training, masks.get("ceo"), update_stats)
    q_firm = F.softmax(f_logits, dim=1)
    pi_ceo = F.softmax(c_logits, dim=1)
    weighted_a = pi_ceo @ p["A"]                                   # structural_model.py:140
    expected = (weighted_a * q_firm).sum(dim=1, keepdim=True)      # structural_model.py:141
    return c_logits, f_logits, expected


def structural_kl_loss(c_logits, f_logits, target_ceo, target_firm) -> torch.Tensor:
    """structural_training.py:48,75-77 — ``KLDivLoss('batchmean')`` on ``log_softmax`` for each side, summed."""
    kl = torch.nn.KLDivLoss(reduction="batchmean")
    return kl(F.log_softmax(c_logits, dim=1), target_ceo) + kl(F.log_softmax(f_logits, dim=1), target_firm)


# --------------------------------------------------------------------------
# all-pairs scoring (analytical_extensions.py:467-483, contrastive.py:296-322)
# --------------------------------------------------------------------------
def allpairs_scores(rows: torch.Tensor, cols: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """analytical_extensions.py:471 — ``torch.mm(rows, cols.t()) * scale`` in the input dtype."""
    return torch.mm(rows, cols.t()) * scale


def allpairs_topk(rows: torch.Tensor, cols: torch.Tensor, k: int, scale: float = 1.0,
                  chunk: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k per row of ``rows @ cols^T * scale`` ordered (score desc, index asc) — the order
    ``np.argsort(-scores, kind='stable')`` gives for the first k entries (analytical_extensions.py:483).
    Scores are formed in float64 from the float32 inputs so the ranking is the *exact* ranking of
    the fp32 operands (free of any summation-order noise); returned scores are float32.
    Returns ``(scores [R,k] f32, idx [R,k] i64)``."""
    R = rows.shape[0]
    k = min(k, cols.shape[0])
    out_s = torch.empty(R, k, dtype=torch.float32)
    out_i = torch.empty(R, k, dtype=torch.int64)
    c64 = cols.double().t().contiguous()
    for r0 in range(0, R, chunk):
        s = (rows[r0:r0 + chunk].double() @ c64) * scale
        order = torch.argsort(-s, dim=1, stable=True)[:, :k]
        out_i[r0:r0 + chunk] = order
        out_s[r0:r0 + chunk] = torch.gather(s, 1, order).float()
    return out_s, out_i


def retrieval_ranks(firm_emb: torch.Tensor, ceo_emb: torch.Tensor) -> np.ndarray:
    """contrastive.py:306-322 — 1-indexed position of the diagonal entry in each row of
    ``firm_emb @ ceo_emb^T`` sorted descending (stable: earlier column wins a tie)."""
    sim = torch.mm(firm_emb, ceo_emb.t())
    n = sim.shape[0]
    order = torch.argsort(-sim, dim=1, stable=True)
    pos = (order == torch.arange(n).unsqueeze(1)).float().argmax(dim=1)
    return (pos + 1).numpy()


def retrieval_metrics(ranks: np.ndarray, top_k: int = 10) -> Dict[str, float]:
    """contrastive.py:326-332."""
    return {
        "recall@1": float(np.mean(ranks <= 1)), "recall@5": float(np.mean(ranks <= 5)),
        "recall@10": float(np.mean(ranks <= top_k)), "MRR": float(np.mean(1.0 / ranks)),
        "median_rank": float(np.median(ranks)),
    }
