"""All-pairs CEO x firm scoring with top-k — the public helper the reference's scattered scoring sites map to
(``analytical_extensions.py:467-483``, ``contrastive.py:296-322``, ``deep_dive.py:277-290`` ...): they all form
``scores = rows @ cols.T * scale`` and then rank every row.  Here the score matrix is never materialised: a
tcgen05/TMEM pass over bf16 copies streams the tiles through a per-row candidate filter, and the surviving
candidates are rescored exactly (fp64 accumulation of the fp32 operands), so the emitted index lists equal the
exact ranking of the fp32 inputs ordered (score desc, index asc).
"""
from typing import Optional, Sequence, Tuple

import torch

from . import _native as N
from . import ops

MAX_K = 128


def _exact_rows(rows: torch.Tensor, cols: torch.Tensor, k: int, scale: float, col_offset: int,
                chunk: int = 65536) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact path for the (rare) rows whose candidate set could not be proven complete: fp64 scores, stable
    ordering.  Plain torch ops on the device; only ever sees flagged rows."""
    r64 = rows.double()
    best_s = torch.full((rows.shape[0], 0), 0.0, dtype=torch.float64, device=rows.device)
    best_i = torch.zeros((rows.shape[0], 0), dtype=torch.int64, device=rows.device)
    for c0 in range(0, cols.shape[0], chunk):
        s = r64 @ cols[c0:c0 + chunk].double().t()
        idx = torch.arange(c0, c0 + s.shape[1], device=rows.device).expand_as(s)
        s, idx = torch.cat([best_s, s], 1), torch.cat([best_i, idx], 1)
        order = torch.argsort(-s, dim=1, stable=True)[:, :k]          # earlier (smaller) index wins a tie
        best_s, best_i = torch.gather(s, 1, order), torch.gather(idx, 1, order)
    return (best_s * scale).float(), best_i + col_offset, best_s * scale


def score_topk(rows: torch.Tensor, cols: torch.Tensor, k: int, scale: float = 1.0, col_offset: int = 0,
               return_flags: bool = False, return_f64: bool = False):
    """Top-``k`` columns of ``rows @ cols.T * scale`` for every row.

    rows [R, D], cols [C, D] float32 on the device (unit-norm tower outputs in the reference's call sites; any
    norm is handled, the filter margin scales with the largest row norms).  Returns ``(scores [R,k] f32,
    indices [R,k] i64)`` ordered by (score desc, index asc); if ``C < k`` the tail is ``(-inf, -1)``.
    ``return_f64`` appends the fp64 scores (needed for an exact merge of per-shard lists).
    Orientation is the caller's choice: pass firms as rows to rank CEOs per firm (analytical_extensions.py:483)
    or CEOs as rows for "top-100 firms per CEO" (BASELINE config 5).
    """
    ops._require_cuda(rows, cols)
    if not 1 <= k <= MAX_K:
        raise ValueError(f"k must be in [1, {MAX_K}]")
    rows, cols = rows.detach().float().contiguous(), cols.detach().float().contiguous()
    R, D = rows.shape
    C = cols.shape[0]
    if cols.shape[1] != D:
        raise ValueError("rows and cols must have the same feature width")
    dev = rows.device
    out_s = torch.empty(R, k, device=dev)
    out_s64 = torch.empty(R, k, dtype=torch.float64, device=dev) if return_f64 else None
    out_i = torch.empty(R, k, dtype=torch.int64, device=dev)
    flags = torch.zeros(R, dtype=torch.int32, device=dev)

    def result():
        out = (out_s, out_i)
        if return_flags:
            out += (flags,)
        if return_f64:
            out += (out_s64,)
        return out

    if R == 0:
        return result()
    if C == 0:
        out_s.fill_(float("-inf")); out_i.fill_(-1)
        if return_f64:
            out_s64.fill_(float("-inf"))
        return result()
    # Filter operands: fp16 copies when the values fit fp16 comfortably (unit-norm tower outputs always do) — 11-bit
    # significands make the filter 4x tighter than bf16 — otherwise bf16 copies (same exponent range as fp32).
    # e = bound on |16-bit-operand score - exact score| (_score_error_bound); the k-th largest filter score and a
    # candidate's filter score are both off by up to e, so 2e separates "certainly in" from "certainly out".
    use_f16, err = _score_error_bound(rows, cols)
    margin = 2.0 * err
    rb, cb = (ops.pack_f16(rows), ops.pack_f16(cols)) if use_f16 else (ops.pack_bf16(rows), ops.pack_bf16(cols))
    chunks = N.lib().cfm_simtile_chunks(R, C)
    rpad = (R + 255) // 256 * 256
    cand = torch.empty(chunks * rpad * N.CFM_TOPK_CAP, 2, dtype=torch.int32, device=dev)    # (score bits, column)
    cand_cnt = torch.empty(chunks * rpad, dtype=torch.int32, device=dev)
    cand_thr = torch.empty(chunks * rpad, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().cfm_allpairs_topk(N.ptr(rows), N.ptr(cols), N.ptr(rb), N.ptr(cb), 1 if use_f16 else 0, R, C, D,
                                          rb.shape[1], k,
                                          float(scale), margin, col_offset, N.ptr(out_s), N.ptr(out_s64), N.ptr(out_i), N.ptr(flags),
                                          N.ptr(cand), N.ptr(cand_cnt), N.ptr(cand_thr),
                                          N.stream_ptr()))
    bad = torch.nonzero(flags, as_tuple=False).flatten()
    if bad.numel():                                   # completeness not provable from the filter: redo exactly
        s, i, s64 = _exact_rows(rows[bad], cols, k, float(scale), col_offset)
        out_s[bad, :s.shape[1]], out_i[bad, :i.shape[1]] = s, i
        if return_f64:
            out_s64[bad, :s.shape[1]] = s64
    return result()


def merge_topk(part_scores: torch.Tensor, part_indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-shard results ``[n_parts, R, k]`` (each ordered (score desc, index asc), indices already global)
    into the global top-k with the same ordering — the cross-GPU merge of the row-sharded scoring path.  Pass the
    float64 scores (``score_topk(..., return_f64=True)``) for a merge that is exact across shards."""
    ops._require_cuda(part_scores, part_indices)
    n_parts, R, k = part_scores.shape
    is64 = part_scores.dtype == torch.float64          # fp64 parts keep the exact cross-shard ordering
    ps = part_scores.contiguous() if is64 else part_scores.float().contiguous()
    pi = part_indices.long().contiguous()
    out_s = torch.empty(R, k, device=ps.device)
    out_i = torch.empty(R, k, dtype=torch.int64, device=ps.device)
    with torch.cuda.device(ps.device):
        N.check(N.lib().cfm_topk_merge(N.ptr(ps), 1 if is64 else 0, N.ptr(pi), n_parts, R, k, N.ptr(out_s), N.ptr(out_i),
                                       N.stream_ptr()))
    return out_s, out_i


def target_ranks(rows: torch.Tensor, cols: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """1-indexed rank of column ``target[i]`` among all columns for row i (fp64 scores of the fp32 operands)."""
    ops._require_cuda(rows, cols, target)
    rows, cols = rows.detach().float().contiguous(), cols.detach().float().contiguous()
    target = target.long().contiguous()
    rank = torch.empty(rows.shape[0], dtype=torch.int64, device=rows.device)
    with torch.cuda.device(rows.device):
        N.check(N.lib().cfm_allpairs_rank(N.ptr(rows), N.ptr(cols), rows.shape[0], cols.shape[0], rows.shape[1],
                                          N.ptr(target), N.ptr(rank), N.stream_ptr()))
    return rank


_RANK_TC_MIN_PAIRS = 1 << 22      # below ~4M pairs the fp64 SIMT kernel is as fast as the filter + rescoring chain
_RANK_BLOCK_ROWS = 65536
_RANK_AMB_CAP = 1 << 25           # listed (row, column) pairs per row block: 256 MB
_RANK_MIN_BLOCK = 2048            # row blocks are split down to this size before the exact kernel takes over


def _score_error_bound(rows: torch.Tensor, cols: torch.Tensor) -> Tuple[bool, float]:
    """(use fp16 operands?, bound on |16-bit-operand score - exact score|): 2u |row||col| by Cauchy-Schwarz over the
    per-element roundings (u = 2^-11 fp16, 2^-9 bf16) plus the fp32 accumulation / fp16-subnormal allowance."""
    norm_bound, amax = torch.stack([rows.norm(dim=1).max() * cols.norm(dim=1).max(),
                                    torch.maximum(rows.abs().max(), cols.abs().max())]).tolist()
    use_f16 = 1e-2 < amax < 1e3
    err = (2.0 ** -10 + 2.0 ** -16) * norm_bound + 4e-6 if use_f16 else (2.0 ** -8 + 2.0 ** -16) * norm_bound
    return use_f16, err


def diagonal_ranks(firm_emb: torch.Tensor, ceo_emb: torch.Tensor, method: str = "auto") -> torch.Tensor:
    """Rank of the true CEO (the diagonal) for every firm — contrastive.py:306-322 without the full sort.

    ``method="tensor"`` (``"auto"`` from ~4M pairs on): the tcgen05 similarity kernel counts, in its epilogue, the
    columns whose 16-bit-operand score lies certainly above the row's exact positive score and lists the few pairs
    too close to call; those are compared exactly (fp64 of the fp32 operands).  ``method="exact"``: the fp64 SIMT
    kernel over every pair.  Both give the same ranks (ties by column index)."""
    ops._require_cuda(firm_emb, ceo_emb)
    n, C = firm_emb.shape[0], ceo_emb.shape[0]
    if method not in ("auto", "tensor", "exact"):
        raise ValueError("method must be 'auto', 'tensor' or 'exact'")
    D = firm_emb.shape[1]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=firm_emb.device)
    if method == "exact" or D > 128 or (method == "auto" and n * C < _RANK_TC_MIN_PAIRS):
        return target_ranks(firm_emb, ceo_emb, torch.arange(n, device=firm_emb.device))
    rows, cols = firm_emb.detach().float().contiguous(), ceo_emb.detach().float().contiguous()
    dev = rows.device
    use_f16, err = _score_error_bound(rows, cols)
    pack = ops.pack_f16 if use_f16 else ops.pack_bf16
    cb = pack(cols)
    out = torch.empty(n, dtype=torch.int64, device=dev)
    nb = min(n, _RANK_BLOCK_ROWS)
    amb = torch.empty(_RANK_AMB_CAP, 2, dtype=torch.int32, device=dev)
    amb_n = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = {}

    def scratch_for(rows_n):
        if rows_n not in scratch:
            rpad = (rows_n + 255) // 256 * 256
            scratch[rows_n] = (torch.empty(N.lib().cfm_simtile_chunks(rows_n, C) * rpad, dtype=torch.int32, device=dev),
                               torch.empty(rows_n, dtype=torch.int32, device=dev),
                               torch.empty(rows_n, dtype=torch.float64, device=dev),
                               torch.empty(2 * rows_n, device=dev))
        return scratch[rows_n]

    # Row blocks whose near-tie list overflows (untrained embeddings: the positive sits inside the bulk of the scores,
    # ~0.5 % of all pairs fall into the window) are split in four and run again; below _RANK_MIN_BLOCK rows the exact
    # kernel takes over.  One host read per round of blocks.
    pending = [(r0, min(nb, n - r0)) for r0 in range(0, n, nb)]
    while pending:
        statuses = []
        with torch.cuda.device(dev):
            for r0, rows_n in pending:
                blk = rows[r0:r0 + rows_n]
                part, extra, diag64, window = scratch_for(rows_n)
                status = torch.zeros(2, dtype=torch.int32, device=dev)
                rbk = pack(blk)
                N.check(N.lib().cfm_allpairs_diag_rank(N.ptr(blk), N.ptr(cols), N.ptr(rbk), N.ptr(cb), 1 if use_f16 else 0,
                                                       rows_n, C, D, cb.shape[1], r0, err, N.ptr(out[r0:r0 + rows_n]),
                                                       N.ptr(status), N.ptr(part), N.ptr(extra), N.ptr(diag64), N.ptr(window),
                                                       N.ptr(amb), _RANK_AMB_CAP, N.ptr(amb_n), N.stream_ptr()))
                statuses.append(status)
        over = torch.stack(statuses)[:, 0].tolist()
        nxt = []
        for (r0, rows_n), flag in zip(pending, over):
            if not flag:
                continue
            if rows_n <= _RANK_MIN_BLOCK:           # still too many near-ties to list (degenerate inputs): exact kernel
                blk = rows[r0:r0 + rows_n]
                out[r0:r0 + rows_n] = target_ranks(blk, cols, torch.arange(r0, r0 + rows_n, device=dev))
            else:
                q = (rows_n + 3) // 4
                nxt += [(a, min(q, r0 + rows_n - a)) for a in range(r0, r0 + rows_n, q)]
        pending = nxt
    return out
