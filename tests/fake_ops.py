"""TEST INFRASTRUCTURE: oracle-backed, CPU-only stand-ins for the tensor-level entry points of
``rag_arc_b200.ops`` so that the HOST logic of the plugin classes (bookkeeping, relevance-score maps,
threshold filters, MMR plumbing, hybrid fusion mapping, persistence) can be exercised against the
reference's golden vectors on a machine without a GPU.  The kernels themselves are covered by the
``-m gpu`` tests, which run the very same test bodies against the real library."""
import contextlib

import numpy as np
import torch

from oracle import dense as odense
from oracle import pool as opool
from oracle import rrf as orrf
from oracle.bm25 import stable_topk


def _normalize_cast(src, dtype=torch.float32, normalize=True, out=None):
    x = src.detach().cpu().numpy().astype(np.float32).copy()
    if normalize:
        odense.normalize_L2(x)
    res = torch.from_numpy(x).to(dtype)
    if out is not None:
        out.copy_(res)
        return out
    return res


def _dense_topk(corpus, queries, k, n_rows=None, path=0, return_path=False, **kw):
    n = corpus.shape[0] if n_rows is None else int(n_rows)
    D, I = odense.flat_ip_search(corpus[:n].float().numpy(), queries.float().numpy(), k)
    res = (torch.from_numpy(D), torch.from_numpy(I))
    return res + (1,) if return_path else res


def _mmr_select(corpus, queries, cand_rows, k, lambda_mult=0.5, *, n_rows=None):
    """numpy restatement of _mmr_select (VectorStore_Faiss.py:16-62) over the stored rows."""
    X = corpus.double().numpy(); Q = queries.double().numpy(); C = cand_rows.numpy()
    out = np.full((C.shape[0], k), -1, np.int32)
    for qi in range(C.shape[0]):
        cand = [int(c) for c in C[qi] if c >= 0]
        remaining = list(range(len(cand)))
        if not remaining:
            continue
        sel = [remaining.pop(0)]
        while len(sel) < k and remaining:
            best, best_score = None, None
            for idx in remaining:
                qs = float(Q[qi] @ X[cand[idx]])
                ms = 0.0
                for s in sel:
                    ms = max(ms, float(X[cand[s]] @ X[cand[idx]]))
                score = lambda_mult * qs - (1 - lambda_mult) * ms
                if best_score is None or score > best_score:
                    best, best_score = idx, score
            sel.append(best); remaining.remove(best)
        out[qi, :len(sel)] = sel
    return torch.from_numpy(out)


def _bm25_scores(index, q_terms, q_len, use_post_val=True):
    qt, ql = q_terms.numpy(), q_len.numpy()
    S = np.zeros((qt.shape[0], index.n_docs))
    for q in range(qt.shape[0]):
        for t in qt[q, :ql[q]]:
            if t >= 0:
                a, b = index.indptr_np[t], index.indptr_np[t + 1]
                S[q, index.post_doc_np[a:b]] += index.idf_np[t] * index.post_val_np[a:b]
    return torch.from_numpy(S)


def _bm25_topk(index, q_terms, q_len, k, use_post_val=True):
    S = _bm25_scores(index, q_terms, q_len).numpy()
    sc = np.full((S.shape[0], k), -np.inf); ids = np.full((S.shape[0], k), -1, np.int64)
    for q in range(S.shape[0]):
        top = stable_topk(S[q], min(k, index.n_docs))
        sc[q, :len(top)] = S[q][top]; ids[q, :len(top)] = top + getattr(index, "id_base", 0)
    return torch.from_numpy(sc), torch.from_numpy(ids)


def _rrf_fuse(ids, top_k, rrf_k=60.0):
    a = ids.numpy()
    L, nq, _ = a.shape
    out_i = np.full((nq, top_k), -1, np.int32); out_s = np.zeros((nq, top_k)); cnt = np.zeros(nq, np.int32)
    for q in range(nq):
        keys, scores = orrf.rrf_fuse_ids([a[l, q].tolist() for l in range(L)], top_k, rrf_k)
        out_i[q, :len(keys)] = keys; out_s[q, :len(keys)] = scores; cnt[q] = len(keys)
    return torch.from_numpy(out_i), torch.from_numpy(out_s), torch.from_numpy(cnt)


def _rrf_fuse_rows(rows, row_to_key, kl, top_k, rrf_k=60.0):
    """Same packed buffer as ops.rrf_fuse_rows, produced by oracle.rrf.rrf_fuse_rows (integer keys stand in
    for the content strings)."""
    live = [r for r in rows if r is not None]
    nq = int(live[0].shape[0])
    tabs = [t.numpy() for t in row_to_key]
    out_i = np.full((nq, top_k), -1, np.int32); out_s = np.zeros((nq, top_k))
    lists = np.full((nq, top_k), -1, np.int32); rws = np.full((nq, top_k), -1, np.int64); cnt = np.zeros(nq, np.int32)
    for q in range(nq):
        per = [([] if r is None else r[q].tolist()) for r in rows]
        pairs, scores = orrf.rrf_fuse_rows(per, tabs, top_k, rrf_k)
        for j, ((l, row), sc) in enumerate(zip(pairs, scores)):
            out_i[q, j] = tabs[l][row]; out_s[q, j] = sc; lists[q, j] = l; rws[q, j] = row
        cnt[q] = len(pairs)
    buf = np.concatenate([rws.reshape(-1).view(np.uint8), lists.reshape(-1).view(np.uint8), cnt.view(np.uint8)])
    return torch.from_numpy(out_i), torch.from_numpy(out_s), torch.from_numpy(buf.copy())


def _pool_normalize(x, mask, mode="mean", normalize=True):
    return torch.from_numpy(opool.pool_normalize(x.float().numpy(), mask.numpy(), mode, normalize).astype(np.float32))


def _l2_aug_dim(d, dtype):
    return d + 1 if dtype == torch.float32 else (d + 3 + 7) // 8 * 8


def _l2_augment(src, dtype, *, is_query, normalize=False, out=None, sqnorm=None):
    """[x | -|x|^2/2] / [q | 1] with the extra term split into three storage-dtype pieces for half
    types - the layout of ragarc_l2_augment, restated with torch on the host."""
    x = _normalize_cast(src, dtype, normalize)
    n, d = x.shape
    da = _l2_aug_dim(d, dtype)
    res = torch.zeros((n, da), dtype=dtype)
    res[:, :d] = x
    xn = (x.float() * x.float()).sum(1, dtype=torch.float32)
    if is_query:
        res[:, d:d + (1 if dtype == torch.float32 else 3)] = 1
    elif dtype == torch.float32:
        res[:, d] = -0.5 * xn
    else:
        h = -0.5 * xn
        h1 = h.to(dtype); r1 = h - h1.float(); h2 = r1.to(dtype); r2 = r1 - h2.float()
        res[:, d], res[:, d + 1], res[:, d + 2] = h1, h2, r2.to(dtype)
    if sqnorm is not None:
        sqnorm.copy_(xn)
    if out is not None:
        out.copy_(res)
        return out
    return res


def _l2_distances(scores, queries_aug, d):
    qn = (queries_aug[:, :d].float() ** 2).sum(1, dtype=torch.float32)
    dist = torch.clamp(qn[:, None] - 2.0 * scores, min=0.0)
    dist[torch.isinf(scores)] = float("inf")
    scores.copy_(dist)
    return scores


@contextlib.contextmanager
def patched():
    from rag_arc_b200 import ops
    fakes = {"normalize_cast": _normalize_cast, "dense_topk": _dense_topk, "mmr_select": _mmr_select,
             "bm25_scores": _bm25_scores, "bm25_topk": _bm25_topk, "rrf_fuse": _rrf_fuse, "rrf_fuse_rows": _rrf_fuse_rows,
             "pool_normalize": _pool_normalize, "l2_aug_dim": _l2_aug_dim, "l2_augment": _l2_augment,
             "l2_distances": _l2_distances}
    saved = {name: getattr(ops, name) for name in fakes}
    try:
        for name, fn in fakes.items():
            setattr(ops, name, fn)
        yield
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
