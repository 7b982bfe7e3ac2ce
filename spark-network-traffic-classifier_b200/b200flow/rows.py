"""Row-level relational steps either side of the hot path: the seeded Bernoulli
randomSplit and stable row compaction (where / handleInvalid="skip" / one randomSplit part), both as
b200flow kernels over row-major device buffers."""
import numpy as np
import torch

from ._lib import call, ptr


def random_split_ids(n, weights, seed, row_offset=0, device="cuda"):
    """DataFrame.randomSplit (kdd99.py:52; A.9 build rule): uint8 split id per row from a uniform keyed by
    (seed, global row index); split = first k with u < cumulative weight."""
    w = np.asarray(weights, np.float64)
    if (w < 0).any() or w.sum() <= 0:
        raise ValueError("Weights must be positive. Found weights: %s" % list(weights))
    cum = np.ascontiguousarray(np.cumsum(w / w.sum()))
    cum[-1] = 1.0
    sid = torch.empty(max(int(n), 1), dtype=torch.uint8, device=device)
    call("b200flow_random_split", int(seed) & 0xFFFFFFFFFFFFFFFF, int(row_offset), int(n), cum.ctypes.data, len(cum), ptr(sid))
    return sid[:n]


def _compact_enqueue(bufs, flag):
    """enqueue the compaction of every buffer; -> (full-size outputs, device kept-count int64[1])."""
    flag = flag.to(torch.uint8).contiguous()
    n = flag.shape[0]
    dev = flag.device
    nb = (n + 1023) // 1024
    scratch = torch.zeros(nb + 1 + (nb + 1) // 2 + 1, dtype=torch.int64, device=dev)
    kept = torch.zeros(1, dtype=torch.int64, device=dev)
    outs = []
    for b in bufs:
        b = b.contiguous()
        rb = b.element_size() * (int(np.prod(b.shape[1:])) if b.dim() > 1 else 1)
        out = torch.empty_like(b)
        call("b200flow_compact_rows", ptr(b), n, rb, ptr(flag), 1, ptr(out), ptr(scratch), ptr(kept))
        outs.append(out)
    return outs, kept


def compact_many(bufs, flag):
    """keep the rows with flag != 0 in every buffer of `bufs` (each [n] or [n, k], contiguous, 4-byte multiple
    row size); order preserved.  Returns (list of compacted tensors, kept count)."""
    n = flag.shape[0]
    if n == 0 or not bufs:
        k = int(flag.sum().item()) if n else 0
        return [b[:k] for b in bufs], k
    outs, kept = _compact_enqueue(bufs, flag)
    k = int(kept.item())
    return [o[:k] for o in outs], k


def split_many(bufs, split_id, n_splits, fetch=None):
    """randomSplit's row movement: the rows of every split, for every buffer, with ONE host sync for all the kept counts.
    -> [(list of compacted tensors, count)] per split.  `fetch`: extra device tensors read in the SAME device->host copy
    (e.g. pending category counts): -> (splits, [host tensors])."""
    n = split_id.shape[0]
    if n == 0 or not bufs:
        out = [compact_many(bufs, split_id == k) for k in range(n_splits)]
        return out if fetch is None else (out, [t.cpu() for t in fetch])
    parts = [_compact_enqueue(bufs, split_id == k) for k in range(n_splits)]
    flat = [kept for _, kept in parts] + [t.reshape(-1).to(torch.int64) for t in (fetch or [])]
    host = torch.cat(flat).cpu()
    counts = host[:n_splits].tolist()
    out = [([o[:int(c)] for o in outs], int(c)) for (outs, _), c in zip(parts, counts)]
    if fetch is None:
        return out
    extra, o = [], n_splits
    for t in fetch:
        extra.append(host[o:o + t.numel()]); o += t.numel()
    return out, extra
