"""Host-side plumbing of the docID-sharded (multi-GPU) path — SURVEY.md 8e.

One rank = one GPU = one IndexSource holding the postings of a contiguous docID range (Trinity's IndexSourcesCollection model,
index_source.h:191-238; exec_query_par + app-side merge, exec.h:56-177).  Every operator is per-docID, so evaluation needs no
data-path collective.  The only exchange is ONE all-gather of per-shard top-k lists (nq*k*(4+4) bytes per rank) followed by a merge
(`trn_merge_topk` on the device); DocumentsOnly results simply concatenate in shard order.

torch.distributed is plumbing here (process group + all_gather); no compute happens in this module.
"""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import numpy as np


def shard_range(ndocs: int, rank: int, world: int) -> Tuple[int, int]:
    """inclusive docID range [lo, hi] of `rank`; ranges are contiguous, disjoint and cover 1..ndocs"""
    if not (0 <= rank < world) or ndocs < world:
        raise ValueError("bad shard request")
    return rank * ndocs // world + 1, (rank + 1) * ndocs // world


def global_document_frequencies(dist, local_df: np.ndarray) -> np.ndarray:
    """df summed over shards == IndexSourcesCollectionBM25Scorer's per-term df over all sources (similarity.h:209-217).
    Every rank must use these (not its local df) for the BM25 weights so that all shards score identically."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(local_df, dtype=np.int64).copy())
    dist.all_reduce(t)  # SUM
    return t.numpy()


def gather_and_merge_topk(dist, docids, scores, k: int, merge: Callable):
    """The one exchange step.  docids/scores: this rank's [nq, k] tensors (padding: score < 0).
    `merge(gathered_docids[world, nq, k], gathered_scores[world, nq, k]) -> (docids[nq, k], scores[nq, k])` —
    on GPUs this is GpuIndexSource.merge_topk (k_topk_merge); CPU tests pass a numpy merge."""
    import torch

    world = dist.get_world_size()
    gd = torch.empty((world,) + tuple(docids.shape), dtype=docids.dtype, device=docids.device)
    gs = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
    dist.all_gather_into_tensor(gd.view(-1), docids.contiguous().view(-1))
    dist.all_gather_into_tensor(gs.view(-1), scores.contiguous().view(-1))
    return merge(gd, gs)


def concat_docs_only(dist, local_docids: np.ndarray) -> np.ndarray:
    """DocumentsOnly: shard order == docID order, so the global result is the concatenation over ranks (variable lengths)."""
    import torch

    world = dist.get_world_size()
    n = torch.tensor([len(local_docids)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    buf = torch.zeros(mx, dtype=torch.int64)
    buf[: len(local_docids)] = torch.from_numpy(local_docids.astype(np.int64))
    parts = [torch.zeros(mx, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([p[: int(s)].numpy() for p, s in zip(parts, sizes)]).astype(np.uint32)


def device_view(ptr: int, n: int, dtype):
    """a torch tensor aliasing `n` 4-byte elements of device memory at `ptr` (no copy): how the engine's own device result buffers
    (trn_last_topk_device) enter torch.distributed collectives without a host round trip"""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4" if dtype == torch.int32 else "<f4", "data": (int(ptr), False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(h, device="cuda")
