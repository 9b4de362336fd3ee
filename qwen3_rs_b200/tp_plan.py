"""Tensor-parallel partition of the Qwen3 checkpoint (host-side mirror of create_impl() in
csrc/q3_engine.cu).  Megatron-style, one exchange per sub-block:

  column-parallel (rows of the [out, in] matrix): Wq by query heads, Wk/Wv by kv heads (GQA groups stay
      intact), W1/W3 by hidden units, lm_head by vocab rows;
  row-parallel (input columns): Wo takes this rank's heads, W2 this rank's hidden slice; each rank
      produces a full-length partial sum and the ranks' partials are added in rank order.

Quantisation groups never straddle a shard (head_dim = 128 and hidden/tp are multiples of the group
size), so every rank computes exactly the int8 activations and per-group int32 dots the single-GPU
run computes for those groups; only the f32 order of the group sum changes.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    size: int
    q_rows: range       # rows of Wq (and entries of q)
    kv_rows: range      # rows of Wk / Wv (and entries of a cache row)
    hidden_rows: range  # rows of W1 / W3, columns of W2
    attn_cols: range    # columns of Wo (= q_rows)
    vocab_rows: range   # rows of lm_head


def shard_plan(cfg, rank: int, size: int) -> ShardPlan:
    """cfg: anything with n_heads, n_kv_heads, head_dim, hidden_dim, vocab_size, group_size."""
    g = lambda k: cfg[k] if isinstance(cfg, dict) else getattr(cfg, k)  # noqa: E731
    n_heads, n_kv, hd, hidden, vocab, gs = (g(k) for k in ("n_heads", "n_kv_heads", "head_dim", "hidden_dim", "vocab_size", "group_size"))
    if not (0 <= rank < size):
        raise ValueError(f"bad tp rank/size {rank}/{size}")
    if n_kv % size or hidden % size or (hidden // size) % gs or vocab % size:
        raise ValueError(f"tp_size {size} does not divide kv heads / hidden groups / vocab")
    ah_l, kv_l, h_l, v_l = n_heads // size * hd, n_kv // size * hd, hidden // size, vocab // size
    return ShardPlan(rank, size, range(rank * ah_l, (rank + 1) * ah_l), range(rank * kv_l, (rank + 1) * kv_l),
                     range(rank * h_l, (rank + 1) * h_l), range(rank * ah_l, (rank + 1) * ah_l),
                     range(rank * v_l, (rank + 1) * v_l))


def prefill_exchange_slices(n4: int, size: int):
    """The batched prefill's row-parallel exchange under TP for size > 2 (csrc/q3_prefill.cuh: k_pf_reduce_scatter /
    k_pf_allgather_resid): rank r sums float4 elements [n4 * r // size, n4 * (r + 1) // size) of all `size` partial blocks in
    rank order, and every rank then takes slice r from rank r.  Returns the slices (they tile [0, n4) without gaps)."""
    return [range(n4 * r // size, n4 * (r + 1) // size) for r in range(size)]
