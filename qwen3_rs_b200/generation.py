"""Host-side generation loop on token ids, mirroring qwen3-inference/src/generation.rs.

The tokenizer / chat templates are out of scope; callers pass token ids.
`generate` keeps the reference's quirk that prompt tokens except the last are never forwarded
(:26-28).  `generate_fast` is the drop-in speed-up: greedy decoding stays on the device
(q3_decode_greedy), only token ids cross PCIe."""
from __future__ import annotations

from typing import List, Optional, Sequence

from .sampler import Sampler


def generate_next_token(transformer, sampler: Sampler, token: int, pos: int) -> int:
    """generation.rs:153-162."""
    logits = transformer.forward(token, pos)  # forward() already returns a copy (logits.to_vec())
    return sampler.sample(logits)


def generate(transformer, sampler: Sampler, prompt_tokens: Sequence[int], max_new: Optional[int] = None,
             bos_token_id: int = -1, eos_token_id: int = -1) -> List[int]:
    """generation.rs:9-48.  Returns the sampled tokens (the terminating bos/eos is not included)."""
    if len(prompt_tokens) == 0:
        raise ValueError("Please provide a prompt")
    seq_len = transformer.get_config().seq_len
    pos, token, out = 0, int(prompt_tokens[0]), []
    while pos < seq_len and (max_new is None or len(out) < max_new):
        if pos < len(prompt_tokens) - 1:
            nxt = int(prompt_tokens[pos + 1])
        else:
            nxt = generate_next_token(transformer, sampler, token, pos)
            if nxt == bos_token_id or nxt == eos_token_id:
                break
            out.append(nxt)
        token = nxt
        pos += 1
    return out


def generate_fast(transformer, prompt_tokens: Sequence[int], max_new: int) -> List[int]:
    """Greedy `generate` with the decode loop resident on the GPU.  Same tokens as
    generate(..., Sampler(temperature=0)) without bos/eos termination."""
    if len(prompt_tokens) == 0:
        raise ValueError("Please provide a prompt")
    seq_len = transformer.get_config().seq_len
    pos0 = len(prompt_tokens) - 1
    n = max(0, min(max_new, seq_len - pos0))
    return transformer.decode_greedy(int(prompt_tokens[-1]), pos0, n)
