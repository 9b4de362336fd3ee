"""Python restatement of the reference's tokenizer and prompt-template exporters (qwen3-export/src/
tokenizer_exporter.rs, chat_template_exporter.rs) for environments without the Rust toolchain: Hugging Face
`tokenizer.json` / `tokenizer_config.json` -> `<out>.tokenizer` + `<out>.template*`, the files
qwen3-inference's Tokenizer (and include/qwen3_transformer.hpp's) reads.  Pinned by the reference's own unit
tests, ported in tests/test_tokenizer_export.py.

Kept on purpose, because the runtime depends on the exact bytes:
  * a token's score is looked up by the TOKEN string in a map keyed by MERGE strings ("a b"), so almost every
    token gets DEFAULT_SCORE (tokenizer_exporter.rs:176) -- the runtime's BPE then merges the leftmost mergeable pair;
  * tokens are written in id order without padding gaps in the id space (:139-160)."""
from __future__ import annotations

import ctypes
import ctypes.util
import json
import math
import os
import struct
from typing import Dict, List, Optional, Tuple

TOKENIZER_FILE_NAME = "tokenizer.json"
DEFAULT_SCORE = -1e6


def _logf(x: float) -> float:
    """f32::ln: the platform libm's logf (what Rust's std calls), not numpy's vectorised log."""
    try:
        libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        libm.logf.restype, libm.logf.argtypes = ctypes.c_float, [ctypes.c_float]
        return float(libm.logf(ctypes.c_float(x)))
    except OSError:  # pragma: no cover
        return struct.unpack("<f", struct.pack("<f", math.log(x)))[0]


def unicode_to_byte_map() -> Dict[str, int]:
    """GPT-2 byte <-> unicode table, inverted (tokenizer_exporter.rs:32-69)."""
    mapping: Dict[str, int] = {}
    for lo, hi in ((33, 126), (161, 172), (174, 255)):
        for b in range(lo, hi + 1):
            mapping[chr(b)] = b
    n = 0
    taken = set(mapping.values())
    for b in range(256):
        if b not in taken:
            mapping[chr(256 + n)] = b
            n += 1
    return mapping


_U2B = unicode_to_byte_map()


def token_to_bytes(token: str) -> bytes:
    """:72-80: mapped characters become their byte, anything else its UTF-8 bytes."""
    out = bytearray()
    for ch in token:
        b = _U2B.get(ch)
        out += bytes([b]) if b is not None else ch.encode("utf-8")
    return bytes(out)


def load_tokenizer_json(model_path: str) -> dict:  # :118-134
    path = os.path.join(model_path, TOKENIZER_FILE_NAME)
    if not os.path.exists(path):
        raise FileNotFoundError(f"tokenizer.json not found in model directory: {model_path}")
    try:
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f)
    except json.JSONDecodeError as e:
        raise ValueError(f"Failed to parse tokenizer.json from {path}") from e


def extract_vocabulary(data: dict) -> Dict[str, int]:  # :190-221
    model_vocab = data.get("model", {}).get("vocab") if isinstance(data.get("model"), dict) else None
    if not isinstance(model_vocab, dict):
        raise ValueError("Could not find vocabulary in tokenizer.json")
    vocab = {tok: int(i) for tok, i in model_vocab.items() if isinstance(i, int) and not isinstance(i, bool) and i >= 0}
    added = data.get("added_tokens")
    if isinstance(added, list):
        for t in added:
            if isinstance(t, dict) and isinstance(t.get("content"), str) and isinstance(t.get("id"), int) and t["id"] >= 0:
                vocab[t["content"]] = int(t["id"])
    return vocab


def extract_merge_ranks(data: dict) -> Dict[str, int]:  # :224-237 (only string merges; ["a","b"] pairs are skipped)
    merges = data.get("model", {}).get("merges") if isinstance(data.get("model"), dict) else None
    if not isinstance(merges, list):
        return {}
    return {m: rank for rank, m in enumerate(merges) if isinstance(m, str)}


def load_token_data(model_path: str) -> Tuple[Dict[str, int], Dict[str, int], int]:  # :103-115
    data = load_tokenizer_json(model_path)
    vocab = extract_vocabulary(data)
    merge_ranks = extract_merge_ranks(data)
    max_token_length = max((len(t.encode("utf-8")) for t in vocab), default=0)  # String::len() = UTF-8 bytes
    return vocab, merge_ranks, max_token_length


def create_ordered_tokens(vocab: Dict[str, int]) -> List[Tuple[int, str]]:  # :137-141
    return sorted(((i, t) for t, i in vocab.items()), key=lambda p: p[0])


def token_score(token: str, merge_ranks: Dict[str, int]) -> float:  # :176
    rank = merge_ranks.get(token)
    return DEFAULT_SCORE if rank is None else -_logf(float(rank + 1))


def export_tokenizer(model_path: str, output_path: str, bos_token_id: int, eos_token_id: int) -> str:  # :88-100, :144-186
    vocab, merge_ranks, max_token_length = load_token_data(model_path)
    out = f"{output_path}.tokenizer"
    with open(out, "wb") as f:
        f.write(struct.pack("<III", max_token_length, bos_token_id, eos_token_id))
        for _, token in create_ordered_tokens(vocab):
            b = token_to_bytes(token)
            f.write(struct.pack("<fI", token_score(token, merge_ranks), len(b)) + b)
    return out


# ---- chat_template_exporter.rs ---------------------------------------------------------------------------
SUFFIXES = {(False, False): ".template", (False, True): ".template.with-thinking",
            (True, False): ".template.with-system", (True, True): ".template.with-system-and-thinking"}


def analyze_template_capabilities(template: str) -> Tuple[str, bool, bool]:
    """-> (template_type, supports_thinking, supports_system)  (chat_template_exporter.rs:71-92)."""
    if "<|im_start|>" in template and "<|im_end|>" in template:
        return "Qwen3", "enable_thinking" in template, ("system" in template and "messages[0].role" in template)
    if "<｜User｜>" in template and "<｜Assistant｜>" in template:
        return "DeepSeek", "think" in template, "system_prompt" in template
    return "Unknown", False, False


def get_template_configs(supports_thinking: bool, supports_system: bool) -> List[Tuple[bool, bool]]:
    """-> [(has_system, enable_thinking)] in the reference's order (:94-141)."""
    configs = [(False, False)]
    if supports_thinking:
        configs.append((False, True))
    if supports_system:
        configs.append((True, False))
        if supports_thinking:
            configs.append((True, True))
    return configs


def render_chat_template(template_type: str, has_system: bool, enable_thinking: bool) -> str:  # :201-265
    if template_type == "Qwen3":
        head = "<|im_start|>system\n%s<|im_end|>\n" if has_system else ""
        tail = "" if enable_thinking else "<think>\n\n</think>\n\n"
        return head + "<|im_start|>user\n%s<|im_end|>\n<|im_start|>assistant\n" + tail
    if template_type == "DeepSeek":
        head = "%s" if has_system else ""
        tail = "" if enable_thinking else "<think>\n</think>"
        return head + "<｜User｜>%s<｜Assistant｜>" + tail
    raise ValueError("Unknown template type, cannot render templates")


def load_chat_template_from_model(model_path: str) -> Optional[str]:  # :143-160
    path = os.path.join(model_path, "tokenizer_config.json")
    if not os.path.exists(path):
        return None
    with open(path, "r", encoding="utf-8") as f:
        cfg = json.load(f)
    t = cfg.get("chat_template")
    return t if isinstance(t, str) else None


def export_templates(model_path: str, output_path: str) -> List[str]:  # :43-69, :162-199
    template = load_chat_template_from_model(model_path)
    if template is None:
        raise ValueError(f"No chat template found in tokenizer_config.json at {model_path}")
    ttype, thinking, system = analyze_template_capabilities(template)
    written = []
    for has_system, enable_thinking in get_template_configs(thinking, system):
        path = f"{output_path}{SUFFIXES[(has_system, enable_thinking)]}"
        with open(path, "w", encoding="utf-8", newline="") as f:
            f.write(render_chat_template(ttype, has_system, enable_thinking))
        written.append(path)
    return written
