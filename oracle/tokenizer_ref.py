"""TEST INFRASTRUCTURE ONLY -- literal Python restatement of qwen3-inference/src/tokenizer.rs and of
render_prompt (generation.rs:188-195), written to check the C++ host mirror (include/qwen3_transformer.hpp).
Deliberately keeps the reference's linear scans; parity unpinned by the reference (it has no tokenizer tests):
pinned here by hand-computed known answers in tests/test_cpp_host.py."""
import struct
from typing import List, Optional


class Tokenizer:
    def __init__(self, checkpoint_path: str, vocab_size: int, enable_thinking: bool):  # tokenizer.rs:40-100
        with open(checkpoint_path + ".tokenizer", "rb") as f:
            data = f.read()
        self.max_token_length, self.bos_token_id, self.eos_token_id = struct.unpack_from("<III", data, 0)
        off = 12
        self.vocab: List[bytes] = []
        self.merge_scores: List[float] = []
        for _ in range(vocab_size):
            if off + 4 > len(data):  # score unreadable: empty token, zero score
                self.vocab.append(b"")
                self.merge_scores.append(0.0)
                off = len(data)
                continue
            self.merge_scores.append(struct.unpack_from("<f", data, off)[0])
            off += 4
            if off + 4 > len(data):
                self.vocab.append(b"")
                off = len(data)
                continue
            (n,) = struct.unpack_from("<I", data, off)
            off += 4
            if off + n > len(data):
                self.vocab.append(b"")
                off = len(data)
                continue
            self.vocab.append(data[off:off + n])
            off += n
        self.vocab_size = vocab_size
        self.prompt_template = self._load_template(checkpoint_path, False, enable_thinking)
        self.system_prompt_template = self._load_template(checkpoint_path, True, enable_thinking)

    @staticmethod
    def _load_template(path: str, with_system: bool, enable_thinking: bool) -> str:  # :103-119
        suffix = {(True, True): ".template.with-system-and-thinking", (True, False): ".template.with-system",
                  (False, True): ".template.with-thinking", (False, False): ".template"}[(with_system, enable_thinking)]
        try:
            with open(path + suffix, "r", encoding="utf-8", newline="") as f:
                return f.read()
        except OSError:
            return ""

    def decode(self, token: int) -> bytes:  # :122-140
        return self.vocab[token] if token < len(self.vocab) else b""

    def _lookup(self, b: bytes) -> Optional[int]:  # :143-151: first position
        for i, t in enumerate(self.vocab):
            if t == b:
                return i
        return None

    def encode(self, text: str) -> List[int]:  # :166-238
        tokens: List[int] = []
        chars = list(text)
        i = 0
        while i < len(chars):
            found_special = False
            if chars[i] == "<":
                end = None
                for j in range(i + 1, min(len(chars), i + self.max_token_length)):
                    if chars[j] == ">":
                        end = j
                        break
                if end is not None:
                    tid = self._lookup("".join(chars[i:end + 1]).encode("utf-8"))
                    if tid is not None:
                        tokens.append(tid)
                        i = end + 1
                        found_special = True
            if not found_special:
                tid = self._lookup(chars[i].encode("utf-8"))
                if tid is not None:
                    tokens.append(tid)
                i += 1
        while True:
            best_score, best_id, best_idx = -1e10, None, None
            for k in range(max(len(tokens) - 1, 0)):
                tid = self._lookup(self.vocab[tokens[k]] + self.vocab[tokens[k + 1]])
                if tid is not None and self.merge_scores[tid] > best_score:
                    best_score, best_id, best_idx = self.merge_scores[tid], tid, k
            if best_id is None:
                break
            tokens[best_idx] = best_id
            del tokens[best_idx + 1]
        return tokens


def render_prompt(pos: int, system_prompt: Optional[str], user_prompt: str, tok: Tokenizer) -> str:  # generation.rs:188-195
    if pos == 0 and system_prompt is not None:
        return tok.system_prompt_template.replace("%s", f"{system_prompt}\n{user_prompt}")
    return tok.prompt_template.replace("%s", user_prompt)


def write_tokenizer_file(path: str, vocab: List[bytes], scores: List[float], max_token_length: int, bos: int, eos: int):
    """The byte layout Tokenizer::new reads (tokenizer.rs:47-87): three u32, then (f32 score, u32 len, bytes) per token."""
    with open(path + ".tokenizer", "wb") as f:
        f.write(struct.pack("<III", max_token_length, bos, eos))
        for b, s in zip(vocab, scores):
            f.write(struct.pack("<fI", s, len(b)) + b)
