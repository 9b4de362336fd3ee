"""Checkpoint exporter: HF-layout Qwen3 directory -> qwen3-rs `.bin` (Q8_0, group-wise int8).

Host-side (offline) tool mirroring `qwen3_export::export_model` for the binary model only
(reference: qwen3-export/src/lib.rs:50-83, model_exporter.rs).  It produces exactly the byte
layout `qwen3-inference` reads:

    off 0   : 13 x i32 LE (magic 0x616a6331, version 1, arch_id, dim, hidden_dim, n_layers,
              n_heads, n_kv_heads, vocab_size, seq_len, head_dim, shared_classifier, group_size),
              zero padded to 256 bytes                           (model_exporter.rs:164-191)
    off 256 : f32 norms: input_layernorm[L], post_attention_layernorm[L], model.norm,
              q_norm[L], k_norm[L] (1.0 x head_dim when absent)  (model_exporter.rs:194-232,
              models/qwen3.rs:16-22)
    then    : each tensor as i8[size] followed by f32[size/gs]: embed, then q,k,v,o,gate,down,up
              projections component-major / layer-minor, then lm_head unless shared
              (model_exporter.rs:235-316, models/qwen3.rs:25-33)

Not on the forward path; the tokenizer / chat-template exporters are out of scope.
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional

import numpy as np

MAGIC_NUMBER = 0x616A6331  # model_exporter.rs:34
VERSION = 1  # :35
HEADER_SIZE = 256  # :36
MIN_GROUP_SIZE = 4  # :37
ARCH_QWEN3 = 1  # models/mod.rs:11

NORM_WEIGHT_LAYERS = (  # qwen3-export/src/models/qwen3.rs:16-22 (name, layered, required)
    ("model.layers.{}.input_layernorm.weight", True, True),
    ("model.layers.{}.post_attention_layernorm.weight", True, True),
    ("model.norm.weight", False, True),
    ("model.layers.{}.self_attn.q_norm.weight", True, False),
    ("model.layers.{}.self_attn.k_norm.weight", True, False),
)
LAYER_COMPONENTS = (  # qwen3-export/src/models/qwen3.rs:25-33
    "self_attn.q_proj",
    "self_attn.k_proj",
    "self_attn.v_proj",
    "self_attn.o_proj",
    "mlp.gate_proj",
    "mlp.down_proj",
    "mlp.up_proj",
)
EMBED_TOKENS_KEY = "model.embed_tokens.weight"
LM_HEAD_KEY = "lm_head.weight"


@dataclass
class ExportConfig:
    """Mirror of qwen3-export's ModelConfig (config_loader.rs:122-190)."""

    dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    vocab_size: int
    max_seq_len: int
    head_dim: int
    norm_eps: float = 1e-6
    bos_token_id: int = 0
    eos_token_id: int = 0

    @staticmethod
    def from_hf_json(path: str) -> "ExportConfig":
        with open(path) as f:
            c = json.load(f)
        archs = c.get("architectures")
        if not archs:
            raise ValueError("Cannot determine architecture")
        if len(archs) != 1:
            raise ValueError(f"Multiple architectures are not supported: {archs}")
        if archs[0] != "Qwen3ForCausalLM":
            raise ValueError(f"Unknown ArchitectureId: {archs[0]}")
        head_dim = c.get("head_dim") or c["hidden_size"] // c["num_attention_heads"]
        return ExportConfig(
            dim=c["hidden_size"],
            hidden_dim=c["intermediate_size"],
            n_layers=c["num_hidden_layers"],
            n_heads=c["num_attention_heads"],
            n_kv_heads=c["num_key_value_heads"],
            vocab_size=c["vocab_size"],
            max_seq_len=c["max_position_embeddings"],
            head_dim=head_dim,
            norm_eps=c.get("rms_norm_eps", 1e-6),
            bos_token_id=c.get("bos_token_id") or 0,
            eos_token_id=c.get("eos_token_id") or 0,
        )


def find_optimal_group_size(hidden_dim: int, requested: int) -> int:
    """model_exporter.rs:48-57 (called with `dim`, :40)."""
    size = min(requested, hidden_dim)
    while size >= MIN_GROUP_SIZE and hidden_dim % size != 0:
        size //= 2
    return max(size, MIN_GROUP_SIZE)


def quantize_q80(weights: np.ndarray, group_size: int):
    """model_exporter.rs:104-161, vectorised.  Returns (int8[n], f32[n/gs], max_error).

    scale = max|w|/127 (1.0 for an all-zero group); q = clamp(rint(w/scale), -127, 127).
    np.rint is round-half-to-even == round_half_to_even() (:321-338) for finite inputs.
    """
    w = np.ascontiguousarray(weights, dtype=np.float32).reshape(-1)
    if w.size % group_size != 0:
        raise ValueError("Weight length is not a multiple of group_size")
    g = w.reshape(-1, group_size)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        gmax = np.fmax.reduce(np.abs(g), axis=1, initial=np.float32(0.0))  # f32::max drops NaN
        scale = np.where(gmax > 0, gmax / np.float32(127.0), np.float32(1.0)).astype(np.float32)
        r = np.rint(g / scale[:, None])
        r = np.clip(r, -127.0, 127.0)
        r = np.where(np.isnan(r), np.float32(0.0), r)  # NaN as i8 -> 0
        q = r.astype(np.int8)
        err = np.abs(q.astype(np.float32) * scale[:, None] - g)
        max_error = float(np.fmax.reduce(err.reshape(-1), initial=np.float32(0.0))) if err.size else 0.0
    return q.reshape(-1), scale, max_error


def write_header(f, cfg: ExportConfig, shared_classifier: bool, group_size: int) -> None:
    """model_exporter.rs:164-191."""
    vals = (
        MAGIC_NUMBER, VERSION, ARCH_QWEN3, cfg.dim, cfg.hidden_dim, cfg.n_layers, cfg.n_heads,
        cfg.n_kv_heads, cfg.vocab_size, cfg.max_seq_len, cfg.head_dim, int(shared_classifier), group_size,
    )
    hdr = struct.pack("<13I", *vals)
    f.write(hdr + b"\0" * (HEADER_SIZE - len(hdr)))


TensorLoader = Callable[[str], Optional[np.ndarray]]
Quantizer = Callable[[np.ndarray, int], tuple]


def detect_shared_classifier(load: TensorLoader) -> bool:
    """qwen3-export/src/models/qwen3.rs:59-74: shared iff lm_head absent, or equal within 1e-6."""
    emb = load(EMBED_TOKENS_KEY)
    head = load(LM_HEAD_KEY)
    if head is not None and emb is not None:
        return head.size == emb.size and bool(
            np.all(np.abs(head.reshape(-1).astype(np.float32) - emb.reshape(-1).astype(np.float32)) < 1e-6)
        )
    if head is None and emb is not None:
        return True
    return False


def export_from_loader(
    load: TensorLoader,
    cfg: ExportConfig,
    output_path: str,
    group_size: int,
    shared_classifier: Optional[bool] = None,
    quantizer: Quantizer = quantize_q80,
) -> Dict[str, float]:
    """BinaryModelExporter::export_binary_model (model_exporter.rs:65-101) over an abstract
    tensor source (`load(name) -> f32 ndarray | None`).  Streams one tensor at a time."""
    gs = find_optimal_group_size(cfg.dim, group_size)
    if shared_classifier is None:
        shared_classifier = detect_shared_classifier(load)
    max_err = 0.0
    with open(output_path, "wb") as f:
        write_header(f, cfg, shared_classifier, gs)
        for pattern, layered, required in NORM_WEIGHT_LAYERS:
            names = [pattern.format(i) for i in range(cfg.n_layers)] if layered else [pattern]
            for name in names:
                t = load(name)
                if t is not None:
                    f.write(np.ascontiguousarray(t, dtype="<f4").tobytes())
                elif not required:
                    f.write(np.ones(cfg.head_dim, dtype="<f4").tobytes())  # :209-213
                else:
                    raise KeyError(f"Missing weight for tensor_name: '{name}'")
        names = [EMBED_TOKENS_KEY]
        for comp in LAYER_COMPONENTS:
            names += [f"model.layers.{i}.{comp}.weight" for i in range(cfg.n_layers)]
        if not shared_classifier:
            names.append(LM_HEAD_KEY)
        for name in names:
            t = load(name)
            if t is None:
                raise KeyError(f"Missing weight tensor: {name}")
            q, s, e = quantizer(t, gs)
            f.write(np.ascontiguousarray(q, dtype=np.int8).tobytes())
            f.write(np.ascontiguousarray(s, dtype="<f4").tobytes())
            max_err = max(max_err, e)
    return {"group_size": gs, "shared_classifier": bool(shared_classifier), "max_error": max_err}


def _safetensors_loader(model_path: str) -> TensorLoader:
    """tensor_reader.rs: every *.safetensors in the directory; F32 and BF16 -> f32 (:86-150)."""
    import torch
    from safetensors import safe_open

    files = sorted(p for p in os.listdir(model_path) if p.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"No safetensors files found in {model_path}")
    handles = [safe_open(os.path.join(model_path, p), framework="pt") for p in files]
    index = {}
    for h in handles:
        for k in h.keys():
            index[k] = h

    def load(name: str):
        h = index.get(name)
        if h is None:
            return None
        t = h.get_tensor(name)
        if t.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"Unsupported tensor dtype {t.dtype} for {name}")
        return t.to(torch.float32).numpy()

    return load


def export_model(model_path: str, output_path: str, group_size: int, quantizer: Quantizer = quantize_q80):
    """qwen3_export::export_model (lib.rs:50-83), binary model part only."""
    cfg = ExportConfig.from_hf_json(os.path.join(model_path, "config.json"))
    return export_from_loader(_safetensors_loader(model_path), cfg, output_path, group_size, quantizer=quantizer)
