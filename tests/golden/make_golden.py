"""Regenerates the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

Everything here comes from the CPU oracle (oracle/q3_oracle.c) on seeded synthetic checkpoints --
the reference itself cannot run in this image (no Rust toolchain) and ships no golden vectors for
its forward path, so these pin the oracle against its own past outputs and give the GPU tests a
fixed target that does not depend on re-running the oracle.

Outputs:
  micro_gs32.bin            the exported checkpoint of the "micro" shape (seed 7, group size 32)
  golden.npz                per-shape logits / greedy tokens / margins / quantize+matmul vectors
  golden_meta.json          sha256 of each regenerated checkpoint, shapes, seeds
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as orc  # noqa: E402
from qwen3_rs_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (shape, group size, seed, prompt, steps)
    ("micro", 32, 7, [3, 17, 5], 24),
    ("tiny", 64, 0, [1], 32),
    ("tiny-untied", 64, 1, [9, 2], 32),
    ("small", 128, 2, [11], 24),
]


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    arrays, meta = {}, {}
    for name, gs, seed, prompt, steps in CASES:
        path = os.path.join(OUT, "micro_gs32.bin") if name == "micro" else f"/tmp/golden_{name}.bin"
        synth.export_synthetic(synth.SHAPES[name], path, gs, seed=seed)
        m = orc.Model(path)
        # pick the last prompt token (among 40 candidates) whose greedy run has the widest minimum
        # top1-top2 margin, so that "identical greedy tokens" is a meaningful, non-fragile check
        best = None
        for cand in range(1, 41):
            m.reset()
            t, mg = m.generate(prompt[:-1] + [cand], steps, with_margins=True)
            if best is None or mg.min() > best[2].min():
                best = (cand, t, mg)
        prompt = prompt[:-1] + [best[0]]
        toks, margins = best[1], best[2]
        # teacher-forced logits along the oracle's own greedy path (fresh cache, every token forwarded)
        m.reset()
        seq = prompt + toks
        logits = np.stack([m.forward(seq[p], p) for p in range(min(len(seq), 12))])
        key = f"{name}_gs{gs}"
        arrays[key + "_greedy"] = np.array(toks, np.int32)
        arrays[key + "_margins"] = margins
        arrays[key + "_prompt"] = np.array(prompt, np.int32)
        arrays[key + "_logits"] = logits.astype(np.float32)
        meta[key] = {"shape": name, "group_size": gs, "seed": seed, "sha256": sha(path),
                     "bytes": os.path.getsize(path), "min_margin": float(margins.min())}
        print(key, "greedy", toks[:12], "min margin", margins.min())
    # operator vectors
    rng = np.random.default_rng(123)
    x = (rng.standard_normal(512) * 3).astype(np.float32)
    x[64:128] = 0.0  # all-zero group -> scale 0, q 0 (tensor.rs:104-116)
    x[130] = 63.5 * (np.abs(x[128:192]).max() / 127)  # lands near a .5 boundary
    for gs in (32, 64, 128):
        q, s = orc.quantize(x, gs)
        arrays[f"quant_q_gs{gs}"], arrays[f"quant_s_gs{gs}"] = q, s
    arrays["quant_x"] = x
    n, d, gs = 256, 48, 64
    wq = rng.integers(-127, 128, size=d * n, dtype=np.int8)
    ws = (rng.random(d * n // gs) * 0.01).astype(np.float32)
    xq, xs = orc.quantize((rng.standard_normal(n)).astype(np.float32), gs)
    arrays.update(mm_wq=wq, mm_ws=ws, mm_xq=xq, mm_xs=xs, mm_out=orc.matmul(xq, xs, wq, ws, n, d, gs),
                  mm_dots=orc.group_dots(xq, wq, n, d, gs))
    np.savez_compressed(os.path.join(OUT, "golden.npz"), **arrays)
    with open(os.path.join(OUT, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
