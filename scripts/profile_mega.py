"""Phase-level timeline of the persistent decode kernel from its in-kernel tagged clock64 stamps (prof_mark in q3_mega.cuh)."""
import os, sys
from collections import defaultdict
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qwen3_rs_b200 import transformer as T

model = sys.argv[1] if len(sys.argv) > 1 else "qwen3-8b"
pos = int(sys.argv[2]) if len(sys.argv) > 2 else 64
GHZ = float(sys.argv[3]) if len(sys.argv) > 3 else 1.965
L = {"qwen3-8b": 36, "qwen3-4b": 36, "qwen3-0.6b": 28}.get(model, 0)
if len(sys.argv) > 4:  # offline: analyse a saved capture
    raw = np.load(sys.argv[4])
else:
    path = bench.bench_checkpoint(model, 64)
    m = T.TransformerBuilder.new(path).with_ctx_length(max(256, pos + 8)).build()
    m.bench_decode(1, 0, 4)
    raw = m.debug_profile(1, pos)
    if os.environ.get("Q3_SAVE_PROFILE_NPY"):
        np.save("gpurun_out/mega_profile_%s_pos%d.npy" % (model, pos), raw)

KINDS = ["qkv", "att", "o", "gu", "dn", "head"]
MAIN = ["pro", "gemv", "sync"]  # prologue (poll + rebuild) / GEMV / wait for this CTA's last warp (head: grid barrier)
FINE = {32: "load", 33: "reduce", 34: "quant", 35: "qknorm", 36: "positions", 37: "arrive", 40: "cta_sync", 41: "grid"}
order = []
acc = defaultdict(list)       # label -> list over CTAs of per-CTA mean us
total = []
NS = raw.shape[0] // 3
for c in range(NS):
    n = int(raw[c, -1])
    t = (raw[c, :n] >> np.uint64(8)).astype(np.int64)
    tag = (raw[c, :n] & np.uint64(255)).astype(np.int64)
    total.append((t[-1] - t[0]) / (GHZ * 1e3))
    per = defaultdict(list)
    pend = []
    seen_layers = 0
    tprev = t[0]
    for i in range(1, n):
        d = (t[i] - tprev) / (GHZ * 1e3)
        if tag[i] >= 48:
            continue  # stage-ready marks: timeline view only
        tprev = t[i]
        if tag[i] >= 32:
            pend.append((FINE.get(int(tag[i]), str(tag[i])), d))
            continue
        k, sub = divmod(int(tag[i]) - 1, 3)
        base = "%s_%s" % (KINDS[k], MAIN[sub])
        if k == 0 and sub == 0:
            seen_layers += 1
        skip = k < 5 and seen_layers <= 2  # first layers: cold start
        for name, dd in pend:
            if not skip: per[base + "." + name].append(dd)
            if c == 0 and base + "." + name not in order: order.append(base + "." + name)
        if not skip: per[base + (".rest" if pend else "")].append(d)
        lab = base + (".rest" if pend else "")
        if c == 0 and lab not in order: order.append(lab)
        pend = []
    for k2, v in per.items():
        acc[k2].append(np.mean(v))
print(f"{model} pos {pos}: kernel {np.mean(total):.1f} us (per-CTA mean)")
print("step                    mean_us  [min .. max over CTAs of the per-CTA mean]")
layer_sum = 0.0
for lab in order:
    x = np.array(acc[lab])
    if not lab.startswith("head"): layer_sum += x.mean()
    print(f"{lab:22s} {x.mean():7.2f}   [{x.min():6.2f} .. {x.max():6.2f}]")
print(f"per layer (sum of the means above, head excluded): {layer_sum:.2f} us")

# merged consumer/producer timeline of one CTA for one mid-model layer
def timeline(c, layer=5):
    rows = []
    for r, who in ((c, "C"), (NS + c, "P"), (2 * NS + c, "      D")):
        n = int(raw[r, -1])
        for e in raw[r, :n]:
            rows.append((int(e >> np.uint64(8)), who, int(e & np.uint64(255))))
    rows.sort()
    # find layer boundaries on the consumer row: tag 1 (qkv_pro done) occurrences
    starts = [t for t, who, tag in rows if who == "C" and tag == 3 + 3 * 4]  # dn_sync done = layer end
    t0, t1 = starts[layer - 1], starts[layer]
    print(f"--- CTA {c}, layer {layer}: timeline (us from the end of the previous layer) ---")
    for t, who, tag in rows:
        if t0 <= t <= t1:
            if tag >= 64: name = "issue slot %d" % (tag - 64)
            elif tag == 58: name = "x loaded"
            elif tag == 59: name = "tile done"
            elif tag >= 48: name = "ready slot %d" % (tag - 48)
            elif tag >= 32: name = FINE.get(tag, str(tag))
            else: name = "%s_%s done" % (KINDS[(tag - 1) // 3], MAIN[(tag - 1) % 3])
            print(f"{(t - t0) / (GHZ * 1e3):8.2f}  {who}  {name}")
timeline(0)
timeline(77)
