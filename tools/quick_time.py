"""Ad-hoc device timing of the head at BASELINE config 2 (not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth, _lib

B, H, W = 32, 128, 128
NSETS = int(os.environ.get("NSETS", "6"))
dev = torch.device("cuda", 0)
base = synth.make_batch(2, B, H, W, "C")
sets = []
for s in range(NSETS):
    d = {}
    for k in ("pix_logits", "link_logits", "pix_lab", "link_lab"):
        a = base[k]
        a = np.roll(a, s, axis=0)  # different image order per set: distinct memory, same statistics
        d[k] = torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    d["out"] = {}
    sets.append(d)
dcfg = head.DecodeConfig(max_boxes=128)
lcfg = head.LossConfig()

def run(mode, d):
    if mode == "loss":
        head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], lcfg, True, False, None, d["out"])
    elif mode == "decode":
        head.decode_raw(d["pix_logits"], d["link_logits"], dcfg, d["out"], want_rects=False)
    elif mode == "fused":
        head.loss_and_decode_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], lcfg, dcfg, d["out"])

for mode in os.environ.get("MODES", "loss,decode,fused").split(","):
    for i in range(10):
        run(mode, sets[i % NSETS])
    torch.cuda.synchronize()
    n = 60
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        run(mode, sets[i % NSETS])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)]) * 1e3
    px = B * H * W
    byts = {"loss": 180, "decode": 76, "fused": 184}[mode] * px
    print("%-6s median %.1f us  p10 %.1f  p90 %.1f  -> %.0f img/s  %.0f GB/s algorithmic" % (
        mode, np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90), B / (np.median(ts) * 1e-6), byts / (np.median(ts) * 1e-6) / 1e9))
print("launches", head.launch_count())
