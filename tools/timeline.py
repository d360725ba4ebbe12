"""Device-side timeline of one head step (needs the -DPLH_TIMELINE build: tensorflow_ocr_b200/libplhead_tl.so).
Prints first-CTA start / last-CTA end of every hot-chain kernel relative to the first kernel, in graph mode."""
import ctypes, os, sys
os.environ["PLH_LIB"] = "libplhead_tl.so"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth, _lib
B, H, W = 32, 128, 128
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.plh_timeline_read.argtypes = [ctypes.c_void_p]
try:
    print("max active clusters of the resident component kernel:", lib.plh_debug_image_cluster_occupancy(H, W))
except AttributeError:
    pass
base = synth.make_batch(2, B, H, W, "C")
sets = []
for s in range(6):
    d = {k: torch.as_tensor(np.ascontiguousarray(np.roll(base[k], s, axis=0))).to(dev) for k in ("pix_logits", "link_logits", "pix_lab", "link_lab")}
    d["out"] = {}
    sets.append(d)
ms = torch.cuda.Stream(dev, priority=int(os.environ.get('BENCH_MAIN_PRIO', '0')))
torch.cuda.set_stream(ms)
lcfg, dcfg = head.LossConfig(), head.DecodeConfig(max_boxes=128, form=os.environ.get("TL_FORM", "auto"))
def step(i):
    d = sets[i % 6]
    if os.environ.get("TL_MODE", "fused") == "decode":
        head.decode_raw(d["pix_logits"], d["link_logits"], dcfg, d["out"], want_rects=False)
    elif os.environ.get("TL_MODE", "fused") == "loss":
        head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], lcfg, True, False, None, d["out"])
    else:
        head.loss_and_decode_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], lcfg, dcfg, d["out"])
for i in range(6):
    step(i)
torch.cuda.synchronize()
graphs = []
for i in range(6):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=ms):
        step(i)
    graphs.append(g)
torch.cuda.synchronize()
torch.cuda.set_stream(ms)
for i in range(30):
    graphs[i % 6].replay()
torch.cuda.synchronize()
names = ["K0 keys (general path)", "K1 select", "K2 counts (split/general)", "K3 main", "D0 flags (standalone)", "D1a tile_cc / cluster A", "D1b cross / cluster B", "D2 flatten / cluster C", "(unused)", "D4 labels / cluster D", "D5 rects", "K1 preamble", "K1 loop", "K1 counted", "K1 posted", "(15)", "img planes loaded", "img runs (CL/CH)", "img edges listed", "img unions", "img flattened", "img sizes", "img latest CTA start", "img latest CTA arrival", "img planes copied"]
acc = np.zeros((len(names), 2))
reps = 20
for r in range(reps):
    lib.plh_timeline_reset()
    graphs[r % 6].replay()
    torch.cuda.synchronize()
    buf = np.zeros(64, np.uint64)
    lib.plh_timeline_read(buf.ctypes.data_as(ctypes.c_void_p))
    t = buf.astype(np.int64).reshape(32, 2)[:len(names)]
    used = t[:, 1] > 0
    t0 = t[used & (t[:, 0] > 0) & (t[:, 0] < 2**62), 0].min()
    acc += np.where(used[:, None], (t - t0) / 1e3, 0.0)
acc /= reps
print("%-26s %9s %9s %9s" % ("kernel", "start us", "end us", "dur us"))
for n, (a, b) in zip(names, acc):
    if b > 0:
        print("%-26s %9.2f %9.2f %9.2f" % (n, max(a, 0.0), b, b - max(a, 0.0)))
print("step span: %.2f us" % (acc[:, 1].max()))
