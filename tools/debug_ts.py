import sys, os, ctypes
os.environ["PLH_DEBUG_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth, _lib
B, H, W = 32, 128, 128
dev = torch.device("cuda", 0)
base = synth.make_batch(2, B, H, W, "C")
pl = torch.as_tensor(base["pix_logits"]).to(dev); ll = torch.as_tensor(base["link_logits"]).to(dev)
lib = _lib.load()
lib.plh_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = {}
for i in range(5):
    head.decode_raw(pl, ll, head.DecodeConfig(max_boxes=128), out, want_rects=False)
torch.cuda.synchronize()
n = 1184
buf = np.zeros((n, 16), np.int64)
print("rc", lib.plh_debug_read(buf.ctypes.data_as(ctypes.c_void_p), n * 16))
done = buf[buf[:, 7] > 0]
tot = done[:, 7] - done[:, 0]
print("ctas with work", len(done), "cycles per item: median %d max %d" % (np.median(tot), tot.max()))
names = ["gather+rank", "start+succ", "follow", "shift", "edge vectors", "calipers", "final"]
for idx in (np.argsort(tot)[len(tot) // 2], np.argmax(tot)):
    r = done[idx]
    print("hull n=%d candidates=%d" % (r[9], r[10]), {nm: int(r[i + 1] - r[i]) for i, nm in enumerate(names)})
print("start spread (cycles, first item of each CTA):", int(done[:, 0].max() - done[:, 0].min()))
