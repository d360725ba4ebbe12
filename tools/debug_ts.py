import sys, os, ctypes
os.environ["PLH_DEBUG_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth, _lib
B, H, W = 32, 128, 128
dev = torch.device("cuda", 0)
base = synth.make_batch(2, B, H, W, "C")
sets = []
for s in range(6):
    d = {k: torch.as_tensor(np.ascontiguousarray(np.roll(base[k], s, axis=0))).to(dev) for k in ("pix_logits", "link_logits", "pix_lab", "link_lab")}
    d["out"] = {}
    sets.append(d)
lib = _lib.load()
lib.plh_debug_ts.argtypes = [ctypes.c_void_p, ctypes.c_int]
for i in range(12):
    d = sets[i % 6]
    head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], head.LossConfig(), True, False, None, d["out"])
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
d = sets[0]
e0.record()
head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], head.LossConfig(), True, False, None, d["out"])
e1.record()
torch.cuda.synchronize()
n = 296
buf = np.zeros((n, 4), np.uint64)
rc = lib.plh_debug_ts(buf.ctypes.data_as(ctypes.c_void_p), n)
st, mid, en = buf[:, 0].astype(np.int64), buf[:, 1].astype(np.int64), buf[:, 2].astype(np.int64)
t0 = st.min()
print("rc", rc, "whole loss call (events): %.1f us" % (e0.elapsed_time(e1) * 1e3))
print("CTA start: min 0, median %.2f us, max %.2f us" % ((np.median(st) - t0) / 1e3, (st.max() - t0) / 1e3))
print("CTA loop end (before fence): min %.2f median %.2f max %.2f us" % ((mid.min() - t0) / 1e3, (np.median(mid) - t0) / 1e3, (mid.max() - t0) / 1e3))
print("CTA after fence: min %.2f median %.2f max %.2f us" % ((en.min() - t0) / 1e3, (np.median(en) - t0) / 1e3, (en.max() - t0) / 1e3))
print("per-CTA loop duration: min %.2f median %.2f max %.2f us" % ((mid - st).min() / 1e3, np.median(mid - st) / 1e3, (mid - st).max() / 1e3))
