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
buf = np.zeros(8, np.int64)
print("rc", lib.plh_debug_read(buf.ctypes.data_as(ctypes.c_void_p), 8))
d = np.diff(buf)
names = ["load", "phase1 runs", "phase2a", "phase2b hier", "phase3 flatten", "phase4 slots", "phase5 labels"]
for n, c in zip(names, d):
    print("%-16s %8d cycles" % (n, c))
print("n_boxes", out["n_boxes"].cpu().numpy().tolist())
