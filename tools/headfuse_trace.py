"""Device-side trace of the tcgen05 logit producer (needs make -C tensorflow_ocr_b200/csrc trace -> libplhead_trace.so):
per 32-channel chunk of CTA 0, when the staging warps began / found their stage free / had stored / had issued the next loads /
had arrived, and when the issuing warp began to wait / found the operands / had issued its MMAs.  Times in us."""
import os, sys, ctypes, numpy as np, torch
os.environ["PLH_LIB"] = "libplhead_trace.so"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_ocr_b200 import head, _lib
dev = torch.device("cuda", 0)
B = 32
x = torch.randn(B, 128, 128, 256, device=dev); w = torch.randn(256, 18, device=dev) / 16; b = torch.randn(18, device=dev)
prev = torch.randn(B, 64, 64, 18, device=dev)
for _ in range(3): head.head_fuse_level_raw([(x, w, None, b, False)], prev=prev, logits=True)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.path.join(ROOT, "tensorflow_ocr_b200", "libplhead_trace.so"))
buf = (ctypes.c_ulonglong * (8 * 512))()
lib.plh_hf_trace_read(buf)
t = np.array(buf, dtype=np.int64).reshape(8, 512)[:, :112]
t0 = t[0, 0]
names = ["w.begin", "w.free", "w.stored", "w.loadissued", "w.arrived", "i.begin", "i.ready", "i.issued"]
d = (t - t0) / 1e3
print("chunk  " + " ".join("%9s" % n for n in names))
for n in list(range(0, 20)) + list(range(96, 112)):
    print("%5d  " % n + " ".join("%9.2f" % d[k, n] for k in range(8)))
seg = lambda a, b2: float(np.mean(t[b2, 16:104] - t[a, 16:104])) / 1e3
print("mean us (chunks 16..103): wait-free %.3f | store %.3f | load-issue %.3f | fence+arrive %.3f | issuer wait %.3f | issue %.3f | per chunk %.3f" % (
    seg(0, 1), seg(1, 2), seg(2, 3), seg(3, 4), seg(5, 6), seg(6, 7), float(t[0, 104] - t[0, 16]) / 88e3))
