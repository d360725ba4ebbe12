"""Ad-hoc: device time of the logit producer (plh_head_fuse_level) at the PixelLink-4s shapes of config 2
(batch 32 at 512x512: conv3_3 128x128x256, conv4_3 64x64x512, conv5_3 32x32x512, fc7 32x32x1024), per level and
for the three launches together, against the bytes each level has to read.  usage: python tools/headfuse_bench.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensorflow_ocr_b200 import head
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)
NS = 3   # rotating input sets: 3 x 1 GB > 126 MB L2
sets = [dict(fc7=R(B, 32, 32, 1024), c5=R(B, 32, 32, 512), c4=R(B, 64, 64, 512), c3=R(B, 128, 128, 256)) for _ in range(NS)]
W = {k: R(n, 18) / n ** 0.5 for k, n in (("fc7", 1024), ("c5", 512), ("c4", 512), ("c3", 256))}
bias = R(18)
w_out, b_out = R(18, 18) / 18 ** 0.5, R(18)
def level1(d): return head.head_fuse_level_raw([(d["fc7"], W["fc7"], None, bias, False), (d["c5"], W["c5"], None, bias, False)])
# PixelLink-4s (no activation on the fuse convolutions): the output matrix is applied to level 2 and folded into level 3's weights
def level2(d, s1): return head.head_fuse_level_raw([(d["c4"], W["c4"], None, bias, False)], prev=s1, w_out=w_out, logits=False)
def level3(d, s2): return head.head_fuse_level_raw([(d["c3"], W["c3"], None, bias, False)], prev=s2, logits=True)
def level3_unfolded(d, s2): return head.head_fuse_level_raw([(d["c3"], W["c3"], None, bias, False)], prev=s2, w_out=w_out, b_out=b_out)
s1 = level1(sets[0]); s2 = level2(sets[0], s1); level3(sets[0], s2)
torch.cuda.synchronize()
def timeit(fn, reps=20):
    for i in range(3): fn(sets[i % NS])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(sets[i % NS])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
peak = 6549.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
res = {}
for name, fn, nbytes in (
        ("level1 fc7+conv5_3 (32x32)", lambda d: level1(d), B * 32 * 32 * (1024 + 512 + 18) * 4),
        ("level2 conv4_3 (64x64)", lambda d: level2(d, s1), B * (64 * 64 * (512 + 18) + 32 * 32 * 18) * 4),
        ("level3 conv3_3 (128x128) + output", lambda d: level3(d, s2), B * (128 * 128 * (256 + 18) + 64 * 64 * 18) * 4),
        ("level3 with the 18x18 output matrix (EAST fork)", lambda d: level3_unfolded(d, s2), B * (128 * 128 * (256 + 18) + 64 * 64 * 18) * 4),
        ("all three", lambda d: level3(d, level2(d, level1(d))), B * (32 * 32 * 1536 + 64 * 64 * 512 + 128 * 128 * 256 + 128 * 128 * 18 + 2 * (32 * 32 + 64 * 64) * 18) * 4)):
    us = timeit(fn)
    res[name] = dict(us=round(us, 1), MB=round(nbytes / 1e6, 1), GBs=round(nbytes / us / 1e3, 1), frac=round(nbytes / us / 1e3 / peak, 3))
    print("%-50s %8.1f us  %8.1f MB  %7.1f GB/s  %.3f of %.0f GB/s" % (name, us, nbytes / 1e6, nbytes / us / 1e3, nbytes / us / 1e3 / peak, peak))
print(json.dumps({"batch": B, "img_per_s_all_three": round(B / res["all three"]["us"] * 1e6), **res}))
