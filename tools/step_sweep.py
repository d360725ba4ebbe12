"""Ad-hoc: device time of the pieces of the head step as CUDA graphs over 6 rotating (cold) input sets.
PLH_LIB selects the library build.  usage: python tools/step_sweep.py [name ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth, _lib
B, H, W, NS = int(os.environ.get("SW_B", 32)), int(os.environ.get("SW_H", 128)), int(os.environ.get("SW_W", 128)), 6
dev = torch.device("cuda", 0)
lib = _lib.load()
base = synth.make_batch(2, B, H, W, "C")
sets = []
for s in range(NS):
    d = {k: torch.as_tensor(np.ascontiguousarray(np.roll(base[k], s, axis=0))).to(dev) for k in ("pix_logits", "link_logits", "pix_lab", "link_lab")}
    d["out"], d["out2"] = {}, {}
    d["ws"] = torch.empty(lib.plh_workspace_bytes(_lib.OP_LOSS, B, H, W, 0), dtype=torch.uint8, device=dev)
    sets.append(d)
ms = torch.cuda.Stream(dev)
torch.cuda.set_stream(ms)
lcfg, only = head.LossConfig(), head.LossConfig(main_only=True)
dc, dt = head.DecodeConfig(max_boxes=128, form="resident"), head.DecodeConfig(max_boxes=128, form="tiled")
A = lambda d: (d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"])
cases = {
    "loss": lambda d: head.pixellink_loss_raw(*A(d), lcfg, True, False, None, d["out2"], d["ws"]),
    "main_only": lambda d: head.pixellink_loss_raw(*A(d), only, True, False, None, d["out2"], d["ws"]),
    "flags": lambda d: head.decode_flags_raw(d["pix_logits"], d["link_logits"], dc, d["out"]),
    "decode_resident": lambda d: head.decode_raw(d["pix_logits"], d["link_logits"], dc, d["out"], want_rects=False),
    "decode_tiled": lambda d: head.decode_raw(d["pix_logits"], d["link_logits"], dt, d["out"], want_rects=False),
    "decode_split": lambda d: head.decode_from_flags_raw(head.decode_flags_raw(d["pix_logits"], d["link_logits"], dt, d["out"])["flags"], dt, d["out"], want_rects=False),
    "step_fork": lambda d: head.loss_and_decode_raw(*A(d), lcfg, dc, d["out"]),
    "step_tile_first": lambda d: head.loss_and_decode_raw(*A(d), lcfg, dt, d["out"]),
    "step_fork_tiled": lambda d: head.loss_and_decode_raw(*A(d), lcfg, dt, d["out"], schedule="fork"),
    "step_serial": lambda d: head.loss_and_decode_raw(*A(d), lcfg, dc, d["out"], parallel=False),
}
names = sys.argv[1:] or list(cases)
res = []
for name in names:
    f = cases[name]
    for d in sets: f(d)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=ms):
        for d in sets: f(d)
    torch.cuda.synchronize(); torch.cuda.set_stream(ms)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 150
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    res.append("%s %.2f" % (name, e0.elapsed_time(e1) * 1e3 / (reps * NS)))
print(os.environ.get("PLH_LIB", "libplhead.so"), "B%d %dx%d us:" % (B, H, W), " | ".join(res))
