"""Ad-hoc: time (a) the main loss pass relaunched alone (cold inputs), (b) the loss chain, (c) the fused head step,
all as CUDA graphs over 6 rotating input sets.  PLH_LIB selects the library build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth, _lib
B, H, W, NS = 32, 128, 128, 6
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("SWEEP_PG"):   # diagnostic: does an initialised process group change the device time?
    import torch.distributed as dist
    dist.init_process_group(os.environ["SWEEP_PG"])
    dist.barrier()
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
lcfg, only, dcfg = head.LossConfig(), head.LossConfig(main_only=True), head.DecodeConfig(max_boxes=128)
def f_main(d): head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], only, True, False, None, d["out2"], d["ws"])
def f_loss(d): head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], lcfg, True, False, None, d["out2"], d["ws"])
def f_dec(d): head.decode_raw(d["pix_logits"], d["link_logits"], dcfg, d["out"], want_rects=False)
def f_step(d): head.loss_and_decode_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], lcfg, dcfg, d["out"])
res = []
for name, f in (("loss", f_loss), ("main_only", f_main), ("decode", f_dec), ("step", f_step)):
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
    res.append("%s %.2f us" % (name, e0.elapsed_time(e1) * 1e3 / (reps * NS)))
print(os.environ.get("PLH_LIB", "libplhead.so"), " | ".join(res))
