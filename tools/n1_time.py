"""Ad-hoc: device time of plh_link_labels (N1) at the benchmark map size, graph of 6 launches over rotating buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head
B, H, W, NS = 32, 128, 128, 6
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
ids = np.zeros((NS, B, H, W), np.uint8)
for s in range(NS):
    for b in range(B):
        for k in range(1, 9):
            y0, x0 = rng.integers(0, H - 8), rng.integers(0, W - 30)
            ids[s, b, y0:y0 + rng.integers(4, 16), x0:x0 + rng.integers(10, 60)] = k
ids = torch.as_tensor(ids).to(dev)
import ctypes as C
from tensorflow_ocr_b200 import _lib
lib = _lib.load()
link = [torch.empty((B, H, W, 8), device=dev) for _ in range(NS)]
pix = [torch.empty((B, H, W), device=dev) for _ in range(NS)]
st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
def run(s):
    lib.plh_link_labels(C.c_void_p(ids[s].data_ptr()), B, H, W, C.c_void_p(link[s].data_ptr()), C.c_void_p(pix[s].data_ptr()), C.c_void_p(st.cuda_stream))
for s in range(NS): run(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=st):
    for s in range(NS): run(s)
torch.cuda.synchronize(); torch.cuda.set_stream(st)
for _ in range(5): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): g.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (200 * NS)
byt = B * H * W * (1 + 32 + 4)
print("plh_link_labels b32 128x128: %.2f us/launch, %.0f GB/s algorithmic (37 B/px)" % (us, byt / us / 1e3))
