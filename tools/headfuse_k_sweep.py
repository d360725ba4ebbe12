import sys, os, torch
sys.path.insert(0, "/root/repo")
from tensorflow_ocr_b200 import head
dev = torch.device("cuda", 0)
def run(K, H, W, B=32):
    xs = [torch.randn(B, H, W, K, device=dev) for _ in range(3)]
    w = torch.randn(K, 18, device=dev) / K ** 0.5
    bias = torch.randn(18, device=dev)
    f = lambda x: head.head_fuse_level_raw([(x, w, None, bias, False)])
    for x in xs: f(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12): f(xs[i % 3])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 12
    nb = B * H * W * (K + 18) * 4
    print("K=%4d %dx%d: %7.1f us %7.1f MB %7.1f GB/s frac %.3f" % (K, H, W, us, nb / 1e6, nb / us / 1e3, nb / us / 1e3 / 6549.4))
run(32, 256, 256); run(64, 256, 128); run(128, 128, 128); run(256, 128, 128); run(1024, 32, 64)
