import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth, _lib
from oracle import decode as D
dev = torch.device("cuda", 0)
B, H, W = 4, 128, 128
inp = synth.make_batch(11, B, H, W, "G", edge_images=True)
pl = torch.as_tensor(inp["pix_logits"]).to(dev); ll = torch.as_tensor(inp["link_logits"]).to(dev)
K = 256
cfg = head.DecodeConfig(max_boxes=K)
for rep in range(3):
    out = head.decode_raw(pl, ll, cfg)
    torch.cuda.synchronize()
    ws = head._workspace(_lib.OP_DECODE, B, H, W, K, dev)
    flags = ws[: B * H * W * 2].view(torch.int16).cpu().numpy().astype(np.int32).reshape(B, H, W) & 0xffff
    for b in range(B):
        P, L = D.thresholds(inp["pix_logits"][b], inp["link_logits"][b])
        Pg = (flags[b] >> 8) & 1
        Lg = np.stack([(flags[b] >> d) & 1 for d in range(8)], -1)
        lab, roots, sizes = D.link_components(P, L, 10)
        gl = out["labels"][b].cpu().numpy()
        nb = int(out["n_boxes"][b])
        diff = np.argwhere(gl != lab)
        print("rep", rep, "image", b, "P mism", int((Pg != P).sum()), "L mism", int((Lg != L).sum()), "P count", int(P.sum()),
              "oracle comps", len(roots), "gpu n_boxes", nb, "label diffs", len(diff))
        if len(diff):
            print("  oracle roots/sizes", list(zip(roots.tolist(), sizes.tolist()))[:20])
            comp = out["comp"][b, :min(nb, K)].cpu().numpy()
            print("  gpu comp", comp.tolist()[:20])
            for (y, x) in diff[:8]:
                print("   at", y, x, "gpu", gl[y, x], "oracle", lab[y, x], "P", P[y, x], "flags", hex(flags[b, y, x]))
            vals, cnt = np.unique(gl[gl != lab], return_counts=True)
            print("  gpu labels at diffs", list(zip(vals.tolist(), cnt.tolist()))[:10])
            vals, cnt = np.unique(lab[gl != lab], return_counts=True)
            print("  oracle labels at diffs", list(zip(vals.tolist(), cnt.tolist()))[:10])
