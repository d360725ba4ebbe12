import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_ocr_b200 import head, synth
from oracle import decode as D
fam, B, H, W = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
inp = synth.make_batch(11, B, H, W, fam, edge_images=(B >= 4))
dev = torch.device("cuda", 0)
pl, ll = torch.as_tensor(inp["pix_logits"]).to(dev), torch.as_tensor(inp["link_logits"]).to(dev)
res = {}
for form in ("auto", "tiled"):
    out = head.decode_raw(pl, ll, head.DecodeConfig(form=form, max_boxes=256))
    torch.cuda.synchronize()
    res[form] = {k: v.cpu().numpy() for k, v in out.items()}
a, t = res["auto"], res["tiled"]
for b in range(B):
    P, L = D.thresholds(inp["pix_logits"][b], inp["link_logits"][b])
    lab, roots, sizes = D.link_components(P, L, 10)
    la, lt = a["labels"][b], t["labels"][b]
    print("image", b, "n_boxes auto/tiled/oracle", a["n_boxes"][b], t["n_boxes"][b], len(roots), "diff px auto-vs-oracle", (la != lab).sum(), "tiled-vs-oracle", (lt != lab).sum())
    if (la != lab).any():
        all_lab, all_roots, all_sizes = D.link_components(P, L, 0)
        ys, xs = np.nonzero(la != lab)
        for y, x in list(zip(ys, xs))[:6]:
            print("  px", y, x, "auto", la[y, x], "oracle", lab[y, x], "P", P[y, x], "comp(all)", all_lab[y, x], "size", all_sizes[list(all_roots).index(all_lab[y, x])] if all_lab[y, x] >= 0 else None)
        bad_roots = np.unique(all_lab[la != lab])
        print("  affected components:", [(int(r), int(all_sizes[list(all_roots).index(r)])) for r in bad_roots[:10]])
        # how does auto split them?
        for r in bad_roots[:3]:
            m = all_lab == r
            print("   comp", r, "auto labels inside:", np.unique(la[m], return_counts=True))
