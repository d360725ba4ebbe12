"""Worker of tests/test_gpu_multi.py (one process per GPU, NCCL): batch shards on the CUDA path against the
unsharded batch.  Exits non-zero on any mismatch; rank 0 prints MULTI_GPU_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflow_ocr_b200 import _lib, head, synth  # noqa: E402
from tensorflow_ocr_b200 import dist as pdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, H, W = 8, 64, 96
    inp = synth.make_batch(61, B, H, W, "G", edge_images=True)
    keys = ("pix_logits", "link_logits", "pix_lab", "link_lab")
    lcfg, dcfg = head.LossConfig(), head.DecodeConfig(max_boxes=128)

    def run(arrs):
        t = {k: torch.as_tensor(np.ascontiguousarray(arrs[k])).to(dev) for k in keys}
        out = {}
        head.pixellink_loss_raw(t["pix_logits"], t["link_logits"], t["pix_lab"], t["link_lab"], lcfg, True, True, None, out)
        head.decode_raw(t["pix_logits"], t["link_logits"], dcfg, out, want_rects=False)
        torch.cuda.synchronize()
        return out

    lo, hi = pdist.shard_bounds(B, world, rank)
    mine = run({k: inp[k][lo:hi] for k in keys})
    def gather(t):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return parts

    assert B % world == 0
    # tower semantics (multigpu_train.py:111-125): every shard normalises by ITS OWN counts, so a rank's result is
    # what the path gives on that sub-batch alone — bit for bit on any GPU of the box (rank 0 recomputes every shard)
    for k in ("stats", "grad_pixel", "grad_link"):
        parts = gather(mine[k])
        if rank == 0:
            for r in range(world):
                a, b = pdist.shard_bounds(B, world, r)
                ref = run({kk: inp[kk][a:b] for kk in keys})
                assert torch.equal(torch.nan_to_num(parts[r], nan=-7.0), torch.nan_to_num(ref[k], nan=-7.0)), (k, r)
    # per-image decisions do not depend on the sharding at all (OHEM and decode are per image): gather the shards'
    # masks / labels / boxes and compare them with the unsharded batch computed on rank 0
    g_mask, g_lab, g_nb, g_boxes = (torch.cat(gather(mine[k]), 0) for k in ("ohem_mask", "labels", "n_boxes", "boxes"))
    if rank == 0:
        full = run(inp)
        assert torch.equal(g_mask, full["ohem_mask"]), "OHEM masks differ between shards and the unsharded batch"
        assert torch.equal(g_lab, full["labels"]), "labels differ"
        assert torch.equal(g_nb, full["n_boxes"]), "box counts differ"
        nb = full["n_boxes"].cpu().numpy()
        for i in range(B):
            assert torch.equal(g_boxes[i, :nb[i]], full["boxes"][i, :nb[i]]), "boxes differ in image %d" % i
    # the collective of the path: sum of the towers' loss scalars == sum of the shard stats computed locally
    red = pdist.allreduce_loss_stats(mine["stats"])
    parts = torch.cat(gather(mine["stats"][:_lib.STATS_FLOATS].reshape(1, -1)), 0)
    if rank == 0:
        s = parts.sum(0)
        cnt = [_lib.ST_N_SEG_POS, _lib.ST_N_SELECTED] + list(range(_lib.ST_SUM_WP, _lib.ST_SUM_WP + 16))
        assert torch.equal(red[cnt], s[cnt]), "all-reduced counts"
        assert torch.allclose(red[_lib.ST_TOTAL] * world, s[_lib.ST_TOTAL], rtol=1e-6, equal_nan=True), "all-reduced loss"
        # and the counts add up to the unsharded batch's (integers: exact)
        assert torch.equal(s[cnt], full["stats"][cnt]), "shard counts do not add up to the batch's"
        print("MULTI_GPU_OK world=%d" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
