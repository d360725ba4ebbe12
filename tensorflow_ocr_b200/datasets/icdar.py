"""Mirror of the reference's ``datasets/icdar.py``: the EAST geometry restore (lines 410-483) and the ground-truth
generator ``generate_rbox`` with its ``valid_link`` (lines 83-105, 486-539, 632-634)."""
from __future__ import annotations

from .. import head

__all__ = ["restore_rectangle_rbox", "restore_rectangle", "rasterize_polygons", "generate_rbox", "generate_rbox_4s"]


def _float_dtype(x):
    """float64 inputs stay float64 (the reference computes in the dtype it is given); everything else
    (float32, float16, integers) is taken as float32."""
    import numpy as np
    import torch
    dt = x.dtype
    return torch.float64 if dt in (torch.float64, np.float64, np.dtype("float64")) else torch.float32


def restore_rectangle_rbox(origin, geometry, return_index=False):
    """datasets/icdar.py:410-479: ``origin`` [N,2], ``geometry`` [N,5] (top,right,bottom,left,theta)
    -> [N,4,2] float64.  Rows come out theta >= 0 first, then theta < 0 (icdar.py:479, quirk
    Q16); ``return_index=True`` (extension) also returns the input row of each output row."""
    o, np_in = head.to_device(origin, dtype=_float_dtype(origin))
    g, _ = head.to_device(geometry, dtype=_float_dtype(geometry), device=o.device)
    if o.dim() != 2 or o.shape[1] != 2 or g.dim() != 2 or g.shape[1] != 5 or g.shape[0] != o.shape[0]:
        raise ValueError("restore_rectangle expects origin [N,2] and geometry [N,5]")
    out, idx = head.restore_rectangle_raw(o, g, want_index=return_index)
    if np_in:
        out = out.cpu().numpy()
        idx = idx.cpu().numpy() if idx is not None else None
    return (out, idx) if return_index else out


def restore_rectangle(origin, geometry):
    """datasets/icdar.py:482-483."""
    return restore_rectangle_rbox(origin, geometry)


def _rasterize_device(im_size, polys, tags, min_text_size):
    import numpy as np
    import torch
    h, w = int(im_size[0]), int(im_size[1])
    polys = np.asarray(polys)
    quads = polys.astype(np.int32).reshape(-1, 4, 2)               # np.int32 truncation, datasets/icdar.py:497
    zero = np.zeros((len(quads),), np.uint8)
    for k, (poly, tag) in enumerate(zip(polys, tags)):             # :499-503 (host: four norms per polygon)
        poly_h = min(np.linalg.norm(poly[0] - poly[3]), np.linalg.norm(poly[1] - poly[2]))
        poly_w = min(np.linalg.norm(poly[0] - poly[1]), np.linalg.norm(poly[2] - poly[3]))
        zero[k] = 1 if (min(poly_h, poly_w) < min_text_size or tag) else 0
    dev = head._require_gpu(None)
    return head.fill_quads_raw(torch.as_tensor(quads).to(dev), [len(quads)], h, w, h, w, mode=0, stride=1,
                               zero_flags=torch.as_tensor(zero).to(dev), want=("last", "first", "training_mask"))


def rasterize_polygons(im_size, polys, tags, min_text_size=10):
    """The drawing half of generate_rbox (datasets/icdar.py:486-514): cv2.fillPoly of every polygon, in order, with
    the same float -> int32 truncation of the vertices — on the GPU (`plh_fill_quads`, bit-identical to OpenCV).
    Returns last_ids (poly_mask: index of the LAST polygon covering a pixel), first_ids (index of the FIRST one:
    which pixels are text once polygons 0..k are in) and training_mask (0 inside polygons that are tagged or
    smaller than min_text_size, FLAGS.min_text_size :25), as numpy arrays."""
    out = _rasterize_device(im_size, polys, tags, min_text_size)
    return out["last"][0].cpu().numpy(), out["first"][0].cpu().numpy(), out["training_mask"][0].cpu().numpy()


def _generate(im_size, polys, tags, min_text_size, stride):
    import numpy as np
    out = _rasterize_device(im_size, polys, tags, min_text_size)
    link, score = head.link_labels_icdar_raw(out["last"], out["first"], stride)
    return (score[0].cpu().numpy().astype(np.uint8), link[0].cpu().numpy(),
            out["training_mask"][0, ::stride, ::stride].cpu().numpy())


def generate_rbox(im_size, polys, tags, min_text_size=10):
    """datasets/icdar.py:486-539: (score_map uint8 [h,w], geo_map float32 [h,w,8] = the 8 link labels,
    training_mask uint8 [h,w]).  The per-pixel valid_link loop (~10^5 interpreted calls per image) runs on the
    GPU; quirks of the original kept (Q17), square maps only."""
    return _generate(im_size, polys, tags, min_text_size, 1)


def generate_rbox_4s(im_size, polys, tags, min_text_size=10):
    """What the batch generator keeps of it (datasets/icdar.py:632-634): score_map[::4, ::4], geo_map[::4, ::4, :],
    training_mask[::4, ::4] — computed at those pixels only."""
    return _generate(im_size, polys, tags, min_text_size, 4)
