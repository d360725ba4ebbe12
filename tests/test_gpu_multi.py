"""2-rank NCCL test on real GPUs: batch shards vs the unsharded batch on the CUDA path.
Skipped on a single-GPU box (the gloo / CPU counterpart is tests/test_host_logic.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_shards_match_unsharded(cuda_dev):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(here, "multi_gpu_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:]


def test_two_devices_one_process(cuda_dev):
    """One process driving two devices (kernel attributes are per device: the > 48 KB shared-memory opt-in must
    be made on each)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tensorflow_ocr_b200 import head, synth
    inp = synth.make_batch(62, 3, 64, 64, "G")
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        t = {k: torch.as_tensor(inp[k]).to(dev) for k in ("pix_logits", "link_logits", "pix_lab", "link_lab")}
        out = head.loss_and_decode_raw(t["pix_logits"], t["link_logits"], t["pix_lab"], t["link_lab"])
        torch.cuda.synchronize(dev)
        outs.append({k: v.cpu().numpy() for k, v in out.items() if k in ("stats", "labels", "n_boxes", "grad_link")})
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k], equal_nan=True), k
