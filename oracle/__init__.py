"""CPU oracle for the PixelLink/EAST head — TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy / OpenCV restatement of the reference's algorithm for the hot path
(SURVEY.md §8a), each function citing the reference file:line it follows.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``tensorflow_ocr_b200`` never
does (tests/test_host_logic.py greps for it).

Parity pinning: the reference has no tests or golden vectors.
The oracle is pinned against ``tests/golden/*.npz``, produced by
``tests/golden/make_golden.py``, which executes the reference's OWN Python
source (read from /root/reference at generation time) against a TensorFlow-1.x
op shim (TF 1.4 cannot be installed offline) and against the container's
OpenCV 4.13.0.  Items that do not exist in the reference at all (focal loss,
EAST geometry loss, locality-aware NMS) are restated from their papers /
upstream and are marked "parity unpinned" where they are defined.
"""
