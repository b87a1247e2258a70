"""CPU oracle for the bugcar perception hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is shipped or measured as the
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as the
checker.  The product package ``bugcar_image_segmentation_b200`` never imports it.

Parity status (see DESIGN.md "Oracle"):
  * pre/post-processing (preprocess, argmax+LUT, BEV warp, occupancy grid): PINNED
    against the reference's own ``models.py`` / ``bev.py`` executed in the build
    container under import stubs (``oracle/refstub.py``) and against OpenCV 4.13;
    golden vectors are committed under ``tests/golden/``.
  * ENet forward: PARITY UNPINNED.  The reference ships neither the layer graph
    (models.py:21-31 only imports a frozen GraphDef) nor any weights
    (.MISSING_LARGE_BLOBS:1-3), and TensorFlow is absent.  ``oracle/enet_oracle.py``
    is a torch-fp32 restatement of canonical ENet (Paszke et al. 2016, PyTorch-ENet
    conventions) with seeded synthetic weights.
"""
