"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).

Nothing under this package is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker (or as the timed
CPU baseline), never as the thing shipped.  The product path
(``livingscenes_b200``) never imports ``oracle`` and fails loudly when its CUDA
extension is missing.

Parity status
-------------
The reference (GradientSpaces/LivingScenes @ f290146) ships NO tests, golden
vectors or known-answer fixtures for the hot path, and the
only native arithmetic on the path lives in an un-vendored third-party package:

    pytorch3d 0.7.4  (install.sh:6, README.md:57)  --  knn_points,
    sample_farthest_points

so the oracle is pinned the only way available: the reference's OWN PyTorch
modules (vec_dgcnn_atten.VecDGCNN_att, model_utils.Shape_Prior / FieldWrapper,
deepsdf_decoder.DeepSDF_Decoder, lib_more.matcher_new, lib_more.pose_estimation)
are imported from /root/reference in the build container (``oracle/ref_loader``)
with the pytorch3d shim below, run on seeded inputs with the shipped checkpoint
and with seeded random weights, and the outputs are committed under
``tests/golden/`` by ``oracle/make_golden.py``.  ``oracle/restatement.py`` (our
own CPU restatement, which travels to the GPU box) is checked against those
fixtures by ``tests/test_oracle.py``.

The pytorch3d boundary itself (kNN tie-breaking / FPS arithmetic order) is
"parity unpinned": its semantics are restated from the published 0.7.4 behaviour
(see ``oracle/p3d_shim.py``), not from a copy of its source or its test vectors.
"""
