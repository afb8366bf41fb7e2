"""CPU oracle for the ReLaX-VQA feature-extraction hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy / torch-CPU fp32)
of the reference algorithm, one function per row of SURVEY.md section 8(a).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package (``relax_vqa_b200``)
never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * fragments.py  (A1-A4, A7, A8): pinned bit-exactly against the reference's shipped
    example PNGs (tests/golden/ref_example) and against the reference's own functions
    imported in the build container (tests/golden/gen_golden.py -> *.npz).
  * resize.py     (A9): pinned bit-exactly against Pillow (the reference's dependency).
  * farneback.py  (A5, A6): pinned against cv2.calcOpticalFlowFarneback / cv2.cvtColor
    (the reference's dependency, present in the image) to the float tolerance stated
    in tests/test_oracle_flow.py, and against the shipped *_residual_of.png fixtures.
  * backbones.py, head.py, pipeline.py (A10-A18): no reference artefact pins these
    numerically (no stored features, heads absent) -> pinned against outputs of the
    reference itself, imported through shims with seeded weights
    (tests/golden/gen_golden.py; fixtures committed under tests/golden/).
"""
