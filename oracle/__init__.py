"""CPU oracle for the climsr generator hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as
the checker (or as the CPU arm that is timed *beside* the CUDA path).  The
product path (``climsr_b200``) never imports this package and raises when
its CUDA extension is missing.

Parity status:
  * generator forward/backward: the reference's own tests pin only the output
    SHAPE (tests/models/test_esrgan.py:7-22) -> values are pinned here by
    golden vectors generated from the imported reference module
    (oracle/make_golden.py, run in the build container where /root/reference
    exists; fixtures committed under tests/golden/).
  * RegressionAccuracy: pinned by the nine known-answer cases of
    tests/metrics/test_regresion_accuracy.py:12-108.
  * PSNR/SSIM/MAE/MSE/RMSE/MAPE/SMAPE/R2: "parity unpinned" - torchmetrics is
    neither vendored in the reference nor installed; the formulas restate the
    torchmetrics 0.5-0.7 API the reference imports (core/task.py:13-21).
"""
