"""CPU oracle for the segmentation training hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch-CPU fp32 + numpy + cv2) of the arithmetic that the
reference `musket-ml/segmentation_training_pipeline` reaches through its un-vendored dependencies
(keras>=2.2.4, tensorflow==1.15, segmentation_models==0.2.1, classification_models, imgaug==0.3.0,
musket_core; reference requires.txt:8-15, setup.py:24).  None of those are installable here and the
reference ships no tests / golden vectors for the path, therefore:

    *** PARITY UNPINNED *** for everything tagged [DEP] (recalled dependency semantics).

What IS pinned in this container:
  * augment sampling indices / pixels: bit-exact against cv2.warpAffine 4.13 (the backend imgaug calls),
    see oracle/augment.py and tests/test_oracle_augment.py;
  * fold splits: sklearn.model_selection.KFold;
  * closed-form known answers for the losses and the Keras-Adam first step.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this package.  The product (`segmentation_training_pipeline_b200`) never does.
"""
