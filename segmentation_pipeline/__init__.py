"""Drop-in import path of the reference package: `from segmentation_pipeline import segmentation` keeps working; the
implementation lives in segmentation_training_pipeline_b200 (B200-native engine behind the same API)."""
