from segmentation_training_pipeline_b200.impl.datasets import *  # noqa: F401,F403
from segmentation_training_pipeline_b200.impl.datasets import PredictionItem, SimplePNGMaskDataSet  # noqa: F401
