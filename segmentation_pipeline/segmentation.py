from segmentation_training_pipeline_b200.segmentation import *  # noqa: F401,F403
from segmentation_training_pipeline_b200.segmentation import (PipelineConfig, custom_models, custom_objects,  # noqa: F401
                                                              dataset_augmenters, extra_train, parse, parse_augmentation,
                                                              parse_loss)
