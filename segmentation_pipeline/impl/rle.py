from segmentation_training_pipeline_b200.impl.rle import *  # noqa: F401,F403
from segmentation_training_pipeline_b200.impl.rle import (masks_as_image, masks_as_images, multi_rle_encode,  # noqa: F401
                                                          rle_decode, rle_encode)
