"""`python src/predict.py [key=value ...]` — same entry point and config keys as the reference's
src/predict.py + configs/predict.yaml, running on the octseg B200 engine
(implementation: oct_segmentation_b200/predict.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oct_segmentation_b200.predict import (MODELS_META, cli, data_processing, load_model, main,  # noqa: E402,F401
                                           pick_device, preprocess_images, save_results, segment)

if __name__ == '__main__':
    cli()
