"""Output folder + logger set-up with the reference's interface (lens/src/loggers.py:7-65).

`model_logger(model)` creates `./lens/output/<ddmmyy-HH-MM-SS>/`, attaches `model.logger`
and `model.output_folder`, and returns the torch device the model runs on.  This package has
no CPU path, so the device is always CUDA and a missing GPU is an error.
"""
import logging
import os
from datetime import datetime

import torch

from .._lib import LensError


def model_logger(model, output_base_folder="./lens/output/"):
    now = datetime.now()
    model.output_folder = os.path.join(output_base_folder, now.strftime("%d%m%y-%H-%M-%S"))
    os.makedirs(model.output_folder, exist_ok=True)
    model.logger = logging.getLogger("LENS")
    if model.logger.hasHandlers():
        model.logger.handlers.clear()
    model.logger.setLevel(logging.DEBUG)
    fh = logging.FileHandler(os.path.join(model.output_folder, "lens.log"), mode="a+")
    fh.setFormatter(logging.Formatter("%(asctime)-15s %(levelname)-8s %(message)s"))
    model.logger.addHandler(fh)
    if not getattr(model, "quiet", False):
        model.logger.addHandler(logging.StreamHandler())
    model.logger.info("")
    model.logger.info("LENS: Locational Encoding with Neuromorphic Systems -- lens_b200 (B200-native inference path)")
    model.logger.info("")
    if not torch.cuda.is_available():
        raise LensError("lens_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    model.logger.info(f"Current device is {torch.cuda.get_device_name(torch.cuda.current_device())}")
    model.logger.info("")
    return torch.device("cuda")
