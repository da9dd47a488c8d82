"""Drop-in module name of the reference (`from src.denoising_toy_utils import *` in main_toy.py)."""
from physicsinformeddiffusionmodels_b200.denoising_toy_utils import *  # noqa: F401,F403
from physicsinformeddiffusionmodels_b200.denoising_toy_utils import device, nn, np, torch, F  # noqa: F401
