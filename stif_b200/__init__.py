"""stif_b200 -- B200-native drop-in for STIF's variogram and kriging hot path.

Same exports as the reference's stif/__init__.py:1-3.
"""
from .data import Data, sinusoidal_feature_transform  # noqa: F401
from .predictor import Predictor  # noqa: F401

__version__ = "0.1.0"
