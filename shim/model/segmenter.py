"""`from model.segmenter import CRIS` (model/segmenter.py:10) -> the sm_100a implementation."""
from cris.pytorch_b200.module import CRIS  # noqa: F401
