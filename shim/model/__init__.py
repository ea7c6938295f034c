"""Drop-in `model` package: put this directory in front of the reference checkout on sys.path / PYTHONPATH and the
reference's own `from model import build_segmenter` (train.py:27, test.py, tools/latency.py:11) resolves to the
sm_100a implementation — same call, same return value ((model, param_list), model/__init__.py:32-49), same
state_dict.  Nothing in the reference tree is edited; `utils`, `engine`, `config/*.yaml` keep coming from it.

    PYTHONPATH=/path/to/cris-b200/shim:/path/to/cris-b200 python train.py --config config/refcoco/cris_r50.yaml
"""
from cris.pytorch_b200.module import CRIS, build_segmenter  # noqa: F401

__all__ = ["build_segmenter", "CRIS"]
