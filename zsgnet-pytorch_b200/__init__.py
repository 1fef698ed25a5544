"""zsg-b200: the ZSGNet per-batch hot path (forward / loss / backward / metric) as hand-written
sm_100a CUDA behind a C ABI, with a host-side mirror of the reference's operator interface.

Reference-facing modules (same names and call signatures as the reference's code/ directory):
    mdl.get_default_net, loss.get_default_loss, evaluator.get_default_eval, dat_loader.get_data
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing at first use)

__all__ = ["_lib"]
