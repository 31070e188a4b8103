"""B200-native ConvLSTM hot path of Smart-NINT (smhassanerfani/nasa-niswan, model.py:196-274)."""
from .model import ConvLSTM, ConvLSTMCell  # noqa: F401
from .engine import Plan  # noqa: F401

__all__ = ["ConvLSTM", "ConvLSTMCell", "Plan"]
