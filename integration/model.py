"""Drop-in `model.py` for a checkout of smhassanerfani/nasa-niswan.

    mv model.py model_reference.py        # keep upstream's UNet / Pix2Pix classes importable
    cp <this repo>/integration/model.py model.py

`train.py:19` (`from model import Generator, UNet, ConvLSTM, initialize_weights`), `test.ipynb:32` and
`dataset_config.ipynb:758` then resolve unchanged: `ConvLSTM` / `ConvLSTMCell` are the B200-native classes (same
constructor, forward contract and state_dict as model.py:196-274), every other public name is re-exported from
upstream's own file.  `NINT_UPSTREAM_MODEL` names that module if it is not `model_reference`.  Without it the other
names still import, and raise only when used: the ConvLSTM path needs nothing from upstream.
"""
import importlib
import os

from nasa_niswan_b200.model import ConvLSTM, ConvLSTMCell  # noqa: F401

_UPSTREAM_NAMES = ("Generator", "GBlock", "Discriminator", "DBlock", "UNet", "Encoder", "Decoder", "conv_block",
                   "initialize_weights")
__all__ = ["ConvLSTM", "ConvLSTMCell"] + list(_UPSTREAM_NAMES)

_why = "the module does not define it"
try:
    _upstream = importlib.import_module(os.environ.get("NINT_UPSTREAM_MODEL", "model_reference"))
except ImportError as exc:
    _upstream, _why = None, f"not importable: {exc}"


def _missing(name):
    def _raise(*args, **kwargs):
        raise ImportError(f"{name} lives in the reference's own model.py, which this shim re-exports from the module "
                          f"'{os.environ.get('NINT_UPSTREAM_MODEL', 'model_reference')}' ({_why}); only "
                          "ConvLSTM / ConvLSTMCell are provided by nasa_niswan_b200")
    _raise.__name__ = name
    return _raise


for _name in _UPSTREAM_NAMES:
    globals()[_name] = getattr(_upstream, _name) if _upstream is not None and hasattr(_upstream, _name) else _missing(_name)
