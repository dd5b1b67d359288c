"""tlsan_b200: B200-native (sm_100a) TLSAN train / scoring hot path behind the reference
``Model`` surface (TLSAN/model.py) and ``input.py`` batch layout.  See DESIGN.md."""
from .input import CsrDataset, DataInput, DataInputTest, PackedBatch  # noqa: F401

__all__ = ["Model", "DataInput", "DataInputTest", "CsrDataset", "PackedBatch"]


def __getattr__(name):
    if name == "Model":
        from .model import Model
        return Model
    raise AttributeError(name)
