from .dataset import StyleTransferDataset  # noqa: F401
