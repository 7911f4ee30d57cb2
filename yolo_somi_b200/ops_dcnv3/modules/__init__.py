from .dcnv3 import DCNv3, DCNv3_pytorch  # noqa: F401
