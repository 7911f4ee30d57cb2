from .dcnv3 import DCNv3  # noqa: F401
