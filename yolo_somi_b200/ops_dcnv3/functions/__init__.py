from .dcnv3_func import DCNv3Function, dcnv3_core  # noqa: F401
