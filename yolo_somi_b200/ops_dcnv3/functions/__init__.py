from .dcnv3_func import DCNv3Function, dcnv3_core, dcnv3_core_pytorch  # noqa: F401
