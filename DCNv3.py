"""Top-level module ``DCNv3``: the import name of the reference's compiled extension
(models/ops_dcnv3/setup.py:66, imported at models/ops_dcnv3/functions/dcnv3_func.py:16).

With this repository on ``sys.path`` the reference's ``functions/dcnv3_func.py`` and
``modules/dcnv3.py`` run unmodified on the sm_100a kernels.  See INTEGRATION.md.
"""
from yolo_somi_b200.dcnv3_ext import dcnv3_backward, dcnv3_forward  # noqa: F401

__all__ = ["dcnv3_forward", "dcnv3_backward"]
