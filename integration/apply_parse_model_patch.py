#!/usr/bin/env python
"""Register the DCNv3 hosting modules in the reference's model builder (the reference bundles ops_dcnv3 but
`parse_model` cannot name it).

    python integration/apply_parse_model_patch.py /path/to/YOLO-SOMI/models/yolo.py [--check]

Edits `models/yolo.py` in place (idempotent; `--check` only reports):
  1. after `from models.common import *` (models/yolo.py:15) import the four hosting modules of this library;
  2. add them to the list of modules whose arguments are rewritten to (c1, c2, ...) (models/yolo.py:1472-1479);
  3. add the two CSP containers to the list that receives the repeat count `n` (models/yolo.py:1488-1492).
After that a model yaml can say e.g. `[-1, 3, C3_DCNv3, [256]]` or `[-1, 1, DCNv3_YOLO, [256, 3, 1]]`.
"""
from __future__ import annotations

import re
import sys

IMPORT = "from yolo_somi_b200.hosting import DCNv3_YOLO, Bottleneck_DCNv3, C3_DCNv3, C2f_DCNv3  # DCNv3 (sm_100a)\n"
CHANNEL_NAMES = "DCNv3_YOLO, Bottleneck_DCNv3, C3_DCNv3, C2f_DCNv3"
REPEAT_NAMES = "C3_DCNv3, C2f_DCNv3"


def patch(text: str) -> str:
    if "yolo_somi_b200.hosting" in text:
        return text                                   # already applied
    m = re.search(r"^from models\.common import \*[^\n]*\n", text, re.M)
    if not m:
        raise ValueError("`from models.common import *` not found")
    text = text[:m.end()] + IMPORT + text[m.end():]
    # (2) the (c1, c2) list: `if m in [Conv, GhostConv, ...]:` followed by `c1, c2 = ch[f], args[0]`
    m = re.search(r"if m in \[\s*Conv\s*,(?P<body>[^\]]*)\]\s*:\s*\n\s*c1, c2 = ch\[f\], args\[0\]", text)
    if not m:
        raise ValueError("channel-handling list of parse_model not found")
    text = text[:m.start("body")] + " " + CHANNEL_NAMES + "," + text[m.start("body"):]
    # (3) the repeat list: `if m in [BottleneckCSP, C3, ...]:` followed by `args.insert(2, n)`
    m = re.search(r"if m in \[\s*BottleneckCSP\s*,(?P<body>[^\]]*)\]\s*:\s*\n\s*args\.insert\(2, n\)", text)
    if not m:
        raise ValueError("repeat-count list of parse_model not found")
    text = text[:m.start("body")] + " " + REPEAT_NAMES + "," + text[m.start("body"):]
    return text


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    path, check = sys.argv[1], "--check" in sys.argv
    src = open(path).read()
    out = patch(src)
    if out == src:
        print("already applied")
    elif check:
        print("would patch", path)
    else:
        open(path, "w").write(out)
        print("patched", path)
