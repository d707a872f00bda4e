"""Writes tests/golden/dit_state_dict_keys.json: the exact state_dict key -> shape map of the REFERENCE's own DiT
registry entries (/root/reference/dit/dit_i23d.py:1665-1697), built here on CPU from the unmodified reference code
(third-party xformers / timm stubbed, _ref_stubs.py).  tests/test_oracle_dit.py checks that the mirror modules have
exactly these keys and shapes, i.e. that a reference checkpoint loads with strict=True.
    python tests/golden/make_dit_keys.py"""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_stubs import *  # noqa: F401,F403,E402
import torch  # noqa: E402
from dit import dit_i23d  # noqa: E402

out = {}
for name, cin in (("DiT-PixArt-PCD-CLAY-B", 3), ("DiT-PixArt-PCD-CLAY-L", 3), ("DiT-PixArt-PCD-CLAY-stage2-L", 10)):
    with contextlib.redirect_stdout(io.StringIO()):
        m = dit_i23d.DiT_models[name](input_size=32, num_classes=0, learn_sigma=False, in_channels=cin, context_dim=1024,
                                      roll_out=True, pooling_ctx_dim=768)
    out[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    print(name, len(out[name]), "keys")
    del m
json.dump(out, open(os.path.join(HERE, "dit_state_dict_keys.json"), "w"), indent=0, sort_keys=True)
