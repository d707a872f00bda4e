"""Time only the DiT legs of bench.py (C3 sampler + deployed DiT-L): python tools/dit_leg.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

dev = torch.device("cuda:0")
print(json.dumps(bench.run_dit_leg(dev)))
print(json.dumps(bench.run_dit_deployed_leg(dev)))
if "n1" in sys.argv:
    print(json.dumps(bench.run_vae_decoder_leg(dev)))
if "c4" in sys.argv:
    print(json.dumps(bench.run_dit_deployed_leg(dev, nfe=10, N=4096)))
