import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from baseline.raster_standin import StandinRasterizer
from tests.helpers import cameras, oracle_view, scene
P, H, W, V = 4000, 160, 144, 2
g = scene(P, 90, 6.0); vs, ps, _, _ = cameras(V, start=1); bg = [1.0, 0.5, 0.2]
dev = torch.device("cuda:0")
r = StandinRasterizer(P, H, W, V)
g13 = torch.tensor(g, device=dev)
color, allmap, radii, nr = r.forward(g13, torch.tensor(vs, device=dev), torch.tensor(ps, device=dev), torch.tensor(bg, device=dev))
o = oracle_view(g, vs[0], ps[0], bg, H, W)
rr = radii[0].cpu().numpy()
print("nr", nr, "oracle", o["num_rendered"], "radii equal", (rr == o["radii"]).mean(), "visible", (rr > 0).sum(), (o["radii"] > 0).sum())
bad = np.nonzero(rr != o["radii"])[0][:10]
print("first mismatches", [(int(i), int(rr[i]), int(o["radii"][i])) for i in bad])
print("dtypes", g13.dtype, g13.shape, g13.is_contiguous())
