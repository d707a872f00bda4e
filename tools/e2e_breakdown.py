"""Where does the end-to-end raster step (bench.py "e2e") spend its time?  Times variants of the step on one GPU."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gaussiananything_b200.gs_surfel import GaussianRenderer2DGS  # noqa: E402
from oracle import surfel_oracle as so  # noqa: E402

dev = torch.device("cuda:0")
P, V, RES = 100000, 6, 512
g = so.synthetic_surfels(P, seed=0)
cams = [so.camera_from_pose25(so.orbit_pose25(60.0 * i, 20.0)) for i in range(V)]
vs = np.stack([c[0] for c in cams]); ps = np.stack([c[1] for c in cams])
rnd = GaussianRenderer2DGS(RES, 3, {})
h_g = torch.tensor(g)[None].pin_memory(); h_vm = torch.tensor(vs)[None].pin_memory(); h_pm = torch.tensor(ps)[None].pin_memory()
h_pos = torch.zeros(1, V, 3).pin_memory(); h_t = torch.rand(1, V, 3, RES, RES).pin_memory()
h_loss = torch.zeros(1).pin_memory(); h_grad = torch.zeros(1, P, 13).pin_memory()


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


d = [h.to(dev) for h in (h_g, h_vm, h_pm, h_pos, h_t)]
print("H2D of the step inputs (24 MB, 5 tensors):      %.3f ms" % timeit(lambda: [h.to(dev, non_blocking=True) for h in (h_g, h_vm, h_pm, h_pos, h_t)]))
print("H2D of the target alone (18.9 MB):              %.3f ms" % timeit(lambda: h_t.to(dev, non_blocking=True)))


def step(loss_kind, d2h):
    gg = d[0].detach().clone().requires_grad_(True)
    out = rnd.render(gg, d[1], d[2], d[3], 0.36)
    if loss_kind == "full":
        loss = ((out["image"] - d[4]) ** 2).mean() + 0.1 * out["dist"].mean() + 0.05 * (1 - out["alpha"]).mean() \
            + 0.01 * out["depth"].mean() + 0.01 * out["rend_normal"].abs().mean()
    else:
        loss = out["image"].sum()
    loss.backward()
    if d2h:
        h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
        h_grad.copy_(gg.grad, non_blocking=True)


print("render+full loss+backward, device-resident:      %.3f ms" % timeit(lambda: step("full", False)))
print("  + D2H of loss and gradient:                    %.3f ms" % timeit(lambda: step("full", True)))
print("render+trivial loss+backward, device-resident:   %.3f ms" % timeit(lambda: step("sum", False)))


def cpu_only():
    t0 = time.perf_counter()
    step("full", False)
    return time.perf_counter() - t0


torch.cuda.synchronize()
ts = [cpu_only() for _ in range(30)]
torch.cuda.synchronize()
print("host time to enqueue one step (median):          %.3f ms" % (np.median(ts) * 1e3))

if "profile" in sys.argv:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(40):
        step("full", True)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(45)
