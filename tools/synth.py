"""Synthetic inputs of the benchmark / parity workloads: numpy only, no oracle, no torch.

Shared by bench.py (both arms), tests/helpers.py and the oracle front-end.  Keeping these builders outside
`oracle/` means the GPU arm of bench.py never imports (or maps) the CPU checker.
"""
import numpy as np

def camera_from_pose25(pose, znear=0.01, zfar=100.0):
    """Restates FlowMatchingEngine.c_to_3dgs_format
    (/root/reference/nsr/lsgm/flow_matching_trainer.py:2174-2228) with
    getWorld2View2 / getProjectionMatrix
    (/root/reference/utils/gs_utils/graphics_utils.py:38-85).
    pose: 25 floats = c2w(16, row major) + K(9, normalised).  Returns
    (cam_view[4,4], cam_view_proj[4,4], cam_pos[3], tanfov) in the reference's
    row-vector (transposed) layout, float32."""
    pose = np.asarray(pose, dtype=np.float32)
    c2w = pose[:16].reshape(4, 4)
    w2c = np.linalg.inv(c2w)
    R = np.transpose(w2c[:3, :3])
    T = w2c[:3, 3]
    fx = float(pose[16])
    fov = 2.0 * np.arctan(1.0 / (2.0 * fx))            # focal2fov(fx, 1)
    tanfov = float(np.tan(fov * 0.5))
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    Rt = np.linalg.inv(np.linalg.inv(Rt))               # getWorld2View2 with trans=0, scale=1
    world_view = np.float32(Rt).T
    th = np.tan(fov / 2.0)
    top = th * znear
    right = th * znear
    Pm = np.zeros((4, 4), np.float32)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    full = (world_view.astype(np.float32) @ Pm.T).astype(np.float32)
    cam_pos = np.linalg.inv(world_view)[3, :3].astype(np.float32)
    return world_view.astype(np.float32), full, cam_pos, tanfov


def orbit_pose25(azim_deg, elev_deg, radius=1.8, fx=1.3889):
    """A look-at-origin camera in the same 25-float layout as
    /root/reference/assets/objv_eval_pose.pt (c2w row-major + normalised K);
    used when that fixture is not on disk (GPU box)."""
    az, el = np.deg2rad(azim_deg), np.deg2rad(elev_deg)
    eye = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
    K = np.array([fx, 0, 0.5, 0, fx, 0.5, 0, 0, 1.0])
    return np.concatenate([c2w.reshape(-1), K]).astype(np.float32)


def synthetic_surfels(P, seed=0, scale_boost=1.0):
    """[P,13] surfels per SURVEY.md 8(d): xyz U(-0.45,0.45)^3, opacity sigmoid(N),
    scales softplus(N(-2.5,1))*0.0045/ln2 clamped to [1e-4,0.05], unit quats,
    rgb 0.5*tanh(N)+0.5 (activations of /root/reference/vit/vit_triplane.py:1289-1313)."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-0.45, 0.45, (P, 3))
    op = 1.0 / (1.0 + np.exp(-rng.standard_normal((P, 1))))
    sc = np.log1p(np.exp(rng.standard_normal((P, 2)) - 2.5)) * (0.0045 / np.log(2.0)) * scale_boost
    sc = np.clip(sc, 1e-4, 0.05)
    q = rng.standard_normal((P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    rgb = 0.5 * np.tanh(rng.standard_normal((P, 3))) + 0.5
    return np.concatenate([xyz, op, sc, q, rgb], 1).astype(np.float32)
