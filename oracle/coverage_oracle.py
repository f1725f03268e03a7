"""CPU oracle for the coverage hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's algorithm (ctu-vras/trajectory_optimization)
with *closed-form* gradients in place of torch autograd.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; nothing under `trajectory_optimization_b200/`
does, and the product path raises when its CUDA library is missing.

Pinning: every function here is checked (tests/test_oracle_golden.py) against
fixtures in `tests/golden/*.npz` that were produced by importing and running the
unmodified reference (`/root/reference/src/model.py`, `tools.py`) on CPU through
`oracle/shims` — see `tests/golden/make_golden.py`.  The reference itself ships
no tests or golden vectors, so those fixtures are the pin.

All functions take a `dtype` (np.float32 mimics the reference's precision,
np.float64 is the "truth" used to judge which of two fp32 answers is closer).
Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

# src/tools.py:320-325 — hard-coded intrinsics of the reference camera.
K_DEFAULT = np.array([[758.03967, 0.0, 621.46572],
                      [0.0, 761.62359, 756.86402],
                      [0.0, 0.0, 1.0]], dtype=np.float32)
IMG_WIDTH, IMG_HEIGHT = 1232.0, 1616.0


def load_intrinsics():
    """src/tools.py:320-325."""
    return K_DEFAULT.copy(), IMG_WIDTH, IMG_HEIGHT


# --------------------------------------------------------------------------------------
# quaternion / frame helpers
# --------------------------------------------------------------------------------------
def normalize_quat(quat, dtype=np.float64):
    """F.normalize(quat) — src/model.py:53 (eps = 1e-12 on the norm)."""
    q = np.asarray(quat, dtype=dtype).reshape(4)
    nrm = np.sqrt((q * q).sum(dtype=dtype))
    return q / max(nrm, dtype(1e-12)), max(nrm, dtype(1e-12))


def rot_from_unit_quat(q):
    """Rotation matrix R(q), q = (w, x, y, z): camera -> world, so that
    quaternion_apply(q^-1, y) == R(q)^T y (src/model.py:54-56)."""
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=q.dtype)


def drot_dquat(q):
    """dR_ab/dq_k of the polynomial above, shape (4, 3, 3)."""
    w, x, y, z = q
    o = q.dtype.type(0)
    d = np.empty((4, 3, 3), dtype=q.dtype)
    d[0] = 2 * np.array([[o, -z, y], [z, o, -x], [-y, x, o]], dtype=q.dtype)
    d[1] = 2 * np.array([[o, y, z], [y, -2 * x, -w], [z, w, -2 * x]], dtype=q.dtype)
    d[2] = 2 * np.array([[-2 * y, x, w], [x, o, z], [-w, z, -2 * y]], dtype=q.dtype)
    d[3] = 2 * np.array([[-2 * z, -w, x], [w, -2 * z, y], [x, y, o]], dtype=q.dtype)
    return d


def to_camera_frame(verts, quat, trans, dtype=np.float64):
    """src/model.py:50-57: c = q^-1 (v - t) q."""
    verts = np.asarray(verts, dtype=dtype)
    q, _ = normalize_quat(quat, dtype)
    R = rot_from_unit_quat(q)
    y = verts - np.asarray(trans, dtype=dtype).reshape(1, 3)
    return y @ R  # rows: R^T y


# --------------------------------------------------------------------------------------
# masks (src/model.py:13-47)
# --------------------------------------------------------------------------------------
def dist_mask(points, min_dist=1.0, max_dist=5.0):
    """src/model.py:13-24 — Gaussian about the camera-frame POINT (mu,mu,mu)."""
    dt = points.dtype.type
    mean = dt((min_dist + max_dist) / 2.0)
    std = dt((max_dist - min_dist) / 2.0)
    d = points - mean
    dist = np.sqrt((d * d).sum(axis=1))
    return np.exp(dt(-0.5) * (dist / std) ** 2)


def fov_mask(points, img_height, img_width, K, eps=1e-6, binary=False):
    """src/model.py:27-47."""
    dt = points.dtype.type
    K = np.asarray(K, dtype=points.dtype)
    h = K @ points.T
    if binary:
        with np.errstate(divide="ignore", invalid="ignore"):
            u = h[0] / h[2]
            v = h[1] / h[2]
        return (h[2] > 0) & (u > 1) & (u < dt(img_width - 1)) & (v > 1) & (v < dt(img_height - 1))
    z = h[2]
    s = dt(1) / (dt(1) + np.exp(-z))
    gw = np.exp(dt(-0.5) * ((h[0] / (z + dt(eps)) - dt(img_width / 2.0)) / dt(img_width)) ** 2)
    gh = np.exp(dt(-0.5) * ((h[1] / (z + dt(eps)) - dt(img_height / 2.0)) / dt(img_height)) ** 2)
    return s * gw * gh


def frustum_cull(points_3xn, img_height, img_width, K, min_dist=1.0, max_dist=10.0):
    """tools.get_cam_frustum_pts — src/tools.py:176-187 (fp32, strict inequalities).
    Returns (culled (M,3), dist_mask (N,), fov_mask (N,))."""
    pts = np.asarray(points_3xn, dtype=np.float32)
    K = np.asarray(K, dtype=np.float32)[:3, :3]
    dmask = (pts[2] > np.float32(min_dist)) & (pts[2] < np.float32(max_dist))
    h = (K @ pts).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = h[0] / h[2]
        v = h[1] / h[2]
    fmask = (h[2] > 0) & (u > 1) & (u < np.float32(img_width - 1)) & (v > 1) & (v < np.float32(img_height - 1))
    return pts[:, dmask & fmask].T.copy(), dmask, fmask


# --------------------------------------------------------------------------------------
# per-pose visibility m_j and dm/dc
# --------------------------------------------------------------------------------------
def visibility(points, trans, quat, K, img_width, img_height, min_dist, max_dist, eps,
               dtype=np.float64, want_grad=False):
    """m = dist_mask * fov_mask in the camera frame of pose (trans, quat)
    (src/model.py:101-110, :219-223).  With want_grad also returns
    (y, g_c, R, q, qnorm): y = x - t (world), g_c = dm/dc (camera frame)."""
    dt = np.dtype(dtype).type
    pts = np.asarray(points, dtype=dtype)
    q, qn = normalize_quat(quat, dtype)
    R = rot_from_unit_quat(q)
    t = np.asarray(trans, dtype=dtype).reshape(1, 3)
    y = pts - t
    c = y @ R
    Kd = np.asarray(K, dtype=dtype)
    d = dist_mask(c, min_dist, max_dist)
    f = fov_mask(c, img_height, img_width, Kd, eps)
    m = d * f
    if not want_grad:
        return m
    mu = dt((min_dist + max_dist) / 2.0)
    sg2 = dt(((max_dist - min_dist) / 2.0) ** 2)
    Wd, Hd = dt(img_width), dt(img_height)
    h = c @ Kd.T
    z = h[:, 2]
    zi = dt(1) / (z + dt(eps))
    u, v = h[:, 0] * zi, h[:, 1] * zi
    s = dt(1) / (dt(1) + np.exp(-z))
    gw = np.exp(dt(-0.5) * ((u - Wd / 2) / Wd) ** 2)
    gh = np.exp(dt(-0.5) * ((v - Hd / 2) / Hd) ** 2)
    # d(dist)/dc
    dd = -d[:, None] * (c - mu) / sg2
    # d(fov)/dh (h = K c), then chain through K
    dgw_du = -gw * (u - Wd / 2) / (Wd * Wd)
    dgh_dv = -gh * (v - Hd / 2) / (Hd * Hd)
    df_dh0 = s * gh * dgw_du * zi
    df_dh1 = s * gw * dgh_dv * zi
    df_dh2 = gw * gh * s * (1 - s) - s * gh * dgw_du * u * zi - s * gw * dgh_dv * v * zi
    df_dh = np.stack([df_dh0, df_dh1, df_dh2], axis=1)
    df = df_dh @ Kd
    g_c = f[:, None] * dd + d[:, None] * df
    return m, y, g_c, R, q, qn


def pose_grads_from_sums(sum_g_c, sum_ygT, R, q, qnorm):
    """Map (sum_j w_j g_c, sum_j w_j y g_c^T) to (dL/dtrans, dL/d(raw quat)).
    c_b = sum_a R_ab y_a  =>  dL/dR_ab = sum_j y_a g_cb ; dL/dt = -R sum g_c;
    F.normalize backward: (I - q q^T)/|q~| (src/model.py:53)."""
    g_t = -(R @ sum_g_c)
    dR = drot_dquat(q)
    g_q = np.einsum("kab,ab->k", dR, sum_ygT)
    g_qraw = (g_q - q * (q @ g_q)) / qnorm
    return g_t, g_qraw


# --------------------------------------------------------------------------------------
# ModelPose (src/model.py:65-127)
# --------------------------------------------------------------------------------------
def pose_objective(points, trans, quat, K, img_width, img_height, min_dist=1.0, max_dist=5.0,
                   eps=1e-6, weight=None, dtype=np.float64):
    """ModelPose.forward + backward.  Returns dict(loss, obs, sum, g_trans, g_quat).
    `weight` is the optional occlusion mask of the hpr=True branch (src/model.py:112-115)."""
    dt = np.dtype(dtype).type
    m, y, g_c, R, q, qn = visibility(points, trans, quat, K, img_width, img_height,
                                     min_dist, max_dist, eps, dtype, want_grad=True)
    if weight is not None:
        wv = np.asarray(weight, dtype=dtype)
        m = m * wv
        g_c = g_c * wv[:, None]
    total = m.sum(dtype=np.float64)
    loss = 1.0 / (total + eps)
    w = -loss * loss
    sum_g = w * g_c.sum(axis=0, dtype=np.float64)
    sum_yg = w * np.einsum("ja,jb->ab", y.astype(np.float64), g_c.astype(np.float64))
    g_t, g_q = pose_grads_from_sums(sum_g, sum_yg, R.astype(np.float64), q.astype(np.float64), float(qn))
    return dict(loss=dt(loss), obs=m, sum=total, g_trans=g_t, g_quat=g_q)


# --------------------------------------------------------------------------------------
# ModelTraj (src/model.py:158-260)
# --------------------------------------------------------------------------------------
def wps_step_from_path(poses0, vis_wps_dist=0.5):
    """src/model.py:214-215 (fp32 mean of fp32 segment norms, then int())."""
    p = np.asarray(poses0, dtype=np.float32)
    seg = p[1:] - p[:-1]
    mean = np.sqrt((seg * seg).sum(axis=1)).mean(dtype=np.float32)
    return int(np.float32(vis_wps_dist) / mean) + 1 if vis_wps_dist != 0 else 1


def length_calc(traj, dtype=np.float64):
    """src/model.py:135-139."""
    t = np.asarray(traj, dtype=dtype)
    seg = t[1:] - t[:-1]
    return np.sqrt((seg * seg).sum(axis=1)).sum(dtype=dtype)


def mean_angle_calc(traj, eps=1e-6, dtype=np.float64):
    """src/model.py:142-155."""
    t = np.asarray(traj, dtype=dtype)
    ab = t[:-2] - t[1:-1]
    ac = t[2:] - t[1:-1]
    cosang = (ab * ac).sum(axis=1) / (np.linalg.norm(ab, axis=1) * np.linalg.norm(ac, axis=1) + dtype(eps))
    return np.arccos(cosang).sum(dtype=dtype) / dtype(len(t) - 2)


def traj_minmax(points, poses, quats, K, img_width, img_height, min_dist=1.0, max_dist=5.0,
                eps=1e-6, dtype=np.float64):
    """Per evaluated pose: a_w = min_j m_jw and max_j m_jw (the two reductions of
    src/model.py:226-227; b_w = max - a_w by monotonicity of rounding)."""
    mins, maxs = [], []
    for w in range(len(poses)):
        m = visibility(points, poses[w], quats[w], K, img_width, img_height, min_dist, max_dist, eps, dtype)
        mins.append(m.min())
        maxs.append(m.max())
    return np.array(mins, dtype=dtype), np.array(maxs, dtype=dtype)


def traj_objective(points, poses, quats, K, img_width, img_height, min_dist=1.0, max_dist=5.0,
                   eps=1e-6, dtype=np.float64, n_total=None, minmax=None, want_grad=True,
                   upstream=None):
    """Visibility term of ModelTraj.forward/criterion for the poses given
    (the caller applies the `range(0, N_wps, wps_step)` selection, src/model.py:217).

    rewards = sigmoid(sum_w logit(clip((m_w - min m_w)/max(m_w - min m_w), .5, 1-eps)))
    vis     = 1 / (mean(rewards) + eps)                      (src/model.py:226-237,246)

    Returns dict(rewards, vis, mean, g_poses (W,3), g_quats (W,4), sums...).
    `minmax` / `n_total` let a caller evaluate one shard of a partitioned cloud with
    globally reduced normalisers; then g_* are this shard's additive partial gradients
    EXCEPT that the min/max-path terms need the global sums — use `traj_partials` +
    `traj_grads_from_partials` for that; this function is the single-shard convenience.
    """
    part = traj_partials(points, poses, quats, K, img_width, img_height, min_dist, max_dist,
                         eps, dtype, minmax=minmax, want_grad=want_grad, upstream=upstream)
    n = len(points) if n_total is None else n_total
    out = traj_grads_from_partials(part, poses, quats, n, eps, want_grad=want_grad,
                                   fused_upstream=upstream is None)
    out["rewards"] = part["rewards"]
    return out


def traj_partials(points, poses, quats, K, img_width, img_height, min_dist=1.0, max_dist=5.0,
                  eps=1e-6, dtype=np.float64, minmax=None, want_grad=True, upstream=None):
    """Shard-additive accumulators of the trajectory objective (what the CUDA pass B
    produces per rank before the SUM all-reduce).

    Per pose w (A.2 of SURVEY.md; ties share evenly like torch's full-reduction
    min()/max() backward; clamp backward gate is inclusive):
      main   : sum_j omega'_jw g_c, sum_j omega'_jw y g_c^T with omega' = G'_j gate/(qc(1-qc))/b
      se,sep : sum_j e'_jw, sum_j e'_jw p_jw                    (e' = G'_j gate/(qc(1-qc)))
      amax   : unweighted sums of g_c, y g_c^T over {j: m_jw - a_w == b_w}, and count
      amin   : same over {j: m_jw == a_w}, and count
    G'_j = r_j (1 - r_j) * upstream_j ; with upstream None the common factor
    c0 = -vis^2/N is applied later (it needs the global mean).
    """
    dt = np.dtype(dtype).type
    W = len(poses)
    pts = np.asarray(points, dtype=dtype)
    N = len(pts)
    if minmax is None:
        mins, maxs = traj_minmax(pts, poses, quats, K, img_width, img_height, min_dist, max_dist, eps, dtype)
    else:
        mins, maxs = (np.asarray(v, dtype=dtype) for v in minmax)
    hi = dt(np.float32(1.0 - eps)) if dtype == np.float32 else dt(1.0 - eps)
    L = np.zeros(N, dtype=dtype)
    keep = []
    for w in range(W):
        if want_grad:
            m, y, g_c, R, q, qn = visibility(pts, poses[w], quats[w], K, img_width, img_height,
                                             min_dist, max_dist, eps, dtype, want_grad=True)
        else:
            m = visibility(pts, poses[w], quats[w], K, img_width, img_height, min_dist, max_dist, eps, dtype)
        a = mins[w]
        pm = m - a
        b = dt(maxs[w] - a)
        with np.errstate(divide="ignore", invalid="ignore"):
            p = pm / b
        qc = np.clip(p, dt(0.5), hi)
        lo = np.log(qc / (dt(1) - qc))
        L = L + lo
        if want_grad:
            keep.append((m, y, g_c, p, qc, pm, a, b))
    r = dt(1) / (dt(1) + np.exp(-L))
    out = dict(rewards=r, sum_r=r.sum(dtype=np.float64), mins=mins, maxs=maxs, n=N)
    if not want_grad:
        return out
    G = (r * (dt(1) - r)).astype(np.float64)
    if upstream is not None:
        G = G * np.asarray(upstream, dtype=np.float64)
    acc = dict(main_g=np.zeros((W, 3)), main_yg=np.zeros((W, 3, 3)), se=np.zeros(W), sep=np.zeros(W),
               amax_g=np.zeros((W, 3)), amax_yg=np.zeros((W, 3, 3)), amax_n=np.zeros(W),
               amin_g=np.zeros((W, 3)), amin_yg=np.zeros((W, 3, 3)), amin_n=np.zeros(W))
    for w, (m, y, g_c, p, qc, pm, a, b) in enumerate(keep):
        y64, g64 = y.astype(np.float64), g_c.astype(np.float64)
        gate = (p >= dt(0.5)) & (p <= hi)
        qc64 = qc.astype(np.float64)
        e = np.where(gate, G / (qc64 * (1.0 - qc64)), 0.0)
        om = e / float(b)
        acc["main_g"][w] = (om[:, None] * g64).sum(axis=0)
        acc["main_yg"][w] = np.einsum("j,ja,jb->ab", om, y64, g64)
        acc["se"][w] = e.sum()
        acc["sep"][w] = (e * p.astype(np.float64)).sum()
        tmax = pm == b
        tmin = m == a
        acc["amax_g"][w] = g64[tmax].sum(axis=0)
        acc["amax_yg"][w] = np.einsum("ja,jb->ab", y64[tmax], g64[tmax])
        acc["amax_n"][w] = tmax.sum()
        acc["amin_g"][w] = g64[tmin].sum(axis=0)
        acc["amin_yg"][w] = np.einsum("ja,jb->ab", y64[tmin], g64[tmin])
        acc["amin_n"][w] = tmin.sum()
    out.update(acc)
    return out


def traj_grads_from_partials(part, poses, quats, n_total, eps=1e-6, want_grad=True, fused_upstream=True):
    """O(W) epilogue: combine (all-reduced) accumulators into mean/vis and pose gradients."""
    mean = part["sum_r"] / n_total
    vis = 1.0 / (mean + eps)
    out = dict(mean=mean, vis=vis)
    if not want_grad:
        return out
    c0 = -vis * vis / n_total if fused_upstream else 1.0
    W = len(poses)
    g_p = np.zeros((W, 3))
    g_q = np.zeros((W, 4))
    for w in range(W):
        q, qn = normalize_quat(quats[w], np.float64)
        R = rot_from_unit_quat(q)
        b = float(part["maxs"][w]) - float(part["mins"][w])
        dLdb = -part["sep"][w] / b
        dLda = -part["se"][w] / b - dLdb
        sg = part["main_g"][w].copy()
        syg = part["main_yg"][w].copy()
        if part["amax_n"][w] > 0:
            sg += dLdb / part["amax_n"][w] * part["amax_g"][w]
            syg += dLdb / part["amax_n"][w] * part["amax_yg"][w]
        if part["amin_n"][w] > 0:
            sg += dLda / part["amin_n"][w] * part["amin_g"][w]
            syg += dLda / part["amin_n"][w] * part["amin_yg"][w]
        gt, gq = pose_grads_from_sums(c0 * sg, c0 * syg, R, q, float(qn))
        g_p[w], g_q[w] = gt, gq
    out.update(g_poses=g_p, g_quats=g_q)
    return out


def traj_criterion_terms(poses, poses0, smoothness_weight=14.0, traj_length_weight=0.02, eps=1e-6,
                         dtype=np.float64):
    """l2 / smooth / length terms of ModelTraj.criterion (src/model.py:248-258)."""
    p, p0 = np.asarray(poses, dtype=dtype), np.asarray(poses0, dtype=dtype)
    l2 = np.sqrt(((p[0] - p0[0]) ** 2).sum())
    smooth = dtype(smoothness_weight) / (mean_angle_calc(p, eps, dtype) + dtype(eps))
    length = dtype(traj_length_weight) * np.abs(length_calc(p, dtype) - length_calc(p0, dtype))
    return dict(l2=l2, smooth=smooth, length=length)


# --------------------------------------------------------------------------------------
# Katz hidden-point removal (src/tools.py:38-85)
# --------------------------------------------------------------------------------------
def spherical_flip(points, param):
    """src/tools.py:38-53, fp32, every op rounded in the reference's order:
    n = ||p||; R = max(n) * 10**param; f = (2 * ((R - n) * p)) / n + p.
    torch's CPU linalg.norm(dim=1) on a contiguous (N,3) fp32 tensor equals
    sqrtf(fmaf(z,z,fmaf(y,y,x*x))) bit-for-bit; the fma chain is
    emulated exactly in float64 (a product of two fp32 is exact in fp64, and the
    double rounding cases are checked against the golden fixture)."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    x, y, z = (p[:, i].astype(np.float64) for i in range(3))
    t = (x * x).astype(np.float32).astype(np.float64)          # fl32(x*x)
    t = (y * y + t).astype(np.float32).astype(np.float64)      # fmaf(y,y,t)
    t = (z * z + t).astype(np.float32)                         # fmaf(z,z,t)
    n = np.sqrt(t, dtype=np.float32)
    radius = np.float32(n.max() * np.float32(10.0 ** param))
    tmp = (radius - n)[:, None] * p
    tmp = np.float32(2) * tmp
    with np.errstate(divide="ignore", invalid="ignore"):
        f = tmp / n[:, None]
    f = f + p
    return f.astype(np.float32), radius, n


def hidden_pts_removal(points, R_param=2):
    """src/tools.py:56-85: Qhull on flipped points + origin; `vertices[:-1]` quirk kept.
    Returns (visible_idx int64 ascending, mask float32 (N,))."""
    from scipy.spatial import ConvexHull
    f, _, _ = spherical_flip(points, R_param)
    allp = np.concatenate([f, np.zeros((1, 3), np.float32)], axis=0)
    hull = ConvexHull(allp)
    vis = np.sort(hull.vertices)[:-1].astype(np.int64)
    mask = np.zeros(len(f), dtype=np.float32)
    mask[vis] = 1
    return vis, mask


# ------------------------------------------------------------------------------------------
# PointCloud2 codec (src/pointcloud_utils.py:22-80, :180-198, :290-338)
# ------------------------------------------------------------------------------------------
PF_SIZES = {1: 1, 2: 1, 3: 2, 4: 2, 5: 4, 6: 4, 7: 4, 8: 8}           # sensor_msgs/PointField datatype -> bytes
PF_NUMPY = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 8: np.float64}


def pc2_to_xyz(data, n_points, point_step, fields, remove_nans=True):
    """`fields` = [(name, offset, datatype)] in message order.  Follows the reference: a structured dtype with one
    uint8 dummy per padding byte (pointcloud2_to_dtype :22-40), parse (:71), NaN/inf filter on x, y, z and copy into an
    (M,3) float64 array (get_xyz_points :180-195)."""
    dtype_list, offset = [], 0
    for name, off, dt in fields:
        while offset < off:
            dtype_list.append(("__%d" % offset, np.uint8))
            offset += 1
        dtype_list.append((name, PF_NUMPY[dt]))
        offset += PF_SIZES[dt]
    while offset < point_step:
        dtype_list.append(("__%d" % offset, np.uint8))
        offset += 1
    arr = np.frombuffer(bytes(data), dtype=np.dtype(dtype_list), count=n_points)
    if remove_nans:
        arr = arr[np.isfinite(arr["x"]) & np.isfinite(arr["y"]) & np.isfinite(arr["z"])]
    out = np.zeros((arr.shape[0], 3), dtype=np.float64)
    out[:, 0], out[:, 1], out[:, 2] = arr["x"], arr["y"], arr["z"]
    return out


def xyz_to_pc2(points):
    """xyz(i)_array_to_pointcloud2 (:290-338): payload bytes, point_step, is_dense."""
    pts = np.asarray(points, np.float32)
    return pts.tobytes(), 4 * pts.shape[1], int(np.isfinite(pts).all())


# ------------------------------------------------------------------------------------------
# voxel-grid filter: restatement of pcl::VoxelGrid<PointXYZ>::applyFilter (PCL 1.8-1.10; third party, not in the
# reference tree, configured by launch/voxels_filtering.launch:8-21).  PARITY UNPINNED: no PCL in this image and the
# reference holds no golden output; PCL's own output is defined only up to the fp32 summation order inside a voxel
# (std::sort is unstable).  This restatement fixes that order to the original point order.
# ------------------------------------------------------------------------------------------
def voxel_grid(points, leaf=0.1, axis=2, limit_min=-2.5, limit_max=2.5):
    p = np.asarray(points, np.float32)
    keep = np.isfinite(p).all(1)
    if axis is not None and axis >= 0:
        keep &= ~((p[:, axis] > np.float32(limit_max)) | (p[:, axis] < np.float32(limit_min)))
    q = p[keep]
    if q.shape[0] == 0:
        return np.zeros((0, 3), np.float32)
    inv = np.float32(1.0) / np.float32(leaf)
    min_b = np.floor(q.min(0) * inv).astype(np.int64)
    max_b = np.floor(q.max(0) * inv).astype(np.int64)
    div_b = max_b - min_b + 1
    if int(div_b[0]) * int(div_b[1]) * int(div_b[2]) > 2 ** 31 - 1:
        raise OverflowError("Leaf size is too small for the input dataset. Integer indices would overflow.")
    ijk = (np.floor(q * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
    order = np.argsort(idx, kind="stable")
    sidx, sq = idx[order], q[order]
    starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
    counts = np.diff(np.r_[starts, len(sidx)])
    acc = np.zeros((len(starts), 3), np.float32)
    for r in range(int(counts.max())):          # r-th member of every voxel: sequential fp32 sums, in point order
        sel = counts > r
        acc[sel] = acc[sel] + sq[starts[sel] + r]
    return acc / counts.astype(np.float32)[:, None]
