"""Seeded synthetic workloads for the LiDAR hot path (SURVEY.md §8(d) "Synthetic inputs").

Street-canyon scene in a world frame with no surface through the origin (avoids the n·x = -1
singularity of the reference plane model, quirk Q2): ground z = -1.8, walls y = ±12,
cross walls every 40 m in x (offset so none passes through x = 0), random boxes.  Points are sampled
uniformly by area + N(0, 0.02 m) along the surface normal + U(-1e-3, 1e-3) jitter (kills kNN distance
ties).  All coordinates are float32, as the reference's pcl::PointXYZI clouds are.

Pure numpy; this is a workload generator, not part of the device path.
"""
import numpy as np

Q_LB = np.array([1.0, 0.0, 0.0, 0.0])          # GLIO/config/config_urban_hk.yaml:90-93
T_LB = np.array([0.0, 0.0, 0.28])              # GLIO/config/config_urban_hk.yaml:94-97
SEED0 = 20260923


def quat_mul(a, b):
    w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3]
    x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2]
    y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3]
    z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]
    return np.array([w, x, y, z])


def quat_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_from_rpy(roll, pitch, yaw):
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    return np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])


def quat_from_rotvec(v):
    a = np.linalg.norm(v)
    if a < 1e-300:
        return np.array([1.0, 0, 0, 0])
    return np.concatenate([[np.cos(a / 2)], np.sin(a / 2) * v / a])


def lidar_pose_in_world(t, q, q_lb=Q_LB, t_lb=T_LB):
    """Q2 = Q * q_lb^-1 ; T2 = T - Q2 * t_lb     (GLIO/src/Estimator.cpp:2216-2217)."""
    q2 = quat_mul(q, quat_conj(q_lb) / np.dot(q_lb, q_lb))
    t2 = t - quat_to_R(q2) @ t_lb
    return t2, q2


class Scene:
    """A set of rectangles: origin o, edge vectors u, v (so area = |u x v|), unit normal n."""

    def __init__(self, x_min, x_max, rng, n_boxes=30, wall_h=14.0):
        rects = []
        L = x_max - x_min
        zg = -1.8
        rects.append(([x_min, -12, zg], [L, 0, 0], [0, 24, 0]))                      # ground
        rects.append(([x_min, -12, zg], [L, 0, 0], [0, 0, wall_h]))                  # wall y=-12
        rects.append(([x_min, 12, zg], [L, 0, 0], [0, 0, wall_h]))                   # wall y=+12
        x = np.floor(x_min / 40.0) * 40.0 + 17.0
        while x < x_max:                                                             # cross walls (upper part only)
            if x > x_min:
                rects.append(([x, -12, zg + 4.0], [0, 24, 0], [0, 0, wall_h - 4.0]))
            x += 40.0
        nb = max(1, int(round(n_boxes * L / 140.0)))
        for _ in range(nb):
            sx, sy, sz = rng.uniform(1.0, 4.0, 3)
            cx = rng.uniform(x_min, x_max - sx)
            if rng.random() < 0.5:                                                   # keep the driving lane free
                cy = rng.uniform(4.0, 11.0 - sy)
            else:
                cy = rng.uniform(-11.0, -4.0 - sy)
            o = np.array([cx, cy, zg])
            rects.append((o + [0, 0, sz], [sx, 0, 0], [0, sy, 0]))                   # top
            rects.append((o, [sx, 0, 0], [0, 0, sz]))
            rects.append((o + [0, sy, 0], [sx, 0, 0], [0, 0, sz]))
            rects.append((o, [0, sy, 0], [0, 0, sz]))
            rects.append((o + [sx, 0, 0], [0, sy, 0], [0, 0, sz]))
        self.o = np.array([r[0] for r in rects], float)
        self.u = np.array([r[1] for r in rects], float)
        self.v = np.array([r[2] for r in rects], float)
        nrm = np.cross(self.u, self.v)
        self.area = np.linalg.norm(nrm, axis=1)
        self.n = nrm / self.area[:, None]

    def sample(self, n, rng, x_lo=None, x_hi=None, noise=0.02, jitter=1e-3):
        """n points, uniform by area, restricted to x in [x_lo, x_hi] (rectangles are clipped along x)."""
        o, u, v, area = self.o.copy(), self.u.copy(), self.v.copy(), self.area.copy()
        if x_lo is not None:
            # clip the x-extent of rectangles whose u edge is along x; drop others outside
            along = np.abs(u[:, 0]) > 0
            x0 = o[:, 0]; x1 = o[:, 0] + u[:, 0]
            nx0 = np.clip(x0, x_lo, x_hi); nx1 = np.clip(x1, x_lo, x_hi)
            keep = np.where(along, nx1 > nx0, (x0 >= x_lo) & (x0 <= x_hi))
            frac = np.where(along, (nx1 - nx0) / np.where(along, u[:, 0], 1.0), 1.0)
            o[:, 0] = np.where(along, nx0, o[:, 0])
            u[:, 0] = np.where(along, nx1 - nx0, u[:, 0])
            area = area * frac * keep
        p = area / area.sum()
        ridx = rng.choice(len(area), size=n, p=p)
        a = rng.random(n)[:, None]; b = rng.random(n)[:, None]
        pts = o[ridx] + a * u[ridx] + b * v[ridx]
        pts += self.n[ridx] * rng.normal(0.0, noise, n)[:, None]
        pts += rng.uniform(-jitter, jitter, (n, 3))
        return pts


def trajectory(K, rng):
    """KF k at x = 1.0 k, y = 0.5 sin(0.1 k), z = 0, yaw = 0.02 sin(0.05 k), roll/pitch ~ N(0, 0.2 deg)."""
    poses = np.zeros((K, 7))
    for k in range(K):
        poses[k, :3] = [1.0 * k, 0.5 * np.sin(0.1 * k), 0.0]
        r, p = rng.normal(0.0, np.deg2rad(0.2), 2)
        poses[k, 3:] = quat_from_rpy(r, p, 0.02 * np.sin(0.05 * k))
    return poses


def perturb(poses, rng, sig_t=0.05, sig_r_deg=0.5):
    """truth (+) N(0, 0.05 m), N(0, 0.5 deg)."""
    out = poses.copy()
    for k in range(len(poses)):
        out[k, :3] += rng.normal(0.0, sig_t, 3)
        dq = quat_from_rotvec(rng.normal(0.0, np.deg2rad(sig_r_deg), 3))
        q = quat_mul(dq, poses[k, 3:])
        out[k, 3:] = q / np.linalg.norm(q)
    return out


def scan_in_lidar_frame(scene, pose, Q, rng, rng_range=50.0, q_lb=Q_LB, t_lb=T_LB):
    t, q = pose[:3], pose[3:]
    pw = scene.sample(Q, rng, t[0] - rng_range, t[0] + rng_range)
    pb = (pw - t) @ quat_to_R(q)                 # R^T (pw - t)
    pl = pb @ quat_to_R(q_lb).T + t_lb           # p_l = q_lb p_b + t_lb
    return pl.astype(np.float32)


def scan_in_body_frame(scene, pose, Q, rng, rng_range=50.0):
    """Batch path applies body poses directly to the stored scan points (quirk Q7)."""
    t, q = pose[:3], pose[3:]
    pw = scene.sample(Q, rng, t[0] - rng_range, t[0] + rng_range)
    return ((pw - t) @ quat_to_R(q)).astype(np.float32)


def window_problem(W=20, Q=100_000, M=1_000_000, seed=SEED0 + 2, n_boxes=30):
    """cfg 1 / cfg 2 / cfg 5 style sliding-window problem."""
    rng = np.random.default_rng(seed)
    scene = Scene(-60.0, W + 60.0, rng, n_boxes=n_boxes)
    truth = trajectory(W, rng)
    init = perturb(truth, rng)
    map_xyz = scene.sample(M, rng, -55.0, W + 55.0).astype(np.float32)
    scans = [scan_in_lidar_frame(scene, truth[k], Q, rng) for k in range(W)]
    return dict(map_xyz=map_xyz, scans=scans, poses_true=truth, poses_init=init, q_lb=Q_LB.copy(),
                t_lb=T_LB.copy(), W=W, Q=Q, M=M, seed=seed)


def batch_problem(K=200, Q=100_000, seed=SEED0 + 3, search_range=6, rng_range=50.0, frames=None):
    """cfg 3 / cfg 4 style batch problem: K keyframes, scans in the body frame.  Every scan has its own RNG stream
    (seed, k), so a rank of a sharded run can generate only the frames it holds (`frames`) and still agree with the
    others; scans[k] is None for frames not requested."""
    rng = np.random.default_rng(seed)
    scene = Scene(-60.0, K + 60.0, rng)
    truth = trajectory(K, rng)
    init = perturb(truth, rng, sig_t=0.03, sig_r_deg=0.2)
    want = set(range(K)) if frames is None else set(int(f) for f in frames)
    scans = [scan_in_body_frame(scene, truth[k], Q, np.random.default_rng([seed, k]), rng_range) if k in want else None for k in range(K)]
    return dict(scans=scans, poses_true=truth, poses_init=init, K=K, Q=Q, search_range=search_range, seed=seed)


def ring_scan(n_rings=32, n_az=1500, seed=SEED0 + 11, fov=(-25.0, 15.0), noise=0.01, drop=0.02):
    """A spinning-LiDAR sweep of the street-canyon scene, ring by ring, for the front end's feature extraction:
    rays are cast from the sensor against the scene's rectangles (nearest hit), range noise N(0, noise), a fraction `drop` of the
    returns missing.  Returns (cloud (n,4) float32: x,y,z,intensity = ring + 0.1 relTime; scan_start, scan_end int32 arrays) laid
    out as Preprocessing::cloudHandler lays out `laserCloud` (GLIO/src/Preprocessing.cpp:529-534): rings concatenated,
    scanStartInd = first index + 5, scanEndInd = last index - 5."""
    rng = np.random.default_rng(seed)
    sc = Scene(-60.0, 60.0, rng)
    origin = np.array([0.3, -0.4, 0.2])
    elev = np.deg2rad(np.linspace(fov[0], fov[1], n_rings))
    clouds, start, end = [], [], []
    total = 0
    for r in range(n_rings):
        az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False) + rng.uniform(0, 1e-3)
        d = np.stack([np.cos(elev[r]) * np.cos(az), np.cos(elev[r]) * np.sin(az), np.full(n_az, np.sin(elev[r]))], 1)
        best = np.full(n_az, np.inf)
        for o, u, v, nrm in zip(sc.o, sc.u, sc.v, sc.n):
            den = d @ nrm
            with np.errstate(divide="ignore", invalid="ignore"):
                t = ((o - origin) @ nrm) / den
            hit = origin + t[:, None] * d - o
            a = (hit @ u) / (u @ u); b = (hit @ v) / (v @ v)
            ok = (np.abs(den) > 1e-9) & (t > 0.5) & (a >= 0) & (a <= 1) & (b >= 0) & (b <= 1)
            best = np.where(ok & (t < best), t, best)
        keep = np.isfinite(best) & (best < 80.0) & (rng.random(n_az) > drop)
        rg = best[keep] + rng.normal(0.0, noise, int(keep.sum()))
        pts = origin + rg[:, None] * d[keep]
        rel = az[keep] / (2 * np.pi)
        cl = np.concatenate([pts, (r + 0.1 * rel)[:, None]], 1).astype(np.float32)
        clouds.append(cl)
        start.append(total + 5); total += len(cl); end.append(total - 6)
    return np.concatenate(clouds), np.asarray(start, np.int32), np.asarray(end, np.int32)
