"""python tools/ik_reference_point_study.py [n] [out.json] -- the open point of the IK restatement (oracle/armsim_oracle.h):
the oracle takes the linear Jacobian at the origin of the EE link FRAME; Bullet's calculateInverseKinematics is believed to
use the link's INERTIAL frame (link 7: 2 cm up the local z axis) while its position error uses the
frame origin.  CPU only (numpy on the oracle's FK / Jacobian).  For n random servo moves from reachable poses of the
reference's workspace it runs Bullet's iteration with both Jacobians and reports what differs:
iteration counts and end points -- i.e. what a step's observation could see if Bullet does it the other way."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as O

COM = np.array([0.0, 0.0, 0.02])          # link-7 inertial origin in the link frame


def rot_err(R_t, R):
    E = R_t @ R.T
    v = 0.5 * np.array([E[2, 1] - E[1, 2], E[0, 2] - E[2, 0], E[1, 0] - E[0, 1]])
    c = 0.5 * (np.trace(E) - 1.0)
    s = np.linalg.norm(v)
    return v * (np.arctan2(s, c) / s if s > 1e-12 else 1.0)


def ik(q, tgt, R_t, com_jacobian, lam=1e-5, iters=20, res=1e-4):
    q = q.copy()
    for it in range(iters):
        p, R, _, _ = O.fk(q)
        J = O.jacobian(q).copy()
        if com_jacobian:                   # v_com = v_frame + omega x (R c)
            r = R @ COM
            for j in range(7):
                J[:3, j] += np.cross(J[3:, j], r)
        e = np.concatenate([tgt - p, rot_err(R_t, R)])
        dq = J.T @ np.linalg.solve(J @ J.T + lam * np.eye(6), e)
        m = np.abs(dq).max()
        if m > np.pi / 4:
            dq *= (np.pi / 4) / m
        q += dq
        if np.linalg.norm(tgt - O.fk(q)[0]) <= res:
            return q, it + 1
    return q, iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    out = sys.argv[2] if len(sys.argv) > 2 else None
    rng = np.random.default_rng(0)
    cfg = O.default_config(O.TASK_REACH, n_envs=1)
    q0 = np.array([cfg.init_q[i] for i in range(7)])
    quat = O.quat_from_euler([cfg.target_rpy[i] for i in range(3)])
    x, y, z, w = quat
    R_t = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                    [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                    [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    lo, hi = np.array([0.2, -0.3, 0.0]), np.array([0.7, 0.3, 0.55])
    q = O.ik(q0, O.fk(q0)[0], quat)[0]     # settle the 90-degree wrist turn of the first call
    d_end, d_q, its_a, its_b = [], [], [], []
    for k in range(n):
        p = O.fk(q)[0]
        tgt = np.clip(p + 0.02 * rng.uniform(-0.7, 0.7, 3), lo, hi)      # rl_reach_env.py:239-242
        qa, ia = ik(q, tgt, R_t, False)
        qb, ib = ik(q, tgt, R_t, True)
        d_end.append(np.linalg.norm(O.fk(qa)[0] - O.fk(qb)[0]))
        d_q.append(np.abs(qa - qb).max())
        its_a.append(ia); its_b.append(ib)
        q = qa                              # random walk along the frame-Jacobian trajectory
    d_end, d_q, its_a, its_b = map(np.array, (d_end, d_q, its_a, its_b))
    res = {"moves": n,
           "end_point_difference_m": {"median": float(np.median(d_end)), "p99": float(np.percentile(d_end, 99)), "max": float(d_end.max())},
           "joint_difference_rad": {"median": float(np.median(d_q)), "p99": float(np.percentile(d_q, 99)), "max": float(d_q.max())},
           "iterations_frame_jacobian": {str(i): int((its_a == i).sum()) for i in sorted(set(its_a))},
           "iterations_inertial_jacobian": {str(i): int((its_b == i).sum()) for i in sorted(set(its_b))},
           "same_iteration_count": float((its_a == its_b).mean()),
           "what": "Bullet's DLS iteration with the linear Jacobian taken at the EE link frame origin (the oracle) vs at the "
                   "link's inertial origin 2 cm up its z axis (believed to be Bullet's), position error at the frame origin in both"}
    print(json.dumps(res, indent=1))
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
