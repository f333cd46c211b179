"""CPU oracle of the loop-closure ICP (TEST INFRASTRUCTURE, like everything under oracle/).

Restates pcl::IterativeClosestPoint as mapOptimization::performSCLoopClosure configures it
(/root/reference/src/mapOptmization.cpp:658-674): max correspondence distance 2 * historyKeyframeSearchRadius, 100 iterations,
transformation / Euclidean fitness epsilons 1e-6, no RANSAC.  PCL is not in this container, so this follows PCL's
published algorithm (registration/impl/icp.hpp computeTransformation, default_convergence_criteria.hpp, Eigen::umeyama without
scaling, Registration::getFitnessScore) — parity unpinned, same status as the VoxelGrid restatement.
Deliberately a different implementation from the product: numpy brute-force neighbours and numpy.linalg.svd instead of a voxel
grid and a Jacobi eigen-solver, so that agreement is a cross-check and not a tautology."""
import numpy as np

F = np.float32


def _nearest(src, tgt, chunk=2048):
    """exact nearest neighbour of every src point in tgt, FLANN L2_Simple arithmetic (((dx*dx) + dy*dy) + dz*dz in fp32), ties
    to the lower target index."""
    idx = np.empty(len(src), np.int64); d2 = np.empty(len(src), F)
    tx, ty, tz = tgt[:, 0][None, :], tgt[:, 1][None, :], tgt[:, 2][None, :]
    for s in range(0, len(src), chunk):
        q = src[s:s + chunk]
        dx = q[:, 0:1] - tx; dy = q[:, 1:2] - ty; dz = q[:, 2:3] - tz
        d = dx * dx; d += dy * dy; d += dz * dz
        j = np.argmin(d, axis=1)
        idx[s:s + chunk] = j; d2[s:s + chunk] = d[np.arange(len(q)), j]
    return idx, d2


def _apply(T, pts):
    """pcl::transformPointCloud with a float Matrix4f: ((t0 x + t1 y) + t2 z) + t3 in fp32."""
    T = T.astype(F)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    out = pts.copy()
    for r in range(3):
        out[:, r] = ((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3]
    return out


def _umeyama(src, dst):
    """Eigen::umeyama(src, dst, with_scaling=false) on the corresponded points, in double, rounded to a float Matrix4f."""
    s = src[:, :3].astype(np.float64); d = dst[:, :3].astype(np.float64)
    ms, md = s.mean(0), d.mean(0)
    sigma = (d - md).T @ (s - ms) / len(s)
    U, S, Vt = np.linalg.svd(sigma)
    D = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2, 2] = -1
    R = U @ D @ Vt
    T = np.eye(4)
    T[:3, :3] = R; T[:3, 3] = md - R @ ms
    return T.astype(F)


def icp(source, target, max_corr_dist, max_iters=100, transformation_epsilon=1e-6, euclidean_fitness_epsilon=1e-6):
    source = np.ascontiguousarray(source, F); target = np.ascontiguousarray(target, F)
    cur = source.copy()
    final = np.eye(4, dtype=F)
    max_d2 = F(max_corr_dist) * F(max_corr_dist)
    prev_mse = np.finfo(np.float64).max
    iterations = 0; converged = False; state = "not_converged"
    while True:
        idx, d2 = _nearest(cur, target)
        keep = d2 <= max_d2
        if keep.sum() < 3:
            state = "no_correspondences"; break
        T = _umeyama(cur[keep], target[idx[keep]])
        cur = _apply(T, cur)
        final = (T @ final).astype(F)
        iterations += 1
        cur_mse = float(d2[keep].astype(np.float64).sum() / keep.sum())
        if iterations >= max_iters:
            converged, state = True, "iterations"; break
        cos_angle = 0.5 * (float(T[0, 0]) + float(T[1, 1]) + float(T[2, 2]) - 1.0)
        tr2 = float(T[0, 3]) ** 2 + float(T[1, 3]) ** 2 + float(T[2, 3]) ** 2
        if cos_angle >= 1.0 - transformation_epsilon and tr2 <= transformation_epsilon:
            converged, state = True, "transform"; break
        if abs(cur_mse - prev_mse) < 1e-12:
            converged, state = True, "abs_mse"; break
        if abs(cur_mse - prev_mse) / prev_mse < euclidean_fitness_epsilon:
            converged, state = True, "rel_mse"; break
        prev_mse = cur_mse
    _, d2 = _nearest(_apply(final, source), target)
    fitness = float(d2.astype(np.float64).mean())
    return dict(converged=converged, state=state, iterations=iterations, transform=final, fitness=fitness)


def loop_find_near_keyframes(kf_clouds, kf_poses, key, search_num, loop_index, leaf, oracle):
    """loopFindNearKeyframes (:821-844): clouds key-search_num..key+search_num, each transformed by pose[loop_index] (or its own
    pose when loop_index == -1), concatenated, VoxelGrid(leaf)."""
    parts = []
    for i in range(-search_num, search_num + 1):
        kn = key + i
        if kn < 0 or kn >= len(kf_clouds):
            continue
        parts.append(oracle.transform_cloud(kf_clouds[kn], kf_poses[loop_index if loop_index != -1 else kn]))
    if not parts:
        return np.zeros((0, 4), F)
    return oracle.voxel_grid(np.concatenate(parts, 0), leaf)[0]
