"""Generate OpenCV golden vectors for the 6x6 LM step (src/mapOptmization.cpp:1237-1271).

Run in the build container (python cv2 4.13, the only OpenCV here):  python tests/golden/make_cv2_golden.py
Writes tests/golden/cv2_lm6.npz.  The reference itself pins no OpenCV version (CMakeLists.txt:27), so these vectors
pin the oracle to OpenCV 4.13's built-in small-matrix paths (hal::QR32f, JacobiImpl_<float>, hal::LU32f, gemm).
"""
import os
import numpy as np
import cv2

rng = np.random.default_rng(20240517)
cases = []
for c in range(64):
    n = int(rng.integers(60, 4000))
    # rows shaped like the LM Jacobian: 3 rotational columns (|.| up to ~50: lever arm x normal) + unit-ish normal columns
    normals = rng.normal(size=(n, 3)); normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    if c % 4 == 1:   # corridor-like: normals concentrated on two axes → near-degenerate translation along x
        normals[:, 0] *= 1e-3
    if c % 4 == 2:   # ground only → strongly degenerate
        normals = np.tile(np.array([[0.0, 0.0, 1.0]]), (n, 1)) + rng.normal(scale=1e-3, size=(n, 3))
    pts = rng.uniform(-60, 60, size=(n, 3)) * np.array([1, 1, 0.1])
    s = rng.uniform(0.2, 1.0, size=(n, 1))
    A = np.concatenate([np.cross(pts, normals) * s, normals * s], axis=1).astype(np.float32)
    B = (rng.normal(scale=0.05, size=(n, 1)) * s).astype(np.float32)
    At = cv2.transpose(A)
    AtA = At @ A if False else cv2.gemm(At, A, 1.0, None, 0.0)
    AtB = cv2.gemm(At, B, 1.0, None, 0.0)
    ok, X = cv2.solve(AtA, AtB, flags=cv2.DECOMP_QR)
    okE, E, V = cv2.eigen(AtA)
    V2 = V.copy(); deg = False
    for i in range(5, -1, -1):
        if E[i, 0] < 100:
            V2[i, :] = 0; deg = True
        else:
            break
    retinv, Vinv = cv2.invert(V, flags=cv2.DECOMP_LU)
    P = cv2.gemm(Vinv, V2, 1.0, None, 0.0)
    X2 = cv2.gemm(P, X, 1.0, None, 0.0)
    cases.append(dict(A=A, B=B, AtA=AtA, AtB=AtB, X=X, ok=ok, E=E, V=V, Vinv=Vinv, P=P, X2=X2, deg=deg))

out = {}
for i, c in enumerate(cases):
    for k, v in c.items():
        if k in ("A", "B") and i >= 8:      # keep the fixture small: raw rows only for the first 8 cases
            continue
        out[f"{k}_{i}"] = np.asarray(v)
out["n_cases"] = np.array(len(cases))
p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv2_lm6.npz")
np.savez_compressed(p, **out)
print("wrote", p, os.path.getsize(p), "bytes; cv2", cv2.__version__)
