"""CPU: how far can Eigen 3.2.5's float `.log()` (src/Frame.cpp:521,553) be from the exact SE(3) logarithm the oracle, the stand-in
header and the product use instead?

Eigen is not in this image, so its code cannot be run or restated operation for operation.  Its PUBLISHED algorithm can:
unsupported/Eigen/src/MatrixFunctions/MatrixLogarithm.h (3.2.x) takes the complex Schur form A = U T U^H and evaluates log(T) by
inverse scaling and squaring -- square roots of the triangular factor until ||T - I||_1 is below the Pade threshold of the scalar
type (float: degree <= 5, 0.5315), one optional extra root when that lowers the degree, then the [m/m] Pade approximant of
log(I + X) in its Gauss-Legendre partial-fraction form, scaled back by 2^s (Higham, "Evaluating Pade approximants of the matrix
logarithm", 2001) -- and transforms back.  This test runs exactly that in SINGLE precision (complex64 throughout) on poses of the
tracker's range and measures the distance to the exact logarithm evaluated in double and rounded to float:

    * the float algorithm lands within a few float ulp of the exact logarithm (measured 3.3e-7 .. 5.1e-7 absolute for pose
      magnitudes 1e-3 .. 0.5 per component), while
    * the oracle's closed-form logarithm is the exact one to the final rounding (<= 1.2e-7),

i.e. replacing `.log()` by the exact logarithm moves a pose by no more than the rounding noise of Eigen's own evaluation.  (For
rotations above 0.1 rad Eigen splits T into eigenvalue clusters and joins the blocks with the Parlett recurrence; one block over
the whole of T, as here, is the same approximant with a different blocking.)"""
import numpy as np
import pytest

scipy_linalg = pytest.importorskip("scipy.linalg")

C = np.complex64
F = np.float32

# maximal ||X||_1 for the degree-m Pade approximant in single precision (MatrixLogarithm.h, getPadeDegree(float))
PADE_MAX_NORM_F32 = {3: 2.5111573934555054e-1, 4: 4.0535837411880493e-1, 5: 5.3149729967117310e-1}


def sqrt_triu_c64(T):
    """Square root of an upper triangular matrix by the column recurrence (Bjorck & Hammarling), complex64 arithmetic."""
    n = T.shape[0]
    R = np.zeros((n, n), C)
    for j in range(n):
        R[j, j] = np.sqrt(C(T[j, j]))
        for i in range(j - 1, -1, -1):
            s = C(0)
            for k in range(i + 1, j):
                s = C(s + C(R[i, k] * R[k, j]))
            R[i, j] = C(C(T[i, j] - s) / C(R[i, i] + R[j, j]))
    return R


def pade_log_c64(X, m):
    """[m/m] Pade approximant of log(I + X): sum_k w_k (I + x_k X)^-1 X with the m-point Gauss-Legendre rule on [0, 1]."""
    nodes, weights = np.polynomial.legendre.leggauss(m)
    nodes, weights = (nodes + 1.0) / 2.0, weights / 2.0
    n = X.shape[0]
    acc = np.zeros((n, n), C)
    eye = np.eye(n, dtype=C)
    for x, w in zip(nodes, weights):
        M = (eye + C(x) * X).astype(C)
        acc = (acc + C(w) * scipy_linalg.solve_triangular(M, X, lower=False).astype(C)).astype(C)
    return acc


def logm_iss_c64(A):
    """Eigen 3.2's MatrixLogarithm on a real 4x4 matrix, single precision: complex Schur, inverse scaling and squaring, Pade."""
    T, U = scipy_linalg.schur(A.astype(C), output="complex")
    T, U = np.triu(T).astype(C), U.astype(C)
    eye = np.eye(4, dtype=C)
    roots, extra = 0, 0
    while True:
        norm = float(np.abs(T - eye).sum(axis=0).max())
        if norm < PADE_MAX_NORM_F32[5]:
            deg = min(m for m in (3, 4, 5) if norm <= PADE_MAX_NORM_F32[m])
            deg2 = min(m for m in (3, 4, 5) if norm / 2 <= PADE_MAX_NORM_F32[m])
            if deg - deg2 <= 1 or extra == 1:
                break
            extra += 1
        T = sqrt_triu_c64(T)
        roots += 1
        assert roots < 40
    L = pade_log_c64((T - eye).astype(C), deg) * C(2.0 ** roots)
    return (U @ L.astype(C) @ U.conj().T).astype(C).real.astype(F)


def pose_of(L):
    return np.array([L[2, 1], L[0, 2], L[1, 0], L[0, 3], L[1, 3], L[2, 3]], np.float64)


def exact_log_f32(T):
    """The exact logarithm in double, rounded once to float: what oracle / stand-in / product compute."""
    return pose_of(np.real(scipy_linalg.logm(T.astype(np.float64)))).astype(F)


def test_float_inverse_scaling_and_squaring_log_stays_within_float_rounding_of_the_exact_log(oracle_mod):
    rng = np.random.default_rng(2024)
    worst_alg, worst_oracle = 0.0, 0.0
    for scale in (1e-3, 1e-2, 0.05, 0.2, 0.5):                  # rotation / translation magnitudes: sub-frame motion .. loop closures
        for _ in range(40):
            pose = (rng.standard_normal(6) * scale).astype(F)
            T = oracle_mod.se3_exp(pose).astype(F).reshape(4, 4)   # exp(hat(pose)) as the tracker builds it (float Pade)
            exact = exact_log_f32(T)
            got = pose_of(logm_iss_c64(T)).astype(F)
            worst_alg = max(worst_alg, float(np.abs(got.astype(np.float64) - exact).max()))
            orc = np.asarray(oracle_mod.se3_log(T.reshape(-1)), F)
            worst_oracle = max(worst_oracle, float(np.abs(orc.astype(np.float64) - exact).max()))
    # the oracle's closed-form double logarithm IS the exact logarithm up to the final rounding ...
    assert worst_oracle <= 2.4e-7, worst_oracle
    # ... and Eigen's algorithm run in float stays within a few float ulp of it (|pose| up to ~1.6: ulp 1.2e-7)
    assert worst_alg <= 1.5e-6, worst_alg
