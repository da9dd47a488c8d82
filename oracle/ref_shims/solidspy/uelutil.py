"""Import shim (test infrastructure): solidspy.uelutil.elast_quad4(coord, params) -> (kloc, mloc),
used at reference residuals_mechanics_K.py:99-103 (init time only).  4-node bilinear quad, plane
stress, 2x2 Gauss, nodes counter-clockwise, dof order (u1x,u1y,...,u4y).  solidspy is not installed
here: parity is pinned by the closed-form Q4 stiffness known-answer test (tests/test_oracle_*.py)."""
import numpy as np


def elast_quad4(coord, params):
    E, nu = float(params[0]), float(params[1])
    C = E / (1.0 - nu ** 2) * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, 0.5 * (1.0 - nu)]])
    gp = 1.0 / np.sqrt(3.0)
    kloc = np.zeros((8, 8))
    mloc = np.zeros((8, 8))
    for r in (-gp, gp):
        for s in (-gp, gp):
            dN = 0.25 * np.array([[-(1 - s), (1 - s), (1 + s), -(1 + s)],
                                  [-(1 - r), -(1 + r), (1 + r), (1 - r)]])
            N = 0.25 * np.array([(1 - r) * (1 - s), (1 + r) * (1 - s), (1 + r) * (1 + s), (1 - r) * (1 + s)])
            J = dN @ coord
            det = np.linalg.det(J)
            dNx = np.linalg.solve(J, dN)
            B = np.zeros((3, 8))
            B[0, 0::2] = dNx[0]
            B[1, 1::2] = dNx[1]
            B[2, 0::2] = dNx[1]
            B[2, 1::2] = dNx[0]
            kloc += det * (B.T @ C @ B)
            Nm = np.zeros((2, 8))
            Nm[0, 0::2] = N
            Nm[1, 1::2] = N
            mloc += det * (Nm.T @ Nm)
    return kloc, mloc
