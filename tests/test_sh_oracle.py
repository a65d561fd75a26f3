"""CPU: the numpy SH restatement (oracle/oracle_np.py::sh_eval_np) is a correct real-SH basis.
There is no reference implementation to pin it to (render.py:82-87 is a placeholder), so it is pinned to the
mathematics: the 16 basis functions are orthonormal on the sphere, band 0 is the constant 1/(2 sqrt(pi)),
band 1 is linear in the direction."""
import numpy as np

from oracle.oracle_np import sh_eval_np


def fibonacci_sphere(n):
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    theta = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], 1)


def basis(dirs, degree=3):
    n = dirs.shape[0]
    K = (degree + 1) ** 2
    Y = np.zeros((n, K))
    for k in range(K):
        c = np.zeros((n, 16, 3)); c[:, k, :] = 0.1
        Y[:, k] = (sh_eval_np(degree, c, dirs, np.zeros(3))[:, 0] - 0.5) / 0.1
    return Y


def test_sh_basis_is_orthonormal():
    d = fibonacci_sphere(40000)
    Y = basis(d)
    G = (Y.T @ Y) * (4 * np.pi / d.shape[0])
    assert np.abs(G - np.eye(16)).max() < 2e-3


def test_sh_low_bands_and_clamp():
    d = fibonacci_sphere(1000)
    Y = basis(d)
    assert np.allclose(Y[:, 0], 0.5 / np.sqrt(np.pi))
    assert np.allclose(Y[:, 1], -0.4886025119029199 * d[:, 1])
    assert np.allclose(Y[:, 2], 0.4886025119029199 * d[:, 2])
    assert np.allclose(Y[:, 3], -0.4886025119029199 * d[:, 0])
    # direction = normalize(mean - campos); degree selects the bands; negative colours clamp at 0
    c = np.zeros((1, 16, 3)); c[0, 0] = [-5.0, 0.0, 1.0]; c[0, 9] = 7.0
    out = sh_eval_np(0, c, np.array([[3.0, 0, 0]]), np.array([1.0, 0, 0]))
    assert np.allclose(out, [[0.0, 0.5, 0.5 + 0.28209479177387814]])
    a = sh_eval_np(2, c, np.array([[3.0, 2.0, 1.0]]), np.zeros(3))
    b = sh_eval_np(3, c, np.array([[3.0, 2.0, 1.0]]), np.zeros(3))
    assert not np.allclose(a, b)
    assert np.allclose(sh_eval_np(3, c, np.array([[6.0, 4.0, 2.0]]), np.zeros(3)), b)  # scale invariant
