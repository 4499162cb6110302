"""Host-side camera helpers (reference: keras_nerf/data/utils.py).  Scalar / 4x4 fp32 math, not a kernel."""
import numpy as np

_f = np.float32


def get_focal_from_fov(field_of_view: float, width: int) -> float:
    """0.5 * width / tan(0.5 * fov) in fp32 (keras_nerf/data/utils.py:5-16)."""
    return float(_f(0.5) * _f(width) / np.tan(_f(0.5 * float(field_of_view))))


def get_translation_t(t):
    """keras_nerf/data/utils.py:19-27"""
    m = np.eye(4, dtype=_f)
    m[2, 3] = _f(t)
    return m


def get_rotation_phi(phi):
    """keras_nerf/data/utils.py:30-38"""
    c, s = np.cos(_f(phi)), np.sin(_f(phi))
    m = np.eye(4, dtype=_f)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    return m


def get_rotation_theta(theta):
    """keras_nerf/data/utils.py:41-49"""
    c, s = np.cos(_f(theta)), np.sin(_f(theta))
    m = np.eye(4, dtype=_f)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
    return m


def pose_spherical(theta, phi, t) -> np.ndarray:
    """camera-to-world for (theta deg, phi deg, radius t) (keras_nerf/data/utils.py:52-63); fp32 [4,4]."""
    c2w = get_translation_t(t)
    c2w = get_rotation_phi(phi / 180.0 * np.pi) @ c2w
    c2w = get_rotation_theta(theta / 180.0 * np.pi) @ c2w
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=_f)
    return (flip @ c2w).astype(_f)
