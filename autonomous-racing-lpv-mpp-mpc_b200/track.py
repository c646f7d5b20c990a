"""Track geometry table used by the scheduling variable kappa (support row of SURVEY.md 2 / 8a-a13).

Host-side mirror of the reference's ``Map`` constructor (Utilities/trackInitialization.py:13-202) and of
``Curvature`` (Utilities/utilities.py:31-50): it only produces the ``PointAndTangent`` table
``[x, y, psi, s_start, length, kappa]`` that the device kernels consume.  Any object with
``.PointAndTangent`` / ``.halfWidth`` (e.g. the reference's own Map) is accepted by the drop-in classes.
"""
import math

import numpy as np

# (length, signed radius) per segment; radius 0 = straight.  trackInitialization.py:28-81
_PI = np.pi
TRACK_SPECS = {
    "3110": [(60 * 0.03, 0), (80 * 0.03, +80 * 0.03 * 2 / _PI), (20 * 0.03, 0), (80 * 0.03, +80 * 0.03 * 2 / _PI),
             (40 * 0.03, -40 * 0.03 * 10 / _PI), (60 * 0.03, +60 * 0.03 * 5 / _PI), (40 * 0.03, -40 * 0.03 * 10 / _PI),
             (80 * 0.03, +80 * 0.03 * 2 / _PI), (20 * 0.03, 0), (80 * 0.03, +80 * 0.03 * 2 / _PI), (80 * 0.03, 0)],
    "oval": [(1.0, 0), (4.5, 4.5 / _PI), (2.0, 0), (4.5, 4.5 / _PI), (1.0, 0)],
    "L_shape": [(1.0, 0), (4.5, 4.5 / _PI), (4.5 / 2, -4.5 / _PI), (4.5, 4.5 / _PI), (4.5 / _PI * 2, 0),
                (4.5 / 2, 4.5 / _PI)],
    "Euge_Track": [(30 * 0.03, +30 * 0.03 * 2 / _PI), (20 * 0.03, -0), (30 * 0.03, -30 * 0.03 * 2 / _PI),
                   (30 * 0.03, +30 * 0.03 * 2 / _PI), (30 * 0.03, +30 * 0.03 * 2 / _PI), (130 * 0.03, 0),
                   (30 * 0.03, +30 * 0.03 * 2 / _PI), (10 * 0.03, -0), (30 * 0.03, +30 * 0.03 * 2 / _PI),
                   (55 * 0.03, -0), (30 * 0.03, -30 * 0.03 * 2 / _PI), (10 * 0.03, -0),
                   (30 * 0.03, +30 * 0.03 * 2 / _PI)],
}
_FIXED_WIDTH = {"3110": (0.6, 0.15), "Euge_Track": (0.4, 0.15)}
_SLACK = {"oval": 0.15, "L_shape": 0.45}


def _wrap(a):
    if a < -np.pi:
        return 2 * np.pi + a
    if a > np.pi:
        return a - 2 * np.pi
    return a


def _sgn(a):
    return 1 if a >= 0 else -1


class Map(object):
    """``Map(track_shape, half_width)``: ``halfWidth`` = launch halfWidth + 0.1 for the oval / L_shape tracks
    (trackInitialization.py:20), fixed for the other two."""

    def __init__(self, track_shape="L_shape", half_width=0.2):
        if track_shape not in TRACK_SPECS:
            raise ValueError("unknown track %r" % (track_shape,))
        if track_shape in _FIXED_WIDTH:
            self.halfWidth, self.slack = _FIXED_WIDTH[track_shape]
        else:
            self.halfWidth, self.slack = half_width + 0.1, _SLACK[track_shape]
        spec = np.array(TRACK_SPECS[track_shape], dtype=np.float64)
        nseg = spec.shape[0]
        pt = np.zeros((nseg + 1, 6))
        for i in range(nseg):
            length, rad = spec[i]
            x_prev, y_prev, ang = (0.0, 0.0, 0.0) if i == 0 else (pt[i - 1, 0], pt[i - 1, 1], pt[i - 1, 2])
            s_start = pt[i, 3] if i == 0 else pt[i - 1, 3] + pt[i - 1, 4]
            if rad == 0.0:
                x = x_prev + length * np.cos(ang)
                y = y_prev + length * np.sin(ang)
                pt[i] = [x, y, ang, s_start, length, 0]
            else:
                direction = 1 if rad >= 0 else -1
                cx = x_prev + np.abs(rad) * np.cos(ang + direction * np.pi / 2)
                cy = y_prev + np.abs(rad) * np.sin(ang + direction * np.pi / 2)
                span = length / np.abs(rad)
                psi = _wrap(ang + span * np.sign(rad))
                normal = _wrap(direction * np.pi / 2 + ang)
                a0 = -(np.pi - np.abs(normal)) * _sgn(normal)
                x = cx + np.abs(rad) * np.cos(a0 + direction * span)
                y = cy + np.abs(rad) * np.sin(a0 + direction * span)
                pt[i] = [x, y, psi, s_start, length, 1 / rad]
        # closing straight back to the origin (trackInitialization.py:188-199)
        xs, ys = pt[-2, 0], pt[-2, 1]
        pt[-1] = [0, 0, 0, pt[-2, 3] + pt[-2, 4], np.sqrt((0 - xs) ** 2 + (0 - ys) ** 2), 0]
        self.PointAndTangent = pt
        self.TrackLength = pt[-1, 3] + pt[-1, 4]


def curvature(s, point_and_tangent):
    """Host Curvature(s) with the reference's failure behaviour (raises when no unique segment holds s)."""
    pt = point_and_tangent
    track_len = pt[-1, 3] + pt[-1, 4]
    if not (s == s) or math.isinf(s):
        raise TypeError("only length-1 arrays can be converted to Python scalars")
    while s > track_len:
        s = s - track_len
    hit = np.nonzero((s >= pt[:, 3]) & (s < pt[:, 3] + pt[:, 4]))[0]
    if hit.size != 1:
        raise TypeError("only length-1 arrays can be converted to Python scalars")
    return pt[int(hit[0]), 5]
