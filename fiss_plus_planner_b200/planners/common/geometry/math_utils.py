"""The two live helpers of the reference's math_utils (planners/common/geometry/math_utils.py:5-35)."""
import numpy as np


def mps2kph(x):
    return x * 3.6


def kph2mps(x):
    return x / 3.6


def unifyAngleRange(angle):
    """Wrap into [-pi, pi] by repeated +-2*pi steps (math_utils.py:28-34)."""
    out = angle
    while out > np.pi:
        out -= 2 * np.pi
    while out < -np.pi:
        out += 2 * np.pi
    return out
