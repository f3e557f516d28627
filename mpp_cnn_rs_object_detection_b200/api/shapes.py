"""Object types of the marked point process: base/shapes/base_shapes.py:11-35 (Point) and
base/shapes/rectangle.py:12-126 (Rectangle, polygon helpers).  Identity semantics are the reference's: two objects
with equal fields are distinct (hash = id, base_shapes.py:16-17)."""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


class Point:
    """Integer pixel centre, x = row, y = column (base_shapes.py:11-20)."""
    __slots__ = ("x", "y", "__weakref__")

    def __init__(self, x: int, y: int):
        self.x = x
        self.y = y

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def get_coord(self) -> np.ndarray:
        return np.array([self.x, self.y])

    def __copy__(self):
        return Point(self.x, self.y)

    def __repr__(self):
        return f"Point(x={self.x}, y={self.y})"


class Rectangle(Point):
    """Oriented rectangle with marks (size, ratio, angle) (rectangle.py:12-37)."""
    __slots__ = ("size", "ratio", "angle")
    PARAMETERS = ["size", "ratio", "angle"]

    def __init__(self, x: int, y: int, size: float, ratio: float, angle: float):
        super().__init__(x, y)
        self.size = size
        self.ratio = ratio
        self.angle = angle

    @property
    def length(self) -> float:  # rectangle.py:20-21
        return (2 * self.size) / (1 + self.ratio)

    @property
    def width(self) -> float:  # rectangle.py:24-25
        return self.ratio * self.length

    @property
    def poly_coord(self) -> np.ndarray:  # rectangle.py:28-30
        return rect_to_poly((self.x, self.y), short=self.length, long=self.width, angle=self.angle + np.pi / 2)

    def __copy__(self):
        return Rectangle(self.x, self.y, self.size, self.ratio, self.angle)

    def __repr__(self):
        return f"Rectangle(x={self.x}, y={self.y}, size={self.size}, ratio={self.ratio}, angle={self.angle})"


def rotation_matrix(alpha) -> np.ndarray:  # rectangle.py:64-66
    c, s = math.cos(alpha), math.sin(alpha)
    return np.array([[c, -s], [s, c]])


def rect_to_poly(center, short: float, long: float, angle: float, dilation: int = 0) -> np.ndarray:
    """(4,2) corner coordinates: local [+-short/2, +-long/2] rotated by `angle`, plus the centre (rectangle.py:69-100)."""
    hs, hl = short / 2 + dilation, long / 2 + dilation
    local = np.array([[hs, hl], [hs, -hl], [-hs, -hl], [-hs, hl]])
    return local @ rotation_matrix(angle).T + np.asarray(center)


def wla_to_sra(a, b, angle):  # rectangle.py:103-104
    return (a + b) / 2, a / b, angle


def sra_to_wla(s, r, angle):  # rectangle.py:107-109
    b = (2 * s) / (1 + r)
    return b * r, b, angle


def polygon_to_abw(poly: np.ndarray) -> Tuple[float, float, float]:
    """(a, b, angle) of a 4-corner polygon, a <= b, angle of the long axis in [0, pi) (rectangle.py:112-126)."""
    assert poly.shape == (4, 2)
    e01 = 0.5 * (np.linalg.norm(poly[0] - poly[1]) + np.linalg.norm(poly[2] - poly[3]))
    e12 = 0.5 * (np.linalg.norm(poly[1] - poly[2]) + np.linalg.norm(poly[3] - poly[0]))
    if e01 < e12:
        a, b = e01, e12
        axis = 0.5 * (poly[2] + poly[1]) - 0.5 * (poly[0] + poly[3])
    else:
        a, b = e12, e01
        axis = 0.5 * (poly[1] + poly[0]) - 0.5 * (poly[3] + poly[2])
    return a, b, math.atan2(axis[1], axis[0]) % math.pi
