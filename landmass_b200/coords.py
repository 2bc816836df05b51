"""Coordinate systems and point-sample distances.

Host-side mirror of the reference's `coords.rs` (crates/landmass/src/coords.rs:6-226):
conversion to/from landmass "standard" coordinates happens on the host, before
and after the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class XYZ:
    """3-D: X right, Y forward, Z up (coords.rs:77-92). Identity conversion."""

    FLIP_POLYGONS = False

    @staticmethod
    def to_landmass(v):
        return np.array([v[0], v[1], v[2]], dtype=np.float32)

    @staticmethod
    def from_landmass(v):
        return np.array([v[0], v[1], v[2]], dtype=np.float32)

    @staticmethod
    def sample_distance_from_agent_radius(radius: float):
        return PointSampleDistance3d.from_agent_radius(radius)


class XY:
    """2-D: X right, Y forward (coords.rs:161-176)."""

    FLIP_POLYGONS = False

    @staticmethod
    def to_landmass(v):
        return np.array([v[0], v[1], 0.0], dtype=np.float32)

    @staticmethod
    def from_landmass(v):
        return np.array([v[0], v[1]], dtype=np.float32)

    @staticmethod
    def sample_distance_from_agent_radius(radius: float):
        # impl FromAgentRadius for f32 (coords.rs:200-205)
        return np.float32(0.2) * np.float32(radius)


@dataclass
class PointSampleDistance3d:
    """coords.rs:96-158."""

    horizontal_distance: float
    distance_above: float
    distance_below: float
    vertical_preference_ratio: float
    animation_link_max_vertical_distance: float

    @classmethod
    def from_agent_radius(cls, radius: float) -> "PointSampleDistance3d":
        r = np.float32(radius)
        return cls(
            horizontal_distance=float(np.float32(0.2) * r),
            distance_above=float(np.float32(0.5) * r),
            distance_below=float(r),
            vertical_preference_ratio=2.0,
            animation_link_max_vertical_distance=float(np.float32(0.5) * r),
        )


@dataclass
class CorePointSampleDistance:
    """coords.rs:209-226: evaluated sample distances."""

    horizontal_distance: float
    distance_above: float
    distance_below: float
    vertical_preference_ratio: float

    @classmethod
    def new(cls, d) -> "CorePointSampleDistance":
        if isinstance(d, PointSampleDistance3d):
            return cls(d.horizontal_distance, d.distance_above, d.distance_below,
                       d.vertical_preference_ratio)
        if isinstance(d, CorePointSampleDistance):
            return d
        # `impl PointSampleDistance for f32` (coords.rs:178-198)
        return cls(float(d), 1.0, 1.0, 1.0)


def animation_link_max_vertical_distance(d) -> float:
    if isinstance(d, PointSampleDistance3d):
        return d.animation_link_max_vertical_distance
    return 1.0
