"""landmass_b200 — B200-native per-frame agent step of the landmass crate.

Host-side mirror of the reference's public API for ONE hot path,
`Archipelago::update` (crates/landmass/src/lib.rs:270-534), on top of the C ABI
in include/landmass_b200.h.  The compute lives in landmass_b200/csrc (CUDA,
sm_100a); importing this package never imports anything from oracle/.
"""
from .archipelago import (Agent, AgentState, Archipelago, ArchipelagoOptions, Character, FindPathError,
                          PathingResult, PathStep, PermittedAnimationLinks, ReachedAnimationLink,
                          SampledPoint, SamplePointError, TargetReachedCondition)
from .anim_links import AnimationLink
from .debug import (AvoidanceDrawer, ConstraintKind, ConstraintLine, DebugDrawError, DebugDrawer,  # noqa: F401
                    LineType, PointType, TriangleType, draw_archipelago_debug, draw_avoidance_data)
from .coords import XY, XYZ, CorePointSampleDistance, PointSampleDistance3d
from .nav_data import Island, NavigationData, SetTypeIndexCostError, Transform
from .nav_mesh import (HeightNavigationMesh, HeightPolygon, NavigationMesh, ValidationError,
                       ValidNavigationMesh)

__all__ = [
    "Agent", "AgentState", "AnimationLink", "Archipelago", "ArchipelagoOptions", "Character", "FindPathError",
    "PathingResult", "PathStep", "PermittedAnimationLinks", "ReachedAnimationLink", "SampledPoint",
    "SamplePointError", "TargetReachedCondition", "XY", "XYZ",
    "CorePointSampleDistance", "PointSampleDistance3d", "Island", "NavigationData",
    "SetTypeIndexCostError", "Transform", "HeightNavigationMesh", "HeightPolygon",
    "NavigationMesh", "ValidationError", "ValidNavigationMesh",
]
