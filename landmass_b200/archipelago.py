"""Host-side mirror of the reference's public API for the per-frame agent step:
`Archipelago`, `ArchipelagoOptions`, `Agent`, `Character`, `AgentState`,
`TargetReachedCondition`, `PathingResult`
(crates/landmass/src/lib.rs:54-547, agent.rs:24-310, character.rs:14-31).

The reference is Rust and no Rust toolchain exists in this image, so the host
side above the C ABI is written here with the same names, argument meaning and
error behaviour; `Archipelago.update` packs the AoS agents into the SoA the C
ABI takes, calls `lmb_step` and scatters the outputs back — exactly what the
Rust `Archipelago::update` shim in INTEGRATION.md does.

There is NO CPU fallback: the default backend is the CUDA library and raises if
it cannot be loaded.  Tests inject the oracle through `backend=`.
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field

import numpy as np

from . import _abi
from .coords import (CorePointSampleDistance, XYZ, animation_link_max_vertical_distance)
from .nav_data import Island, NavigationData, SetTypeIndexCostError, Transform  # noqa: F401
from .slotmap import DenseSlotMap, Key


class AgentState(enum.IntEnum):
    """agent.rs:24-47 — codes match `lmb_agent_state`."""

    Idle = 0
    ReachedTarget = 1
    ReachedAnimationLink = 2
    UsingAnimationLink = 3
    Moving = 4
    AgentNotOnNavMesh = 5
    TargetNotOnNavMesh = 6
    NoPath = 7
    Paused = 8


@dataclass
class TargetReachedCondition:
    """agent.rs:136-160."""

    kind: int = _abi.REACH_DISTANCE
    distance: float | None = None

    @classmethod
    def Distance(cls, d=None):
        return cls(_abi.REACH_DISTANCE, d)

    @classmethod
    def VisibleAtDistance(cls, d=None):
        return cls(_abi.REACH_VISIBLE_AT_DISTANCE, d)

    @classmethod
    def StraightPathDistance(cls, d=None):
        return cls(_abi.REACH_STRAIGHT_PATH_DISTANCE, d)


@dataclass
class PermittedAnimationLinks:
    """agent.rs:170-187: `All` or `Kinds(set)`."""

    kinds: frozenset | None = None  # None = All

    def is_permitted(self, kind: int) -> bool:
        return self.kinds is None or kind in self.kinds


@dataclass
class ReachedAnimationLink:
    """agent.rs:112-119."""

    link_id: int
    start_point: np.ndarray
    end_point: np.ndarray


class NotReachedAnimationLinkError(Exception):
    pass


class NotUsingAnimationLinkError(Exception):
    pass


class Agent:
    """agent.rs:50-310."""

    def __init__(self, position, velocity, radius, desired_speed, max_speed, cs=XYZ):
        self.cs = cs
        self.position = position
        self.velocity = velocity
        self.radius = float(radius)
        self.desired_speed = float(desired_speed)
        self.max_speed = float(max_speed)
        self.current_target = None
        self.target_reached_condition = TargetReachedCondition.Distance(None)
        self.animation_link_reached_distance = None
        self.permitted_animation_links = PermittedAnimationLinks()
        self.paused = False
        # feature `debug-avoidance` (agent.rs:87-90, 105-108): keep the avoidance constraints of each update
        self.keep_avoidance_data = False
        self._avoidance_data = None  # ("Satisfied", constraints) | ("Fallback", original, fallback); lines (n, 4)
        self._override_type_index_to_cost: dict[int, float] = {}
        self._current_desired_move = cs.from_landmass(np.zeros(3, dtype=np.float32))
        self._state = AgentState.Idle
        self._current_animation_link: ReachedAnimationLink | None = None
        self._using_animation_link = False

    @classmethod
    def create(cls, position, velocity, radius, desired_speed, max_speed, cs=XYZ):
        return cls(position, velocity, radius, desired_speed, max_speed, cs)

    def override_type_index_cost(self, type_index: int, cost: float) -> bool:
        if cost <= 0.0:
            return False
        self._override_type_index_to_cost[type_index] = float(cost)
        return True

    def remove_overridden_type_index_cost(self, type_index: int) -> bool:
        return self._override_type_index_to_cost.pop(type_index, None) is not None

    def get_type_index_cost_overrides(self):
        return list(self._override_type_index_to_cost.items())

    def get_desired_velocity(self):
        return self._current_desired_move

    def state(self) -> AgentState:
        return self._state

    def reached_animation_link(self):
        return self._current_animation_link

    def start_animation_link(self):
        if self._current_animation_link is None:
            raise NotReachedAnimationLinkError()
        self._using_animation_link = True

    def end_animation_link(self):
        if not self._using_animation_link:
            raise NotUsingAnimationLinkError()
        self._using_animation_link = False

    def is_using_animation_link(self) -> bool:
        return self._using_animation_link


@dataclass
class Character:
    """character.rs:14-31."""

    position: object = None
    velocity: object = None
    radius: float = 0.0


@dataclass
class ArchipelagoOptions:
    """lib.rs:63-93."""

    point_sample_distance: object
    neighbourhood: float
    avoidance_time_horizon: float
    obstacle_avoidance_time_horizon: float
    reached_destination_avoidance_responsibility: float

    @classmethod
    def from_agent_radius(cls, radius: float, cs=XYZ) -> "ArchipelagoOptions":
        return cls(
            point_sample_distance=cs.sample_distance_from_agent_radius(radius),
            neighbourhood=float(np.float32(10.0) * np.float32(radius)),
            avoidance_time_horizon=0.5,
            obstacle_avoidance_time_horizon=0.25,
            reached_destination_avoidance_responsibility=0.1,
        )

    def to_abi(self) -> _abi.Options:
        core = CorePointSampleDistance.new(self.point_sample_distance)
        o = _abi.Options()
        o.horizontal_distance = core.horizontal_distance
        o.distance_above = core.distance_above
        o.distance_below = core.distance_below
        o.vertical_preference_ratio = core.vertical_preference_ratio
        o.neighbourhood = self.neighbourhood
        o.avoidance_time_horizon = self.avoidance_time_horizon
        o.obstacle_avoidance_time_horizon = self.obstacle_avoidance_time_horizon
        o.reached_destination_avoidance_responsibility = \
            self.reached_destination_avoidance_responsibility
        return o


class SamplePointError(Exception):
    """query.rs:57-66: kind is "OutOfRange" or "NavDataDirty"."""

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind

    def __eq__(self, other):
        return isinstance(other, SamplePointError) and other.kind == self.kind

    __hash__ = Exception.__hash__


class FindPathError(Exception):
    """query.rs:96-103: ("NonPositiveTypeIndexCost", type_index, cost) or ("NoPathFound",)."""

    def __init__(self, kind: str, *args):
        super().__init__(kind, *args)
        self.kind = kind
        self.args_ = args

    def __eq__(self, other):
        return isinstance(other, FindPathError) and (other.kind, other.args_) == (self.kind, self.args_)

    __hash__ = Exception.__hash__


class SampledPoint:
    """query.rs:15-55: a point on the navigation meshes."""

    def __init__(self, point, node_ref, type_index, landmass_point):
        self._point = point
        self._node_ref = node_ref          # (island key, polygon index)
        self._type_index = type_index
        self._landmass_point = landmass_point

    def point(self):
        return self._point.copy()

    def island(self) -> Key:
        return self._node_ref[0]

    def type_index(self) -> int:
        return self._type_index


class PathStep:
    """query.rs:105-123: `Waypoint(point)` or `AnimationLink{start_point, end_point, link_id}`."""

    def __init__(self, kind, point=None, start_point=None, end_point=None, link_id=None):
        self.kind, self.point, self.start_point, self.end_point, self.link_id = kind, point, start_point, end_point, link_id

    @classmethod
    def Waypoint(cls, point):
        return cls("Waypoint", point=np.asarray(point, dtype=np.float32))

    @classmethod
    def AnimationLink(cls, start_point, end_point, link_id):
        return cls("AnimationLink", start_point=np.asarray(start_point, dtype=np.float32),
                   end_point=np.asarray(end_point, dtype=np.float32), link_id=link_id)

    def _key(self):
        t = lambda v: None if v is None else tuple(float(x) for x in v)
        return self.kind, t(self.point), t(self.start_point), t(self.end_point), self.link_id

    def __eq__(self, other):
        return isinstance(other, PathStep) and self._key() == other._key()

    def __repr__(self):
        k = self._key()
        return f"Waypoint({k[1]})" if self.kind == "Waypoint" else f"AnimationLink({k[2]} -> {k[3]}, {k[4]})"


@dataclass
class PathingResult:
    """lib.rs:538-547."""

    agent: Key
    success: bool
    explored_nodes: int


class Archipelago:
    """lib.rs:54-535."""

    def __init__(self, archipelago_options: ArchipelagoOptions, cs=XYZ, backend=None,
                 max_agents: int = 1024, no_avoidance: bool = False):
        self.cs = cs
        self.archipelago_options = archipelago_options
        self.nav_data = NavigationData(cs)
        self.agents = DenseSlotMap()
        self.characters = DenseSlotMap()
        self._pathing_results: list[PathingResult] = []
        self._uploaded_generation = -1
        self._flat = None
        self._last_step_keys = []
        self._device_versions: dict[int, int] = {}
        self._max_agents = max_agents
        if backend is None:
            from ._native import NativeBackend  # raises loudly if the CUDA library is missing

            # device slot = DenseSlotMap key index, and index 0 is the slot map's sentinel: one slot more than agents
            backend = NativeBackend(max_agents=max_agents + 1,
                                    flags=_abi.CONFIG_NO_AVOIDANCE if no_avoidance else 0)
        self.backend = backend

    # --- agents / characters / islands (lib.rs:106-205) ---
    def add_agent(self, agent: Agent) -> Key:
        if len(self.agents) >= self._max_agents:
            raise ValueError(f"Archipelago was created for at most {self._max_agents} agents")
        return self.agents.insert(agent)

    def remove_agent(self, agent_id: Key):
        if self.agents.remove(agent_id) is None:
            raise KeyError("Agent should be present in the archipelago")

    def get_agent(self, agent_id: Key):
        return self.agents.get(agent_id)

    get_agent_mut = get_agent

    def get_agent_ids(self):
        return self.agents.keys()

    def add_character(self, character: Character) -> Key:
        return self.characters.insert(character)

    def remove_character(self, character_id: Key):
        if self.characters.remove(character_id) is None:
            raise KeyError("Character should be present in the archipelago")

    def get_character(self, character_id: Key):
        return self.characters.get(character_id)

    get_character_mut = get_character

    def get_character_ids(self):
        return self.characters.keys()

    def add_island(self, island: Island) -> Key:
        return self.nav_data.add_island(island)

    def remove_island(self, island_id: Key):
        self.nav_data.remove_island(island_id)

    def get_island(self, island_id: Key):
        return self.nav_data.get_island(island_id)

    def get_island_mut(self, island_id: Key):
        # IslandMut marks dirty on mutable access (nav_data.rs:1211-1216)
        island = self.nav_data.get_island(island_id)
        if island is not None:
            self.nav_data.dirty = True
        return island

    def get_island_ids(self):
        return self.nav_data.get_island_ids()

    def add_animation_link(self, link) -> Key:
        return self.nav_data.add_animation_link(link)

    def remove_animation_link(self, link_id: Key):
        self.nav_data.remove_animation_link(link_id)

    def get_animation_link(self, link_id: Key):
        return self.nav_data.get_animation_link(link_id)

    def get_animation_link_ids(self):
        return self.nav_data.get_animation_link_ids()

    def set_type_index_cost(self, type_index: int, cost: float):
        self.nav_data.set_type_index_cost(type_index, cost)

    def get_type_index_cost(self, type_index: int):
        return self.nav_data.get_type_index_cost(type_index)

    def get_type_index_costs(self):
        return list(self.nav_data.type_index_to_cost.items())

    def get_pathing_results(self):
        return self._pathing_results

    # --- helpers ---
    def node_id(self, island_id: Key, polygon_index: int) -> int:
        """Global node id used across the C ABI."""
        self._sync_navdata()
        return int(self._flat["island_poly_begin"][self._flat["_island_key_to_index"][island_id]]) + polygon_index

    def node_ref(self, node_id: int):
        """Inverse of `node_id`: (island key, polygon index)."""
        begin = self._flat["island_poly_begin"]
        i = int(np.searchsorted(begin, node_id, side="right") - 1)
        key = [k for k, v in self._flat["_island_key_to_index"].items() if v == i][0]
        return key, int(node_id - begin[i])

    def _sync_navdata(self):
        self.nav_data.update(0.01, animation_link_max_vertical_distance(
            self.archipelago_options.point_sample_distance))
        if self._uploaded_generation != self.nav_data.generation:
            flat = self.nav_data.flatten()
            self.backend.navdata_upload(flat, self._nav_remap(self._flat, flat))
            self._flat = flat
            self._uploaded_generation = self.nav_data.generation

    def _nav_remap(self, old, new):
        """What `NavigationData::update` hands `Archipelago::update` (lib.rs:274-281; nav_data.rs:326-403)
        — the invalidated islands (deleted or dirty) and the dropped off-mesh links (an end on such an
        island, or their animation link deleted) — expressed as maps from the previous upload's island /
        link indices to the new ones, so the backend can keep every path `Path::is_valid` keeps."""
        changed = self.nav_data.take_changed_islands()
        if old is None:
            return None
        new_index = new["_island_key_to_index"]
        island_map = np.array([new_index[k] if (k in new_index and k not in changed) else _abi.NONE
                               for k in old["_island_keys"]], dtype=np.uint32)
        new_link = {ident: i for i, ident in enumerate(new["_link_identity"])}
        link_map = np.array([new_link.get(ident, _abi.NONE) if (ident[0] not in changed and ident[2] not in changed)
                             else _abi.NONE for ident in old["_link_identity"]], dtype=np.uint32)
        return {"island_map": island_map, "link_map": link_map}

    def sample_point(self, point, point_sample_distance=None) -> "SampledPoint":
        """`Archipelago::sample_point` (lib.rs:235-246, query.rs:70-94) through the batched device entry
        point.  Raises SamplePointError("NavDataDirty") when the navigation data was mutated since the last
        `update`, SamplePointError("OutOfRange") when no node is within the sample distance."""
        if self.nav_data.dirty or any(i.dirty for i in self.nav_data.islands.values()):
            raise SamplePointError("NavDataDirty")
        self._sync_navdata()
        opts = self.archipelago_options.to_abi()
        if point_sample_distance is not None:
            core = CorePointSampleDistance.new(point_sample_distance)
            opts.horizontal_distance = core.horizontal_distance
            opts.distance_above = core.distance_above
            opts.distance_below = core.distance_below
            opts.vertical_preference_ratio = core.vertical_preference_ratio
        pts, nodes = self.backend.sample_points(
            self.cs.to_landmass(point).reshape(1, 3), opts)
        if nodes[0] == _abi.NONE:
            raise SamplePointError("OutOfRange")
        island, polygon = self.node_ref(int(nodes[0]))
        return SampledPoint(self.cs.from_landmass(pts[0]), (island, polygon),
                            int(self.nav_data.get_island(island).nav_mesh.poly_type[polygon]), pts[0].copy())

    def find_path(self, start_point: "SampledPoint", end_point: "SampledPoint", override_type_index_costs=None,
                  permitted_animation_links: PermittedAnimationLinks | None = None):
        """`Archipelago::find_path` (lib.rs:248-268, query.rs:185-273): the straight path between two
        sampled points as a list of `PathStep`.  Raises FindPathError("NonPositiveTypeIndexCost", type, cost)
        or FindPathError("NoPathFound")."""
        assert not (self.nav_data.dirty or any(i.dirty for i in self.nav_data.islands.values())), \
            "The navigation data has been mutated, but we have SampledPoints, so this should be impossible."
        overrides = dict(override_type_index_costs or {})
        for type_index, cost in overrides.items():
            if cost <= 0.0:
                raise FindPathError("NonPositiveTypeIndexCost", type_index, cost)
        self._sync_navdata()
        kinds = None if permitted_animation_links is None else permitted_animation_links.kinds
        succ, begin, kind, pts, link = self.backend.find_straight_paths(
            [self.node_id(*start_point._node_ref)], [start_point._landmass_point],
            [self.node_id(*end_point._node_ref)], [end_point._landmass_point], [overrides], [kinds])
        if not succ[0]:
            raise FindPathError("NoPathFound")
        link_keys = {k.idx: k for k in self.nav_data.get_animation_link_ids()}
        steps = []
        for k in range(int(begin[0]), int(begin[1])):
            a, b = self.cs.from_landmass(pts[k, :3]), self.cs.from_landmass(pts[k, 3:])
            steps.append(PathStep.Waypoint(a) if kind[k] == 0 else PathStep.AnimationLink(a, b, link_keys[int(link[k])]))
        # the sampled points are returned as given (query.rs:224, 227), not after a to/from_landmass round trip
        steps[0] = PathStep.Waypoint(start_point.point())
        if steps[-1].kind == "Waypoint":
            steps[-1] = PathStep.Waypoint(end_point.point())
        return steps

    # --- the hot path ---
    def update(self, delta_time: float):
        """`Archipelago::update` (lib.rs:270-534): pack -> lmb_step -> scatter."""
        cs = self.cs
        self._sync_navdata()
        items = self.agents.items()
        n = len(items)
        ag = {
            "n": n,
            "slot": np.zeros(n, dtype=np.uint32),
            "flags": np.zeros(n, dtype=np.uint32),
            "position": np.zeros((n, 3), dtype=np.float32),
            "velocity": np.zeros((n, 3), dtype=np.float32),
            "target": np.zeros((n, 3), dtype=np.float32),
            "radius": np.zeros(n, dtype=np.float32),
            "desired_speed": np.zeros(n, dtype=np.float32),
            "max_speed": np.zeros(n, dtype=np.float32),
            "reach_condition": np.zeros(n, dtype=np.uint32),
            "reach_distance": np.full(n, -1.0, dtype=np.float32),
            "anim_link_reach_distance": np.full(n, -1.0, dtype=np.float32),
        }
        ov_begin = np.zeros(n + 1, dtype=np.uint32)
        ov_type, ov_cost = [], []
        pm_begin = np.zeros(n + 1, dtype=np.uint32)
        pm_kind = []
        live_slots = {}
        for i, (key, a) in enumerate(items):
            flags = 0
            if self._device_versions.get(key.idx) != key.version:
                flags |= _abi.AGENT_RESET
            live_slots[key.idx] = key.version
            if a.current_target is not None:
                flags |= _abi.AGENT_HAS_TARGET
                ag["target"][i] = cs.to_landmass(a.current_target)
            if a.paused:
                flags |= _abi.AGENT_PAUSED
            if a._using_animation_link:
                flags |= _abi.AGENT_USING_ANIMATION_LINK
            if a.permitted_animation_links.kinds is None:
                flags |= _abi.AGENT_PERMIT_ALL_LINKS
            else:
                pm_kind.extend(sorted(a.permitted_animation_links.kinds))
            if a.keep_avoidance_data or a._avoidance_data is not None:
                # (an agent that stopped asking still needs to learn whether avoidance ran, to drop its old data)
                flags |= _abi.AGENT_KEEP_AVOIDANCE_DATA
            pm_begin[i + 1] = len(pm_kind)
            ag["slot"][i] = key.idx
            ag["flags"][i] = flags
            ag["position"][i] = cs.to_landmass(a.position)
            ag["velocity"][i] = cs.to_landmass(a.velocity)
            ag["radius"][i] = a.radius
            ag["desired_speed"][i] = a.desired_speed
            ag["max_speed"][i] = a.max_speed
            ag["reach_condition"][i] = a.target_reached_condition.kind
            if a.target_reached_condition.distance is not None:
                ag["reach_distance"][i] = a.target_reached_condition.distance
            if a.animation_link_reached_distance is not None:
                ag["anim_link_reach_distance"][i] = a.animation_link_reached_distance
            for t, c in sorted(a._override_type_index_to_cost.items()):
                ov_type.append(t)
                ov_cost.append(c)
            ov_begin[i + 1] = len(ov_type)
        self._device_versions = live_slots
        self._last_step_keys = [key for key, _ in items]
        ag["override_begin"] = ov_begin
        ag["override_type"] = np.array(ov_type, dtype=np.uint32)
        ag["override_cost"] = np.array(ov_cost, dtype=np.float32)
        ag["permitted_begin"] = pm_begin
        ag["permitted_kind"] = np.array(pm_kind, dtype=np.uint32)

        chars = self.characters.values()
        ch = {
            "n": len(chars),
            "position": np.array([cs.to_landmass(c.position) for c in chars], dtype=np.float32).reshape(-1, 3),
            "velocity": np.array([cs.to_landmass(c.velocity) for c in chars], dtype=np.float32).reshape(-1, 3),
            "radius": np.array([c.radius for c in chars], dtype=np.float32),
        }
        out = self.backend.step(ag, ch, self.archipelago_options.to_abi(), float(delta_time))

        link_keys = None
        for i, (key, a) in enumerate(items):
            a._current_desired_move = cs.from_landmass(out["desired_velocity"][i])
            a._state = AgentState(int(out["state"][i]))
            if out["reached_link_id"][i] != _abi.NONE:
                if link_keys is None:  # AnimationLinkId of the reached link, like find_path's steps
                    link_keys = {k.idx: k for k in self.nav_data.get_animation_link_ids()}
                a._current_animation_link = ReachedAnimationLink(
                    link_keys.get(int(out["reached_link_id"][i]), int(out["reached_link_id"][i])),
                    cs.from_landmass(out["reached_link_start"][i]),
                    cs.from_landmass(out["reached_link_end"][i]))
            else:
                a._current_animation_link = None
            if a.keep_avoidance_data or a._avoidance_data is not None:
                # agent.avoidance_data = agent.keep_avoidance_data.then_some(debug_data), only for the agents
                # avoidance visited (avoidance.rs:83-87, 161-172)
                ran, original, fallback, fell_back = self.backend.agent_avoidance_constraints(i)
                if ran:
                    if not a.keep_avoidance_data:
                        a._avoidance_data = None
                    elif fell_back:
                        a._avoidance_data = ("Fallback", original, fallback)
                    else:
                        a._avoidance_data = ("Satisfied", original)
        self._pathing_results = [
            PathingResult(items[int(r[0])][0], bool(r[1]), int(r[2]))
            for r in self.backend.pathing_results()
        ]

    def agent_path(self, agent_id: Key):
        """`agent.current_path` read-back: (nodes, portals) or None."""
        nodes, portals = self.backend.agent_path(agent_id.idx)
        if len(nodes) == 0:
            return None
        return nodes, portals
