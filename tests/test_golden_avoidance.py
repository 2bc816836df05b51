"""Obstacle sets and avoidance sanity — transcribed from crates/landmass/src/avoidance_test.rs.

Obstacle sets (avoidance_test.rs:68-387) are exact as sets (order / rotation tolerant, like the
reference's `assert_obstacles_match!`): they pin landmass's border flood + the visibility-set
restatement.  They run on the oracle only (the C ABI exposes no obstacle read-back); the CUDA
kernel is compared with the oracle on velocities in test_parity_synth.py."""
import numpy as np
import pytest

from landmass_b200 import Island, Transform
from tests.golden import meshes
from tests.helpers import new_archipelago, valid_mesh


def obstacle_matches(left, right):  # avoidance_test.rs:16-44
    (lc, lv), (rc, rv) = left, right
    if lc != rc or len(lv) != len(rv):
        return False
    lv = [tuple(map(float, v)) for v in lv]
    rv = [tuple(map(float, v)) for v in rv]
    if lc:
        return any(lv[o:] + lv[:o] == rv for o in range(len(lv)))
    return lv == rv


def assert_obstacles_match(left, right):  # avoidance_test.rs:46-66
    right = list(right)
    for lo in left:
        for i, ro in enumerate(right):
            if obstacle_matches(lo, ro):
                right.pop(i)
                break
        else:
            raise AssertionError(f"unmatched obstacle {lo}; remaining {right}")
    assert not right, f"unmatched expected obstacles {right}"


@pytest.fixture
def oracle_arch():
    from oracle.oracle_py import OracleBackend

    b = OracleBackend()
    yield new_archipelago(b)
    b.close()


def obstacles(arch, point, node, limit):
    arch._sync_navdata()
    return arch.backend.borders_to_obstacles(point, arch.node_id(*node), limit)


def test_computes_obstacle_for_box(oracle_arch):  # avoidance_test.rs:68-113
    arch = oracle_arch
    off = np.array([130.0, -50.0, 20.0], dtype=np.float32)
    i = arch.add_island(Island(Transform(tuple(off), 0.0), valid_mesh(meshes.AVOID_BOX)))
    got = obstacles(arch, np.array([1.5, 1.5, 0.0], dtype=np.float32) + off, (i, 0), 10.0)
    o = off[:2]
    exp = [(True, [np.float32([1, 1]) + o, np.float32([1, 2]) + o, np.float32([2, 2]) + o, np.float32([2, 1]) + o])]
    assert_obstacles_match(got, exp)


def test_dead_end_makes_open_obstacle(oracle_arch):  # avoidance_test.rs:115-235
    arch = oracle_arch
    i = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(meshes.DEAD_END)))
    assert_obstacles_match(obstacles(arch, (1.5, 1.5, 0.0), (i, 0), 10.0),
                           [(False, [(4, 3), (4, 2), (4, 1), (3, 1), (2, 1), (1, 1), (1, 2), (2, 2), (3, 2)])])
    assert_obstacles_match(obstacles(arch, (3.5, 3.5, 0.0), (i, 4), 10.0),
                           [(False, [(3, 2), (3, 3), (3, 4), (4, 4), (4, 3), (4, 2), (4, 1), (3, 1), (2, 1)])])
    assert_obstacles_match(obstacles(arch, (3.5, 3.5, 0.0), (i, 4), 1.0),
                           [(False, [(3, 2), (3, 3), (3, 4), (4, 4), (4, 3), (4, 2)])])
    assert_obstacles_match(obstacles(arch, (3.5, 1.5, 0.0), (i, 2), 10.0),
                           [(True, [(1, 1), (1, 2), (2, 2), (3, 2), (3, 3), (3, 4), (4, 4), (4, 3), (4, 2), (4, 1),
                                    (3, 1), (2, 1)])])


def test_split_borders(oracle_arch):  # avoidance_test.rs:237-327
    arch = oracle_arch
    i = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(meshes.SPLIT_BORDERS)))
    assert_obstacles_match(
        obstacles(arch, (3.0, 0.9, 0.0), (i, 0), 10.0),
        [(False, [(1, 1), (2, 1), (3, 1), (4, 1), (5, 1)]),
         (False, [(6, 2), (6, 1), (6, 0), (5, 0), (4, 0), (3, 0), (2, 0), (1, 0), (0, 0), (0, 1), (0, 2)])])


def test_creates_obstacles_across_boundary_link(oracle_arch):  # avoidance_test.rs:329-387
    arch = oracle_arch
    mesh = valid_mesh(meshes.RAISED_BOX)
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = arch.add_island(Island(Transform((1.0, 0.0, 0.0), 0.0), mesh))
    assert_obstacles_match(
        obstacles(arch, (2.5, 1.5, 1.0), (i2, 0), 10.0),
        [(False, [(2, 1), (1, 1), (1, 2), (2, 2)]), (False, [(2, 2), (3, 2), (3, 1), (2, 1)])])


# ---- velocity sanity (avoidance_test.rs:389-665): the reference pins these to +-0.05 only ----
RECT = dict(vertices=[(-1.0, -1.0, 0.0), (13.0, -1.0, 0.0), (13.0, 3.0, 0.0), (-1.0, 3.0, 0.0)],
            polygons=[[0, 1, 2, 3]], polygon_type_indices=[0])


def _run_avoidance(make_backend, agents, characters, neighbourhood, horizon, dt):
    """The reference tests call apply_avoidance_to_agents with preset desired moves; here the same
    scene goes through the whole step: each agent's target lies straight ahead, so phase C yields
    the preset desired move (direction * desired_speed)."""
    from landmass_b200 import Agent, Character

    arch = new_archipelago(make_backend())
    arch.archipelago_options.neighbourhood = neighbourhood
    arch.archipelago_options.avoidance_time_horizon = horizon
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(RECT)))
    ids = []
    for pos, vel, radius, target in agents:
        a = Agent.create(pos, vel, radius, 1.0, 1.0)
        a.current_target = target
        ids.append(arch.add_agent(a))
    for pos, vel, radius in characters:
        arch.add_character(Character(pos, vel, radius))
    arch.update(dt)
    return [np.asarray(arch.get_agent(i).get_desired_velocity()) for i in ids]


def test_applies_no_avoidance_for_far_agents(make_backend):  # avoidance_test.rs:389-488
    v = _run_avoidance(make_backend,
                       [((1.0, 1.0, 0.0), (0.0, 0.0, 0.0), 0.01, (12.0, 1.0, 0.0)),
                        ((11.0, 1.0, 0.0), (0.0, 0.0, 0.0), 0.01, (0.0, 1.0, 0.0))], [], 5.0, 0.5, 0.01)
    assert tuple(v[0]) == (1.0, 0.0, 0.0) and tuple(v[1]) == (-1.0, 0.0, 0.0)


def test_applies_avoidance_for_two_agents(make_backend):  # avoidance_test.rs:490-583
    v = _run_avoidance(make_backend,
                       [((1.0, 1.0, 0.0), (1.0, 0.0, 0.0), 1.0, (12.5, 1.0, 0.0)),
                        ((11.0, 1.01, 0.0), (-1.0, 0.0, 0.0), 1.0, (-0.5, 1.01, 0.0))], [], 15.0, 15.0, 0.0)
    assert np.allclose(v[0], (0.98, -0.2, 0.0), atol=0.05), v[0]
    assert np.allclose(v[1], (-0.98, 0.2, 0.0), atol=0.05), v[1]


def test_agent_avoids_character(make_backend):  # avoidance_test.rs:585-665
    v = _run_avoidance(make_backend, [((1.0, 1.0, 0.0), (1.0, 0.0, 0.0), 1.0, (12.5, 1.0, 0.0))],
                       [((11.0, 1.01, 0.0), (-1.0, 0.0, 0.0), 1.0)], 15.0, 15.0, 0.01)
    assert np.allclose(v[0], (np.sqrt(1.0 - 0.4 * 0.4), -0.4, 0.0), atol=0.05), v[0]


def test_agent_speeds_up_to_avoid_character(make_backend):  # lib_test.rs:426-486
    """An agent heading for the point a character is also heading for, from the same distance:
    alone it moves at its desired speed (exact); with the character it speeds up (max_speed 2) to
    pass first instead of braking."""
    from landmass_b200 import XY, Agent, Character

    square = dict(vertices=[(-10.0, -10.0), (10.0, -10.0), (10.0, 10.0), (-10.0, 10.0)], polygons=[[0, 1, 2, 3]],
                  polygon_type_indices=[0])
    arch = new_archipelago(make_backend(), cs=XY)
    arch.archipelago_options.avoidance_time_horizon = 100.0
    arch.archipelago_options.neighbourhood = 10.0
    arch.add_island(Island(Transform(), valid_mesh(square, cs=XY)))
    a = Agent.create((5.0, 0.0), (-1.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    a.current_target = (-5.0, 0.0)
    agent = arch.add_agent(a)
    arch.update(0.01)
    assert tuple(np.asarray(arch.get_agent(agent).get_desired_velocity(), dtype=np.float32)) == (-1.0, 0.0)
    arch.add_character(Character((0.0, 5.0), (0.0, -1.0), 0.5))
    arch.update(0.01)
    v = np.asarray(arch.get_agent(agent).get_desired_velocity(), dtype=np.float32)
    assert float(np.hypot(v[0], v[1])) > 1.1, v


def test_agent_speeds_up_to_avoid_character_preset_move(make_backend):  # avoidance_test.rs:667-763
    """The avoidance_test.rs variant: current velocity (-1, 0) but desired move (+1, 0) (here: a target
    straight ahead in +x), time horizon 15.  Alone the desired move is kept exactly; with the
    character on the symmetric collision course the result is faster than 1.1."""
    from landmass_b200 import XY, Agent, Character

    square = dict(vertices=[(-10.0, -10.0), (10.0, -10.0), (10.0, 10.0), (-10.0, 10.0)], polygons=[[0, 1, 2, 3]],
                  polygon_type_indices=[0])

    def run(with_character):
        arch = new_archipelago(make_backend(), cs=XY)
        arch.archipelago_options.neighbourhood = 15.0
        arch.archipelago_options.avoidance_time_horizon = 15.0
        arch.add_island(Island(Transform(), valid_mesh(square, cs=XY)))
        a = Agent.create((5.0, 0.0), (-1.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
        a.current_target = (9.0, 0.0)
        agent = arch.add_agent(a)
        if with_character:
            arch.add_character(Character((0.0, 5.0), (0.0, -1.0), 0.5))
        arch.update(0.01)
        return np.asarray(arch.get_agent(agent).get_desired_velocity(), dtype=np.float32)

    assert tuple(run(False)) == (1.0, 0.0)
    v = run(True)
    assert float(np.hypot(v[0], v[1])) > 1.1, v


def test_reached_target_agent_has_different_avoidance(make_backend):  # avoidance_test.rs:765-853
    """35 frames of a moving agent passing one that has reached its target: with
    reached_destination_avoidance_responsibility = 1/3 the standing agent takes 1/4 of the
    avoidance and the moving one 3/4 (dodgy_2d's responsibility weighting)."""
    from landmass_b200 import XY, Agent

    square = dict(vertices=[(-10.0, -10.0), (10.0, -10.0), (10.0, 10.0), (-10.0, 10.0)], polygons=[[0, 1, 2, 3]],
                  polygon_type_indices=[0])
    arch = new_archipelago(make_backend(), cs=XY)
    arch.add_island(Island(Transform(), valid_mesh(square, cs=XY)))
    a1 = Agent.create((0.0, 0.0), (0.0, 0.0), 0.5, 1.0, 1.0, cs=XY)
    a1.current_target = (0.0, 0.0)
    a2 = Agent.create((0.0, -3.0), (0.0, 1.0), 0.5, 1.0, 1.0, cs=XY)
    a2.current_target = (0.0, 3.0)
    id1, id2 = arch.add_agent(a1), arch.add_agent(a2)
    arch.archipelago_options.avoidance_time_horizon = 100.0
    arch.archipelago_options.obstacle_avoidance_time_horizon = 0.1
    arch.archipelago_options.reached_destination_avoidance_responsibility = 1.0 / 3.0
    for _ in range(35):
        arch.update(0.1)
        for i in (id1, id2):
            ag = arch.get_agent_mut(i)
            ag.velocity = np.asarray(ag.get_desired_velocity(), dtype=np.float32)
            ag.position = np.asarray(ag.position, dtype=np.float32) + ag.velocity * np.float32(0.1)
    x1 = float(arch.get_agent(id1).position[0])
    x2 = float(arch.get_agent(id2).position[0])
    # the reference asserts one-sidedly: (x - expected) < 0.01
    assert x1 - 0.25 < 0.01 and x2 - 0.75 < 0.01, (x1, x2)
    # and both did move aside, the mover three times as far as the stander
    assert abs(x1) > 0.1 and abs(x2) > 0.3 and abs(abs(x2) / abs(x1) - 3.0) < 0.5, (x1, x2)


def test_switching_nav_mesh_to_fewer_vertices_does_not_result_in_panic(make_backend):  # avoidance_test.rs:855-925
    """Two copies of a hexagon with redundant vertices, the second shifted so that its border is slightly
    misaligned (the boundary link clips new vertices into both nodes: `ModifiedNode`s, which avoidance walks);
    then the first island's mesh is replaced by one with fewer vertices and the archipelago updated again."""
    from landmass_b200 import XY, Agent

    redundant = valid_mesh(dict(vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 2.0), (0.0, 2.0), (0.0, 0.6), (0.0, 0.2)],
                                polygons=[[0, 1, 2, 3, 4, 5]], polygon_type_indices=[0]), cs=XY)
    simplified = valid_mesh(dict(vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 2.0), (0.0, 2.0)],
                                 polygons=[[0, 1, 2, 3]], polygon_type_indices=[0]), cs=XY)
    arch = new_archipelago(make_backend(), cs=XY)
    island_1 = arch.add_island(Island(Transform((0.0, 0.0), 0.0), redundant))
    arch.add_island(Island(Transform((1.0, 0.25), 0.0), redundant))
    a = Agent.create((0.5, 1.25), (0.0, 0.0), 0.5, 1.0, 1.0, cs=XY)
    a.current_target = (1.5, 1.25)
    agent = arch.add_agent(a)
    arch.update(0.01)
    # "ensures that pathing + avoidance is running"
    assert tuple(np.asarray(arch.get_agent(agent).get_desired_velocity())) == (1.0, 0.0)
    arch.get_island_mut(island_1).set_nav_mesh(simplified)
    arch.update(0.01)
    assert tuple(np.asarray(arch.get_agent(agent).get_desired_velocity())) == (1.0, 0.0)
