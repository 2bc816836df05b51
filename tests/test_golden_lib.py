"""End-to-end `Archipelago::update` known answers — transcribed from
crates/landmass/src/lib_test.rs:156-423 (`computes_and_follows_path`: states, desired velocities
and pathing results over three frames, through the host mirror -> C ABI, all four phases)."""
import numpy as np
import pytest

from landmass_b200 import (Agent, AgentState, Archipelago, ArchipelagoOptions, Island,
                           PointSampleDistance3d, Transform)
from tests.helpers import valid_mesh

L_MESH = dict(
    vertices=[(1.0, 1.0, 1.0), (2.0, 1.0, 1.0), (3.0, 1.0, 1.0), (4.0, 1.0, 1.0), (4.0, 2.0, 1.0),
              (4.0, 3.0, 1.0), (4.0, 4.0, 1.0), (3.0, 4.0, 1.0), (3.0, 3.0, 1.0), (3.0, 2.0, 1.0),
              (2.0, 2.0, 1.0), (1.0, 2.0, 1.0)],
    polygons=[[0, 1, 10, 11], [1, 2, 9, 10], [2, 3, 4, 9], [4, 5, 8, 9], [5, 6, 7, 8]],
    polygon_type_indices=[0] * 5)


def norm2(x, y, s):
    v = np.array([x, y, 0.0], dtype=np.float32)
    return v / np.float32(np.sqrt(np.float32(x * x + y * y))) * np.float32(s)


def test_computes_and_follows_path(make_backend):
    opts = ArchipelagoOptions.from_agent_radius(0.5)
    opts.point_sample_distance = PointSampleDistance3d(0.1, 0.1, 0.1, 1.0, 1.0)
    arch = Archipelago(opts, backend=make_backend())
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(L_MESH)))
    arch.archipelago_options.neighbourhood = 0.0
    arch.archipelago_options.obstacle_avoidance_time_horizon = 0.01

    a1 = arch.add_agent(Agent.create((1.5, 1.5, 1.09), (0, 0, 0), 0.5, 2.0, 2.0))
    a2 = arch.add_agent(Agent.create((3.5, 3.5, 0.95), (0, 0, 0), 0.5, 2.0, 2.0))
    off = arch.add_agent(Agent.create((1.5, 2.5, 1.0), (0, 0, 0), 0.5, 2.0, 2.0))
    high = arch.add_agent(Agent.create((1.5, 1.5, 1.11), (0, 0, 0), 0.5, 2.0, 2.0))
    for a in (a1, off, high):
        arch.get_agent_mut(a).current_target = (3.5, 3.5, 0.95)
    arch.get_agent_mut(a2).current_target = (1.5, 1.5, 1.09)

    for a in (a1, a2, off, high):  # nothing has happened yet
        assert arch.get_agent(a).state() == AgentState.Idle
        assert not np.any(arch.get_agent(a).get_desired_velocity())

    def vel(a):
        return np.asarray(arch.get_agent(a).get_desired_velocity())

    arch.update(0.01)
    assert arch.get_agent(a1).state() == AgentState.Moving
    assert arch.get_agent(a2).state() == AgentState.Moving
    assert np.allclose(vel(a1), norm2(1.5, 0.5, 2.0), atol=1e-2)
    assert np.allclose(vel(a2), norm2(-0.5, -1.5, 2.0), atol=1e-2)
    for a in (off, high):
        assert arch.get_agent(a).state() == AgentState.AgentNotOnNavMesh
        assert not np.any(vel(a))
    pr = arch.get_pathing_results()
    assert len(pr) == 2 and pr[0].success and pr[1].success
    assert pr[0].explored_nodes > 0 and pr[1].explored_nodes > 0
    assert pr[0].agent == a1 and pr[1].agent == a2

    arch.get_agent_mut(a1).position = (2.5, 1.5, 1.0)
    arch.update(0.01)
    assert np.allclose(vel(a1), norm2(0.5, 0.5, 2.0), atol=1e-7)
    assert np.allclose(vel(a2), norm2(-0.5, -1.5, 2.0), atol=1e-2)
    assert arch.get_agent(a1).state() == AgentState.Moving
    assert arch.get_agent(a2).state() == AgentState.Moving
    assert len(arch.get_pathing_results()) == 0  # paths are reused
    for a in (off, high):
        assert arch.get_agent(a).state() == AgentState.AgentNotOnNavMesh
        assert not np.any(vel(a))

    arch.get_agent_mut(a1).position = (3.4, 3.4, 1.0)
    arch.get_agent_mut(a2).position = (3.5, 2.5, 1.0)
    arch.update(0.01)
    assert arch.get_agent(a1).state() == AgentState.ReachedTarget
    assert arch.get_agent(a2).state() == AgentState.Moving
    assert not np.any(vel(a1))
    assert np.allclose(vel(a2), norm2(-0.5, -0.5, 2.0), atol=1e-2)
    for a in (off, high):
        assert arch.get_agent(a).state() == AgentState.AgentNotOnNavMesh
        assert not np.any(vel(a))


def test_readme_two_agents_swap(make_backend):
    """crates/landmass/README.md:55-131 (doctest): two agents of radius 1 swap places in a 15x15
    room over 300 frames of dt = 0.1 and end within 0.1 of their targets."""
    from landmass_b200 import TargetReachedCondition

    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (15.0, 0.0, 0.0), (15.0, 15.0, 0.0), (0.0, 15.0, 0.0)],
        polygons=[[0, 1, 2, 3]], polygon_type_indices=[0]))))
    ids = []
    for pos, tgt in (((1.0, 1.0, 0.0), (11.0, 1.1, 0.0)), ((11.0, 1.1, 0.0), (1.0, 1.0, 0.0))):
        a = Agent.create(np.array(pos, dtype=np.float32), np.zeros(3, dtype=np.float32), 1.0, 1.0, 2.0)
        a.current_target = tgt
        a.target_reached_condition = TargetReachedCondition.Distance(0.01)
        ids.append(arch.add_agent(a))
    dt = np.float32(0.1)
    for _ in range(300):
        arch.update(float(dt))
        for i in ids:
            a = arch.get_agent_mut(i)
            a.velocity = np.asarray(a.get_desired_velocity(), dtype=np.float32).copy()
            a.position = (np.asarray(a.position, dtype=np.float32) + a.velocity * dt).astype(np.float32)
    assert np.allclose(arch.get_agent(ids[0]).position, (11.0, 1.1, 0.0), atol=0.1)
    assert np.allclose(arch.get_agent(ids[1]).position, (1.0, 1.0, 0.0), atol=0.1)
