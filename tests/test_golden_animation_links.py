"""Animation links — transcribed from crates/landmass/src/pathfinding_test.rs:1187-1511 (A* through
off-mesh animation links, cost, permitted kinds) and lib_test.rs:1013-1123
(`agent_can_use_animation_link`: Moving -> ReachedAnimationLink -> UsingAnimationLink -> Moving with
exact desired velocities).  Link creation runs on the host (landmass_b200/anim_links.py)."""
import math

import numpy as np
import pytest

from landmass_b200 import (XY, Agent, AgentState, AnimationLink, Archipelago, ArchipelagoOptions,
                           Island, PermittedAnimationLinks, Transform, _abi)
from tests.helpers import find_path, new_archipelago, valid_mesh

GAP_MESH = dict(
    vertices=[(0.0, 0.0), (2.0, 0.0), (3.0, 0.0), (0.0, 1.0), (2.0, 1.0), (3.0, 1.0), (0.0, 2.0), (2.0, 2.0),
              (3.0, 2.0), (0.0, 3.0), (2.0, 3.0), (3.0, 3.0)],
    polygons=[[0, 1, 4, 3], [1, 2, 5, 4], [4, 5, 8, 7], [7, 8, 11, 10], [6, 7, 10, 9]],
    polygon_type_indices=[0] * 5)


def gap_arch(make_backend, cost):
    arch = new_archipelago(make_backend(), cs=XY)
    i = arch.add_island(Island(Transform((10.0, 10.0), 0.0), valid_mesh(GAP_MESH, cs=XY)))
    link = arch.add_animation_link(AnimationLink(start_edge=((10.0, 11.0), (11.0, 11.0)),
                                                 end_edge=((10.0, 12.0), (11.0, 12.0)), cost=cost, kind=0,
                                                 bidirectional=False))
    return arch, i, link


def off_mesh_link_for(arch, link_key):
    arch._sync_navdata()
    f = arch._flat
    hits = [l for l in range(f["n_links"]) if f["link_kind"][l] == _abi.LINK_ANIMATION
            and f["link_anim_id"][l] == link_key.idx]
    assert len(hits) == 1
    return hits[0]


def test_animation_link_is_used(make_backend):  # pathfinding_test.rs:1187-1284
    arch, i, link = gap_arch(make_backend, 1.0)
    l = off_mesh_link_for(arch, link)
    path, _ = find_path(arch, (i, 0), (10.5, 10.5, 0.0), (i, 4), (10.5, 12.5, 0.0))
    assert path == ([(i, [0], []), (i, [4], [])], [((i, 0), (i, 4), l)])


def test_animation_link_is_used_if_cheaper(make_backend):  # pathfinding_test.rs:1286-1423
    arch, i, link = gap_arch(make_backend, 1.5)
    sp, ep = (12.5, 10.5, 0.0), (10.5, 12.5, 0.0)
    path, _ = find_path(arch, (i, 1), sp, (i, 4), ep)
    assert path == ([(i, [1, 2, 3, 4], [2, 2, 3])], [])
    arch.remove_animation_link(link)
    link = arch.add_animation_link(AnimationLink(((10.0, 11.0), (11.0, 11.0)), ((10.0, 12.0), (11.0, 12.0)),
                                                 kind=0, cost=0.75, bidirectional=False))
    l = off_mesh_link_for(arch, link)
    path, _ = find_path(arch, (i, 1), sp, (i, 4), ep)
    assert path == ([(i, [1, 0], [3]), (i, [4], [])], [((i, 0), (i, 4), l)])


def test_animation_link_is_not_used_if_not_permitted(make_backend):  # pathfinding_test.rs:1425-1511
    arch, i, link = gap_arch(make_backend, 1.0)
    a = Agent.create((10.5, 10.5), (0.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    a.current_target = (10.5, 12.5)
    a.permitted_animation_links = PermittedAnimationLinks(frozenset())
    aid = arch.add_agent(a)
    arch.update(0.01)
    nodes, portals = arch.agent_path(aid)
    assert [arch.node_ref(int(n))[1] for n in nodes] == [0, 1, 2, 3, 4]
    assert [int(p) for p in portals[:-1]] == [1, 2, 2, 3] and portals[-1] == _abi.NONE
    # with every kind permitted the same agent takes the link
    b = Agent.create((10.5, 10.5), (0.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    b.current_target = (10.5, 12.5)
    bid = arch.add_agent(b)
    arch.update(0.01)
    nodes, portals = arch.agent_path(bid)
    assert [arch.node_ref(int(n))[1] for n in nodes] == [0, 4]
    assert int(portals[0]) == (_abi.PORTAL_LINK | off_mesh_link_for(arch, link))


def test_agent_can_use_animation_link(make_backend):  # lib_test.rs:1013-1123
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=make_backend())
    mesh = valid_mesh(dict(
        vertices=[(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 1.0), (2.0, 2.0), (1.0, 2.0)],
        polygons=[[0, 1, 2, 3], [2, 1, 4], [2, 4, 5, 6]], polygon_type_indices=[0] * 3), cs=XY)
    arch.add_island(Island(Transform(), mesh))
    arch.add_island(Island(Transform((3.0, 5.0), float(np.float32(math.pi))), mesh))
    link = arch.add_animation_link(AnimationLink(((1.0, 1.9), (2.0, 1.9)), ((1.0, 3.1), (2.0, 3.1)),
                                                 cost=1.0, kind=0, bidirectional=False))
    a = Agent.create((0.25, 0.75), (0.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    a.current_target = (2.75, 4.25)
    aid = arch.add_agent(a)

    def normalize(v):
        v = np.array(v, dtype=np.float32)
        return v * (np.float32(1.0) / np.float32(np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1]))))

    arch.update(1.0)
    agent = arch.get_agent_mut(aid)
    assert agent.state() == AgentState.Moving
    assert np.array_equal(agent.get_desired_velocity(),
                          normalize(np.float32([1.0, 1.0]) - np.float32([0.25, 0.75])))
    assert not agent.is_using_animation_link() and agent.reached_animation_link() is None

    agent.position = (1.25, 0.75)
    arch.update(1.0)
    assert agent.state() == AgentState.Moving
    assert np.array_equal(agent.get_desired_velocity(), np.float32([0.0, 1.0]))
    assert agent.reached_animation_link() is None

    agent.position = (1.25, 1.5)
    arch.update(1.0)
    assert agent.state() == AgentState.ReachedAnimationLink
    assert np.array_equal(agent.get_desired_velocity(), np.float32([0.0, 1.0]))
    assert not agent.is_using_animation_link()
    r = agent.reached_animation_link()
    assert r is not None and r.link_id == link
    assert np.array_equal(r.start_point, np.float32([1.25, 1.9]))
    assert np.array_equal(r.end_point, np.float32([1.25, 3.1]))
    agent.start_animation_link()

    agent.position = (100.0, 200.0)
    arch.update(1.0)
    assert agent.state() == AgentState.UsingAnimationLink
    assert agent.is_using_animation_link() and agent.reached_animation_link() is None
    agent.end_animation_link()

    agent.position = (1.75, 3.1)
    arch.update(1.0)
    assert agent.state() == AgentState.Moving
    assert np.array_equal(agent.get_desired_velocity(),
                          normalize(np.float32([2.0, 4.0]) - np.float32([1.75, 3.1])))
    assert not agent.is_using_animation_link() and agent.reached_animation_link() is None
