"""`Agent::has_reached_target` known answers (agent.rs:330-404), end to end through
`Archipelago.update` — transcribed from crates/landmass/src/agent_test.rs:157-363 (plus 381-544, 593-774 further down)
(`long_detour_reaches_target_in_different_ways`): a U-shaped corridor on a rotated island where
the agent starts one unit from its target as the crow flies but 21 units along the path.

The reference test calls `has_reached_target` directly with a hand-built path and funnel result;
here the same agent positions / conditions go through the whole step (sampling, A*, funnel, reached
test), whose corridor [0, 1, 2] and first waypoint are the ones the reference test assumes.  The
cases whose agent position lies off the mesh (agent_test.rs:287-301, 348-362) cannot be sampled
by a real update and are left out."""
import math

import numpy as np
import pytest

from landmass_b200 import (Agent, AgentState, Archipelago, ArchipelagoOptions, Island,
                           TargetReachedCondition, Transform)
from tests.helpers import valid_mesh

U_MESH = dict(
    vertices=[(1.0, 1.0, 0.0), (1.0, 11.0, 0.0), (0.0, 12.0, 0.0), (2.0, 11.0, 0.0), (3.0, 12.0, 0.0),
              (2.0, 1.0, 0.0)],
    polygons=[[0, 1, 2], [2, 1, 3, 4], [4, 3, 5]],
    polygon_type_indices=[0, 0, 0])

TRANSLATION = (2.0, 4.0, 3.0)
ROTATION = math.pi * -0.85


def apply(p):
    """Transform::apply (util.rs:358-367) in float64 — only used to place agents; the sampled node
    does not depend on the last bits."""
    c, s = math.cos(ROTATION), math.sin(ROTATION)
    return (c * p[0] - s * p[1] + TRANSLATION[0], s * p[0] + c * p[1] + TRANSLATION[1], p[2] + TRANSLATION[2])


CASES = [
    # (condition, local agent position, reached?)                       agent_test.rs lines
    (TargetReachedCondition.Distance(1.1), (1.0, 1.0, 0.0), True),              # 200-219
    (TargetReachedCondition.Distance(1.1), (1.0, 2.0, 0.0), False),             # 221-237
    (TargetReachedCondition.VisibleAtDistance(15.0), (1.0, 1.0, 0.0), False),   # 244-258
    (TargetReachedCondition.VisibleAtDistance(15.0), (1.0, 10.0, 0.0), False),  # 260-275
    (TargetReachedCondition.VisibleAtDistance(15.0), (2.0, 11.0, 0.0), True),   # 277-285 (rounded the corner)
    (TargetReachedCondition.StraightPathDistance(15.0), (1.0, 1.0, 0.0), False),  # 308-323 (21 units left)
    (TargetReachedCondition.StraightPathDistance(15.0), (1.0, 10.0, 0.0), True),  # 325-346 (12 units left)
]


@pytest.mark.parametrize("condition,local_pos,reached", CASES)
def test_long_detour_reaches_target_in_different_ways(make_backend, condition, local_pos, reached):
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())
    arch.add_island(Island(Transform(TRANSLATION, ROTATION), valid_mesh(U_MESH)))
    agent = arch.add_agent(Agent.create(apply(local_pos), (0, 0, 0), 0.0, 0.0, 0.0))
    arch.get_agent_mut(agent).current_target = apply((2.0, 1.0, 0.0))
    arch.get_agent_mut(agent).target_reached_condition = condition
    arch.update(0.01)
    state = arch.get_agent(agent).state()
    assert state == (AgentState.ReachedTarget if reached else AgentState.Moving)
    pr = arch.get_pathing_results()
    assert len(pr) == 1 and pr[0].success
    if not reached:
        # the corridor and portal edges the reference test builds by hand (agent_test.rs:189-199)
        nodes, portals = arch.backend.agent_path(agent.idx)
        assert [int(n) for n in nodes] == [0, 1, 2][-len(nodes):]
        if len(nodes) == 3:
            assert [int(p) for p in portals[:2]] == [1, 2]


@pytest.mark.parametrize("target,corridor", [((12.5, 12.5), [0, 4, 3]), ((10.5, 12.5), [0, 4])])
@pytest.mark.parametrize("condition,reached", [
    (TargetReachedCondition.Distance(10.0), True),               # agent_test.rs:467-487
    (TargetReachedCondition.VisibleAtDistance(10.0), False),     # agent_test.rs:489-508
    (TargetReachedCondition.StraightPathDistance(10.0), False),  # agent_test.rs:510-529
])
def test_using_animation_link_does_not_reach_target(make_backend, condition, reached, target, corridor):
    """agent_test.rs:381-544: an animation link between the agent and its target keeps the
    VisibleAtDistance / StraightPathDistance conditions from firing however large the distance;
    plain Distance ignores the path.  The reference test hands `has_reached_target` a hand-built
    path and next-waypoint index; here the whole update runs (A* picks the link: it costs 0.75).

    Second target (in the link's end node): `find_next_point_in_straight_path` returns the index
    AFTER the link (path.rs:352-356), which is then the target's own index, so a real update takes
    the `next_waypoint.0 == target_waypoint.0` exit of StraightPathDistance (agent.rs:365-367) and
    reports ReachedTarget — unlike the unit test, whose hand-written index (0, 1) is one the funnel
    never returns.  The oracle and the kernel follow the code."""
    from landmass_b200 import XY, AnimationLink
    from tests.test_golden_animation_links import GAP_MESH
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=make_backend())
    arch.add_island(Island(Transform((10.0, 10.0), 0.0), valid_mesh(GAP_MESH, cs=XY)))
    arch.add_animation_link(AnimationLink(start_edge=((10.0, 11.0), (11.0, 11.0)),
                                          end_edge=((10.0, 12.0), (11.0, 12.0)), cost=0.75, kind=0,
                                          bidirectional=False))
    arch.update(1.0)
    a = Agent.create((10.5, 10.5), (0.0, 0.0), 0.0, 0.0, 0.0, cs=XY)
    a.current_target = target
    a.target_reached_condition = condition
    # lib.rs:500-510: without this the 10-unit condition distance doubles as the distance at which the
    # link itself counts as reached (agent.rs:312-323)
    a.animation_link_reached_distance = 0.1
    aid = arch.add_agent(a)
    arch.update(0.01)
    if len(corridor) == 2 and condition.kind == TargetReachedCondition.StraightPathDistance().kind:
        reached = True
    assert arch.get_agent(aid).state() == (AgentState.ReachedTarget if reached else AgentState.Moving)
    if not reached:
        nodes, portals = arch.agent_path(aid)
        assert [arch.node_ref(int(n))[1] for n in nodes] == corridor


# A row of five unit squares whose polygon ids read 2, 3, 4, 1, 0 from left to right (the corridor of
# agent_test.rs:700-709), plus polygon 5 above the second square and polygon 6 above the fourth.
def _chain_mesh():
    v = lambda x, y: y * 6 + x
    sq = lambda k, row: [v(k, row), v(k + 1, row), v(k + 1, row + 1), v(k, row + 1)]
    geom = {0: sq(4, 0), 1: sq(3, 0), 2: sq(0, 0), 3: sq(1, 0), 4: sq(2, 0), 5: sq(1, 1), 6: sq(3, 1)}
    return dict(vertices=[(float(x), float(y), 0.0) for y in range(3) for x in range(6)],
                polygons=[geom[i] for i in range(7)], polygon_type_indices=[0] * 7)


CHAIN_CENTRE = {2: (0.5, 0.5, 0.0), 3: (1.5, 0.5, 0.0), 4: (2.5, 0.5, 0.0), 1: (3.5, 0.5, 0.0), 0: (4.5, 0.5, 0.0),
                5: (1.5, 1.5, 0.0), 6: (3.5, 1.5, 0.0)}


@pytest.mark.parametrize("agent_poly,target_poly,touch_island,repaths", [
    (3, 1, True, True),    # agent_test.rs:711-721 invalidated island
    (5, 1, False, True),   # 723-733 agent node not on the path
    (3, 6, False, True),   # 735-745 target node not on the path
    (1, 3, False, True),   # 747-757 agent and target in the wrong order
    (3, 1, False, False),  # 759-774 FollowPath(index 1, index 3): the path is kept
])
def test_repaths_for_invalid_path_or_nodes_off_path(make_backend, agent_poly, target_poly, touch_island, repaths):
    """agent_test.rs:671-774 (`does_agent_need_repath`, agent.rs:426-475) through whole updates: the
    first frame finds the corridor [2, 3, 4, 1, 0]; the second frame moves the agent and the target
    and either re-plans (a pathing result is reported) or keeps following the old corridor."""
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())
    island = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(_chain_mesh())))
    aid = arch.add_agent(Agent.create(CHAIN_CENTRE[2], (0, 0, 0), 0.25, 1.0, 1.0))
    arch.get_agent_mut(aid).current_target = CHAIN_CENTRE[0]
    arch.update(0.01)
    assert len(arch.get_pathing_results()) == 1                            # 687-698 no path yet
    nodes, _ = arch.agent_path(aid)
    assert [int(n) for n in nodes] == [2, 3, 4, 1, 0]

    arch.get_agent_mut(aid).position = CHAIN_CENTRE[agent_poly]
    arch.get_agent_mut(aid).current_target = CHAIN_CENTRE[target_poly]
    if touch_island:
        arch.get_island_mut(island).set_transform(Transform((0.0, 0.0, 0.0), 0.0))
    arch.update(0.01)
    assert len(arch.get_pathing_results()) == (1 if repaths else 0)
    assert arch.get_agent(aid).state() == AgentState.Moving
    nodes, _ = arch.agent_path(aid)
    if repaths:
        assert int(nodes[0]) == agent_poly and int(nodes[-1]) == target_poly
    else:
        assert [int(n) for n in nodes] == [2, 3, 4, 1, 0]
        v = arch.get_agent(aid).get_desired_velocity()
        np.testing.assert_allclose(v, (1.0, 0.0, 0.0), atol=1e-6)


def test_clears_path_for_no_target_or_missing_nodes(make_backend):
    """agent_test.rs:593-669: with a path but no target the path is cleared (Idle); with the agent
    or the target off the mesh the path is cleared and the state says which."""
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(_chain_mesh())))
    aid = arch.add_agent(Agent.create(CHAIN_CENTRE[2], (0, 0, 0), 0.25, 1.0, 1.0))
    arch.update(0.01)
    assert arch.get_agent(aid).state() == AgentState.Idle and not arch.get_pathing_results()  # DoNothing
    for break_it, want in [(lambda a: setattr(a, "current_target", None), AgentState.Idle),
                           (lambda a: setattr(a, "position", (0.5, -30.0, 0.0)), AgentState.AgentNotOnNavMesh),
                           (lambda a: setattr(a, "current_target", (0.5, 40.0, 0.0)), AgentState.TargetNotOnNavMesh)]:
        a = arch.get_agent_mut(aid)
        a.position, a.current_target = CHAIN_CENTRE[2], CHAIN_CENTRE[0]
        arch.update(0.01)
        assert len(arch.agent_path(aid)[0]) == 5
        break_it(arch.get_agent_mut(aid))
        arch.update(0.01)
        assert arch.get_agent(aid).state() == want
        assert arch.agent_path(aid) is None and not arch.get_pathing_results()
        assert tuple(arch.get_agent(aid).get_desired_velocity()) == (0.0, 0.0, 0.0)


def test_agent_host_api():
    """Host-side Agent bookkeeping, no update involved: agent_test.rs:20-64 (type-index cost
    overrides; zero / negative costs are refused) and 775-799 (start / end of an animation link need
    a reached link / a started link)."""
    from landmass_b200 import XY
    from landmass_b200.archipelago import (NotReachedAnimationLinkError, NotUsingAnimationLinkError,
                                           ReachedAnimationLink)
    agent = Agent.create((0.0, 0.0), (0.0, 0.0), 1.0, 1.0, 1.0, cs=XY)
    assert agent.override_type_index_cost(1, 3.0) and agent.override_type_index_cost(2, 0.5)
    assert sorted(agent.get_type_index_cost_overrides()) == [(1, 3.0), (2, 0.5)]
    agent.override_type_index_cost(1, 5.0)
    assert sorted(agent.get_type_index_cost_overrides()) == [(1, 5.0), (2, 0.5)]
    assert not agent.override_type_index_cost(0, 0.0) and not agent.override_type_index_cost(0, -0.5)
    assert agent.remove_overridden_type_index_cost(1) and not agent.remove_overridden_type_index_cost(1)

    agent = Agent.create((0.0, 0.0), (0.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    with pytest.raises(NotReachedAnimationLinkError):
        agent.start_animation_link()
    with pytest.raises(NotUsingAnimationLinkError):
        agent.end_animation_link()
    agent._current_animation_link = ReachedAnimationLink((0.0, 0.0), (1.0, 0.0), None)
    with pytest.raises(NotUsingAnimationLinkError):
        agent.end_animation_link()
    agent.start_animation_link()
    assert agent.is_using_animation_link()
    agent.end_animation_link()
    assert not agent.is_using_animation_link()


@pytest.mark.parametrize("condition", [TargetReachedCondition.Distance(2.0),
                                       TargetReachedCondition.StraightPathDistance(2.0),
                                       TargetReachedCondition.VisibleAtDistance(2.0)])
@pytest.mark.parametrize("target,reached", [((2.5, 1.0), True), ((3.5, 1.0), False), ((2.0, 2.0), True),
                                            ((2.5, 2.5), False)])
def test_has_reached_target_at_end_node(make_backend, condition, target, reached):
    """agent_test.rs:70-154: agent and target in the same node, so the next waypoint IS the target and
    all three conditions reduce to the distance (1.5, 2.5, sqrt 2, 2.12 against 2.0).  The reference
    test moves the waypoint in x/z over an empty mesh; here the same offsets lie in the plane of one
    5x5 polygon on the reference test's transformed island."""
    tr, rot = (2.0, 3.0, 4.0), math.pi * 0.85

    def world(p):
        c, s = math.cos(rot), math.sin(rot)
        return (c * p[0] - s * p[1] + tr[0], s * p[0] + c * p[1] + tr[1], tr[2])

    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())
    arch.add_island(Island(Transform(tr, rot), valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (5.0, 0.0, 0.0), (5.0, 5.0, 0.0), (0.0, 5.0, 0.0)], polygons=[[0, 1, 2, 3]],
        polygon_type_indices=[0]))))
    agent = arch.add_agent(Agent.create(world((1.0, 1.0)), (0, 0, 0), 0.0, 0.0, 0.0))
    arch.get_agent_mut(agent).current_target = world(target)
    arch.get_agent_mut(agent).target_reached_condition = condition
    arch.update(0.01)
    assert arch.get_agent(agent).state() == (AgentState.ReachedTarget if reached else AgentState.Moving)


def test_uses_sampled_point_for_reaching_target(make_backend):
    """agent_test.rs:546-591: the distance is taken from the agent's SAMPLED point.  An agent hovering
    0.45 above the mesh is 0.54 from a target 0.3 away on the ground, but its sampled point is 0.3
    from it: reached at Distance(0.5).  From 0.6 away horizontally it is not."""
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())
    arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(dict(
        vertices=[(0.0, 0.0, 0.0), (5.0, 0.0, 0.0), (5.0, 5.0, 0.0), (0.0, 5.0, 0.0)], polygons=[[0, 1, 2, 3]],
        polygon_type_indices=[0]))))
    near = arch.add_agent(Agent.create((1.0, 1.0, 0.45), (0, 0, 0), 0.0, 0.0, 0.0))
    far = arch.add_agent(Agent.create((1.0, 1.0, 0.45), (0, 0, 0), 0.0, 0.0, 0.0))
    for key, tx in ((near, 1.3), (far, 1.6)):
        arch.get_agent_mut(key).current_target = (tx, 1.0, 0.0)
        arch.get_agent_mut(key).target_reached_condition = TargetReachedCondition.Distance(0.5)
    arch.update(0.01)
    assert arch.get_agent(near).state() == AgentState.ReachedTarget
    assert arch.get_agent(far).state() == AgentState.Moving
