"""`Agent::has_reached_target` known answers (agent.rs:330-404), end to end through
`Archipelago.update` — transcribed from crates/landmass/src/agent_test.rs:157-363
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
