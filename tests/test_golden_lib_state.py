"""`Archipelago::update` state-machine known answers (lib.rs:343-523), transcribed from
crates/landmass/src/lib_test.rs: `agent_overrides_node_costs` (670-745, exact `normalize`),
`paused_agent_does_not_repath` (776-906), `paused_agent_path_is_removed_when_invalid` (908-1003),
`agent_path_cleared_when_clearing_target` (1146-1173), `agent_path_cleared_on_bad_target`
(1175-1204), `agent_could_not_find_path` (1206-1232).  2-D `XY` coordinates like the reference
tests; both backends (oracle here, CUDA on the GPU box)."""
import numpy as np

from landmass_b200 import XY, Agent, AgentState, Archipelago, ArchipelagoOptions, Island, Transform
from tests.helpers import valid_mesh


def mesh_xy(vertices, polygons, types=None):
    return valid_mesh(dict(vertices=vertices, polygons=polygons,
                           polygon_type_indices=types or [0] * len(polygons)), cs=XY)


def _t(x, y):
    return Transform((x, y), 0.0)


def new_arch(backend):
    return Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=backend)


def path_start_and_end(arch, agent):
    """(first node, last node) of `agent.current_path` as (island key, polygon) pairs, or None."""
    p = arch.agent_path(agent)
    if p is None:
        return None
    nodes, _ = p
    return arch.node_ref(int(nodes[0])), arch.node_ref(int(nodes[-1]))


ONE_NODE = ([(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0)], [[0, 1, 2, 3]])
TWO_NODE = ([(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (1.0, 2.0), (0.0, 2.0)], [[0, 1, 2, 3], [3, 2, 4, 5]])


def test_agent_overrides_node_costs(make_backend):  # lib_test.rs:670-745
    arch = new_arch(make_backend())
    mesh = mesh_xy(
        [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (2.0, 0.0), (2.0, 1.0), (3.0, 0.0), (3.0, 1.0),
         (2.0, 11.0), (3.0, 11.0), (2.0, 12.0), (3.0, 12.0), (1.0, 12.0), (1.0, 11.0), (0.0, 12.0), (0.0, 11.0)],
        [[0, 1, 2, 3], [2, 1, 4, 5], [5, 4, 6, 7], [5, 7, 9, 8], [8, 9, 11, 10], [8, 10, 12, 13],
         [13, 12, 14, 15], [3, 2, 13, 15]],
        [0, 0, 0, 0, 0, 0, 0, 1])
    arch.set_type_index_cost(1, 1.0)
    arch.add_island(Island(_t(0.0, 0.0), mesh))
    agent = Agent.create((0.5, 0.5), (0.0, 0.0), 0.5, 1.0, 1.0, cs=XY)
    assert agent.override_type_index_cost(1, 10.0)
    agent.current_target = (0.5, 11.5)
    agent_id = arch.add_agent(agent)
    arch.update(1.0)
    # the agent could go straight up, but with its overridden cost the detour to the right is cheaper
    v = np.float32([1.5, 0.5])
    expected = v * (np.float32(1.0) / np.sqrt(np.float32(v[0] * v[0] + v[1] * v[1])))  # glam normalize
    got = np.asarray(arch.get_agent(agent_id).get_desired_velocity(), dtype=np.float32)
    assert np.array_equal(got, expected), (got, expected)




def test_paused_agent_does_not_repath(make_backend):  # lib_test.rs:776-906
    arch = new_arch(make_backend())
    mesh = mesh_xy([(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (1.0, 2.0), (0.0, 2.0), (1.0, 3.0), (0.0, 3.0)],
                   [[0, 1, 2, 3], [3, 2, 4, 5], [5, 4, 6, 7]])
    island = arch.add_island(Island(_t(0.0, 0.0), mesh))
    a = Agent.create((0.5, 0.5), (0.0, 0.0), 0.5, 1.0, 1.0, cs=XY)
    a.current_target = (0.5, 1.5)
    agent = arch.add_agent(a)

    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island, 0), (island, 1))

    # pause the agent and move it — and its target — around
    am = arch.get_agent_mut(agent)
    am.paused = True
    am.position = (0.5, 1.5)
    am.current_target = (0.5, 2.5)
    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island, 0), (island, 1))  # unchanged
    assert arch.get_agent(agent).state() == AgentState.Paused

    # agent and target completely off the nav mesh
    am = arch.get_agent_mut(agent)
    am.position = (3.5, 1.5)
    am.current_target = (3.5, 2.5)
    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island, 0), (island, 1))

    # back onto the path, unpaused
    am = arch.get_agent_mut(agent)
    am.position = (0.5, 0.5)
    am.current_target = (0.5, 1.5)
    am.paused = False
    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island, 0), (island, 1))
    assert arch.get_agent(agent).state() == AgentState.Moving

    # paused and moved off the path
    am = arch.get_agent_mut(agent)
    am.position = (0.5, 1.5)
    am.current_target = (0.5, 2.5)
    am.paused = True
    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island, 0), (island, 1))
    assert arch.get_agent(agent).state() == AgentState.Paused

    # unpaused: the path finally changes
    arch.get_agent_mut(agent).paused = False
    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island, 1), (island, 2))
    assert arch.get_agent(agent).state() == AgentState.Moving


def test_paused_agent_path_is_removed_when_invalid(make_backend):  # lib_test.rs:908-1003
    arch = new_arch(make_backend())
    mesh = mesh_xy(*ONE_NODE)
    island_1 = arch.add_island(Island(_t(0.0, 0.0), mesh))
    island_2 = arch.add_island(Island(_t(0.0, 1.0), mesh))
    arch.add_island(Island(_t(1.0, 0.0), mesh))
    a = Agent.create((0.5, 0.5), (0.0, 0.0), 0.5, 1.0, 1.0, cs=XY)
    a.current_target = (0.5, 1.5)
    agent = arch.add_agent(a)

    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island_1, 0), (island_2, 0))
    assert arch.get_agent(agent).state() == AgentState.Moving

    am = arch.get_agent_mut(agent)
    am.paused = True
    am.position = (1.5, 0.5)
    arch.update(1.0)
    assert path_start_and_end(arch, agent) == ((island_1, 0), (island_2, 0))  # the path did not change
    assert arch.get_agent(agent).state() == AgentState.Paused

    # still paused, but an invalid path is removed
    arch.remove_island(island_2)
    arch.update(1.0)
    assert arch.agent_path(agent) is None
    assert arch.get_agent(agent).state() == AgentState.Paused


def test_agent_path_cleared_when_clearing_target(make_backend):  # lib_test.rs:1146-1173
    arch = new_arch(make_backend())
    arch.add_island(Island(_t(0.0, 0.0), mesh_xy(*TWO_NODE)))
    a = Agent.create((0.5, 0.5), (0.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    a.current_target = (0.5, 1.5)
    agent = arch.add_agent(a)
    arch.update(1.0)
    assert arch.agent_path(agent) is not None
    assert arch.get_agent(agent).state() == AgentState.Moving
    arch.get_agent_mut(agent).current_target = None
    arch.update(1.0)
    assert arch.agent_path(agent) is None
    assert arch.get_agent(agent).state() == AgentState.Idle


def test_agent_path_cleared_on_bad_target(make_backend):  # lib_test.rs:1175-1204
    arch = new_arch(make_backend())
    arch.add_island(Island(_t(0.0, 0.0), mesh_xy(*TWO_NODE)))
    a = Agent.create((0.5, 0.5), (0.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    a.current_target = (0.5, 1.5)
    agent = arch.add_agent(a)
    arch.update(1.0)
    assert arch.agent_path(agent) is not None
    assert arch.get_agent(agent).state() == AgentState.Moving
    arch.get_agent_mut(agent).current_target = (1.5, 1.5)  # off the nav mesh
    arch.update(1.0)
    assert arch.agent_path(agent) is None
    assert arch.get_agent(agent).state() == AgentState.TargetNotOnNavMesh


def test_agent_could_not_find_path(make_backend):  # lib_test.rs:1206-1232
    arch = new_arch(make_backend())
    mesh = mesh_xy(*ONE_NODE)
    arch.add_island(Island(_t(0.0, 0.0), mesh))
    arch.add_island(Island(_t(2.0, 0.0), mesh))
    a = Agent.create((0.5, 0.5), (0.0, 0.0), 0.5, 1.0, 2.0, cs=XY)
    a.current_target = (2.5, 0.5)  # not connected to the agent
    agent = arch.add_agent(a)
    arch.update(1.0)
    assert arch.agent_path(agent) is None
    assert arch.get_agent(agent).state() == AgentState.NoPath


def test_add_and_remove_agents_characters_islands():
    """lib_test.rs:22-94 (agents), 96-153 (characters), 488-526 (islands), 528-563 (island dirty flag):
    slot-map bookkeeping on the host; no backend is needed until `update`."""
    from landmass_b200 import Character
    from landmass_b200.archipelago import Archipelago, ArchipelagoOptions
    from oracle.oracle_py import OracleBackend
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=OracleBackend())
    zero = (0.0, 0.0, 0.0)
    for add, get, ids, remove, make in [
        (arch.add_agent, arch.get_agent, arch.get_agent_ids, arch.remove_agent,
         lambda r: Agent.create(zero, zero, r, 0.0, 0.0)),
        (arch.add_character, arch.get_character, arch.get_character_ids, arch.remove_character,
         lambda r: Character(position=zero, velocity=zero, radius=r)),
    ]:
        k1, k2, k3 = add(make(1.0)), add(make(2.0)), add(make(3.0))
        assert sorted(ids(), key=repr) == sorted([k1, k2, k3], key=repr)
        assert [get(k).radius for k in (k1, k2, k3)] == [1.0, 2.0, 3.0]
        remove(k2)
        assert sorted(ids(), key=repr) == sorted([k1, k3], key=repr)
        assert [get(k).radius for k in (k1, k3)] == [1.0, 3.0]
        remove(k3)
        assert list(ids()) == [k1] and get(k1).radius == 1.0
        remove(k1)
        assert list(ids()) == []

    empty = valid_mesh(dict(vertices=[], polygons=[], polygon_type_indices=[]))
    i1, i2, i3 = (arch.add_island(Island(Transform(zero, 0.0), empty)) for _ in range(3))
    assert sorted(arch.get_island_ids(), key=repr) == sorted([i1, i2, i3], key=repr)
    arch.remove_island(i2)
    assert sorted(arch.get_island_ids(), key=repr) == sorted([i1, i3], key=repr)

    assert arch.get_island(i1).dirty
    arch.update(0.01)
    assert not arch.get_island(i1).dirty
    arch.get_island_mut(i1).set_transform(Transform(zero, 0.0))
    assert arch.get_island(i1).dirty
    arch.update(0.01)
    assert not arch.get_island(i1).dirty
