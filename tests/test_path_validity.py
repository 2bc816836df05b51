"""`Path::is_valid` across navigation-data changes (path.rs:381-400; lib.rs:274-281, 350-358;
agent.rs:449-456), through whole updates and `lmb_navdata_update`.

The reference keeps an agent's path when the islands it crosses were neither deleted nor dirty and
none of its off-mesh links was dropped; every other path is re-planned (or, for a paused agent,
cleared).  path_test.rs:842-917 (`path_not_valid_for_invalidated_islands_or_off_mesh_links`) checks
that on a hand-built path of three island segments and two links; here the same shape of path
(islands 1 -> 2 -> 3 over boundary links) is planned by A*, the world then changes, and the next
update either reports a new pathing result or does not.  Node ids and link indices of the stored
path are re-expressed in the new flattening, which the read-back corridor shows."""
import numpy as np
import pytest

from landmass_b200 import (XY, Agent, AgentState, AnimationLink, Archipelago, ArchipelagoOptions, Island,
                           Transform, _abi)
from tests.helpers import valid_mesh
from tests.test_golden_animation_links import GAP_MESH

ONE = dict(vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0)], polygons=[[0, 1, 2, 3]],
           polygon_type_indices=[0])
TWO = dict(vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (2.0, 0.0, 0.0), (2.0, 1.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0)],
           polygons=[[0, 1, 4, 5], [1, 2, 3, 4]], polygon_type_indices=[0, 1])


def corridor(arch, aid):
    path = arch.agent_path(aid)
    if path is None:
        return None
    nodes, portals = path
    out = []
    for n, p in zip(nodes, portals):
        p = int(p)
        if p != _abi.NONE and p & _abi.PORTAL_LINK:
            l = p & 0x7FFFFFFF
            p = ("link", arch.node_ref(int(arch._flat["link_dst_node"][l])))
        out.append((arch.node_ref(int(n)), p))
    return out


def three_islands(make_backend, bystander_first):
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=make_backend())
    keys = {}
    if bystander_first:  # its removal then shifts every later island's node ids
        keys[4] = arch.add_island(Island(Transform((0.0, 10.0, 0.0), 0.0), valid_mesh(TWO)))
    keys[1] = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), valid_mesh(ONE)))
    keys[2] = arch.add_island(Island(Transform((1.0, 0.0, 0.0), 0.0), valid_mesh(TWO)))
    keys[3] = arch.add_island(Island(Transform((3.0, 0.0, 0.0), 0.0), valid_mesh(ONE)))
    if not bystander_first:
        keys[4] = arch.add_island(Island(Transform((0.0, 10.0, 0.0), 0.0), valid_mesh(TWO)))
    a = Agent.create((0.5, 0.5, 0.0), (0.0, 0.0, 0.0), 0.25, 1.0, 1.0)
    a.current_target = (3.5, 0.5, 0.0)
    aid = arch.add_agent(a)
    arch.update(0.01)
    assert len(arch.get_pathing_results()) == 1
    want = [((keys[1], 0), ("link", (keys[2], 0))), ((keys[2], 0), 1), ((keys[2], 1), ("link", (keys[3], 0))),
            ((keys[3], 0), _abi.NONE)]
    assert corridor(arch, aid) == want
    return arch, keys, aid, want


@pytest.mark.parametrize("bystander_first", [False, True])
@pytest.mark.parametrize("change", ["touch 4", "remove 4", "add 5", "type cost", "touch 1", "touch 2", "touch 3",
                                    "move 3 away"])
def test_path_survives_unrelated_changes_only(make_backend, change, bystander_first):
    arch, keys, aid, want = three_islands(make_backend, bystander_first)
    replans, reachable = False, True
    if change == "touch 4":
        arch.get_island_mut(keys[4]).set_transform(Transform((0.0, 10.0, 0.0), 0.0))
    elif change == "remove 4":
        arch.remove_island(keys[4])
    elif change == "add 5":
        arch.add_island(Island(Transform((-5.0, -5.0, 0.0), 0.0), valid_mesh(TWO)))
    elif change == "type cost":
        arch.set_type_index_cost(1, 3.0)
    elif change == "move 3 away":
        arch.get_island_mut(keys[3]).set_transform(Transform((30.0, 0.0, 0.0), 0.0))
        replans, reachable = False, False  # the target is off the mesh now: ClearPathBadTarget, no search
    else:  # path_test.rs:892-903: each island of the path, invalidated (dirty even without a real change)
        which = int(change.split()[1])
        t = {1: (0.0, 0.0, 0.0), 2: (1.0, 0.0, 0.0), 3: (3.0, 0.0, 0.0)}[which]
        arch.get_island_mut(keys[which]).set_transform(Transform(t, 0.0))
        replans = True
    arch.update(0.01)
    assert len(arch.get_pathing_results()) == (1 if replans else 0)
    if reachable:
        assert arch.get_agent(aid).state() == AgentState.Moving
        assert corridor(arch, aid) == want  # same corridor — kept and re-indexed, or found again
        np.testing.assert_allclose(arch.get_agent(aid).get_desired_velocity(), (1.0, 0.0, 0.0), atol=1e-6)
    else:
        assert arch.get_agent(aid).state() == AgentState.TargetNotOnNavMesh
        assert arch.agent_path(aid) is None


def gap_arch(make_backend):
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=make_backend())
    i = arch.add_island(Island(Transform((10.0, 10.0), 0.0), valid_mesh(GAP_MESH, cs=XY)))
    far = arch.add_island(Island(Transform((50.0, 10.0), 0.0), valid_mesh(GAP_MESH, cs=XY)))
    l = arch.add_animation_link(AnimationLink(start_edge=((10.0, 11.0), (11.0, 11.0)),
                                              end_edge=((10.0, 12.0), (11.0, 12.0)), cost=1.0, kind=0,
                                              bidirectional=False))
    a = Agent.create((10.5, 10.5), (0.0, 0.0), 0.25, 1.0, 1.0, cs=XY)
    a.current_target = (10.5, 12.5)
    a.animation_link_reached_distance = 0.01
    aid = arch.add_agent(a)
    arch.update(0.01)
    assert len(arch.get_pathing_results()) == 1
    assert corridor(arch, aid) == [((i, 0), ("link", (i, 4))), ((i, 4), _abi.NONE)]
    return arch, i, far, l, aid


def far_link():
    return AnimationLink(start_edge=((50.0, 11.0), (51.0, 11.0)), end_edge=((50.0, 12.0), (51.0, 12.0)), cost=1.0, kind=0,
                         bidirectional=False)


def test_dropped_animation_link_invalidates_the_path(make_backend):
    """path_test.rs:905-916: an invalidated off-mesh link with all islands intact.  Only a deleted
    animation link does that (nav_data.rs:380-387); the agent re-plans around the gap."""
    arch, i, far, l, aid = gap_arch(make_backend)
    arch.remove_animation_link(l)
    arch.update(0.01)
    assert len(arch.get_pathing_results()) == 1
    assert [n for n, _ in corridor(arch, aid)] == [(i, p) for p in (0, 1, 2, 3, 4)]


def test_other_animation_links_do_not_invalidate_the_path(make_backend):
    """A second animation link elsewhere is created (its off-mesh link sorts before nothing of ours but
    shifts link indices when removed) and removed again: the path through the first link is kept."""
    arch, i, far, l, aid = gap_arch(make_backend)
    other = arch.add_animation_link(far_link())
    arch.update(0.01)
    assert arch.get_pathing_results() == []
    assert corridor(arch, aid) == [((i, 0), ("link", (i, 4))), ((i, 4), _abi.NONE)]
    arch.remove_animation_link(other)
    arch.update(0.01)
    assert arch.get_pathing_results() == []
    assert corridor(arch, aid) == [((i, 0), ("link", (i, 4))), ((i, 4), _abi.NONE)]
    assert arch.get_agent(aid).state() == AgentState.Moving


def test_link_indices_are_remapped(make_backend):
    """The far island's link exists first (lower source node... higher, so ours is link 0); adding a
    link on an island that sorts BEFORE ours moves our link to a higher index."""
    arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5, cs=XY), cs=XY, backend=make_backend())
    first = arch.add_island(Island(Transform((50.0, 10.0), 0.0), valid_mesh(GAP_MESH, cs=XY)))
    i = arch.add_island(Island(Transform((10.0, 10.0), 0.0), valid_mesh(GAP_MESH, cs=XY)))
    arch.add_animation_link(AnimationLink(start_edge=((10.0, 11.0), (11.0, 11.0)), end_edge=((10.0, 12.0), (11.0, 12.0)),
                                          cost=1.0, kind=0, bidirectional=False))
    a = Agent.create((10.5, 10.5), (0.0, 0.0), 0.25, 1.0, 1.0, cs=XY)
    a.current_target = (10.5, 12.5)
    a.animation_link_reached_distance = 0.01
    aid = arch.add_agent(a)
    arch.update(0.01)
    _, portals = arch.agent_path(aid)
    assert int(portals[0]) == _abi.PORTAL_LINK | 0
    arch.add_animation_link(far_link())
    arch.update(0.01)
    assert arch.get_pathing_results() == []
    _, portals = arch.agent_path(aid)
    assert int(portals[0]) == _abi.PORTAL_LINK | 1
    assert corridor(arch, aid) == [((i, 0), ("link", (i, 4))), ((i, 4), _abi.NONE)]
    # and the agent still walks to the link and reaches it
    for _ in range(200):
        ag = arch.get_agent(aid)
        if ag.state() == AgentState.ReachedAnimationLink:
            break
        ag.position = tuple(np.asarray(ag.position) + np.asarray(ag.get_desired_velocity()) * 0.01)
        arch.update(0.01)
    assert arch.get_agent(aid).state() == AgentState.ReachedAnimationLink


@pytest.mark.parametrize("touch", [4, 2])
def test_paused_agent_keeps_a_valid_path_and_loses_an_invalid_one(make_backend, touch):
    """lib.rs:350-358.  (lib_test.rs:908-1003 is the reference's own test of the invalid half.)"""
    arch, keys, aid, want = three_islands(make_backend, bystander_first=True)
    arch.get_agent_mut(aid).paused = True
    t = {4: (0.0, 10.0, 0.0), 2: (1.0, 0.0, 0.0)}[touch]
    arch.get_island_mut(keys[touch]).set_transform(Transform(t, 0.0))
    arch.update(0.01)
    assert arch.get_pathing_results() == []
    assert arch.get_agent(aid).state() == AgentState.Paused
    assert corridor(arch, aid) == (want if touch == 4 else None)
    arch.get_agent_mut(aid).paused = False
    arch.update(0.01)
    assert len(arch.get_pathing_results()) == (0 if touch == 4 else 1)
    assert corridor(arch, aid) == want


def test_no_target_with_an_invalid_path_goes_idle(make_backend):
    """agent.rs:432-438 comes before the validity check: ClearPathNoTarget (state Idle), not DoNothing."""
    arch, keys, aid, want = three_islands(make_backend, bystander_first=False)
    arch.get_agent_mut(aid).current_target = None
    arch.get_island_mut(keys[2]).set_transform(Transform((1.0, 0.0, 0.0), 0.0))
    arch.update(0.01)
    assert arch.get_agent(aid).state() == AgentState.Idle and arch.agent_path(aid) is None
    assert arch.get_pathing_results() == []


@pytest.mark.gpu
def test_changing_world_matches_the_oracle_frame_by_frame():
    """Two archipelagos — CUDA and oracle — get the same edits: a 4x4 tiling of 6x6-cell islands joined
    by boundary links, 300 agents walking to random targets, and every third frame an island is
    touched, removed or put back.  States, pathing results (agent, success, explored_nodes),
    corridors and desired velocities must agree on every frame; the pathing results also show that
    only agents whose corridor crossed a changed island re-plan."""
    from landmass_b200 import synth
    from landmass_b200._native import NativeBackend
    from oracle.oracle_py import OracleBackend

    mesh = synth.grid_mesh(6, 6)
    rng = np.random.Generator(np.random.PCG64(5))
    worlds = []
    for backend in (NativeBackend(max_agents=512), OracleBackend()):
        arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=backend)
        islands = {(ix, iy): arch.add_island(Island(Transform((6.0 * ix, 6.0 * iy, 0.0), 0.0), mesh))
                   for iy in range(4) for ix in range(4)}
        worlds.append((arch, islands, []))
    starts = rng.uniform(0.2, 23.8, size=(300, 2))
    targets = rng.uniform(0.2, 23.8, size=(300, 2))
    for arch, _, agents in worlds:
        for s, t in zip(starts, targets):
            a = Agent.create((s[0], s[1], 0.0), (0.0, 0.0, 0.0), 0.3, 1.5, 2.0)
            a.current_target = (t[0], t[1], 0.0)
            agents.append(arch.add_agent(a))
    removed = None
    replans = []
    for frame in range(24):
        if frame and frame % 3 == 0:
            kind = (frame // 3) % 3
            cells = sorted(worlds[0][1])
            cell = cells[int(rng.integers(len(cells)))]
            for arch, islands, _ in worlds:
                where = Transform((6.0 * cell[0], 6.0 * cell[1], 0.0), 0.0)
                if kind == 0:    # dirty without a real change
                    arch.get_island_mut(islands[cell]).set_transform(where)
                elif kind == 1:  # gone
                    arch.remove_island(islands.pop(cell))
                else:            # the island removed three frames ago comes back (as a new island)
                    back = Transform((6.0 * removed[0], 6.0 * removed[1], 0.0), 0.0)
                    islands[removed] = arch.add_island(Island(back, mesh))
            if kind == 1:
                removed = cell
        for arch, _, _ in worlds:
            arch.update(1.0 / 30.0)
        (g, _, ga), (c, _, ca) = worlds
        gr = [(ga.index(r.agent), r.success, r.explored_nodes) for r in g.get_pathing_results()]
        cr = [(ca.index(r.agent), r.success, r.explored_nodes) for r in c.get_pathing_results()]
        assert gr == cr, frame
        replans.append(len(gr))
        for ka, kb in zip(ga, ca):
            a, b = g.get_agent(ka), c.get_agent(kb)
            assert a.state() == b.state(), frame
            np.testing.assert_allclose(a.get_desired_velocity(), b.get_desired_velocity(), rtol=1e-4, atol=1e-6)
            pa, pb = g.agent_path(ka), c.agent_path(kb)
            assert (pa is None) == (pb is None)
            if pa is not None:
                assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
        # both worlds move by the CUDA velocities, so their inputs stay identical
        vels = [np.asarray(g.get_agent(k).get_desired_velocity()) for k in ga]
        for arch, _, agents in worlds:
            for k, v in zip(agents, vels):
                ag = arch.get_agent_mut(k)
                ag.position = tuple(np.asarray(ag.position, dtype=np.float32) + v * np.float32(1.0 / 30.0))
                ag.velocity = tuple(v)
    assert replans[0] == 300
    assert all(0 < r < 300 for r in replans[3::3]), replans  # some, never all, agents re-plan after an edit
