"""The whole host API driven by a random script on two archipelagos at once — one on the CUDA
library, one on the oracle — and compared after every update.  The script does what a game does
between frames: agents appear and disappear (slot reuse), get new targets or lose them, are paused
and resumed, take animation links (`start_animation_link` / `end_animation_link` with a teleport to
the link's end point), change their reach conditions and cost overrides; characters wander; islands
are touched, removed and re-added; type costs change; animation links come and go.  After each
update: every agent's state, desired velocity (1e-4 relative), reached animation link and stored
corridor, and the frame's pathing results (agent, success, explored_nodes) must agree."""
import numpy as np
import pytest

from landmass_b200 import (Agent, AgentState, AnimationLink, Archipelago, ArchipelagoOptions, Character, Island,
                           PermittedAnimationLinks, TargetReachedCondition, Transform, synth)

pytestmark = pytest.mark.gpu

CELL = 8.0  # island size


class World:
    def __init__(self, backend, mesh):
        self.arch = Archipelago(ArchipelagoOptions.from_agent_radius(0.5), backend=backend)
        self.mesh = mesh
        self.islands = {}
        self.agents = []      # keys, in creation order (parallel in both worlds)
        self.chars = []
        self.links = []

    def add_island(self, cell):
        self.islands[cell] = self.arch.add_island(
            Island(Transform((CELL * cell[0], CELL * cell[1], 0.0), 0.0), self.mesh))


def apply_both(worlds, fn):
    for w in worlds:
        fn(w)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_script_in_lockstep(seed):
    from landmass_b200 import _abi
    from landmass_b200._native import NativeBackend
    from oracle.oracle_py import OracleBackend

    # seed 4 replays seed 1 with the agents taken in work order (LMB_CONFIG_ORDER_ALWAYS: longest corridor first in
    # k_follow, grid cell order in k_avoidance — what large agent counts run)
    flags = _abi.CONFIG_ORDER_ALWAYS if seed == 4 else 0
    seed = 1 if seed == 4 else seed

    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    mesh = synth.irregular_mesh(8, 8, seed=seed, straight_border=True, hole_fraction=0.06, relief=0.3)
    worlds = [World(NativeBackend(max_agents=256, flags=flags), mesh), World(OracleBackend(), mesh)]
    cells = [(x, y) for y in range(3) for x in range(3)]
    for c in cells:
        apply_both(worlds, lambda w, c=c: w.add_island(c))

    def rand_point():
        p = rng.uniform(0.3, 3 * CELL - 0.3, size=2)
        return (float(p[0]), float(p[1]), 0.0)

    def add_agent():
        pos, tgt = rand_point(), rand_point()
        radius, speed = float(rng.uniform(0.2, 0.45)), float(rng.uniform(0.8, 1.5))
        cond = [TargetReachedCondition.Distance, TargetReachedCondition.VisibleAtDistance,
                TargetReachedCondition.StraightPathDistance][int(rng.integers(3))](
            None if rng.uniform() < 0.5 else float(rng.uniform(0.5, 2.0)))
        permitted = None if rng.uniform() < 0.7 else frozenset(int(k) for k in rng.integers(0, 2, size=1))
        override = None if rng.uniform() < 0.7 else (int(rng.integers(3)), float(rng.uniform(0.5, 4.0)))

        def fn(w):
            a = Agent.create(pos, (0.0, 0.0, 0.0), radius, speed, 2.0)
            a.current_target = tgt
            a.target_reached_condition = cond
            a.permitted_animation_links = PermittedAnimationLinks(permitted)
            if override:
                a.override_type_index_cost(*override)
            w.agents.append(w.arch.add_agent(a))
        apply_both(worlds, fn)

    for _ in range(120):
        add_agent()
    for _ in range(4):
        pos = rand_point()
        apply_both(worlds, lambda w, pos=pos: w.chars.append(
            w.arch.add_character(Character(position=pos, velocity=(0.3, -0.2, 0.0), radius=0.5))))

    def add_link():
        a, b = rand_point(), rand_point()
        kind, cost, bidir = int(rng.integers(2)), float(rng.uniform(0.5, 3.0)), bool(rng.integers(2))
        link = AnimationLink(start_edge=(a, (a[0] + 1.2, a[1] + 0.2, 0.0)), end_edge=(b, (b[0] + 1.2, b[1] - 0.2, 0.0)),
                             cost=cost, kind=kind, bidirectional=bidir)
        apply_both(worlds, lambda w: w.links.append(w.arch.add_animation_link(link)))

    for _ in range(6):
        add_link()
    removed_cell = None
    dt = 0.1
    seen_states = set()
    total_results = 0
    for frame in range(40):
        # ---- edits between frames ----
        for _ in range(int(rng.integers(0, 4))):
            op = int(rng.integers(10))
            live = [i for i, k in enumerate(worlds[0].agents) if k is not None]
            i = live[int(rng.integers(len(live)))]
            if op == 0 and len(live) > 60:
                def fn(w, i=i):
                    w.arch.remove_agent(w.agents[i])
                    w.agents[i] = None
                apply_both(worlds, fn)
            elif op == 1:
                add_agent()
            elif op == 2:
                t = rand_point()
                apply_both(worlds, lambda w, i=i, t=t: setattr(w.arch.get_agent_mut(w.agents[i]), "current_target", t))
            elif op == 3:
                apply_both(worlds, lambda w, i=i: setattr(w.arch.get_agent_mut(w.agents[i]), "current_target", None))
            elif op == 4:
                apply_both(worlds, lambda w, i=i: setattr(w.arch.get_agent_mut(w.agents[i]), "paused",
                                                          not w.arch.get_agent(w.agents[i]).paused))
            elif op == 5:
                cell = cells[int(rng.integers(len(cells)))]
                if cell in worlds[0].islands:
                    apply_both(worlds, lambda w, cell=cell: w.arch.get_island_mut(w.islands[cell]).set_transform(
                        Transform((CELL * cell[0], CELL * cell[1], 0.0), 0.0)))
            elif op == 6:
                if removed_cell is None:
                    cell = cells[int(rng.integers(len(cells)))]
                    apply_both(worlds, lambda w, cell=cell: w.arch.remove_island(w.islands.pop(cell)))
                    removed_cell = cell
                else:
                    apply_both(worlds, lambda w, cell=removed_cell: w.add_island(cell))
                    removed_cell = None
            elif op == 7:
                t, c = int(rng.integers(3)), float(rng.uniform(0.5, 3.0))
                apply_both(worlds, lambda w, t=t, c=c: w.arch.set_type_index_cost(t, c))
            elif op == 8:
                if len(worlds[0].links) > 3 and rng.uniform() < 0.5:
                    j = int(rng.integers(len(worlds[0].links)))
                    apply_both(worlds, lambda w, j=j: w.arch.remove_animation_link(w.links.pop(j)))
                else:
                    add_link()
            else:
                pos = rand_point()
                apply_both(worlds, lambda w, i=i, pos=pos: setattr(w.arch.get_agent_mut(w.agents[i]), "position", pos))

        # ---- one frame ----
        for w in worlds:
            w.arch.update(dt)
        g, c = worlds
        gr = [(g.agents.index(r.agent), r.success, r.explored_nodes) for r in g.arch.get_pathing_results()]
        cr = [(c.agents.index(r.agent), r.success, r.explored_nodes) for r in c.arch.get_pathing_results()]
        assert gr == cr, f"frame {frame}: pathing results"
        total_results += len(gr)
        moves = []
        for i, (ka, kb) in enumerate(zip(g.agents, c.agents)):
            if ka is None:
                moves.append(None)
                continue
            a, b = g.arch.get_agent(ka), c.arch.get_agent(kb)
            assert a.state() == b.state(), f"frame {frame} agent {i}"
            seen_states.add(a.state())
            np.testing.assert_allclose(a.get_desired_velocity(), b.get_desired_velocity(), rtol=1e-4, atol=1e-6,
                                       err_msg=f"frame {frame} agent {i}")
            ra, rb = a.reached_animation_link(), b.reached_animation_link()
            assert (ra is None) == (rb is None)
            if ra is not None:
                assert ra.link_id == rb.link_id
                assert np.array_equal(ra.start_point, rb.start_point) and np.array_equal(ra.end_point, rb.end_point)
            pa, pb = g.arch.agent_path(ka), c.arch.agent_path(kb)
            assert (pa is None) == (pb is None), f"frame {frame} agent {i}"
            if pa is not None:
                assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1]), f"frame {frame} agent {i}"
            moves.append((np.asarray(a.get_desired_velocity(), dtype=np.float32), ra))
        # ---- the game moves its agents (by the CUDA world's velocities, for both) ----
        for w in worlds:
            for k, m in zip(w.agents, moves):
                if k is None:
                    continue
                ag = w.arch.get_agent_mut(k)
                v, reached = m
                if ag.is_using_animation_link():
                    ag.end_animation_link()
                elif reached is not None and ag.state() == AgentState.ReachedAnimationLink:
                    ag.start_animation_link()
                    ag.position = tuple(np.asarray(reached.end_point, dtype=np.float32))
                elif not ag.paused:
                    ag.position = tuple(np.asarray(ag.position, dtype=np.float32) + v * np.float32(dt))
                    ag.velocity = tuple(v)
            for k in w.chars:
                ch = w.arch.get_character(k)
                ch.position = tuple(np.asarray(ch.position, dtype=np.float32) + np.asarray(ch.velocity, dtype=np.float32) * np.float32(dt))
    assert {AgentState.Moving, AgentState.Paused, AgentState.Idle} <= seen_states
    assert total_results > 150
