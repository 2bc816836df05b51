"""A* core known answers — transcribed from crates/landmass/src/astar_test.rs:38-218.
These pin the oracle's model of Rust's BinaryHeap + the comparator at astar.rs:56-83."""
from oracle import oracle_py

W = 5
THREE = [([(10.0, 1, 1), (3.0, 2, 2)], 0.0), ([(10.0, 3, 0), (3.0, 4, 2)], 0.0),
         ([(4.0, 5, 1), (3.0, 6, 0)], 0.0)]


def grid(perfect):
    states = []
    for y in range(W):
        for x in range(W):
            i = len(states)
            adj = []
            if x > 0:
                adj.append((1.0, 1, i - 1))
            if x < W - 1:
                adj.append((1.0, 2, i + 1))
            if y > 0:
                adj.append((1.0, 3, i - W))
            if y < W - 1:
                adj.append((1.0, 4, i + W))
            states.append((adj, float(W * 2 - (x + y) - 2) if perfect else 0.0))
    return states


def test_finds_simple_path_no_heuristic():  # astar_test.rs:38-62
    path, explored = oracle_py.astar_adjacency(THREE, 0, 1)
    assert path == [2, 5] and explored == 3


def test_finds_no_path():  # astar_test.rs:64-89 (state 3 does not exist -> unreachable)
    path, explored = oracle_py.astar_adjacency(THREE + [([], 0.0)], 0, 3)
    assert path is None and explored == 3


def test_grid_no_heuristic():  # astar_test.rs:91-144
    path, explored = oracle_py.astar_adjacency(grid(False), 0, W * W - 1)
    assert sorted(path) == [2, 2, 2, 2, 4, 4, 4, 4]
    assert explored == W * W


def test_grid_perfect_heuristic():  # astar_test.rs:146-186 — pins the tie-break quirk
    path, explored = oracle_py.astar_adjacency(grid(True), 0, W * W - 1)
    assert path == [4, 4, 4, 4, 2, 2, 2, 2]
    assert explored == W * 2 - 1


def test_dead_end_wrong_heuristic():  # astar_test.rs:188-218
    states = [([(1.0, 1, 1), (2.0, 3, 2)], 3.0), ([(1.0, 2, 0)], 1.0),
              ([(1.0, 4, 0), (1.0, 1, 3)], 1.0), ([(1.0, 2, 2)], 0.0)]
    path, explored = oracle_py.astar_adjacency(states, 0, 3)
    assert path == [3, 1] and explored == 4
