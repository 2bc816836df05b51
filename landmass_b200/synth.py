"""Synthetic navmeshes and agent populations for the BASELINE.json configs
(SURVEY.md §8d).  Data only: deterministic generators shared by bench.py, smoke() and
the parity tests, so the CUDA path and the CPU oracle see identical inputs.

Common conventions (SURVEY.md §8d "Common inputs"): XYZ coordinates, z = 0, island
rotation 0, unit cell = 1.0, quads CCW [v(x,y), v(x+1,y), v(x+1,y+1), v(x,y+1)],
polygon index = row-major over open cells, seed 0x1A2B3C4D.
"""
from __future__ import annotations

import numpy as np

from .nav_mesh import NavigationMesh, ValidNavigationMesh

SEED = 0x1A2B3C4D


def _quads_from_open(open_mask: np.ndarray, cell: float = 1.0, types=None) -> NavigationMesh:
    """open_mask[y, x] -> one quad per open cell, shared vertices on the full lattice."""
    h, w = open_mask.shape
    ys, xs = np.nonzero(open_mask)  # row-major
    vx, vy = np.meshgrid(np.arange(w + 1, dtype=np.float32), np.arange(h + 1, dtype=np.float32))
    verts = np.stack([vx.ravel() * np.float32(cell), vy.ravel() * np.float32(cell),
                      np.zeros((h + 1) * (w + 1), dtype=np.float32)], axis=1)
    v00 = ys * (w + 1) + xs
    polys = np.stack([v00, v00 + 1, v00 + (w + 1) + 1, v00 + (w + 1)], axis=1)
    t = np.zeros(len(polys), dtype=np.int64) if types is None else np.asarray(types)[ys, xs]
    return _FastQuadMesh(verts, polys, t)


class _FastQuadMesh(NavigationMesh):
    """NavigationMesh whose polygons are held as an (P,4) array (validate() accepts any
    sequence of sequences; this just avoids building 500k Python lists)."""

    def __init__(self, verts, polys, types):
        super().__init__(vertices=verts, polygons=polys, polygon_type_indices=types)


def grid_mesh(nx: int, ny: int, cell: float = 1.0, types=None) -> ValidNavigationMesh:
    """Open nx x ny quad grid (configs 1, 3, 4)."""
    return _quads_from_open(np.ones((ny, nx), dtype=bool), cell, types).validate()


def maze_open_mask(rooms: int, seed: int = SEED) -> np.ndarray:
    """Perfect maze with rooms x rooms rooms on a (2*rooms+1)^2 fine grid: rooms at odd
    coordinates, walls carved by a randomised iterative DFS (configs 2, 5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = 2 * rooms + 1
    open_mask = np.zeros((g, g), dtype=bool)
    visited = np.zeros((rooms, rooms), dtype=bool)
    # pre-draw random direction orders in bulk for speed
    perms = np.array([[0, 1, 2, 3], [0, 1, 3, 2], [0, 2, 1, 3], [0, 2, 3, 1], [0, 3, 1, 2], [0, 3, 2, 1],
                      [1, 0, 2, 3], [1, 0, 3, 2], [1, 2, 0, 3], [1, 2, 3, 0], [1, 3, 0, 2], [1, 3, 2, 0],
                      [2, 0, 1, 3], [2, 0, 3, 1], [2, 1, 0, 3], [2, 1, 3, 0], [2, 3, 0, 1], [2, 3, 1, 0],
                      [3, 0, 1, 2], [3, 0, 2, 1], [3, 1, 0, 2], [3, 1, 2, 0], [3, 2, 0, 1], [3, 2, 1, 0]])
    draws = rng.integers(0, 24, size=rooms * rooms * 2 + 16)
    di = 0
    dx = (1, -1, 0, 0)
    dy = (0, 0, 1, -1)
    stack = [(0, 0)]
    visited[0, 0] = True
    open_mask[1, 1] = True
    while stack:
        x, y = stack[-1]
        moved = False
        order = perms[draws[di % len(draws)]]
        di += 1
        for d in order:
            nx_, ny_ = x + dx[d], y + dy[d]
            if 0 <= nx_ < rooms and 0 <= ny_ < rooms and not visited[ny_, nx_]:
                visited[ny_, nx_] = True
                open_mask[2 * y + 1 + dy[d], 2 * x + 1 + dx[d]] = True
                open_mask[2 * ny_ + 1, 2 * nx_ + 1] = True
                stack.append((nx_, ny_))
                moved = True
                break
        if not moved:
            stack.pop()
    return open_mask


def maze_mesh(rooms: int, seed: int = SEED) -> ValidNavigationMesh:
    return _quads_from_open(maze_open_mask(rooms, seed)).validate()


def loopy_maze_mesh(rooms: int, extra_openings: int, seed: int = SEED) -> ValidNavigationMesh:
    """A perfect maze with `extra_openings` more walls removed (each closes one cycle), so a search
    can enter some polygons through more than one edge."""
    mask = maze_open_mask(rooms, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    walls = [(2 * y + 1, 2 * x + 2) for y in range(rooms) for x in range(rooms - 1)]
    walls += [(2 * y + 2, 2 * x + 1) for y in range(rooms - 1) for x in range(rooms)]
    closed = [w for w in walls if not mask[w]]
    for i in rng.permutation(len(closed))[:extra_openings]:
        mask[closed[i]] = True
    return _quads_from_open(mask).validate()


def rosette_mesh(n_sides: int = 12, rings: int = 3) -> ValidNavigationMesh:
    """A regular n-gon surrounded by `rings` rings of n quads (all convex, CCW): a polygon with more
    than 8 edges forces the CSR state ids of the A* kernel."""
    ang = 2.0 * np.pi * np.arange(n_sides) / n_sides
    verts = []
    for r in range(rings + 1):
        rad = 2.0 + 2.0 * r
        verts += [(rad * np.cos(a), rad * np.sin(a), 0.0) for a in ang]
    polys = [list(range(n_sides))]
    types = [0]
    for r in range(rings):
        a0, b0 = r * n_sides, (r + 1) * n_sides
        for i in range(n_sides):
            j = (i + 1) % n_sides
            polys.append([a0 + i, b0 + i, b0 + j, a0 + j])
            types.append(0)
    return NavigationMesh(vertices=np.array(verts, dtype=np.float32), polygons=polys,
                          polygon_type_indices=types).validate()


def irregular_mesh(nx: int, ny: int, seed: int = SEED, jitter: float = 0.2, hole_fraction: float = 0.08,
                   relief: float = 0.6, straight_border: bool = False) -> ValidNavigationMesh:
    """A mesh none of whose arithmetic is exact: an nx x ny lattice with every vertex moved by up to
    `jitter` cells, heights from a smooth bump field (up to `relief`), about `hole_fraction` of the
    cells removed, the others kept as quads or cut into two triangles along a random diagonal, and
    polygon types 0..2 in blotches.  With `straight_border` the outer ring of vertices is not moved
    (and lies at z = 0), so copies of the mesh tile with coinciding borders."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gx, gy = np.meshgrid(np.arange(nx + 1, dtype=np.float64), np.arange(ny + 1, dtype=np.float64))
    dx = rng.uniform(-jitter, jitter, size=gx.shape)
    dy = rng.uniform(-jitter, jitter, size=gx.shape)
    z = relief * 0.5 * (np.sin(gx * 0.37 + 0.3) * np.cos(gy * 0.29 - 0.2) + 0.35 * np.sin(gx * 0.9 + gy * 0.7))
    if straight_border:
        edge = (gx == 0) | (gx == nx) | (gy == 0) | (gy == ny)
        dx[edge] = dy[edge] = 0.0
        z[edge] = 0.0
    verts = np.stack([(gx + dx).ravel(), (gy + dy).ravel(), z.ravel()], axis=1).astype(np.float32)
    hole = rng.uniform(size=(ny, nx)) < hole_fraction
    hole[0, :] = hole[-1, :] = hole[:, 0] = hole[:, -1] = False
    cut = rng.integers(0, 3, size=(ny, nx))  # 0 quad, 1 / 2 the two diagonals
    blotch = (np.floor(gx[:-1, :-1] / 3.0) + 2 * np.floor(gy[:-1, :-1] / 4.0)).astype(np.int64) % 3
    polys, types = [], []
    for y in range(ny):
        for x in range(nx):
            if hole[y, x]:
                continue
            v00 = y * (nx + 1) + x
            a, b, c, d = v00, v00 + 1, v00 + (nx + 1) + 1, v00 + (nx + 1)
            t = int(blotch[y, x])
            if cut[y, x] == 0:
                polys.append([a, b, c, d])
                types.append(t)
            elif cut[y, x] == 1:
                polys.extend([[a, b, c], [a, c, d]])
                types.extend([t, t])
            else:
                polys.extend([[a, b, d], [b, c, d]])
                types.extend([t, t])
    return NavigationMesh(vertices=verts, polygons=polys, polygon_type_indices=types).validate()


def random_points_in_polygons(mesh: ValidNavigationMesh, n: int, rng, lift: float = 0.0):
    """Uniform over polygons of any shape: a random convex combination of the polygon's first
    triangle fan, `lift` above the surface.  Returns (points (n,3) f32, polygon index (n,))."""
    p = rng.integers(0, mesh.n_polys, size=n)
    pts = np.zeros((n, 3), dtype=np.float32)
    for i, poly in enumerate(p):
        v = mesh.vertices[mesh.polygon_vertices(int(poly))]
        k = int(rng.integers(1, len(v) - 1))
        w = rng.dirichlet([1.0, 1.0, 1.0])
        w = 0.9 * w + 0.1 / 3.0  # stay off the edges
        pts[i] = (w[0] * v[0] + w[1] * v[k] + w[2] * v[k + 1]).astype(np.float32)
    pts[:, 2] += np.float32(lift)
    return pts, p


def random_points_on_mesh(mesh: ValidNavigationMesh, n: int, rng, lo: float = 0.1, hi: float = 0.9):
    """Uniform over polygons (axis-aligned quads): polygon min corner + uniform offset in
    [lo, hi]^2 of its extent.  Returns (points (n,3) f32, polygon index (n,))."""
    p = rng.integers(0, mesh.n_polys, size=n)
    b = mesh.poly_bounds[p]
    u = rng.uniform(lo, hi, size=(n, 2)).astype(np.float32)
    pts = np.zeros((n, 3), dtype=np.float32)
    pts[:, 0] = b[:, 0] + (b[:, 3] - b[:, 0]) * u[:, 0]
    pts[:, 1] = b[:, 1] + (b[:, 4] - b[:, 1]) * u[:, 1]
    pts[:, 2] = b[:, 2]
    return pts, p


def make_agents(mesh: ValidNavigationMesh, n: int, seed: int = SEED, radius: float = 0.5,
                desired_speed: float = 1.0, max_speed: float = 2.0, slot_base: int = 0) -> dict:
    """Agent SoA (the `lmb_agents_in` arrays) with uniform positions and targets."""
    from . import _abi

    rng = np.random.Generator(np.random.PCG64(seed))
    pos, _ = random_points_on_mesh(mesh, n, rng)
    tgt, _ = random_points_on_mesh(mesh, n, rng)
    return {
        "n": n,
        "slot": np.arange(slot_base, slot_base + n, dtype=np.uint32),
        "flags": np.full(n, _abi.AGENT_HAS_TARGET | _abi.AGENT_PERMIT_ALL_LINKS | _abi.AGENT_RESET,
                         dtype=np.uint32),
        "position": pos,
        "velocity": np.zeros((n, 3), dtype=np.float32),
        "target": tgt,
        "radius": np.full(n, radius, dtype=np.float32),
        "desired_speed": np.full(n, desired_speed, dtype=np.float32),
        "max_speed": np.full(n, max_speed, dtype=np.float32),
    }


def single_island_flat(mesh: ValidNavigationMesh) -> dict:
    """Flattened navdata (dict for _abi.make_navdata_desc) of one island at the origin."""
    from .nav_data import Island, NavigationData, Transform

    nd = NavigationData()
    nd.add_island(Island(Transform(), mesh))
    nd.update(0.01, 0.25)
    return nd.flatten()


def default_options(radius: float = 0.5):
    from .archipelago import ArchipelagoOptions

    return ArchipelagoOptions.from_agent_radius(radius).to_abi()


def tiled_archipelago_flat(k: int, g: int, type_block: int = 16, costs=((0, 1.0), (1, 2.0), (2, 5.0)),
                           cell: float = 1.0) -> dict:
    """BASELINE config 4's world: k x k islands, each a g x g unit-quad grid, tiled by pure translations so that
    borders coincide (boundary links at edge_link_distance 0.01), polygon types (x/type_block + y/type_block) mod 3
    with costs {0: 1, 1: 2, 2: 5}.  Returns the flattened navigation data (island linking done on the host,
    nav_data.rs:326-514)."""
    from .nav_data import Island, NavigationData, Transform

    types = ((np.arange(g)[None, :] // type_block + np.arange(g)[:, None] // type_block) % 3)
    mesh = grid_mesh(g, g, cell=cell, types=types)
    nd = NavigationData()
    for iy in range(k):
        for ix in range(k):
            nd.add_island(Island(Transform(translation=(ix * g * cell, iy * g * cell, 0.0), rotation=0.0), mesh))
    for t, c in costs:
        nd.set_type_index_cost(t, c)
    nd.update(0.01, 0.25)
    return nd.flatten()


def uniform_agents(n: int, span_x: float, span_y: float, seed: int = SEED, target_reach: float | None = None,
                   radius: float = 0.5) -> dict:
    """n agents uniform over [0.1, span - 0.1]^2 with uniform targets (or targets within `target_reach` of the
    agent); same per-agent parameters as make_agents."""
    from . import _abi

    rng = np.random.Generator(np.random.PCG64(seed))
    pos = np.zeros((n, 3), dtype=np.float32)
    tgt = np.zeros((n, 3), dtype=np.float32)
    hi = np.array([span_x - 0.1, span_y - 0.1])
    pos[:, :2] = rng.uniform([0.1, 0.1], hi, size=(n, 2))
    if target_reach is None:
        tgt[:, :2] = rng.uniform([0.1, 0.1], hi, size=(n, 2))
    else:
        tgt[:, :2] = np.clip(pos[:, :2] + rng.uniform(-target_reach, target_reach, size=(n, 2)), 0.1, hi)
    return {
        "n": n,
        "slot": np.arange(n, dtype=np.uint32),
        "flags": np.full(n, _abi.AGENT_HAS_TARGET | _abi.AGENT_RESET | _abi.AGENT_PERMIT_ALL_LINKS, dtype=np.uint32),
        "position": pos,
        "velocity": np.zeros((n, 3), dtype=np.float32),
        "target": tgt,
        "radius": np.full(n, radius, dtype=np.float32),
        "desired_speed": np.full(n, 1.0, dtype=np.float32),
        "max_speed": np.full(n, 2.0, dtype=np.float32),
    }
