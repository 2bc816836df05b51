"""Point sampling known answers — transcribed from the reference's tests:
crates/landmass/src/nav_mesh_test.rs:427-840 and nav_data_test.rs:32-132.
Run against the oracle (CPU) and, marked gpu, against the CUDA kernel through the C ABI."""
import math

import numpy as np
import pytest

from landmass_b200 import Island, Transform, _abi
from tests.golden import meshes
from tests.helpers import core_options, new_archipelago, valid_mesh


def sample(arch, point, opts):
    arch._sync_navdata()
    pts, nodes = arch.backend.sample_points(np.array([point], dtype=np.float32), opts)
    if nodes[0] == _abi.NONE:
        return None
    return tuple(float(x) for x in pts[0]), int(nodes[0])


def single(make_backend, spec):
    arch = new_archipelago(make_backend())
    arch.add_island(Island(Transform(), valid_mesh(spec)))
    return arch


def test_sample_point_returns_none_for_far_point(make_backend):  # nav_mesh_test.rs:426-510
    arch = single(make_backend, meshes.ORGANIC)
    o = core_options(0.1, 0.1, 0.1, 1.0)
    for p in [(-3.0, 0.0, 0.0), (6.0, 0.0, 0.0), (0.0, 0.0, -3.0), (0.0, 0.0, 2.0)]:
        assert sample(arch, p, o) is None


def test_sample_point_in_nodes(make_backend):  # nav_mesh_test.rs:512-612
    arch = single(make_backend, meshes.ORGANIC)
    o1 = core_options(1.0, 1.0, 1.0, 1.0)
    assert sample(arch, (1.5, 1.5, 0.95), o1) == ((1.5, 1.5, 0.0), 0)
    assert sample(arch, (1.5, 4.0, -0.95), o1) == ((1.5, 4.0, 0.0), 1)
    o5 = core_options(5.0, 5.0, 5.0, 1.0)
    assert sample(arch, (2.5, 3.5, -0.55), o5) == ((2.5, 3.5, -1.0), 3)
    p, n = sample(arch, (2.5, 4.5, 0.1), o5)
    assert n == 2 and tuple(round(x * 1000.0) / 1000.0 for x in p) == (2.5, 4.5, 0.5)


def test_sample_point_near_node(make_backend):  # nav_mesh_test.rs:614-702
    arch = single(make_backend, meshes.ORGANIC)
    o1 = core_options(1.0, 1.0, 1.0, 1.0)
    assert sample(arch, (-0.5, 1.5, 0.25), o1) == ((0.0, 1.5, 0.0), 0)
    assert sample(arch, (0.5, 5.5, -0.25), o1) == ((1.0, 5.0, 0.0), 1)
    assert sample(arch, (4.5, 1.5, -0.25), o1) == ((4.0, 1.5, 0.0), 0)
    o5 = core_options(5.0, 5.0, 5.0, 1.0)
    assert sample(arch, (2.5, 5.5, 0.5), o5) == ((2.5, 5.0, 0.5), 2)


def test_sample_ignores_closer_horizontal(make_backend):  # nav_mesh_test.rs:719-755
    arch = single(make_backend, meshes.STACKED)
    o = core_options(0.25, 5.0, 5.0, 1.0)
    p, n = sample(arch, (1.5, 0.5, 0.0), o)
    assert n == 1 and p == (1.5, 0.5, float(np.float32(-1.9)))


def test_sample_favoring_vertical(make_backend):  # nav_mesh_test.rs:757-791
    arch = single(make_backend, meshes.STACKED)
    o = core_options(5.0, 5.0, 5.0, 2.0)
    p, n = sample(arch, (2.0, 0.5, 0.0), o)
    assert n == 1 and p == (2.0, 0.5, float(np.float32(-1.9)))


def test_sample_filters_vertical_points_differently(make_backend):  # nav_mesh_test.rs:793-840
    arch = single(make_backend, meshes.UNIT_QUAD)
    o = core_options(100.0, 2.0, 1.0, 1.0)  # distance_above 1.0, distance_below 2.0
    assert sample(arch, (0.5, 0.5, -0.5), o) == ((0.5, 0.5, 0.0), 0)
    assert sample(arch, (0.5, 0.5, -1.5), o) is None
    assert sample(arch, (0.5, 0.5, 0.5), o) == ((0.5, 0.5, 0.0), 0)
    assert sample(arch, (0.5, 0.5, 1.5), o) == ((0.5, 0.5, 0.0), 0)
    assert sample(arch, (0.5, 0.5, 2.5), o) is None


def test_sample_point_uses_height_mesh_if_available(make_backend):  # nav_mesh_test.rs:1109-1163
    from tests.helpers import create_height_mesh

    hm = create_height_mesh(
        [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (1.0, 1.0, 0.0), (0.0, 2.0, 1.0),
         (1.0, 2.0, 1.0), (0.0, 3.0, 1.0), (1.0, 3.0, 1.0)],
        [[[0, 1, 3, 2], [2, 3, 5, 4], [4, 5, 7, 6]]])
    spec = dict(vertices=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 3.0, 1.0), (0.0, 3.0, 1.0)],
                polygons=[[0, 1, 2, 3]], polygon_type_indices=[0], height_mesh=hm)
    arch = single(make_backend, spec)
    o = core_options(100.0, 0.1, 0.1, 1.0)
    assert sample(arch, (0.5, 1.0, 0.0), o) == ((0.5, 1.0, 0.0), 0)
    assert sample(arch, (0.5, 2.0, 1.0), o) == ((0.5, 2.0, 1.0), 0)


def test_samples_points_multi_island(make_backend):  # nav_data_test.rs:32-132
    spec = dict(
        vertices=[(1.0, 1.0, 1.0), (2.0, 1.0, 1.0), (3.0, 1.0, 1.0), (4.0, 1.0, 1.0), (4.0, 2.0, 1.0),
                  (3.0, 2.0, 1.0), (2.0, 2.0, 1.0), (1.0, 2.0, 1.0)],
        polygons=[[0, 1, 6, 7], [1, 2, 5, 6], [2, 3, 4, 5]], polygon_type_indices=[0, 0, 0])
    mesh = valid_mesh(spec)
    arch = new_archipelago(make_backend())
    i1 = arch.add_island(Island(Transform((0.0, 0.0, 0.0), 0.0), mesh))
    i2 = arch.add_island(Island(Transform((5.0, 0.0, 0.1), math.pi * 0.5), mesh))
    o = core_options(0.1, 0.1, 0.1, 1.0)
    assert sample(arch, (1.5, 1.5, 1.09), o) == ((1.5, 1.5, 1.0), arch.node_id(i1, 0))
    assert sample(arch, (0.95, 0.95, 0.95), o) == ((1.0, 1.0, 1.0), arch.node_id(i1, 0))
    assert sample(arch, (3.5, 1.5, 1.04), o) == ((3.5, 1.5, 1.0), arch.node_id(i1, 2))
    p, n = sample(arch, (3.5, 1.5, 1.06), o)
    assert n == arch.node_id(i2, 0)
    assert tuple(round(x * 1e6) / 1e6 for x in p) == (3.5, 1.5, 1.1)
