"""`NavigationMesh::validate` known answers — transcribed from
crates/landmass/src/nav_mesh_test.rs:16-424 (bounds, derived polygons, every ValidationError,
connectivity, boundary edges, regions), 703-717 (edge indices) and 885-1106 (height-mesh errors).
Validation is host-side (landmass_b200/nav_mesh.py); its output arrays are what the kernels read."""
import numpy as np
import pytest

from landmass_b200 import XYZ, HeightNavigationMesh, HeightPolygon, NavigationMesh, ValidationError
from tests.helpers import create_height_mesh

TWO_TRIANGLES = [(0.0, 0.0, 0.0), (1.0, 0.0, 1.0), (2.0, 1.0, 0.0), (0.5, 3.0, 0.5), (0.75, 4.0, -0.25),
                 (0.25, 4.0, 0.0)]
TRIANGLE = [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0)]


class FlippedXYZ(XYZ):  # nav_mesh_test.rs:137-152
    FLIP_POLYGONS = True


def error_of(**mesh):
    with pytest.raises(ValidationError) as e:
        NavigationMesh(**mesh).validate()
    return e.value.kind, e.value.args_


def test_validation_computes_bounds():  # nav_mesh_test.rs:16-38, 40-59
    m = NavigationMesh(TWO_TRIANGLES, [[0, 1, 2], [3, 4, 5]], [0, 0]).validate()
    assert m.mesh_bounds.tolist() == [0.0, 0.0, -0.25, 2.0, 4.0, 1.0]
    m = NavigationMesh([(-1.0, -1.0, -1.0), (1.0, 0.0, 1.0), (2.0, 1.0, 0.0)], [[0, 1, 2]], [0]).validate()
    assert m.mesh_bounds.tolist() == [-1.0, -1.0, -1.0, 2.0, 1.0, 1.0]


def test_polygons_derived_and_vertices_copied():  # nav_mesh_test.rs:61-106
    m = NavigationMesh(TWO_TRIANGLES, [[0, 1, 2], [3, 4, 5]], [1337, 123]).validate()
    assert np.array_equal(m.vertices, np.array(TWO_TRIANGLES, dtype=np.float32))
    assert [m.polygon_vertices(p).tolist() for p in range(2)] == [[0, 1, 2], [3, 4, 5]]
    assert [m.connectivity(p) for p in range(2)] == [[None] * 3, [None] * 3]
    assert m.poly_region.tolist() == [0, 1] and m.poly_type.tolist() == [1337, 123]
    assert m.poly_bounds.tolist() == [[0.0, 0.0, 0.0, 2.0, 1.0, 1.0], [0.25, 3.0, -0.25, 0.75, 4.0, 0.5]]
    three = np.float32(3.0)
    assert np.array_equal(m.poly_center, np.array([[3.0, 1.0, 1.0], [1.5, 11.0, 0.25]], dtype=np.float32) / three)


def test_validation_errors():  # nav_mesh_test.rs:108-134, 154-178, 198-308
    assert error_of(vertices=[(0.0, 0.0, 0.0), (1.0, 1.0, 0.0), (1.0, 0.0, 0.0)], polygons=[[0, 1, 2]],
                    polygon_type_indices=[0, 0]) == ("TypeIndicesHaveWrongLength", (1, 2))
    assert error_of(vertices=[(0.0, 0.0, 0.0), (1.0, 1.0, 0.0), (1.0, 0.0, 0.0)], polygons=[[0, 1, 2]],
                    polygon_type_indices=[0]) == ("ConcavePolygon", (0,))
    assert error_of(vertices=[(0.0, 0.0, 0.0), (1.0, 1.0, 0.0)], polygons=[[0, 1]],
                    polygon_type_indices=[0]) == ("NotEnoughVerticesInPolygon", (0,))
    assert error_of(vertices=TRIANGLE, polygons=[[0, 1, 3]],
                    polygon_type_indices=[0]) == ("InvalidVertexIndexInPolygon", (0,))
    assert error_of(vertices=TRIANGLE, polygons=[[0, 1, 1, 2]],
                    polygon_type_indices=[0]) == ("DegenerateEdgeInPolygon", (0,))
    assert error_of(vertices=TRIANGLE + [(2.0, 0.0, 0.0), (2.0, 1.0, 0.0), (2.0, 0.0, 1.0), (2.0, 1.0, 1.0)],
                    polygons=[[0, 1, 2], [1, 3, 4, 2], [1, 5, 6, 2]],
                    polygon_type_indices=[0, 0, 0]) == ("DoublyConnectedEdge", (1, 2))


def test_no_error_on_concave_polygon_with_flipped_coordinate_system():  # nav_mesh_test.rs:180-195
    NavigationMesh([(0.0, 0.0, 0.0), (1.0, 1.0, 0.0), (1.0, 0.0, 0.0)], [[0, 1, 2]], [0], cs=FlippedXYZ).validate()


def test_derives_connectivity_and_boundary_edges():  # nav_mesh_test.rs:310-385
    m = NavigationMesh(
        TRIANGLE + [(2.0, 0.0, 0.0), (2.0, 1.0, 0.0), (3.0, 0.0, 1.0), (3.0, 1.0, 1.0), (1.0, 2.0, 1.0), (2.0, 2.0, 1.0)],
        [[0, 1, 2], [1, 3, 4, 2], [3, 5, 6, 4], [2, 4, 8, 7]], [0, 0, 0, 0]).validate()
    assert [m.connectivity(p) for p in range(4)] == [
        [None, (1, 3), None],
        [None, (2, 3), (3, 0), (0, 1)],
        [None, None, None, (1, 1)],
        [(1, 2), None, None, None],
    ]
    assert sorted(map(tuple, m.boundary_edges.tolist())) == [(0, 0), (0, 2), (1, 0), (2, 0), (2, 1), (2, 2), (3, 1),
                                                             (3, 2), (3, 3)]


def test_finds_regions():  # nav_mesh_test.rs:387-424
    m = NavigationMesh(
        [(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (1.0, 1.0, 0.0), (0.0, 1.0, 0.0), (1.0, 2.0, 0.0), (0.0, 2.0, 0.0),
         (1.0, 3.0, 0.0), (0.0, 3.0, 0.0),
         (2.0, 0.0, 0.0), (3.0, 0.0, 0.0), (3.0, 1.0, 0.0), (2.0, 1.0, 0.0), (3.0, 2.0, 0.0), (2.0, 2.0, 0.0)],
        [[0, 1, 2, 3], [3, 2, 4, 5], [5, 4, 6, 7], [8, 9, 10, 11], [11, 10, 12, 13]], [0] * 5).validate()
    assert m.poly_region.tolist() == [0, 0, 0, 1, 1]


def test_valid_polygon_gets_edge_indices():  # nav_mesh_test.rs:703-717
    verts = [(np.cos(a), np.sin(a), 0.0) for a in np.linspace(0.0, 2 * np.pi, 11)[:10]]
    # a convex pentagon over vertices 1, 3, 5, 7, 9 of a decagon (the reference's index list
    # [1, 3, 9, 2, 7] is not a valid polygon; only the (left, right) = (next, this) rule is under test)
    m = NavigationMesh(verts, [[1, 3, 5, 7, 9]], [0]).validate()
    assert m.get_edge_indices(0, 0) == (3, 1)
    assert m.get_edge_indices(0, 2) == (7, 5)
    assert m.get_edge_indices(0, 4) == (1, 9)


def test_height_mesh_errors():  # nav_mesh_test.rs:884-1081
    square = TRIANGLE + [(0.0, 1.0, 0.0)]
    tri_mesh = dict(vertices=TRIANGLE, polygons=[[0, 1, 2]], polygon_type_indices=[0])
    assert error_of(**tri_mesh, height_mesh=create_height_mesh(square, [[[0, 1, 2]], [[2, 3, 0]]])) == \
        ("IncorrectNumberOfPolygons", (1, 2))
    assert error_of(**tri_mesh, height_mesh=HeightNavigationMesh(
        vertices=TRIANGLE, triangles=[[0, 1, 2]], polygons=[HeightPolygon(0, 3, 0, 2)])) == \
        ("InvalidPolygonIndices", (0,))
    assert error_of(**tri_mesh, height_mesh=HeightNavigationMesh(
        vertices=TRIANGLE, triangles=[[0, 1, 2]], polygons=[HeightPolygon(1, 3, 0, 1)])) == \
        ("InvalidIndexInTriangle", (0, 3))
    assert error_of(**tri_mesh, height_mesh=HeightNavigationMesh(
        vertices=square, triangles=[[0, 1, 2], [2, 3, 0]], polygons=[HeightPolygon(0, 3, 0, 2)])) == \
        ("InvalidIndexInTriangle", (1, 3))
    assert error_of(**tri_mesh, height_mesh=create_height_mesh(TRIANGLE, [[[2, 1, 0]]])) == \
        ("ClockwiseTriangle", (0,))


def test_no_error_on_concave_height_polygon_with_flipped_coordinate_system():  # nav_mesh_test.rs:1083-1106
    NavigationMesh(TRIANGLE, [[2, 1, 0]], [0], height_mesh=create_height_mesh(TRIANGLE, [[[2, 1, 0]]]),
                   cs=FlippedXYZ).validate()


# --- ValidNavigationMesh::sample_edge (feeds animation-link creation) — nav_mesh_test.rs:1164-1345 ---

def sampled(mesh, a, b, max_vertical_distance):
    from landmass_b200.anim_links import sample_edge
    edge = (np.array(a, dtype=np.float32), np.array(b, dtype=np.float32))
    return sorted((node, (float(t0), float(t1))) for node, (t0, t1) in sample_edge(mesh, edge, max_vertical_distance))


F32 = lambda x: float(np.float32(x))


def test_samples_edge_on_square():  # nav_mesh_test.rs:1164-1214
    m = NavigationMesh(TRIANGLE + [(0.0, 1.0, 0.0)], [[0, 1, 2, 3]], [0]).validate()
    assert sampled(m, (0.5, 0.0, 0.0), (1.0, 0.5, 0.0), 1.0) == [(0, (0.0, 1.0))]
    assert sampled(m, (0.0, 0.5, 0.0), (2.0, 0.5, 0.0), 1.0) == [(0, (0.0, 0.5))]
    assert sampled(m, (-1.0, 0.5, 0.0), (3.0, 0.5, 0.0), 1.0) == [(0, (0.25, 0.5))]
    assert sampled(m, (1.0, 0.5, 0.0), (2.0, 0.5, 0.0), 1.0) == []


def test_samples_edge_for_multiple_nodes():  # nav_mesh_test.rs:1216-1256
    from landmass_b200 import XY
    m = NavigationMesh([(0.0, 0.0), (1.0, 0.0), (2.0, 0.0), (0.0, 1.0), (1.0, 1.0), (2.0, 1.0), (0.0, 2.0), (1.0, 2.0),
                        (2.0, 2.0)], [[0, 1, 4, 3], [1, 2, 5, 4], [3, 4, 7, 6], [4, 5, 8, 7]], [0] * 4, cs=XY).validate()
    assert sampled(m, (0.75, 0.5, 0.0), (1.5, 1.25, 0.0), 1.0) == [
        (0, (0.0, F32(0.33333334))), (1, (F32(0.33333334), F32(0.6666667))), (3, (F32(0.6666667), 1.0))]


def test_samples_edge_for_multiple_floors():  # nav_mesh_test.rs:1258-1297
    m = NavigationMesh(TRIANGLE + [(0.0, 1.0, 0.0), (0.0, 0.0, 2.0), (1.0, 0.0, 2.0), (1.0, 1.0, 2.0), (0.0, 1.0, 2.0)],
                       [[0, 1, 2, 3], [4, 5, 6, 7]], [0] * 2).validate()
    assert sampled(m, (0.0, 0.1, 0.5), (1.0, 0.1, 0.5), 1.0) == [(0, (0.0, 1.0))]
    assert sampled(m, (0.0, 0.1, 1.5), (1.0, 0.1, 1.5), 1.0) == [(1, (0.0, 1.0))]


def test_samples_edge_along_stairs():  # nav_mesh_test.rs:1299-1345
    m = NavigationMesh(TRIANGLE + [(0.0, 1.0, 0.0), (0.0, 2.0, 2.0), (1.0, 2.0, 2.0), (0.0, 3.0, 2.0), (1.0, 3.0, 2.0)],
                       [[0, 1, 2, 3], [3, 2, 5, 4], [4, 5, 7, 6]], [0] * 3).validate()
    assert sampled(m, (0.5, 0.5, 1.0), (0.5, 2.5, 1.0), 0.5) == [(1, (0.375, 0.625))]
    assert sampled(m, (0.5, 0.5, 1.0), (0.5, 2.5, 1.0), 1.1) == [(0, (0.0, 0.25)), (1, (0.25, 0.75)), (2, (0.75, 1.0))]
