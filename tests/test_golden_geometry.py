"""Host-side geometry helpers behind island linking and animation-link creation — known answers
transcribed from crates/landmass/src/geometry_test.rs:8-77 (`edge_intersection`, used by
landmass_b200/linking.py) and 79-545 (`clip_edge_to_triangle`, used by landmass_b200/anim_links.py).
All values are compared as f32, exactly, except where the reference itself uses approx_eq."""
import numpy as np
import pytest

from landmass_b200.anim_links import clip_edge_to_triangle
from landmass_b200.linking import edge_intersection

f32 = np.float32


def V(*p):
    return np.array(p, dtype=f32)


def intersect(e1, e2, d):
    ok, p0, p1 = edge_intersection(V(*e1[0])[None], V(*e1[1])[None], V(*e2[0])[None], V(*e2[1])[None],
                                   f32(f32(d) * f32(d)))
    return (p0[0].tolist(), p1[0].tolist()) if ok[0] else None


def L(*p):
    return V(*p).tolist()


def test_edge_intersects_when_on_same_line():  # geometry_test.rs:8-29
    third = f32(5.0) / f32(3.0)
    assert intersect(((1, 1, 1), (4, 2, 2.5)), ((7, 3, 4), (3, third, 2)), 0.1) == (L(3, third, 2), L(4, 2, 2.5))
    assert intersect(((1, 1, 1), (4, 2, 2.5)), ((7, 3, 4), (5, f32(7.0) / f32(3.0), 3)), 0.1) is None


def test_edge_intersects_only_when_nearby():  # geometry_test.rs:31-52
    assert intersect(((7, 0, 0), (3, 0, 0)), ((4, 0, 0.09), (6, 0, 0.09)), 0.1) == (L(6, 0, 0.045), L(4, 0, 0.045))
    assert intersect(((7, 0, 0), (3, 0, 0)), ((4, 0, 0.11), (6, 0, 0.11)), 0.1) is None


def test_crossed_edges_dont_intersect_unless_close():  # geometry_test.rs:54-77
    assert intersect(((5, 0, 0), (10, 0, 0)), ((14, 0, 4.5), (9, 0, -0.5)), 0.51) == (L(9, 0, -0.25), L(9.875, 0, 0.125))
    assert intersect(((5, 0, 0), (10, 0, 0)), ((14, 0, 4.5), (9, 0, -0.5)), 0.49) is None


def clip(tri, a, b, d):
    r = clip_edge_to_triangle(tuple(V(*p) for p in tri), (V(*a), V(*b)), d)
    return None if r is None else (float(r[0]), float(r[1]))


def F(*x):
    return tuple(float(f32(v)) for v in x)


def flat_tri(z):
    return ((1, 1, z), (11, 1, z), (6, 11, z))


def test_clips_edge_to_triangle_through():  # geometry_test.rs:79-121
    t = flat_tri(7)
    assert clip(t, (0, 6, 7), (12, 6, 7), 1.0) == F(0.29166666, 0.70833333)
    assert clip(t, (12, 6, 7), (0, 6, 7), 1.0) == F(0.29166666, 0.70833333)
    twelfth = F(f32(1.0) / f32(12.0), f32(11.0) / f32(12.0))
    assert clip(t, (6, 0, 7), (6, 12, 7), 1.0) == twelfth
    assert clip(t, (6, 12, 7), (6, 0, 7), 1.0) == twelfth


def test_clips_edge_to_triangle_one_sided():  # geometry_test.rs:123-200
    t = flat_tri(13)
    for outer in [(1, 6, 13), (11, 6, 13), (6, -4, 13), (6, 16, 13)]:
        assert clip(t, (6, 6, 13), outer, 1.0) == (0.0, 0.5)
        assert clip(t, outer, (6, 6, 13), 1.0) == (0.5, 1.0)


def test_clips_edge_entirely_outside():  # geometry_test.rs:202-277
    t = flat_tri(2)
    for a, b in [((11, 6, 2), (15, 6, 2)), ((0, 6, 2), (3, 6, 2)), ((6, 0, 2), (6, -5, 2)), ((6, 12, 2), (6, 15, 2))]:
        assert clip(t, a, b, 1.0) is None and clip(t, b, a, 1.0) is None


def test_clips_edge_to_triangle_against_all_sides():  # geometry_test.rs:279-367
    t = (V(1, 1, 10), V(11, 1, 13), V(6, 11, 20))
    half = f32(0.5)
    c = [(t[0] + t[1]) * half, (t[1] + t[2]) * half, (t[2] + t[0]) * half]
    d = [c[1] - c[0], c[2] - c[1], c[0] - c[2]]
    for k in range(3):
        a, b = c[k] - d[k] * half, c[(k + 1) % 3] + d[k] * half
        assert clip(t, a, b, 1.0) == (0.25, 0.75)
        assert clip(t, b, a, 1.0) == (0.25, 0.75)


def test_clips_whole_edge_if_too_far_flat():  # geometry_test.rs:369-412
    t = flat_tri(5)
    assert clip(t, (1, 6, 1), (11, 6, 1), 5.0) == (0.25, 0.75)
    assert clip(t, (1, 6, 1), (11, 6, 1), 3.0) is None
    assert clip(t, (1, 6, 9), (11, 6, 9), 5.0) == (0.25, 0.75)
    assert clip(t, (1, 6, 11), (11, 6, 11), 5.0) is None


def test_clips_whole_edge_if_too_far_sloped():  # geometry_test.rs:414-455
    t = flat_tri(5)
    assert clip(t, (1, 6, 5), (11, 6, 15), 2.0) is None
    assert clip(t, (1, 6, 15), (11, 6, 5), 2.0) is None
    assert clip(((1, 1, 5), (11, 1, 15), (6, 11, 10)), (1, 6, 5), (11, 6, 5), 2.0) is None


def test_clips_edge_part_way_when_vertical_distance_too_great():  # geometry_test.rs:457-501
    t = flat_tri(5)
    assert clip(t, (6, 0, 0), (6, 12, 6), 2.0) == pytest.approx((0.5, 11.0 / 12.0), rel=1e-5)
    assert clip(t, (6, 0, 6), (6, 12, 0), 2.0) == pytest.approx((1.0 / 12.0, 0.5), rel=1e-5)
    assert clip(t, (6, 0, 11), (6, 12, -1), 3.0) == (0.25, 0.75)
    assert clip(t, (6, 0, -1), (6, 12, 11), 3.0) == (0.25, 0.75)


def test_clips_edge_part_way_estimating_triangle_height():  # geometry_test.rs:503-526
    t = ((1, 1, 0), (11, 1, 10), (6, 11, 5))
    assert clip(t, (1, 6, 5), (11, 6, 5), 2.0) == F(0.3, 0.7)
    assert clip(t, (11, 6, 5), (1, 6, 5), 2.0) == F(0.3, 0.7)


def test_clips_edge_height_correctly_when_extending_line():  # geometry_test.rs:528-545
    assert clip(((0, 0, 2), (1, 0, 2), (1, 1, 2)), (0, 0.1, 1.5), (0.5, 0.1, 1.5), 1.0) == F(0.20000005, 1.0)


# --- util_test.rs: the two BoundingBox operations the host-side link sampling uses ---

def test_bounding_box_detects_ray_segment_intersection():  # util_test.rs:182-237
    from landmass_b200.anim_links import intersects_ray_segment
    bmin, bmax = V(-1, -2, -3), V(5, 4, 3)

    def check(v0, v1, expected):
        assert intersects_ray_segment(bmin, bmax, V(*v0), V(*v1)) == expected
        assert intersects_ray_segment(bmin, bmax, V(*v1), V(*v0)) == expected

    for axis in range(3):
        def on_axis(t):
            p = [2.0, 2.0, 2.0]
            p[axis] = t
            return tuple(p)
        # the reference's six cases per axis: inside, through, entering, leaving, before, after
        for a, b, expected in [(2, 4, True), (-3, 6, True), (-10, 2, True), (2, 10, True), (-10, -3, False),
                               (6, 10, False)]:
            check(on_axis(a), on_axis(b), expected)
    check((-2, -3, -4), (6, 5, 4), True)
    check((6, -3, -4), (-2, 5, 4), True)
    check((-2, 5, -4), (6, -3, 4), True)
    check((-2, -3, 4), (6, 5, -4), True)
    check((8, -13, -4), (-12, 7, -4), False)
    check((16, -5, 4), (-4, 15, 4), False)


def test_transforms_bounds():  # util_test.rs:239-270
    import math
    from types import SimpleNamespace
    from landmass_b200.anim_links import transformed_bounds
    t, r = V(1, 2, 3), f32(0.75)
    assert transformed_bounds(SimpleNamespace(mesh_bounds=None), t, r) is None  # Empty stays Empty
    mesh = SimpleNamespace(mesh_bounds=np.array([1, 3, 2, 6, 4, 5], dtype=f32))
    mn, mx = transformed_bounds(mesh, V(-4, 1, -3), f32(math.pi * -0.75))
    root_2 = math.sqrt(2.0)
    np.testing.assert_allclose(mn, (-3.0 / root_2 - 4.0, -10.0 / root_2 + 1.0, -1.0), atol=1e-6)
    np.testing.assert_allclose(mx, (3.0 / root_2 - 4.0, -4.0 / root_2 + 1.0, 2.0), atol=1e-6)
