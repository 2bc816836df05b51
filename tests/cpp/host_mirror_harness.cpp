// Test harness for include/landmass_b200.hpp (the C++ host mirror): runs a scripted game loop through
// `BasicArchipelago` on a backend bound at run time (function pointers handed in by the Python test: the
// CPU oracle here, the CUDA library on a GPU box) and records a per-frame trace that the test compares with
// the Python mirror running the same script.  TEST INFRASTRUCTURE.
#include <cstdint>
#include <cstring>

#include "landmass_b200.hpp"
#include "landmass_b200_anim_links.hpp"
#include "landmass_b200_linking.hpp"

namespace lb = landmass_b200;

struct FnBackend {
  using Ctx = void;
  static inline int32_t (*p_upload)(void*, const lmb_navdata_desc*) = nullptr;
  static inline int32_t (*p_step)(void*, const lmb_agents_in*, const lmb_characters_in*, const lmb_options*, float,
                                  const lmb_agents_out*) = nullptr;
  static inline int32_t (*p_results)(void*, lmb_pathing_result*, uint32_t, uint32_t*) = nullptr;
  static inline int32_t (*p_update)(void*, const lmb_navdata_desc*, const lmb_nav_remap*) = nullptr;
  static int32_t navdata_upload(Ctx* c, const lmb_navdata_desc* d) { return p_upload(c, d); }
  static int32_t navdata_update(Ctx* c, const lmb_navdata_desc* d, const lmb_nav_remap* r) { return p_update(c, d, r); }
  static int32_t step(Ctx* c, const lmb_agents_in* a, const lmb_characters_in* ch, const lmb_options* o, float dt,
                      const lmb_agents_out* out) {
    return p_step(c, a, ch, o, dt, out);
  }
  static int32_t pathing_results(Ctx* c, lmb_pathing_result* out, uint32_t cap, uint32_t* n) {
    return p_results(c, out, cap, n);
  }
  static inline int32_t (*p_path)(void*, uint32_t, uint32_t*, uint32_t*, uint32_t, uint32_t*) = nullptr;
  static inline int32_t (*p_constraints)(void*, uint32_t, uint32_t, float*, uint32_t*, uint32_t*, uint32_t*) = nullptr;
  static int32_t agent_path(Ctx* c, uint32_t slot, uint32_t* nodes, uint32_t* portals, uint32_t cap, uint32_t* len) {
    return p_path(c, slot, nodes, portals, cap, len);
  }
  static int32_t agent_avoidance_constraints(Ctx* c, uint32_t i, uint32_t cap, float* lines, uint32_t* no, uint32_t* nf,
                                             uint32_t* ran) {
    // bound only by the tests that exercise the debug export; everywhere else nobody asks
    if (!p_constraints) {
      *no = *nf = *ran = 0;
      return LMB_OK;
    }
    return p_constraints(c, i, cap, lines, no, nf, ran);
  }
  static inline int32_t (*p_sample)(void*, uint32_t, const float*, const lmb_options*, float*, uint32_t*) = nullptr;
  static inline int32_t (*p_straight)(void*, uint32_t, const uint32_t*, const float*, const uint32_t*, const float*,
                                      const uint32_t*, const uint32_t*, const float*, uint32_t*, uint32_t*, uint32_t*,
                                      float*, uint32_t*, uint32_t, const uint32_t*, const uint32_t*,
                                      const uint32_t*) = nullptr;
  static int32_t sample_points(Ctx* c, uint32_t n, const float* points, const lmb_options* o, float* out_points,
                               uint32_t* out_nodes) {
    return p_sample(c, n, points, o, out_points, out_nodes);
  }
  static int32_t find_straight_paths(Ctx* c, uint32_t n, const uint32_t* sn, const float* sp, const uint32_t* en,
                                     const float* ep, const uint32_t* ob, const uint32_t* ot, const float* oc,
                                     uint32_t* succ, uint32_t* begin, uint32_t* kind, float* pts, uint32_t* link,
                                     uint32_t cap, const uint32_t* pa, const uint32_t* pb, const uint32_t* pk) {
    return p_straight(c, n, sn, sp, en, ep, ob, ot, oc, succ, begin, kind, pts, link, cap, pa, pb, pk);
  }
  static std::string last_error(Ctx*) { return ""; }
};

extern "C" {

// The script (mirrored line by line in tests/test_host_mirror_cpp.py):
//   agents A (1,1)->(11,1.1) r=1, B (11,1.1)->(1,1) r=1, C (6,8)->(6,2) r=0.5 with a cost override and
//   StraightPathDistance(0.25); one character at (6,4) drifting in +x.  Frame 10: A is removed; frame 12: D is
//   added (re-uses A's slot index with a new version); frames 20-24: B is paused.
// Trace per frame: [4 agents][position xyz, desired velocity xyz, state] (absent: -1000 / 99), then the number
// of pathing results and the sum of their explored nodes, then for each result its agent key idx + version.
int32_t hm_run_script(void* ctx, void* fn_upload, void* fn_step, void* fn_results, const lmb_navdata_desc* desc,
                      uint32_t max_agents, uint32_t frames, float dt, float* trace_f, uint32_t* trace_u,
                      char* error, uint32_t error_cap, void* fn_path, void* fn_constraints, uint32_t* extras_u,
                      float* extras_f) {
  // optional debug read-backs (fn_path / fn_constraints may be null): agent C keeps its avoidance data from the
  // start; after the last frame extras_u = [corridor length of B, its first node, C's original / fallback constraint
  // counts, C fell back] and extras_f = C's constraint lines (4 floats each)
  FnBackend::p_path = reinterpret_cast<decltype(FnBackend::p_path)>(fn_path);
  FnBackend::p_constraints = reinterpret_cast<decltype(FnBackend::p_constraints)>(fn_constraints);
  FnBackend::p_upload = reinterpret_cast<decltype(FnBackend::p_upload)>(fn_upload);
  FnBackend::p_step = reinterpret_cast<decltype(FnBackend::p_step)>(fn_step);
  FnBackend::p_results = reinterpret_cast<decltype(FnBackend::p_results)>(fn_results);
  try {
    lb::BasicArchipelago<FnBackend> arch(ctx, lb::ArchipelagoOptions::from_agent_radius(0.5f), max_agents);
    // desc == nullptr: the 15 x 15 room is validated and flattened by the C++ mirror itself
    lb::NavigationMesh room;
    room.vertices = {{0.0f, 0.0f, 0.0f}, {15.0f, 0.0f, 0.0f}, {15.0f, 15.0f, 0.0f}, {0.0f, 15.0f, 0.0f}};
    room.polygons = {{0, 1, 2, 3}};
    room.polygon_type_indices = {0};
    lb::SingleIslandNavigationData own(lb::Transform{}, room.validate());
    arch.set_navigation_data(desc ? *desc : own.desc());
    auto make = [](lb::Vec3 pos, lb::Vec3 target, float radius) {
      lb::Agent a = lb::Agent::create(pos, {0.0f, 0.0f, 0.0f}, radius, 1.0f, 2.0f);
      a.current_target = target;
      a.target_reached_condition = lb::TargetReachedCondition::Distance(0.01f);
      return a;
    };
    lb::AgentId ids[4];
    bool alive[4] = {true, true, true, false};
    ids[0] = arch.add_agent(make({1.0f, 1.0f, 0.0f}, {11.0f, 1.1f, 0.0f}, 1.0f));
    ids[1] = arch.add_agent(make({11.0f, 1.1f, 0.0f}, {1.0f, 1.0f, 0.0f}, 1.0f));
    {
      lb::Agent c = make({6.0f, 8.0f, 0.0f}, {6.0f, 2.0f, 0.0f}, 0.5f);
      c.target_reached_condition = lb::TargetReachedCondition::StraightPathDistance(0.25f);
      c.override_type_index_cost(0, 3.0f);
      c.keep_avoidance_data = fn_constraints != nullptr;
      ids[2] = arch.add_agent(c);
    }
    lb::CharacterId ch = arch.add_character({{6.0f, 4.0f, 0.0f}, {0.2f, 0.0f, 0.0f}, 0.5f});
    for (uint32_t f = 0; f < frames; f++) {
      if (f == 10) {
        arch.remove_agent(ids[0]);
        alive[0] = false;
      }
      if (f == 12) {
        ids[3] = arch.add_agent(make({2.0f, 12.0f, 0.0f}, {12.0f, 12.0f, 0.0f}, 0.5f));
        alive[3] = true;
      }
      if (f == 20) arch.get_agent_mut(ids[1])->paused = true;
      if (f == 25) arch.get_agent_mut(ids[1])->paused = false;
      arch.update(dt);
      float* tf = trace_f + (size_t)f * 4 * 6;
      uint32_t* tu = trace_u + (size_t)f * (4 + 2 + 2 * 4);
      for (int k = 0; k < 4; k++) {
        if (!alive[k]) {
          for (int c = 0; c < 6; c++) tf[6 * k + c] = -1000.0f;
          tu[k] = 99u;
          continue;
        }
        lb::Agent* a = arch.get_agent_mut(ids[k]);
        a->velocity = a->get_desired_velocity();
        for (int c = 0; c < 3; c++) a->position[c] = a->position[c] + a->velocity[c] * dt;
        for (int c = 0; c < 3; c++) tf[6 * k + c] = a->position[c], tf[6 * k + 3 + c] = a->velocity[c];
        tu[k] = (uint32_t)a->state();
      }
      lb::Character* c = arch.get_character_mut(ch);
      for (int k = 0; k < 3; k++) c->position[k] = c->position[k] + c->velocity[k] * dt;
      const auto& pr = arch.get_pathing_results();
      uint32_t explored = 0;
      for (const auto& r : pr) explored += r.explored_nodes;
      tu[4] = (uint32_t)pr.size();
      tu[5] = explored;
      for (int k = 0; k < 4; k++) {
        tu[6 + 2 * k] = k < (int)pr.size() ? pr[(size_t)k].agent.idx : 0u;
        tu[7 + 2 * k] = k < (int)pr.size() ? pr[(size_t)k].agent.version : 0u;
      }
    }
    if (fn_path && extras_u) {
      const auto path = arch.agent_path(ids[1]);
      extras_u[0] = (uint32_t)path.first.size();
      extras_u[1] = path.first.empty() ? LMB_NONE : path.first[0];
    }
    if (fn_constraints && extras_u) {
      const lb::Agent::AvoidanceData* data = arch.get_agent(ids[2])->avoidance_data();
      extras_u[2] = data ? (uint32_t)data->original.size() : LMB_NONE;
      extras_u[3] = data ? (uint32_t)data->fallback.size() : LMB_NONE;
      extras_u[4] = data && data->fell_back ? 1u : 0u;
      if (data) {
        size_t w = 0;
        for (const auto& l : data->original)
          for (float x : l) extras_f[w++] = x;
        for (const auto& l : data->fallback)
          for (float x : l) extras_f[w++] = x;
      }
    }
  } catch (const std::exception& e) {
    std::strncpy(error, e.what(), error_cap - 1);
    error[error_cap - 1] = 0;
    return -1;
  }
  return 0;
}

// `sample_point` + `find_path` for n (start, end) pairs; query q with override cost `override_cost[q]` on type 0
// when that is > 0.  out_n_steps[q] = number of steps, or -1 / -2 (start / end OutOfRange), -3 (NoPathFound);
// steps are appended to out_steps as 7 floats (is_link, point xyz, end point xyz), at most step_cap of them.
int32_t hm_queries(void* ctx, void* fn_upload, void* fn_sample, void* fn_straight, const lmb_navdata_desc* desc,
                   uint32_t n, const float* starts, const float* ends, const float* override_cost, int32_t* out_n_steps,
                   float* out_steps, uint32_t step_cap, char* error, uint32_t error_cap) {
  FnBackend::p_upload = reinterpret_cast<decltype(FnBackend::p_upload)>(fn_upload);
  FnBackend::p_sample = reinterpret_cast<decltype(FnBackend::p_sample)>(fn_sample);
  FnBackend::p_straight = reinterpret_cast<decltype(FnBackend::p_straight)>(fn_straight);
  try {
    lb::BasicArchipelago<FnBackend> arch(ctx, lb::ArchipelagoOptions::from_agent_radius(0.5f), 4);
    arch.set_navigation_data(*desc);
    uint32_t used = 0;
    for (uint32_t q = 0; q < n; q++) {
      auto a = arch.sample_point({starts[3 * q], starts[3 * q + 1], starts[3 * q + 2]});
      auto b = arch.sample_point({ends[3 * q], ends[3 * q + 1], ends[3 * q + 2]});
      if (!a || !b) {
        out_n_steps[q] = !a ? -1 : -2;
        continue;
      }
      std::vector<std::pair<uint32_t, float>> ov;
      if (override_cost[q] > 0.0f) ov.emplace_back(0u, override_cost[q]);
      auto steps = arch.find_path(*a, *b, ov);
      if (!steps) {
        out_n_steps[q] = -3;
        continue;
      }
      out_n_steps[q] = (int32_t)steps->size();
      for (const lb::PathStep& st : *steps) {
        if (used >= step_cap) throw std::runtime_error("step capacity");
        float* o = out_steps + 7 * (size_t)used++;
        o[0] = st.is_animation_link ? 1.0f : 0.0f;
        for (int c = 0; c < 3; c++) o[1 + c] = st.point[c], o[4 + c] = st.is_animation_link ? st.end_point[c] : 0.0f;
      }
    }
  } catch (const std::exception& e) {
    std::strncpy(error, e.what(), error_cap - 1);
    error[error_cap - 1] = 0;
    return -1;
  }
  return 0;
}

// NavigationMesh::validate + SingleIslandNavigationData on caller-provided input; the flattened arrays are
// copied out for comparison with the Python mirror.  Returns 0, or 1 + ValidationError::Kind with the
// payload in out_err[0..1].
static int32_t validate_flatten(uint32_t n_vertices, const float* vertices, uint32_t n_polys, const uint32_t* poly_begin,
                                const uint32_t* poly_verts, uint32_t n_types, const uint32_t* types,
                                const float* translation, float rotation, uint64_t* out_err, uint32_t* out_u32,
                                float* out_f32, float* out_island, const lb::HeightNavigationMesh* height,
                                uint32_t* out_hu, float* out_hf);

int32_t hm_validate_flatten(uint32_t n_vertices, const float* vertices, uint32_t n_polys, const uint32_t* poly_begin,
                            const uint32_t* poly_verts, uint32_t n_types, const uint32_t* types,
                            const float* translation, float rotation, uint64_t* out_err, uint32_t* out_u32,
                            float* out_f32, float* out_island) {
  return validate_flatten(n_vertices, vertices, n_polys, poly_begin, poly_verts, n_types, types, translation, rotation,
                          out_err, out_u32, out_f32, out_island, nullptr, nullptr, nullptr);
}

// The same with a height mesh: hpolys [n_hpolys*4], hverts [n_hverts*3], htris [n_htris*3].  On success out_hu
// receives n_hverts, n_htris, hpoly [P*4], htris [HT*3] (one word each) and out_hf the height vertices.
int32_t hm_validate_flatten_height(uint32_t n_vertices, const float* vertices, uint32_t n_polys, const uint32_t* poly_begin,
                                   const uint32_t* poly_verts, uint32_t n_types, const uint32_t* types,
                                   const float* translation, float rotation, uint32_t n_hpolys, const uint32_t* hpolys,
                                   uint32_t n_hverts, const float* hverts, uint32_t n_htris, const uint32_t* htris,
                                   uint64_t* out_err, uint32_t* out_u32, float* out_f32, float* out_island,
                                   uint32_t* out_hu, float* out_hf) {
  lb::HeightNavigationMesh h;
  for (uint32_t p = 0; p < n_hpolys; p++) h.polygons.push_back({hpolys[4 * p], hpolys[4 * p + 1], hpolys[4 * p + 2], hpolys[4 * p + 3]});
  for (uint32_t v = 0; v < n_hverts; v++) h.vertices.push_back({hverts[3 * v], hverts[3 * v + 1], hverts[3 * v + 2]});
  for (uint32_t t = 0; t < n_htris; t++) h.triangles.push_back({htris[3 * t], htris[3 * t + 1], htris[3 * t + 2]});
  return validate_flatten(n_vertices, vertices, n_polys, poly_begin, poly_verts, n_types, types, translation, rotation,
                          out_err, out_u32, out_f32, out_island, &h, out_hu, out_hf);
}

static int32_t validate_flatten(uint32_t n_vertices, const float* vertices, uint32_t n_polys, const uint32_t* poly_begin,
                                const uint32_t* poly_verts, uint32_t n_types, const uint32_t* types,
                                const float* translation, float rotation, uint64_t* out_err, uint32_t* out_u32,
                                float* out_f32, float* out_island, const lb::HeightNavigationMesh* height,
                                uint32_t* out_hu, float* out_hf) {
  lb::NavigationMesh mesh;
  if (height) mesh.height_mesh = *height;
  for (uint32_t v = 0; v < n_vertices; v++) mesh.vertices.push_back({vertices[3 * v], vertices[3 * v + 1], vertices[3 * v + 2]});
  for (uint32_t p = 0; p < n_polys; p++)
    mesh.polygons.emplace_back(poly_verts + poly_begin[p], poly_verts + poly_begin[p + 1]);
  mesh.polygon_type_indices.assign(types, types + n_types);
  try {
    lb::SingleIslandNavigationData nd(lb::Transform{{translation[0], translation[1], translation[2]}, rotation},
                                      mesh.validate());
    const lmb_navdata_desc& d = nd.desc();
    // u32: poly_edge_begin [P+1] | edge_vertex [E] | edge_neighbor [E] | edge_reverse [E] | poly_type [P] |
    //      poly_region [P] | node_link_begin [P+1] | node_region_number [P] | island_poly_begin[2] | island_vert_begin[2]
    uint32_t* u = out_u32;
    auto put = [&](const uint32_t* src, size_t n) {
      std::memcpy(u, src, n * sizeof(uint32_t));
      u += n;
    };
    put(d.poly_edge_begin, d.n_polys + 1), put(d.edge_vertex, d.n_edges), put(d.edge_neighbor, d.n_edges);
    put(d.edge_reverse, d.n_edges), put(d.poly_type, d.n_polys), put(d.poly_region, d.n_polys);
    put(d.node_link_begin, d.n_polys + 1), put(d.node_region_number, d.n_polys);
    put(d.island_poly_begin, 2), put(d.island_vert_begin, 2);
    // f32: vertices [V*3] | poly_bounds [P*6] | poly_center [P*3]
    float* f = out_f32;
    std::memcpy(f, d.vertices, (size_t)d.n_vertices * 3 * sizeof(float)), f += (size_t)d.n_vertices * 3;
    std::memcpy(f, d.poly_bounds, (size_t)d.n_polys * 6 * sizeof(float)), f += (size_t)d.n_polys * 6;
    std::memcpy(f, d.poly_center, (size_t)d.n_polys * 3 * sizeof(float));
    // island: translation xyz, rotation, bounds[6]
    std::memcpy(out_island, d.island_translation, 3 * sizeof(float));
    out_island[3] = d.island_rotation[0];
    std::memcpy(out_island + 4, d.island_bounds, 6 * sizeof(float));
    out_err[0] = d.n_edges, out_err[1] = d.n_links + d.n_modified + d.n_region_numbers + d.n_possible_links;
    if (height) {
      if (!d.island_has_height || !d.island_has_height[0]) return -2;
      uint32_t* hu = out_hu;
      *hu++ = d.n_hverts, *hu++ = d.n_htris;
      std::memcpy(hu, d.hpoly, (size_t)d.n_polys * 4 * sizeof(uint32_t)), hu += (size_t)d.n_polys * 4;
      for (size_t k = 0; k < 3 * (size_t)d.n_htris; k++) *hu++ = d.htris[k];
      std::memcpy(out_hf, d.hverts, 3 * (size_t)d.n_hverts * sizeof(float));
    }
    return 0;
  } catch (const lb::ValidationError& e) {
    out_err[0] = e.a, out_err[1] = e.b;
    return 1 + (int32_t)e.kind;
  }
}

// A changing world through BasicIslandArchipelago (mirrored line by line in tests/test_host_mirror_cpp.py): three 6 x 6
// grid islands in a row, A B C; agents 0-2 walk from A to C, agent 3 from C to A.  Frame 8: island D is added above
// B; frame 14: C is removed; frame 18: B is touched (same mesh set again: its paths are invalidated); frame 22: a
// new island takes C's place; frame 25: type 0 gets cost 2.  Trace per frame: [4 agents][position xyz, velocity xyz]
// and [4 states, number of pathing results, sum of explored nodes, number of links].
int32_t hm_run_world_script(void* ctx, void* fn_upload, void* fn_update, void* fn_step, void* fn_results,
                            uint32_t frames, float dt, float* trace_f, uint32_t* trace_u, char* error,
                            uint32_t error_cap) {
  FnBackend::p_upload = reinterpret_cast<decltype(FnBackend::p_upload)>(fn_upload);
  FnBackend::p_update = reinterpret_cast<decltype(FnBackend::p_update)>(fn_update);
  FnBackend::p_step = reinterpret_cast<decltype(FnBackend::p_step)>(fn_step);
  FnBackend::p_results = reinterpret_cast<decltype(FnBackend::p_results)>(fn_results);
  try {
    lb::NavigationMesh grid;
    for (int y = 0; y <= 6; y++)
      for (int x = 0; x <= 6; x++) grid.vertices.push_back({(float)x, (float)y, 0.0f});
    for (uint32_t y = 0; y < 6; y++)
      for (uint32_t x = 0; x < 6; x++) {
        const uint32_t v = y * 7 + x;
        grid.polygons.push_back({v, v + 1, v + 8, v + 7});
        grid.polygon_type_indices.push_back(0);
      }
    const lb::ValidNavigationMesh mesh = grid.validate();
    auto island_at = [&](float x, float y) { return lb::Island{lb::Transform{{x, y, 0.0f}, 0.0f}, mesh}; };
    lb::BasicIslandArchipelago<FnBackend> arch(ctx, lb::ArchipelagoOptions::from_agent_radius(0.5f), 16);
    arch.add_island(island_at(0.0f, 0.0f));
    const lb::IslandId b = arch.add_island(island_at(6.0f, 0.0f));
    lb::IslandId c = arch.add_island(island_at(12.0f, 0.0f));
    auto make = [](lb::Vec3 pos, lb::Vec3 target) {
      lb::Agent a = lb::Agent::create(pos, {0.0f, 0.0f, 0.0f}, 0.4f, 2.0f, 3.0f);
      a.current_target = target;
      return a;
    };
    lb::AgentId ids[4] = {arch.add_agent(make({0.7f, 0.8f, 0.0f}, {17.2f, 5.1f, 0.0f})),
                          arch.add_agent(make({1.5f, 3.1f, 0.0f}, {16.5f, 2.2f, 0.0f})),
                          arch.add_agent(make({0.9f, 5.2f, 0.0f}, {13.4f, 0.8f, 0.0f})),
                          arch.add_agent(make({16.8f, 3.3f, 0.0f}, {0.6f, 2.9f, 0.0f}))};
    for (uint32_t f = 0; f < frames; f++) {
      if (f == 8) arch.add_island(island_at(6.0f, 6.0f));
      if (f == 14) arch.remove_island(c);
      if (f == 18) arch.get_island_mut(b)->mesh = mesh;
      if (f == 22) c = arch.add_island(island_at(12.0f, 0.0f));
      if (f == 25) arch.set_type_index_cost(0, 2.0f);
      arch.update(dt);
      float* tf = trace_f + (size_t)f * 4 * 6;
      uint32_t* tu = trace_u + (size_t)f * 7;
      for (int k = 0; k < 4; k++) {
        lb::Agent* a = arch.get_agent_mut(ids[k]);
        a->velocity = a->get_desired_velocity();
        for (int x = 0; x < 3; x++) a->position[x] = a->position[x] + a->velocity[x] * dt;
        for (int x = 0; x < 3; x++) tf[6 * k + x] = a->position[x], tf[6 * k + 3 + x] = a->velocity[x];
        tu[k] = (uint32_t)a->state();
      }
      uint32_t explored = 0;
      for (const auto& r : arch.get_pathing_results()) explored += r.explored_nodes;
      tu[4] = (uint32_t)arch.get_pathing_results().size();
      tu[5] = explored;
      tu[6] = (uint32_t)arch.navigation_data()->n_links();
    }
  } catch (const std::exception& e) {
    std::strncpy(error, e.what(), error_cap - 1);
    error[error_cap - 1] = 0;
    return -1;
  }
  return 0;
}

// clip_edge_to_triangle on n (triangle, edge) pairs: tri [n*9], edge [n*6]; out [n*2], ok [n]
void hm_clip_edges(uint32_t n, const float* tri, const float* edge, float max_vertical_distance, float* out,
                   uint32_t* ok) {
  for (uint32_t i = 0; i < n; i++) {
    const float* t = tri + 9 * (size_t)i;
    const float* e = edge + 6 * (size_t)i;
    const lb::Vec3 tv[3] = {{t[0], t[1], t[2]}, {t[3], t[4], t[5]}, {t[6], t[7], t[8]}};
    const lb::Vec3 ev[2] = {{e[0], e[1], e[2]}, {e[3], e[4], e[5]}};
    const auto r = lb::clip_edge_to_triangle(tv, ev, max_vertical_distance);
    ok[i] = r ? 1u : 0u;
    out[2 * (size_t)i] = r ? r->first : 0.0f;
    out[2 * (size_t)i + 1] = r ? r->second : 0.0f;
  }
}

// validate (optionally with a height mesh) + sample_edge; returns the number of portions (<= cap) or -1
int32_t hm_sample_edge(uint32_t n_vertices, const float* vertices, uint32_t n_polys, const uint32_t* poly_begin,
                       const uint32_t* poly_verts, uint32_t n_hpolys, const uint32_t* hpolys, uint32_t n_hverts,
                       const float* hverts, uint32_t n_htris, const uint32_t* htris, const float* edge,
                       float max_vertical_distance, uint32_t cap, uint32_t* out_nodes, float* out_t) {
  try {
    lb::NavigationMesh mesh;
    for (uint32_t v = 0; v < n_vertices; v++) mesh.vertices.push_back({vertices[3 * v], vertices[3 * v + 1], vertices[3 * v + 2]});
    for (uint32_t p = 0; p < n_polys; p++) {
      mesh.polygons.emplace_back(poly_verts + poly_begin[p], poly_verts + poly_begin[p + 1]);
      mesh.polygon_type_indices.push_back(0);
    }
    if (n_hpolys) {
      lb::HeightNavigationMesh h;
      for (uint32_t p = 0; p < n_hpolys; p++) h.polygons.push_back({hpolys[4 * p], hpolys[4 * p + 1], hpolys[4 * p + 2], hpolys[4 * p + 3]});
      for (uint32_t v = 0; v < n_hverts; v++) h.vertices.push_back({hverts[3 * v], hverts[3 * v + 1], hverts[3 * v + 2]});
      for (uint32_t t = 0; t < n_htris; t++) h.triangles.push_back({htris[3 * t], htris[3 * t + 1], htris[3 * t + 2]});
      mesh.height_mesh = h;
    }
    const lb::ValidNavigationMesh valid = mesh.validate();
    const lb::Vec3 ev[2] = {{edge[0], edge[1], edge[2]}, {edge[3], edge[4], edge[5]}};
    const auto portions = lb::sample_edge(valid, ev, max_vertical_distance);
    if (portions.size() > cap) return -1;
    for (size_t k = 0; k < portions.size(); k++) {
      out_nodes[k] = portions[k].node;
      out_t[2 * k] = portions[k].t0, out_t[2 * k + 1] = portions[k].t1;
    }
    return (int32_t)portions.size();
  } catch (const std::exception&) {
    return -1;
  }
}

// Several islands -> validate each -> link (boundary links, modified nodes, regions) -> flatten, all in the C++
// mirror; the descriptor's arrays are streamed out in a fixed order for comparison with nav_data.flatten.
// Island i owns vertices [vert_begin[i], vert_begin[i+1]) and polygons [poly_begin[i], poly_begin[i+1]) whose vertex
// indices are island-local.  Returns 0 or -1 (error text in `error`); *n_u / *n_f = words written.
int32_t hm_link_flatten(uint32_t n_islands, const uint32_t* vert_begin, const float* vertices, const uint32_t* poly_begin,
                        const uint32_t* poly_edge_begin, const uint32_t* poly_verts, const uint32_t* types,
                        const float* translations, const float* rotations, uint32_t* out_u, uint32_t cap_u,
                        float* out_f, uint32_t cap_f, uint32_t* n_u, uint32_t* n_f, char* error, uint32_t error_cap) {
  try {
    std::vector<lb::Island> islands;
    for (uint32_t i = 0; i < n_islands; i++) {
      lb::NavigationMesh mesh;
      for (uint32_t v = vert_begin[i]; v < vert_begin[i + 1]; v++)
        mesh.vertices.push_back({vertices[3 * v], vertices[3 * v + 1], vertices[3 * v + 2]});
      for (uint32_t p = poly_begin[i]; p < poly_begin[i + 1]; p++) {
        mesh.polygons.emplace_back(poly_verts + poly_edge_begin[p], poly_verts + poly_edge_begin[p + 1]);
        mesh.polygon_type_indices.push_back(types[p]);
      }
      islands.push_back({lb::Transform{{translations[3 * i], translations[3 * i + 1], translations[3 * i + 2]}, rotations[i]},
                         mesh.validate()});
    }
    lb::NavigationData nd(std::move(islands));
    const lmb_navdata_desc& d = nd.desc();
    uint32_t wu = 0, wf = 0;
    auto pu = [&](const uint32_t* src, size_t n) {
      if (wu + n > cap_u) throw std::runtime_error("u32 capacity");
      if (n) std::memcpy(out_u + wu, src, n * sizeof(uint32_t));
      wu += (uint32_t)n;
    };
    auto pf = [&](const float* src, size_t n) {
      if (wf + n > cap_f) throw std::runtime_error("f32 capacity");
      if (n) std::memcpy(out_f + wf, src, n * sizeof(float));
      wf += (uint32_t)n;
    };
    const uint32_t head[6] = {d.n_polys, d.n_edges, d.n_links, d.n_modified, d.n_region_numbers, d.n_vertices};
    pu(head, 6);
    pu(d.island_poly_begin, d.n_islands + 1), pu(d.island_vert_begin, d.n_islands + 1);
    pu(d.poly_edge_begin, d.n_polys + 1), pu(d.edge_vertex, d.n_edges), pu(d.edge_neighbor, d.n_edges);
    pu(d.edge_reverse, d.n_edges), pu(d.poly_type, d.n_polys), pu(d.poly_region, d.n_polys);
    pu(d.link_dst_node, d.n_links), pu(d.link_dst_type, d.n_links), pu(d.link_kind, d.n_links);
    pu(d.link_reverse, d.n_links), pu(d.node_link_begin, d.n_polys + 1), pu(d.node_links, d.n_links);
    pu(d.modified_node, d.n_modified), pu(d.modified_edge_begin, d.n_modified + 1);
    pu(d.modified_edges, 2 * (size_t)d.modified_edge_begin[d.n_modified]);
    pu(d.modified_vert_begin, d.n_modified + 1);
    pu(d.node_region_number, d.n_polys), pu(d.region_root, d.n_region_numbers);
    pf(d.vertices, 3 * (size_t)d.n_vertices), pf(d.poly_bounds, 6 * (size_t)d.n_polys);
    pf(d.poly_center, 3 * (size_t)d.n_polys), pf(d.island_bounds, 6 * (size_t)d.n_islands);
    pf(d.link_portal, 6 * (size_t)d.n_links), pf(d.modified_verts, 2 * (size_t)d.modified_vert_begin[d.n_modified]);
    *n_u = wu, *n_f = wf;
  } catch (const std::exception& e) {
    std::strncpy(error, e.what(), error_cap - 1);
    error[error_cap - 1] = 0;
    return -1;
  }
  return 0;
}

}  // extern "C"
