// Kernel 4 — local avoidance (phase D of `Archipelago::update`).
//
// Replaces `apply_avoidance_to_agents` and `nav_mesh_borders_to_dodgy_obstacles`
// (crates/landmass/src/avoidance.rs:16-471) and the `dodgy_2d` calls they make
// (`VisibilitySet`, `Agent::compute_avoiding_velocity`).
//
//   * neighbour search: the reference rebuilds a 3-D kd-tree of all sampled agents every
//     frame; here agents are counting-sorted into a uniform xy grid (count -> scan ->
//     scatter) and each agent gathers candidates from the cells its query disc touches,
//     keeps those passing the reference's two distance tests and orders them by
//     (squared 3-D distance, agent index) — kdtree::within's ascending order.
//   * obstacles: best-first flood over polygons from the agent's node with an exact
//     angular visibility set (disjoint cones in a division-only pseudo-angle, each owned
//     by the nearest line), then the reference's loop stitching.
//   * ORCA: obstacle half-planes, one half-plane per neighbour (with dodgy_2d's
//     avoidance_responsibility), incremental 2-D LP + 3-D fallback.
//
// PARITY NOTE: dodgy_2d's source is not in /root/reference.  The arithmetic here mirrors
// oracle/avoidance_oracle.cpp op for op (RVO2 restatement); see that file's header.
//
// One thread per agent; per-agent variable-length state lives in a global scratch block.
#include <type_traits>

#include "lmb_device.cuh"
#include "lmb_kernels.h"

namespace lmb {

namespace {

constexpr float RVO_EPSILON = 0.00001f;
constexpr float F_INF = __builtin_huge_valf();

struct Cone {
  float lo, hi;
  uint32_t line;
};
struct LoopNode {
  uint32_t vid, next, prev;
};
struct Line {
  V2 point, direction;
};

// Per-agent variable-length arrays, INTERLEAVED by agent: element i of the agent with chunk-local index t lives
// at base[i * stride + t] (stride = agents per chunk, a multiple of 32).  The threads of a warp walk their
// arrays in near lock-step, so element i of 32 neighbouring agents is one contiguous run instead of 32 sectors
// several KB apart (one private block per agent, as before, made every access its own 32-byte sector).
template <class T>
struct Strided {
  T* p;
  uint32_t stride;
  __device__ __forceinline__ T& operator[](uint32_t i) const { return p[(size_t)i * stride]; }
  __device__ __forceinline__ explicit operator bool() const { return p != nullptr; }
};

struct Scratch {
  Strided<float2> nb;         // (d2, index bits)
  Strided<uint2> heap;        // (node, score bits)
  Strided<uint32_t> explored;      // explored set of the flood: open-addressing table of node ids, 4 x max_explore
                                   // slots, EMPTY (LMB_NONE) between agents (the host fills the scratch with 0xFF
                                   // whenever its layout changes; every flood un-inserts what it inserted)
  Strided<uint32_t> explored_slot; // the slots this flood filled, in insertion order (2 x max_explore)
  Strided<uint32_t> explored_list; // the explored node ids in insertion order (2 x max_explore): the whole set while
                                   // it is small (coalesced across the agents of a warp, unlike hashed slots)
  Strided<float4> vis;        // line a.xy, b.xy (relative to the agent, oriented CCW)
  Strided<uint2> border;      // (left vertex id, right vertex id) per line
  Strided<Cone> cones[2];
  Strided<float2> new_verts;
  Strided<LoopNode> lnodes;
  Strided<uint2> loops;       // (head, tail); LMB_NONE head = removed
  Strided<uint2> finished;
  Strided<Line> lines;
  Strided<Line> proj;
  uint32_t max_nb, max_explore, max_border, max_cones, max_lines, max_loops;
};

__device__ __forceinline__ Scratch carve(const AvoidScratch& s, uint32_t t) {
  Scratch r;
  uint8_t* p = s.base;
  const uint32_t stride = s.agent_stride;
  r.max_nb = s.max_neighbours;
  r.max_explore = s.max_explore;
  r.max_border = s.max_border_edges;
  r.max_cones = s.max_cones;
  r.max_lines = s.max_lines;
  r.max_loops = s.max_border_edges;
  auto take = [&](auto& arr, size_t count) {
    typedef typename std::remove_reference<decltype(arr[0])>::type T;
    arr.p = reinterpret_cast<T*>(p) + t;
    arr.stride = stride;
    p += count * stride * sizeof(T);  // every element size divides 16 or is 12: regions stay 4-byte aligned,
                                      // and 16-byte ones 16-byte aligned because stride is a multiple of 32
  };
  take(r.vis, r.max_border);          // 16-byte element types first (alignment)
  take(r.lines, r.max_lines);
  take(r.proj, r.max_lines);
  take(r.nb, r.max_nb);
  take(r.heap, r.max_explore);
  take(r.border, r.max_border);
  take(r.new_verts, r.max_border);
  take(r.loops, r.max_loops);
  take(r.finished, r.max_loops);
  take(r.explored, (size_t)r.max_explore * 4);
  take(r.explored_slot, (size_t)r.max_explore * 2);
  take(r.explored_list, (size_t)r.max_explore * 2);
  take(r.cones[0], r.max_cones);
  take(r.cones[1], r.max_cones);
  take(r.lnodes, (size_t)r.max_border * 2);
  return r;
}

// ---------------------------------------------------------------------------
// visibility set (mirrors oracle/avoidance_oracle.cpp VisibilitySet)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pangle(V2 v) {
  float d = fadd(fabsf(v.x), fabsf(v.y));
  float t = fdiv(v.x, d);
  return v.y >= 0.0f ? fsub(1.0f, t) : fadd(3.0f, t);
}
__device__ __forceinline__ V2 pdir(float g) {
  if (g < 2.0f) {
    float t = fsub(1.0f, g);
    return {t, fsub(1.0f, fabsf(t))};
  }
  float t = fsub(g, 3.0f);
  return {t, fsub(fabsf(t), 1.0f)};
}
__device__ __forceinline__ float ray_depth(V2 a, V2 b, V2 d) {
  return fdiv(perp_dot(a, b), perp_dot(d, b - a));
}

struct VisSet {
  Scratch* s;
  uint32_t n_lines;
  uint32_t n_cones;
  int cur;  // active cone buffer
  bool overflow;
};

// Tests (add_id < 0) or inserts the span [lo,hi] of line (a,b).  When inserting, the new
// cone list is streamed into the other buffer and *n_out is its length.
__device__ bool vis_process_piece(VisSet& vs, V2 a, V2 b, float lo, float hi, int add_id, uint32_t* n_out) {
  const Strided<Cone> cones = vs.s->cones[vs.cur];
  const Strided<Cone> out = add_id >= 0 ? vs.s->cones[1 - vs.cur] : Strided<Cone>{nullptr, 0u};
  uint32_t no = 0;
  bool visible = false;
  float cur = lo;
  bool tail_done = false;
  auto emit = [&](float l, float h, uint32_t line) {
    if (!out || !(l < h)) return;
    if (no > 0 && out[no - 1].line == line && out[no - 1].hi == l) {
      out[no - 1].hi = h;
    } else if (no < vs.s->max_cones) {
      out[no++] = Cone{l, h, line};
    } else {
      vs.overflow = true;
    }
  };
  for (uint32_t ci = 0; ci < vs.n_cones; ci++) {
    Cone c = cones[ci];
    if (!(c.hi > lo && c.lo < hi)) {
      if (c.lo >= hi && !tail_done) {
        tail_done = true;
        if (cur < hi) {
          visible = true;
          if (add_id >= 0) emit(cur, hi, (uint32_t)add_id);
        }
      }
      emit(c.lo, c.hi, c.line);
      continue;
    }
    if (c.lo > cur) {
      visible = true;
      if (add_id >= 0) emit(cur, c.lo, (uint32_t)add_id);
      cur = c.lo;
    }
    float olo = fmaxf(cur, c.lo), ohi = fminf(hi, c.hi);
    if (c.lo < olo) emit(c.lo, olo, c.line);
    if (olo < ohi) {
      V2 d = pdir(fmul(fadd(olo, ohi), 0.5f));
      float4 ol = vs.s->vis[c.line];
      float un = ray_depth(a, b, d), uc = ray_depth(V2{ol.x, ol.y}, V2{ol.z, ol.w}, d);
      if (un < uc) {
        visible = true;
        emit(olo, ohi, add_id >= 0 ? (uint32_t)add_id : c.line);
      } else {
        emit(olo, ohi, c.line);
      }
    }
    if (ohi < c.hi) emit(ohi, c.hi, c.line);
    if (c.hi > cur) cur = c.hi;
  }
  if (!tail_done && cur < hi) {
    visible = true;
    if (add_id >= 0) emit(cur, hi, (uint32_t)add_id);
  }
  if (n_out) *n_out = no;
  return visible;
}

// returns visibility; when `add`, *out_id is the new line id (valid only if visible)
__device__ bool vis_process(VisSet& vs, V2 a, V2 b, bool add, uint32_t* out_id) {
  float det = perp_dot(a, b);
  if (det < 0.0f) {
    V2 t = a;
    a = b;
    b = t;
  }
  if (!(det != 0.0f)) {
    // line through the origin: never an occluder; visible as a portal iff the origin is on it
    return !add && det == 0.0f && dot(a, b) <= 0.0f;
  }
  float lo = pangle(a), hi = pangle(b);
  if (lo == hi) return false;
  if (!add) {
    if (lo < hi) return vis_process_piece(vs, a, b, lo, hi, -1, nullptr);
    return vis_process_piece(vs, a, b, lo, 4.0f, -1, nullptr) ||
           vis_process_piece(vs, a, b, 0.0f, hi, -1, nullptr);
  }
  if (vs.n_lines >= vs.s->max_border) {
    vs.overflow = true;
    return false;
  }
  const int id = (int)vs.n_lines;
  vs.s->vis[id] = make_float4(a.x, a.y, b.x, b.y);  // registered so pieces can compare against it
  bool visible;
  uint32_t no = 0;
  if (lo < hi) {
    visible = vis_process_piece(vs, a, b, lo, hi, id, &no);
    if (visible) {
      vs.cur = 1 - vs.cur;
      vs.n_cones = no;
    }
  } else {
    // span crosses the cut at pseudo-angle 0/4: two pieces, each committed if visible
    // (inserting an invisible piece would reproduce the list unchanged)
    bool v1 = vis_process_piece(vs, a, b, lo, 4.0f, id, &no);
    if (v1) {
      vs.cur = 1 - vs.cur;
      vs.n_cones = no;
    }
    bool v2 = vis_process_piece(vs, a, b, 0.0f, hi, id, &no);
    if (v2) {
      vs.cur = 1 - vs.cur;
      vs.n_cones = no;
    }
    visible = v1 || v2;
  }
  if (visible) {
    vs.n_lines++;
    if (out_id) *out_id = (uint32_t)id;
  }
  return visible;
}

// ---------------------------------------------------------------------------
// std BinaryHeap<ExploreNode> (avoidance.rs:200-226): a <= b  <=>  b.score <= a.score
// ---------------------------------------------------------------------------
struct ExploreHeap {
  Strided<uint2> data;
  uint32_t len;
  __device__ __forceinline__ static bool le(uint2 a, uint2 b) {
    return __uint_as_float(b.y) <= __uint_as_float(a.y);
  }
  __device__ void sift_up(uint32_t start, uint32_t pos, uint2 hole) {
    while (pos > start) {
      uint32_t parent = (pos - 1) >> 1;
      uint2 p = data[parent];
      if (le(hole, p)) break;
      data[pos] = p;
      pos = parent;
    }
    data[pos] = hole;
  }
  __device__ void push(uint2 e) { sift_up(0, len++, e); }
  __device__ uint2 pop() {
    uint2 item = data[--len];
    if (len > 0) {
      uint2 hole = item;
      item = data[0];
      uint32_t end = len, pos = 0, child = 1;
      while (end >= 2 && child <= end - 2) {
        uint2 x = data[child], y = data[child + 1];
        bool right = le(x, y);
        data[pos] = right ? y : x;
        pos = child + (right ? 1u : 0u);
        child = 2 * pos + 1;
      }
      if (child == end - 1) {
        data[pos] = data[child];
        pos = child;
      }
      sift_up(0, pos, hole);
    }
    return item;
  }
};

__device__ __forceinline__ V2 world_xy(const DevNav& nav, uint32_t gv) {
  return {__ldg(nav.verts_world + 3 * (size_t)gv), __ldg(nav.verts_world + 3 * (size_t)gv + 1)};
}

// vertex id -> obstacle vertex position (world xy); avoidance.rs:441-448
__device__ __forceinline__ V2 vid_point(const DevNav& nav, const Scratch& s, uint32_t vid) {
  if (vid & 0x80000000u) {
    float2 v = s.new_verts[vid & 0x7fffffffu];
    return {v.x, v.y};
  }
  V2 w = world_xy(nav, vid);
  return {fsub(w.x, 0.0f), fsub(w.y, 0.0f)};
}

// ---------------------------------------------------------------------------
// ORCA (mirrors oracle/avoidance_oracle.cpp)
// ---------------------------------------------------------------------------
struct DodgyAgent {
  V2 position, velocity;
  float radius, responsibility;
};

__device__ Line line_for_neighbour(const DodgyAgent& self, const DodgyAgent& other, float time_horizon,
                                   float time_step) {
  V2 relative_position = other.position - self.position;
  V2 relative_velocity = self.velocity - other.velocity;
  float dist_sq = length_squared(relative_position);
  float combined_radius = fadd(self.radius, other.radius);
  float combined_radius_sq = fmul(combined_radius, combined_radius);
  Line line;
  V2 u;
  bool inside_vo;
  if (dist_sq > combined_radius_sq) {
    float inv_time_horizon = fdiv(1.0f, time_horizon);
    V2 w = relative_velocity - relative_position * inv_time_horizon;
    float w_length_sq = length_squared(w);
    float dot1 = dot(w, relative_position);
    if (dot1 < 0.0f && fmul(dot1, dot1) > fmul(combined_radius_sq, w_length_sq)) {
      float w_length = fsqrt(w_length_sq);
      V2 unit_w = w / w_length;
      line.direction = {unit_w.y, -unit_w.x};
      float cutoff_radius = fmul(combined_radius, inv_time_horizon);
      u = unit_w * fsub(cutoff_radius, w_length);
      inside_vo = w_length_sq < fmul(cutoff_radius, cutoff_radius);
    } else {
      float leg = fsqrt(fsub(dist_sq, combined_radius_sq));
      // dodgy_2d: `determinant(..).signum()` — +0.0 is positive, an exactly symmetric approach takes the
      // left leg (RVO2's `> 0` would go right); pinned by lib_test.rs:426-486
      if (!signbit(perp_dot(relative_position, w))) {
        line.direction = V2{fsub(fmul(relative_position.x, leg), fmul(relative_position.y, combined_radius)),
                            fadd(fmul(relative_position.x, combined_radius), fmul(relative_position.y, leg))} /
                         dist_sq;
      } else {
        line.direction = -(V2{fadd(fmul(relative_position.x, leg), fmul(relative_position.y, combined_radius)),
                              fadd(fmul(-relative_position.x, combined_radius), fmul(relative_position.y, leg))} /
                           dist_sq);
      }
      float dot2 = dot(relative_velocity, line.direction);
      u = line.direction * dot2 - relative_velocity;
      inside_vo = perp_dot(line.direction, relative_velocity) < 0.0f;
    }
  } else {
    float inv_time_step = fdiv(1.0f, time_step);
    V2 w = relative_velocity - relative_position * inv_time_step;
    float w_length = length(w);
    V2 unit_w = w / w_length;
    line.direction = {unit_w.y, -unit_w.x};
    float cutoff_radius = fmul(combined_radius, inv_time_step);
    u = unit_w * fsub(cutoff_radius, w_length);
    inside_vo = w_length < cutoff_radius;
  }
  float responsibility =
      inside_vo ? fdiv(self.responsibility, fadd(self.responsibility, other.responsibility)) : 1.0f;
  line.point = self.velocity + u * responsibility;
  return line;
}

struct ObVertex {
  V2 point;
  bool convex, has_prev, has_next;
  V2 unit_dir, prev_unit_dir;
};

// vertex `i` of an obstacle whose vertices are the loop walked backwards from `tail`
struct ObstacleView {
  const DevNav* nav;
  const Scratch* s;
  uint32_t head, tail, n;
  bool closed;
};

__device__ ObVertex make_ob_vertex(const ObstacleView& ob, uint32_t node, bool is_first, bool is_last) {
  // obstacle order = reversed loop order: next in obstacle = loop prev, prev in obstacle = loop next
  const Scratch& s = *ob.s;
  ObVertex v;
  LoopNode ln = s.lnodes[node];
  v.point = vid_point(*ob.nav, s, ln.vid);
  v.has_prev = ob.closed || !is_first;
  v.has_next = ob.closed || !is_last;
  uint32_t prev_node = ln.next != LMB_NONE ? ln.next : ob.head;  // wrap for closed loops
  uint32_t next_node = ln.prev != LMB_NONE ? ln.prev : ob.tail;
  V2 prev = vid_point(*ob.nav, s, s.lnodes[prev_node].vid);
  V2 next = vid_point(*ob.nav, s, s.lnodes[next_node].vid);
  v.unit_dir = {0.0f, 0.0f};
  v.prev_unit_dir = {0.0f, 0.0f};
  if (v.has_next) v.unit_dir = normalize(next - v.point);
  if (v.has_prev) v.prev_unit_dir = normalize(v.point - prev);
  v.convex = (v.has_prev && v.has_next && ob.n > 2) ? perp_dot(prev - next, v.point - prev) >= 0.0f : true;
  return v;
}

__device__ void lines_for_obstacle(const DodgyAgent& self, const ObstacleView& ob, float margin,
                                   float time_horizon, Strided<Line> lines, uint32_t& n_lines, uint32_t max_lines,
                                   bool& overflow) {
  if (ob.n < 2) return;
  const Scratch& s = *ob.s;
  // clearance from obstacle edges = obstacle margin alone (see oracle/avoidance_oracle.cpp)
  const float radius = margin;
  const float inv_th = fdiv(1.0f, time_horizon);
  const uint32_t n_edges = ob.closed ? ob.n : ob.n - 1;
  uint32_t node = ob.tail;  // obstacle vertex 0 = last loop element
  ObVertex o1v = make_ob_vertex(ob, node, true, ob.n == 1);
  for (uint32_t e = 0; e < n_edges; e++) {
    uint32_t next_node = s.lnodes[node].prev != LMB_NONE ? s.lnodes[node].prev : ob.tail;
    ObVertex o2v = make_ob_vertex(ob, next_node, (e + 1) % ob.n == 0, (e + 1) % ob.n == ob.n - 1);
    const ObVertex* o1 = &o1v;
    const ObVertex* o2 = &o2v;
    auto push = [&](const Line& l) {
      if (n_lines < max_lines)
        lines[n_lines++] = l;
      else
        overflow = true;
    };
    do {
      V2 rel1 = o1->point - self.position, rel2 = o2->point - self.position;
      if (!(perp_dot(o2->point - o1->point, self.position - o1->point) < 0.0f)) break;
      bool covered = false;
      for (uint32_t j = 0; j < n_lines; j++) {
        Line l = lines[j];
        if (fsub(perp_dot(rel1 * inv_th - l.point, l.direction), fmul(inv_th, radius)) >= -RVO_EPSILON &&
            fsub(perp_dot(rel2 * inv_th - l.point, l.direction), fmul(inv_th, radius)) >= -RVO_EPSILON) {
          covered = true;
          break;
        }
      }
      if (covered) break;
      float dist_sq1 = length_squared(rel1), dist_sq2 = length_squared(rel2);
      float radius_sq = fmul(radius, radius);
      V2 obstacle_vector = o2->point - o1->point;
      float sp = fdiv(dot(-rel1, obstacle_vector), length_squared(obstacle_vector));
      float dist_sq_line = length_squared(-rel1 - obstacle_vector * sp);
      Line line;
      if (sp < 0.0f && dist_sq1 <= radius_sq) {
        if (o1->convex) {
          line.point = {0.0f, 0.0f};
          line.direction = normalize(V2{-rel1.y, rel1.x});
          push(line);
        }
        break;
      } else if (sp > 1.0f && dist_sq2 <= radius_sq) {
        if (o2->convex && (!o2->has_next || perp_dot(rel2, o2->unit_dir) >= 0.0f)) {
          line.point = {0.0f, 0.0f};
          line.direction = normalize(V2{-rel2.y, rel2.x});
          push(line);
        }
        break;
      } else if (sp >= 0.0f && sp <= 1.0f && dist_sq_line <= radius_sq) {
        line.point = {0.0f, 0.0f};
        line.direction = -o1->unit_dir;
        push(line);
        break;
      }
      V2 left_leg, right_leg;
      V2 edge_dir = o1->unit_dir;
      if (sp < 0.0f && dist_sq_line <= radius_sq) {
        if (!o1->convex) break;
        o2 = o1;
        float leg1 = fsqrt(fsub(dist_sq1, radius_sq));
        left_leg = V2{fsub(fmul(rel1.x, leg1), fmul(rel1.y, radius)), fadd(fmul(rel1.x, radius), fmul(rel1.y, leg1))} / dist_sq1;
        right_leg = V2{fadd(fmul(rel1.x, leg1), fmul(rel1.y, radius)), fadd(fmul(-rel1.x, radius), fmul(rel1.y, leg1))} / dist_sq1;
      } else if (sp > 1.0f && dist_sq_line <= radius_sq) {
        if (!o2->convex) break;
        o1 = o2;
        float leg2 = fsqrt(fsub(dist_sq2, radius_sq));
        left_leg = V2{fsub(fmul(rel2.x, leg2), fmul(rel2.y, radius)), fadd(fmul(rel2.x, radius), fmul(rel2.y, leg2))} / dist_sq2;
        right_leg = V2{fadd(fmul(rel2.x, leg2), fmul(rel2.y, radius)), fadd(fmul(-rel2.x, radius), fmul(rel2.y, leg2))} / dist_sq2;
      } else {
        if (o1->convex) {
          float leg1 = fsqrt(fsub(dist_sq1, radius_sq));
          left_leg = V2{fsub(fmul(rel1.x, leg1), fmul(rel1.y, radius)), fadd(fmul(rel1.x, radius), fmul(rel1.y, leg1))} / dist_sq1;
        } else {
          left_leg = -edge_dir;
        }
        if (o2->convex) {
          float leg2 = fsqrt(fsub(dist_sq2, radius_sq));
          right_leg = V2{fadd(fmul(rel2.x, leg2), fmul(rel2.y, radius)), fadd(fmul(-rel2.x, radius), fmul(rel2.y, leg2))} / dist_sq2;
        } else {
          right_leg = edge_dir;
        }
      }
      rel1 = o1->point - self.position;
      rel2 = o2->point - self.position;
      bool left_foreign = false, right_foreign = false;
      if (o1->convex && o1->has_prev && perp_dot(left_leg, -o1->prev_unit_dir) >= 0.0f) {
        left_leg = -o1->prev_unit_dir;
        left_foreign = true;
      }
      if (o2->convex && o2->has_next && perp_dot(right_leg, o2->unit_dir) <= 0.0f) {
        right_leg = o2->unit_dir;
        right_foreign = true;
      }
      V2 left_cutoff = rel1 * inv_th, right_cutoff = rel2 * inv_th;
      V2 cutoff_vec = right_cutoff - left_cutoff;
      bool same = (o1 == o2);
      float t = same ? 0.5f : fdiv(dot(self.velocity - left_cutoff, cutoff_vec), length_squared(cutoff_vec));
      float t_left = dot(self.velocity - left_cutoff, left_leg);
      float t_right = dot(self.velocity - right_cutoff, right_leg);
      if ((t < 0.0f && t_left < 0.0f) || (same && t_left < 0.0f && t_right < 0.0f)) {
        V2 unit_w = normalize(self.velocity - left_cutoff);
        line.direction = {unit_w.y, -unit_w.x};
        line.point = left_cutoff + unit_w * fmul(radius, inv_th);
        push(line);
        break;
      } else if (t > 1.0f && t_right < 0.0f) {
        V2 unit_w = normalize(self.velocity - right_cutoff);
        line.direction = {unit_w.y, -unit_w.x};
        line.point = right_cutoff + unit_w * fmul(radius, inv_th);
        push(line);
        break;
      }
      float dist_sq_cutoff = (t < 0.0f || t > 1.0f || same)
                                 ? F_INF
                                 : length_squared(self.velocity - (left_cutoff + cutoff_vec * t));
      float dist_sq_left = t_left < 0.0f ? F_INF : length_squared(self.velocity - (left_cutoff + left_leg * t_left));
      float dist_sq_right = t_right < 0.0f ? F_INF : length_squared(self.velocity - (right_cutoff + right_leg * t_right));
      if (dist_sq_cutoff <= dist_sq_left && dist_sq_cutoff <= dist_sq_right) {
        line.direction = -edge_dir;
        line.point = left_cutoff + V2{-line.direction.y, line.direction.x} * fmul(radius, inv_th);
        push(line);
      } else if (dist_sq_left <= dist_sq_right) {
        if (left_foreign) break;
        line.direction = left_leg;
        line.point = left_cutoff + V2{-line.direction.y, line.direction.x} * fmul(radius, inv_th);
        push(line);
      } else {
        if (right_foreign) break;
        line.direction = -right_leg;
        line.point = right_cutoff + V2{-line.direction.y, line.direction.x} * fmul(radius, inv_th);
        push(line);
      }
    } while (false);
    node = next_node;
    o1v = o2v;
  }
}

__device__ bool linear_program_1(Strided<Line> lines, uint32_t line_no, float radius, V2 opt_velocity,
                                 bool direction_opt, V2& result) {
  const Line ln = lines[line_no];
  float dot_product = dot(ln.point, ln.direction);
  float discriminant = fsub(fadd(fmul(dot_product, dot_product), fmul(radius, radius)), length_squared(ln.point));
  if (discriminant < 0.0f) return false;
  float sqrt_discriminant = fsqrt(discriminant);
  float t_left = fsub(-dot_product, sqrt_discriminant);
  float t_right = fadd(-dot_product, sqrt_discriminant);
  for (uint32_t i = 0; i < line_no; i++) {
    Line li = lines[i];
    float denominator = perp_dot(ln.direction, li.direction);
    float numerator = perp_dot(li.direction, ln.point - li.point);
    if (fabsf(denominator) <= RVO_EPSILON) {
      if (numerator < 0.0f) return false;
      continue;
    }
    float t = fdiv(numerator, denominator);
    if (denominator >= 0.0f)
      t_right = fminf(t_right, t);
    else
      t_left = fmaxf(t_left, t);
    if (t_left > t_right) return false;
  }
  if (direction_opt) {
    if (dot(opt_velocity, ln.direction) > 0.0f)
      result = ln.point + ln.direction * t_right;
    else
      result = ln.point + ln.direction * t_left;
  } else {
    float t = dot(ln.direction, opt_velocity - ln.point);
    if (t < t_left)
      result = ln.point + ln.direction * t_left;
    else if (t > t_right)
      result = ln.point + ln.direction * t_right;
    else
      result = ln.point + ln.direction * t;
  }
  return true;
}

__device__ uint32_t linear_program_2(Strided<Line> lines, uint32_t n, float radius, V2 opt_velocity,
                                     bool direction_opt, V2& result) {
  if (direction_opt) {
    result = opt_velocity * radius;
  } else if (length_squared(opt_velocity) > fmul(radius, radius)) {
    result = normalize(opt_velocity) * radius;
  } else {
    result = opt_velocity;
  }
  for (uint32_t i = 0; i < n; i++) {
    Line li = lines[i];
    if (perp_dot(li.direction, li.point - result) > 0.0f) {
      V2 temp = result;
      if (!linear_program_1(lines, i, radius, opt_velocity, direction_opt, result)) {
        result = temp;
        return i;
      }
    }
  }
  return n;
}

// Returns the size of the last projected constraint set left in `proj` (the fallback constraints
// the debug-avoidance export reports, debug.rs:470-481).
__device__ uint32_t linear_program_3(Strided<Line> lines, uint32_t n, uint32_t num_obst_lines, uint32_t begin_line,
                                     float radius, V2& result, Strided<Line> proj) {
  float distance_ = 0.0f;
  uint32_t last_np = 0;
  for (uint32_t i = begin_line; i < n; i++) {
    Line li = lines[i];
    if (perp_dot(li.direction, li.point - result) > distance_) {
      uint32_t np = 0;
      for (uint32_t j = 0; j < num_obst_lines; j++) proj[np++] = lines[j];
      for (uint32_t j = num_obst_lines; j < i; j++) {
        Line lj = lines[j];
        Line line;
        float determinant = perp_dot(li.direction, lj.direction);
        if (fabsf(determinant) <= RVO_EPSILON) {
          if (dot(li.direction, lj.direction) > 0.0f) continue;
          line.point = (li.point + lj.point) * 0.5f;
        } else {
          line.point = li.point + li.direction * fdiv(perp_dot(lj.direction, li.point - lj.point), determinant);
        }
        line.direction = normalize(lj.direction - li.direction);
        proj[np++] = line;
      }
      V2 temp = result;
      if (linear_program_2(proj, np, radius, V2{-li.direction.y, li.direction.x}, true, result) < np) result = temp;
      distance_ = perp_dot(li.direction, li.point - result);
      last_np = np;
    }
  }
  return last_np;
}

}  // namespace

// ---------------------------------------------------------------------------
// records + grid
// ---------------------------------------------------------------------------
__global__ void k_build_records(AgentArrays ag, AgentPersist persist, lmb_options o, AvoidRecord* records,
                                float* max_radius_bits) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ag.n) return;
  AvoidRecord r;
  uint32_t node = ag.agent_node[i];
  r.valid = node != LMB_NONE ? 1u : 0u;
  r.px = ag.agent_point[3 * (size_t)i];
  r.py = ag.agent_point[3 * (size_t)i + 1];
  r.pz = ag.agent_point[3 * (size_t)i + 2];
  r.vx = ag.velocity[3 * (size_t)i];
  r.vy = ag.velocity[3 * (size_t)i + 1];
  r.radius = ag.radius[i];
  r.responsibility = persist.state[ag.slot[i]] == LMB_STATE_REACHED_TARGET
                         ? o.reached_destination_avoidance_responsibility
                         : 1.0f;
  records[i] = r;
}

__global__ void k_character_records(uint32_t n, const float* velocity, const float* radius, const float* points,
                                    const uint32_t* nodes, AvoidRecord* records) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AvoidRecord r;
  r.valid = nodes[i] != LMB_NONE ? 1u : 0u;
  r.px = points[3 * (size_t)i];
  r.py = points[3 * (size_t)i + 1];
  r.pz = points[3 * (size_t)i + 2];
  r.vx = velocity[3 * (size_t)i];
  r.vy = velocity[3 * (size_t)i + 1];
  r.radius = radius[i];
  r.responsibility = 0.0f;
  records[i] = r;
}

// max radius over valid records (agent_max_radius, avoidance.rs:33,57): radii are >= 0 so
// the float bit pattern orders like an unsigned integer.
__global__ void k_max_radius(uint32_t n, const AvoidRecord* records, uint32_t* max_bits) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  float r = 0.0f;
  if (i < n && records[i].valid) r = fmaxf(records[i].radius, 0.0f);
  for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  if ((threadIdx.x & 31) == 0 && r > 0.0f) atomicMax(max_bits, __float_as_uint(r));
}

__device__ __forceinline__ uint32_t avoid_cell(const AvoidGrid& g, float x, float y) {
  int32_t cx = (int32_t)floorf((x - g.ox) * g.inv_cell), cy = (int32_t)floorf((y - g.oy) * g.inv_cell);
  cx = cx < 0 ? 0 : (cx >= (int32_t)g.nx ? (int32_t)g.nx - 1 : cx);
  cy = cy < 0 ? 0 : (cy >= (int32_t)g.ny ? (int32_t)g.ny - 1 : cy);
  return (uint32_t)cy * g.nx + (uint32_t)cx;
}

__global__ void k_grid_count(AvoidGrid g, uint32_t n, const AvoidRecord* records) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !records[i].valid) return;
  atomicAdd(&g.cell_count[avoid_cell(g, records[i].px, records[i].py)], 1u);
}
__global__ void k_grid_fill(AvoidGrid g, uint32_t n, const AvoidRecord* records) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !records[i].valid) return;
  uint32_t c = avoid_cell(g, records[i].px, records[i].py);
  uint32_t k = atomicAdd(&g.cell_count[c], 1u);  // cell_count was reset to 0 after the scan
  AvoidRecord r = records[i];
  r.valid = i;  // every record in the grid is valid: the word carries the record's index instead
  g.sorted_rec[g.cell_start[c] + k] = r;
}

// ---------------------------------------------------------------------------
// the per-agent avoidance kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(64, 14)
k_avoidance(DevNav nav, AgentArrays ag, lmb_options o, float delta_time, uint32_t first_agent_global,
            uint32_t chunk_begin, uint32_t chunk_n, uint32_t n_records, const AvoidRecord* records,
            uint32_t n_characters, const AvoidRecord* char_records, AvoidGrid grid, AvoidScratch scratch,
            const uint32_t* max_radius_bits, float* out_move, AvoidDebug dbg, AvoidOrder ord) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= chunk_n) return;
  uint32_t i = chunk_begin + t;  // local dense agent index
  if (ord.mode == 1) {
    // every record is this rank's own: the agents in the grid's cell order (AvoidOrder)
    if (i >= grid.cell_start[(size_t)grid.nx * grid.ny]) return;  // positions beyond the valid records
    i = grid.sorted_rec[i].valid;                                   // (the word carries the record's index)
  } else if (ord.mode == 2) {
    if (i >= ord.pos[n_records]) return;
    i = ord.order[i];
  }
  const uint32_t gi = first_agent_global + i;  // index into the (gathered) record array
  const AvoidRecord me = records[gi];
  if (!me.valid) return;
  Scratch s = carve(scratch, t);

  bool overflow = false;
  if (delta_time == 0.0f) delta_time = 1.0f;

  const float agent_max_radius = __uint_as_float(*max_radius_bits);
  const float neighbourhood = fadd(agent_max_radius, o.neighbourhood);
  const float neighbourhood_squared = fmul(neighbourhood, neighbourhood);

  // ---- neighbours: agents (ascending (d2, index)), then characters ----
  // (measured and not kept: the first 32 entries of the list in shared memory — config 3: 5.7 -> 7.4 ms; the
  // interleaved HBM scratch already hits L1 and the two-tier accessor costs more than it saves)
  auto nb_get = [&](uint32_t k) -> float2 { return s.nb[k]; };
  auto nb_set = [&](uint32_t k, float2 v) { s.nb[k] = v; };
  uint32_t n_nb = 0;
  {
    int32_t cx0 = (int32_t)floorf((me.px - neighbourhood - grid.ox) * grid.inv_cell) - 1;
    int32_t cx1 = (int32_t)floorf((me.px + neighbourhood - grid.ox) * grid.inv_cell) + 1;
    int32_t cy0 = (int32_t)floorf((me.py - neighbourhood - grid.oy) * grid.inv_cell) - 1;
    int32_t cy1 = (int32_t)floorf((me.py + neighbourhood - grid.oy) * grid.inv_cell) + 1;
    cx0 = cx0 < 0 ? 0 : cx0;
    cy0 = cy0 < 0 ? 0 : cy0;
    cx1 = cx1 >= (int32_t)grid.nx ? (int32_t)grid.nx - 1 : cx1;
    cy1 = cy1 >= (int32_t)grid.ny ? (int32_t)grid.ny - 1 : cy1;
    for (int32_t cy = cy0; cy <= cy1; cy++) {
      {
        // Cells cx0..cx1 of a grid row are consecutive cell ids, and k_grid_fill left a COPY of every record in
        // cell order: the candidates of one row are one contiguous run of 32-byte records (config 3: ~10 per
        // row) instead of a cell_start pair per cell and a random sector per record (5.7 -> 4.8 ms per M agents).
        const uint32_t c0 = (uint32_t)cy * grid.nx + (uint32_t)cx0, c1 = (uint32_t)cy * grid.nx + (uint32_t)cx1;
        const uint32_t kb = grid.cell_start[c0], ke = grid.cell_start[c1 + 1];
        // Four candidates at a time: the first half of each record (position) is loaded for all four before any is
        // used (independent loads in flight together); the second half (radius, index) only for the ~1 in 3 that
        // passes the first distance test.
        const float4* rec4 = reinterpret_cast<const float4*>(grid.sorted_rec);
        for (uint32_t k0 = kb; k0 < ke; k0 += 4) {
          float4 pa[4];
#pragma unroll
          for (uint32_t u = 0; u < 4; u++) pa[u] = __ldg(rec4 + 2 * (size_t)min(k0 + u, ke - 1u));
#pragma unroll
          for (uint32_t u = 0; u < 4; u++) {
            if (k0 + u >= ke) continue;
            float dx = fsub(pa[u].x, me.px), dy = fsub(pa[u].y, me.py), dz = fsub(pa[u].z, me.pz);
            float d2 = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
            if (!(d2 <= neighbourhood_squared)) continue;
            const float4 pb = __ldg(rec4 + 2 * (size_t)(k0 + u) + 1);  // {vy, radius, responsibility, index}
            const uint32_t j = __float_as_uint(pb.w);                  // the record's index (k_grid_fill)
            if (j == gi) continue;
            float nh = fadd(o.neighbourhood, pb.y);
            if (!(d2 < fmul(nh, nh))) continue;
            // appended unsorted: the (d2, index) order is produced where it is consumed (the ORCA lines below)
            if (n_nb >= s.max_nb) {
              overflow = true;
              continue;
            }
            nb_set(n_nb++, make_float2(d2, __uint_as_float(j)));
          }
        }
      }
    }
  }
  const uint32_t n_agent_nb = n_nb;
  for (uint32_t j = 0; j < n_characters; j++) {
    const AvoidRecord r = char_records[j];
    if (!r.valid) continue;
    float dx = fsub(r.px, me.px), dy = fsub(r.py, me.py), dz = fsub(r.pz, me.pz);
    float d2 = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    if (!(d2 <= neighbourhood_squared) || !(d2 < neighbourhood_squared)) continue;
    if (n_nb >= s.max_nb) {
      overflow = true;
      continue;
    }
    nb_set(n_nb++, make_float2(d2, __uint_as_float(j | 0x80000000u)));
  }

  // ---- obstacles: nav_mesh_borders_to_dodgy_obstacles (avoidance.rs:189-471) ----
  const V2 agent_point{me.px, me.py};
  const float distance_limit = fmul(o.neighbourhood, o.neighbourhood);
  VisSet vs;
  vs.s = &s;
  vs.n_lines = 0;
  vs.n_cones = 0;
  vs.cur = 0;
  vs.overflow = false;
  uint32_t n_new_verts = 0;
  {
    ExploreHeap heap;
    heap.data = s.heap;
    heap.len = 0;
    uint32_t n_explored = 0;
    const uint32_t ex_mask = 4u * s.max_explore - 1u;  // max_explore is a power of two (64 << k)
    constexpr uint32_t EX_LINEAR = 8;
    // slot of `node` in the hash table (true) or the empty slot where it belongs (false)
    auto ex_probe = [&](uint32_t node, uint32_t& slot) -> bool {
      for (uint32_t hh = (node * 2654435761u) >> 7;; hh++) {
        const uint32_t v = s.explored[hh & ex_mask];
        if (v == node || v == LMB_NONE) {
          slot = hh & ex_mask;
          return v == node;
        }
      }
    };
    heap.push(make_uint2(ag.agent_node[i], __float_as_uint(0.0f)));
    while (heap.len > 0) {
      const uint32_t node = heap.pop().x;
      // `explored.insert(node)` (avoidance.rs:228-232).  Up to EX_LINEAR nodes the set is a list scanned linearly
      // (a coarse mesh: a handful of polygons per flood, and entry k of 32 neighbouring agents is one line).  On a
      // fine mesh a flood visits hundreds of polygons and pops each several times — the linear scan was half of
      // the kernel's instructions there (ncu, config 1) — so beyond that the set is a hash table with linear
      // probing, at most half full.
      bool seen = false;
      uint32_t h = 0;
      if (n_explored <= EX_LINEAR) {
        for (uint32_t k = 0; k < n_explored; k++)
          if (s.explored_list[k] == node) {
            seen = true;
            break;
          }
      } else {
        seen = ex_probe(node, h);
      }
      if (seen) continue;
      if (n_explored >= s.max_explore * 2) {
        overflow = true;
        break;
      }
      s.explored_list[n_explored] = node;
      if (n_explored == EX_LINEAR) {  // the list outgrows the linear scan: hash what it holds
        for (uint32_t k = 0; k < EX_LINEAR; k++) {
          uint32_t hk;
          ex_probe(s.explored_list[k], hk);
          s.explored[hk] = s.explored_list[k];
          s.explored_slot[k] = hk;
        }
        ex_probe(node, h);
      }
      if (n_explored >= EX_LINEAR) {
        s.explored[h] = node;
        s.explored_slot[n_explored] = h;
      }
      n_explored++;
      const uint32_t eb = nav.poly_edge_begin[node], n = nav.poly_edge_begin[node + 1] - eb;
      auto consider = [&](V2 v1, V2 v2, uint32_t dst) {
        if (!vis_process(vs, v1, v2, false, nullptr)) return;
        float score = fminf(length_squared(v1), length_squared(v2));
        if (score < distance_limit) {
          if (heap.len >= s.max_explore) {
            overflow = true;
            return;
          }
          heap.push(make_uint2(dst, __float_as_uint(score)));
        }
      };
      for (uint32_t e = 0; e < n; e++) {
        uint32_t nb = nav.edge_neighbor[eb + e];
        if (nb == LMB_NONE) continue;
        uint32_t l = nav.edge_vertex[eb + (e == n - 1 ? 0 : e + 1)], r = nav.edge_vertex[eb + e];
        consider(world_xy(nav, r) - agent_point, world_xy(nav, l) - agent_point, nb);
      }
      for (uint32_t k = nav.node_link_begin[node]; k < nav.node_link_begin[node + 1]; k++) {
        uint32_t lk = nav.node_links[k];
        if (nav.link_kind[lk] != LMB_LINK_BOUNDARY) continue;
        const float* pp = nav.link_portal + 6 * (size_t)lk;
        consider(V2{pp[3], pp[4]} - agent_point, V2{pp[0], pp[1]} - agent_point, nav.link_dst_node[lk]);
      }
      const uint32_t mod = nav.node_modified[node];
      if (mod != LMB_NONE) {
        const DevIsland& is = nav.islands[nav.poly_island[node]];
        const uint32_t island_verts = is.vert_end - is.vert_begin;
        const uint32_t vb = nav.modified_vert_begin[mod];
        for (uint32_t k = nav.modified_edge_begin[mod]; k < nav.modified_edge_begin[mod + 1]; k++) {
          uint32_t idx[2] = {nav.modified_edges[2 * k], nav.modified_edges[2 * k + 1]};
          V2 pt[2];
          uint32_t vid[2];
          for (int q = 0; q < 2; q++) {
            if (idx[q] >= island_verts) {
              uint32_t v = vb + (idx[q] - island_verts);
              V2 nv{nav.modified_verts[2 * (size_t)v], nav.modified_verts[2 * (size_t)v + 1]};
              if (n_new_verts >= s.max_border) {
                overflow = true;
                vid[q] = 0x80000000u;
              } else {
                vid[q] = 0x80000000u | n_new_verts;
                s.new_verts[n_new_verts++] = make_float2(nv.x, nv.y);
              }
              pt[q] = nv - agent_point;
            } else {
              vid[q] = is.vert_begin + idx[q];
              pt[q] = world_xy(nav, vid[q]) - agent_point;
            }
          }
          uint32_t id;
          if (vis_process(vs, pt[0], pt[1], true, &id)) s.border[id] = make_uint2(vid[0], vid[1]);
        }
      } else {
        for (uint32_t e = 0; e < n; e++) {
          if (nav.edge_neighbor[eb + e] != LMB_NONE) continue;
          uint32_t bv2 = nav.edge_vertex[eb + (e == n - 1 ? 0 : e + 1)], bv1 = nav.edge_vertex[eb + e];
          uint32_t id;
          if (vis_process(vs, world_xy(nav, bv1) - agent_point, world_xy(nav, bv2) - agent_point, true, &id))
            s.border[id] = make_uint2(bv1, bv2);
        }
      }
    }
    // leave the explored table empty for the next agent that gets this scratch
    if (n_explored > EX_LINEAR)
      for (uint32_t k = 0; k < n_explored; k++) s.explored[s.explored_slot[k]] = LMB_NONE;
  }
  overflow = overflow || vs.overflow;

  // ---- stitch visible lines into loops (avoidance.rs:389-439), ascending line id ----
  uint32_t n_loops = 0, n_finished = 0, n_lnodes = 0;
  if (!overflow) {
    const Strided<Cone> cones = s.cones[vs.cur];
    for (uint32_t line_id = 0; line_id < vs.n_lines; line_id++) {
      bool vis = false;
      for (uint32_t c = 0; c < vs.n_cones; c++)
        if (cones[c].line == line_id) {
          vis = true;
          break;
        }
      if (!vis) continue;
      const uint2 edge = s.border[line_id];
      int left_loop = -1, right_loop = -1;
      for (uint32_t li = 0; li < n_loops; li++) {
        uint2 lp = s.loops[li];
        if (left_loop < 0 && s.lnodes[lp.y].vid == edge.x) {
          left_loop = (int)li;
          if (right_loop >= 0) break;
        }
        if (right_loop < 0 && s.lnodes[lp.x].vid == edge.y) {
          right_loop = (int)li;
          if (left_loop >= 0) break;
        }
      }
      auto new_node = [&](uint32_t vid) {
        uint32_t k = n_lnodes++;
        s.lnodes[k] = LoopNode{vid, LMB_NONE, LMB_NONE};
        return k;
      };
      if (left_loop < 0 && right_loop < 0) {
        uint32_t a = new_node(edge.x), b = new_node(edge.y);
        s.lnodes[a].next = b;
        s.lnodes[b].prev = a;
        s.loops[n_loops++] = make_uint2(a, b);
      } else if (left_loop >= 0 && right_loop < 0) {
        uint32_t b = new_node(edge.y), tail = s.loops[left_loop].y;
        s.lnodes[tail].next = b;
        s.lnodes[b].prev = tail;
        s.loops[left_loop].y = b;
      } else if (left_loop < 0) {
        uint32_t a = new_node(edge.x), head = s.loops[right_loop].x;
        s.lnodes[a].next = head;
        s.lnodes[head].prev = a;
        s.loops[right_loop].x = a;
      } else if (left_loop == right_loop) {
        s.finished[n_finished++] = s.loops[left_loop];
        for (uint32_t k = (uint32_t)left_loop; k + 1 < n_loops; k++) s.loops[k] = s.loops[k + 1];  // Vec::remove
        n_loops--;
      } else {
        // swap_remove(left), swap_remove(right'), push(left ++ right)
        uint2 left = s.loops[left_loop];
        s.loops[left_loop] = s.loops[n_loops - 1];
        n_loops--;
        uint32_t ri = ((uint32_t)right_loop == n_loops) ? (uint32_t)left_loop : (uint32_t)right_loop;
        uint2 right = s.loops[ri];
        s.loops[ri] = s.loops[n_loops - 1];
        n_loops--;
        s.lnodes[left.y].next = right.x;
        s.lnodes[right.x].prev = left.y;
        s.loops[n_loops++] = make_uint2(left.x, right.y);
      }
    }
  }

  // ---- ORCA lines + LP ----
  DodgyAgent self{V2{me.px, me.py}, V2{me.vx, me.vy}, me.radius, me.responsibility};
  uint32_t n_lines = 0;
  if (!overflow) {
    for (uint32_t pass = 0; pass < 2; pass++) {
      const Strided<uint2> list = pass == 0 ? s.finished : s.loops;
      const uint32_t cnt = pass == 0 ? n_finished : n_loops;
      for (uint32_t k = 0; k < cnt; k++) {
        ObstacleView ob;
        ob.nav = &nav;
        ob.s = &s;
        ob.head = list[k].x;
        ob.tail = list[k].y;
        ob.closed = pass == 0;
        uint32_t cntv = 1;
        for (uint32_t q = ob.head; q != ob.tail; q = s.lnodes[q].next) cntv++;
        ob.n = cntv;
        lines_for_obstacle(self, ob, 0.0f, o.obstacle_avoidance_time_horizon, s.lines, n_lines, s.max_lines,
                           overflow);
      }
    }
  }
  const uint32_t obstacle_line_count = n_lines;
  if (!overflow) {
    // Neighbours in kdtree `within` order — agents by ascending (squared distance, index), then characters the
    // same way (avoidance.rs:90-130) — by SELECTION: for every output position one pass over the unsorted list
    // for the smallest key above the previous one.  n^2 independent loads instead of the n^2 / 4 DEPENDENT
    // load-compare-store steps of an insertion sort in the HBM scratch (ncu, config 3: 31 % of the kernel's stall
    // samples sat on that one load).  Keys are unique (distinct indices).
    // The record of neighbour k (a random 32-byte sector) is requested BEFORE the selection pass for neighbour
    // k + 1 runs and used after it, so the pass hides the round trip.
    auto select_next = [&](uint32_t k, float last_d2, uint32_t last_j, float& out_d2) -> uint32_t {
      const uint32_t pb = k < n_agent_nb ? 0u : n_agent_nb, pe = k < n_agent_nb ? n_agent_nb : n_nb;
      if (k == n_agent_nb) last_d2 = -1.0f, last_j = 0;
      float bd2 = F_INF;
      uint32_t j = 0xffffffffu;
      for (uint32_t m = pb; m < pe; m++) {
        const float2 q = nb_get(m);
        const uint32_t qj = __float_as_uint(q.y);
        const bool after = q.x > last_d2 || (q.x == last_d2 && qj > last_j);
        const bool better = q.x < bd2 || (q.x == bd2 && qj < j);
        if (after && better) bd2 = q.x, j = qj;
      }
      out_d2 = bd2;
      return j;
    };
    float cur_d2 = -1.0f;
    uint32_t cur_j = n_nb ? select_next(0, -1.0f, 0u, cur_d2) : 0u;
    for (uint32_t k = 0; k < n_nb; k++) {
      const uint32_t j = cur_j;
      const AvoidRecord r = (j & 0x80000000u) ? char_records[j & 0x7fffffffu] : records[j];
      if (k + 1 < n_nb) {
        float nd2;
        cur_j = select_next(k + 1, cur_d2, j, nd2);
        cur_d2 = nd2;
      }
      DodgyAgent other{V2{r.px, r.py}, V2{r.vx, r.vy}, r.radius, r.responsibility};
      if (n_lines >= s.max_lines) {
        overflow = true;
        break;
      }
      s.lines[n_lines++] = line_for_neighbour(self, other, o.avoidance_time_horizon, delta_time);
    }
  }
  if (overflow) {
    atomicAdd(scratch.overflow, 1u);
    return;
  }
  const uint32_t slot = ag.slot[i];
  (void)slot;
  V2 preferred{out_move[3 * (size_t)i], out_move[3 * (size_t)i + 1]};
  V2 result;
  uint32_t fail = linear_program_2(s.lines, n_lines, ag.max_speed[i], preferred, false, result);
  uint32_t n_fallback = 0;
  if (fail < n_lines)
    n_fallback = linear_program_3(s.lines, n_lines, obstacle_line_count, fail, ag.max_speed[i], result, s.proj);
  if (dbg.row_of) {
    // debug-avoidance export (avoidance.rs:161-172): the constraint lines of agents that asked for them
    const uint32_t row = dbg.row_of[i];
    if (row != LMB_NONE) {
      uint32_t* h = dbg.header + 4 * (size_t)row;
      float4* out = dbg.lines + (size_t)row * dbg.lines_per_row;
      const uint32_t no = n_lines < dbg.lines_per_row ? n_lines : dbg.lines_per_row;
      const uint32_t nf = n_fallback < dbg.lines_per_row - no ? n_fallback : dbg.lines_per_row - no;
      for (uint32_t k = 0; k < no; k++)
        out[k] = make_float4(s.lines[k].point.x, s.lines[k].point.y, s.lines[k].direction.x, s.lines[k].direction.y);
      for (uint32_t k = 0; k < nf; k++)
        out[no + k] = make_float4(s.proj[k].point.x, s.proj[k].point.y, s.proj[k].direction.x, s.proj[k].direction.y);
      h[0] = 1u;  // avoidance ran for this agent in this step
      h[1] = no;
      h[2] = nf;
      h[3] = fail < n_lines ? 1u : 0u;  // DebugData::Fallback
    }
  }
  out_move[3 * (size_t)i] = result.x;
  out_move[3 * (size_t)i + 1] = result.y;
  out_move[3 * (size_t)i + 2] = 0.0f;
}

// desired_move (by slot) -> staging (dense), and back
__global__ void k_stage_moves(AgentArrays ag, AgentPersist persist, float* staged) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ag.n) return;
  uint32_t s = ag.slot[i];
  for (int k = 0; k < 3; k++) staged[3 * (size_t)i + k] = persist.desired_move[3 * (size_t)s + k];
}
// (skipped when some agent's scratch overflowed: the host then re-runs the phase with larger capacities and
// nothing of the aborted attempt may have been published)
__global__ void k_commit_moves(AgentArrays ag, AgentPersist persist, const float* staged, const uint32_t* overflow) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ag.n || *overflow) return;
  uint32_t s = ag.slot[i];
  for (int k = 0; k < 3; k++) persist.desired_move[3 * (size_t)s + k] = staged[3 * (size_t)i + k];
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline uint32_t blocks_for(uint32_t n, uint32_t b) { return (n + b - 1) / b; }

void launch_build_records(const AgentArrays& ag, const AgentPersist& persist, const lmb_options& o,
                          AvoidRecord* records, cudaStream_t stream) {
  if (ag.n) k_build_records<<<blocks_for(ag.n, 256), 256, 0, stream>>>(ag, persist, o, records, nullptr);
}
void launch_character_records(uint32_t n, const float* velocity, const float* radius, const float* points,
                              const uint32_t* nodes, AvoidRecord* records, cudaStream_t stream) {
  if (n) k_character_records<<<blocks_for(n, 256), 256, 0, stream>>>(n, velocity, radius, points, nodes, records);
}

size_t avoid_scratch_bytes_per_agent(const AvoidScratch& s) {
  size_t b = 0;
  b += (size_t)s.max_neighbours * 8;
  b += (size_t)s.max_explore * 8;
  b += (size_t)s.max_explore * 32;  // explored table (4 x) + the slots a flood filled (2 x) + the node list (2 x)
  b += (size_t)s.max_border_edges * 16;
  b += (size_t)s.max_border_edges * 8;
  b += (size_t)s.max_cones * 12 * 2;
  b += (size_t)s.max_border_edges * 8;
  b += (size_t)s.max_border_edges * 2 * 12;
  b += (size_t)s.max_border_edges * 8 * 2;
  b += (size_t)s.max_lines * 16 * 2;
  return (b + 127) & ~(size_t)127;
}

void launch_avoid_grid(uint32_t n_records, const AvoidRecord* records, AvoidGrid& grid, uint32_t* max_radius_bits,
                       cudaStream_t stream) {
  uint32_t cells = grid.nx * grid.ny;
  cudaMemsetAsync(grid.cell_count, 0, ((size_t)cells + 1) * sizeof(uint32_t), stream);
  cudaMemsetAsync(max_radius_bits, 0, sizeof(uint32_t), stream);
  if (n_records) {
    k_max_radius<<<blocks_for(n_records, 256), 256, 0, stream>>>(n_records, records, max_radius_bits);
    k_grid_count<<<blocks_for(n_records, 256), 256, 0, stream>>>(grid, n_records, records);
  }
  launch_exclusive_scan(grid.cell_count, grid.cell_start, cells + 1, grid.scan_tmp, stream);
  cudaMemsetAsync(grid.cell_count, 0, ((size_t)cells + 1) * sizeof(uint32_t), stream);
  if (n_records) k_grid_fill<<<blocks_for(n_records, 256), 256, 0, stream>>>(grid, n_records, records);
}

// AvoidOrder mode 2: this rank's records [first, first + n) picked out of the cell order
__global__ void k_order_flags(AvoidGrid g, uint32_t n_records, uint32_t first, uint32_t n, uint32_t* flags) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n_records) return;
  const uint32_t n_sorted = g.cell_start[(size_t)g.nx * g.ny];
  flags[p] = (p < n_sorted && g.sorted_rec[p].valid - first < n) ? 1u : 0u;
}
__global__ void k_order_scatter(AvoidGrid g, uint32_t n_records, uint32_t first, const uint32_t* flags, const uint32_t* pos,
                                uint32_t* order) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_records || !flags[p]) return;
  order[pos[p]] = g.sorted_rec[p].valid - first;
}
void launch_avoid_order(const AvoidGrid& grid, uint32_t n_records, uint32_t first, uint32_t n, const AvoidOrder& ord,
                        cudaStream_t stream) {
  if (ord.mode != 2 || !n_records) return;
  k_order_flags<<<blocks_for(n_records + 1, 256), 256, 0, stream>>>(grid, n_records, first, n, ord.flags);
  launch_exclusive_scan(ord.flags, ord.pos, n_records + 1, ord.scan_tmp, stream);
  k_order_scatter<<<blocks_for(n_records, 256), 256, 0, stream>>>(grid, n_records, first, ord.flags, ord.pos, ord.order);
}

void launch_stage_moves(const AgentArrays& ag, const AgentPersist& persist, float* staged, cudaStream_t stream) {
  if (ag.n) k_stage_moves<<<blocks_for(ag.n, 256), 256, 0, stream>>>(ag, persist, staged);
}
void launch_commit_moves(const AgentArrays& ag, const AgentPersist& persist, const float* staged,
                         const uint32_t* overflow, cudaStream_t stream) {
  if (ag.n) k_commit_moves<<<blocks_for(ag.n, 256), 256, 0, stream>>>(ag, persist, staged, overflow);
}

void launch_avoidance_chunk(const DevNav& nav, const AgentArrays& ag, const lmb_options& o, float delta_time,
                            uint32_t first_agent_global, uint32_t chunk_begin, uint32_t chunk_n,
                            uint32_t n_records, const AvoidRecord* records, uint32_t n_characters,
                            const AvoidRecord* char_records, const AvoidGrid& grid, const AvoidScratch& scratch,
                            const uint32_t* max_radius_bits, float* staged_moves, const AvoidDebug& dbg,
                            const AvoidOrder& ord, cudaStream_t stream) {
  if (!chunk_n) return;
  // Few agents: small blocks, down to ONE agent per block (= per warp).  One thread walks its agent's scratch
  // serially (flood heap, cone list, loop nodes: tens of thousands of dependent instructions on a fine mesh) and
  // the agents of a warp diverge — their instruction streams run one after the other (3.4 of 8 lanes active per
  // instruction at 8 agents per warp) — so while the GPU has warp slots to spare, every agent gets a warp of its
  // own: a block holds 1 agent up to 8 agents per SM, 2 up to 16, ... 64 beyond 256 per SM.  Measured
  // (scripts/small_frame.py, 64 x 64 grid): 100 agents 1.48 -> 0.62 ms, 1 000 1.66 -> 0.75, 4 000 3.18 -> 1.98;
  // from 16 000 agents on the old and new sizes agree.
  // (measured and not kept: ONE persistent launch for all chunks, warps taking 32 agents at a time off a counter —
  // config 3 3.27 -> 3.31 ms per M agents, 16 000 agents on the 64 x 64 grid 12.4 -> 16.8 ms: the chunk tails were
  // not the cost, and the loop around the body cost 120 bytes more stack)
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
#ifndef LMB_AVOID_MIN_THREADS
#define LMB_AVOID_MIN_THREADS 1
#define LMB_AVOID_SPREAD 32
#endif
  uint32_t threads = 64;
  while (threads > LMB_AVOID_MIN_THREADS && (uint64_t)chunk_n * 8 <= (uint64_t)sm_count * threads * LMB_AVOID_SPREAD) threads >>= 1;
  k_avoidance<<<blocks_for(chunk_n, threads), threads, 0, stream>>>(nav, ag, o, delta_time, first_agent_global, chunk_begin,
                                                          chunk_n, n_records, records, n_characters, char_records,
                                                          grid, scratch, max_radius_bits, staged_moves, dbg, ord);
}

}  // namespace lmb
