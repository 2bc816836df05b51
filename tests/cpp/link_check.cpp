// Link check for include/landmass_b200.hpp bound to the CUDA library (`landmass_b200::Archipelago` =
// BasicArchipelago<CudaBackend>): every lmb_* entry point the header calls must resolve against
// liblandmass_b200.so.  Never run without a GPU: main() refuses unless asked, and then only creates a context.
#include <cstdio>
#include <cstring>

#include "landmass_b200.hpp"

namespace lb = landmass_b200;

int main(int argc, char** argv) {
  if (argc < 2 || std::strcmp(argv[1], "--run") != 0) {
    std::printf("abi %u\n", lmb_abi_version());
    return 0;
  }
  lmb_config cfg{};
  cfg.struct_size = sizeof(cfg);
  cfg.max_agents = 16;
  lmb_ctx* ctx = nullptr;
  if (lmb_create(&cfg, &ctx) != LMB_OK) return 1;  // no CUDA device: the library has no CPU fallback
  {
    lb::Archipelago arch(ctx, lb::ArchipelagoOptions::from_agent_radius(0.5f), 16);
    lb::NavigationMesh room;
    room.vertices = {{0.0f, 0.0f, 0.0f}, {15.0f, 0.0f, 0.0f}, {15.0f, 15.0f, 0.0f}, {0.0f, 15.0f, 0.0f}};
    room.polygons = {{0, 1, 2, 3}};
    room.polygon_type_indices = {0};
    lb::SingleIslandNavigationData nav(lb::Transform{}, room.validate());
    arch.set_navigation_data(nav.desc());
    lb::Agent a = lb::Agent::create({1.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, 0.5f, 1.0f, 2.0f);
    a.current_target = lb::Vec3{11.0f, 1.1f, 0.0f};
    const lb::AgentId id = arch.add_agent(a);
    arch.update(0.1f);
    const lb::Vec3& v = arch.get_agent(id)->get_desired_velocity();
    std::printf("desired velocity %.3f %.3f, %zu pathing result(s)\n", v[0], v[1], arch.get_pathing_results().size());
    auto p = arch.sample_point({3.0f, 3.0f, 0.0f});
    auto q = arch.sample_point({12.0f, 9.0f, 0.0f});
    if (p && q) std::printf("find_path: %zu step(s)\n", arch.find_path(*p, *q)->size());
  }
  lmb_destroy(ctx);
  return 0;
}
