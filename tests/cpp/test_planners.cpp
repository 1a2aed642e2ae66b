// GPU test of the C++ host mirror (include/meshnav_b200/planners.hpp) against the oracle.
// Mirrors how the reference's gtest (mesh_layers/test/inflation_layer_test.cpp) drives the classes directly,
// without a ROS graph.  Built and run by tests/test_gpu_cpp_host.py (g++, links libmeshnav_b200.so + liboracle.so).
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "../../include/meshnav_b200/planners.hpp"

extern "C" {   // oracle (test infrastructure)
void* orc_mesh_create(uint32_t V, uint32_t F, const float* pos, const uint32_t* faces, const uint32_t* edges, uint32_t E);
void orc_mesh_destroy(void* h);
void orc_edge_distances(void* h, float* out);
uint32_t orc_dijkstra(void* h, const float* edge_weights, const float* vertex_costs, const uint8_t* invalid, uint32_t seed_vertex,
                      int64_t robot_vertex, double cost_limit, double goal_dist_offset, int canonical_ties, float* distances,
                      uint32_t* predecessors, double* stats);
int32_t orc_cvp_backtrack(void* h, const float* vector_map, const float start[3], uint32_t start_face, const float goal[3],
                          uint32_t goal_face, double step_width, uint32_t max_points, float* path_pos, uint32_t* path_face, uint32_t* n_points);
uint32_t orc_cvp(void* h, const float* edge_weights, const float* vertex_costs, const uint8_t* invalid, uint32_t seed_face,
                 const float* seed_pos, int64_t robot_face, double cost_limit, double goal_dist_offset, int canonical_ties,
                 float* distances, uint32_t* predecessors, float* direction, int32_t* cutting_faces, double* stats);
void orc_obstacle_update(void* h, uint32_t n, const float* points, const float* tf, const float* axis, double max_obstacle_dist,
                         double robot_height, uint8_t* lethal_mask, uint8_t* changed_mask);
}

using namespace meshnav_b200;
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static uint64_t splitmix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

int main() {
  const int n = 80; const float h = 0.1f;
  std::vector<float> pos(3 * n * n); std::vector<uint32_t> faces;
  for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
    const uint64_t r = splitmix((uint64_t)(j * n + i) ^ 42ull);
    pos[3 * (j * n + i)] = i * h + ((float)((r >> 11) & 0xffff) / 65535.0f - 0.5f) * 0.04f;
    pos[3 * (j * n + i) + 1] = j * h + ((float)((r >> 31) & 0xffff) / 65535.0f - 0.5f) * 0.04f;
    pos[3 * (j * n + i) + 2] = 0.3f * std::sin(0.7f * i * h) * std::cos(0.5f * j * h);
  }
  for (int j = 0; j + 1 < n; ++j) for (int i = 0; i + 1 < n; ++i) {
    const uint32_t v00 = j * n + i, v10 = v00 + 1, v01 = v00 + n, v11 = v01 + 1;
    faces.insert(faces.end(), {v00, v10, v11, v00, v11, v01});
  }
  auto map = std::make_shared<MeshMap>(pos, faces);
  void* om = orc_mesh_create(n * n, (uint32_t)faces.size() / 3, pos.data(), faces.data(), nullptr, 0);
  const uint32_t V = map->numVertices(), E = map->numEdges();
  std::vector<float> ed(E); orc_edge_distances(om, ed.data());
  CHECK(std::memcmp(ed.data(), map->edgeDistances().data(), sizeof(float) * E) == 0);

  PoseStamped robot, goal;
  robot.position = map->vertexPosition(15 * n + 12); goal.position = map->vertexPosition(60 * n + 65);

  // ---- DijkstraMeshPlanner::makePlan ----
  DijkstraMeshPlanner dj; CHECK(dj.initialize("dijkstra", map));
  std::vector<PoseStamped> plan; double cost = 0; std::string msg;
  CHECK(dj.makePlan(robot, goal, 0.1, plan, cost, msg) == SUCCESS);
  CHECK(plan.size() > 10 && cost > 6.0 && cost < 12.0);
  std::vector<float> od(V); std::vector<uint32_t> op(V);
  const uint32_t seed_v = (uint32_t)map->getNearestVertexHandle(goal.position), robot_v = (uint32_t)map->getNearestVertexHandle(robot.position);
  CHECK(orc_dijkstra(om, ed.data(), map->vertexCosts().data(), nullptr, seed_v, robot_v, 1.0, 0.3, 1, od.data(), op.data(), nullptr) == 0);
  CHECK(std::memcmp(od.data(), dj.potential().data(), sizeof(float) * V) == 0);          // bit-identical distances
  CHECK(std::memcmp(op.data(), dj.predecessors().data(), sizeof(uint32_t) * V) == 0);    // bit-exact predecessors
  CHECK(std::fabs(cost - (double)od[robot_v]) < 0.05 * cost + 0.3);                      // path length ~ potential at the robot
  // pose orientation (dijkstra_mesh_planner.cpp:93-113: calculatePoseFromPosition with the vertex normals): unit quaternions,
  // x axis along the step, z axis within a few degrees of the up direction of the nearly flat test terrain
  for (size_t i = 0; i + 1 < plan.size(); ++i) {
    const Quaternion& q = plan[i].orientation;
    CHECK(std::fabs(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w - 1.0) < 1e-9);
    const double xx = 1 - 2 * (q.y * q.y + q.z * q.z), xy = 2 * (q.x * q.y + q.z * q.w), xz = 2 * (q.x * q.z - q.y * q.w);
    const double zz = 1 - 2 * (q.x * q.x + q.y * q.y);
    const Vector d = plan[i].direction;
    CHECK(xx * d.x + xy * d.y + xz * d.z > 0.9 && zz > 0.5);
  }
  // robot == goal vertex -> SUCCESS with an empty vertex path (dijkstra_mesh_planner.cpp:252-255)
  plan.clear(); CHECK(dj.makePlan(goal, goal, 0.1, plan, cost, msg) == SUCCESS);
  // lethal wall -> NO_PATH_FOUND (:358-362)
  for (uint32_t v = 0; v < V; ++v) if (pos[3 * v] > 3.0f && pos[3 * v] < 3.3f) map->vertexCosts()[v] = 2.0f;
  plan.clear(); CHECK(dj.makePlan(robot, goal, 0.1, plan, cost, msg) == NO_PATH_FOUND && !msg.empty());

  // ---- CVPMeshPlanner::makePlan / waveFrontPropagation ----
  CVPMeshPlanner cvp; CHECK(cvp.initialize("cvp", map));
  plan.clear(); msg.clear();
  CHECK(cvp.makePlan(robot, goal, 0.1, plan, cost, msg) == NO_PATH_FOUND);
  std::fill(map->vertexCosts().begin(), map->vertexCosts().end(), 0.0f);
  plan.clear();
  CHECK(cvp.makePlan(robot, goal, 0.1, plan, cost, msg) == SUCCESS && cost > 6.0 && cost < 8.5);
  // vector-field back-tracking (cvp:920-951): robot first, goal last, steps of step_width, smoother than the edge path
  CHECK(plan.size() > 12 && length(plan.front().position - robot.position) == 0.0f && length(plan.back().position - goal.position) == 0.0f);
  for (size_t i = 0; i + 2 < plan.size(); ++i) CHECK(length(plan[i + 1].position - plan[i].position) < 0.65f);
  for (size_t i = 0; i + 1 < plan.size(); ++i) {
    // pose orientation (util.cpp:267-298): unit quaternion whose x axis is the path direction projected into the face plane
    const Quaternion& q = plan[i].orientation;
    CHECK(std::fabs(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w - 1.0) < 1e-9);
    const double xx = 1 - 2 * (q.y * q.y + q.z * q.z), xy = 2 * (q.x * q.y + q.z * q.w), xz = 2 * (q.x * q.z - q.y * q.w);
    const Vector d = plan[i].direction;
    CHECK(xx * d.x + xy * d.y + xz * d.z > 0.9);
  }
  {
    // the same walk by the oracle on the GPU's vector map is bit-identical
    const int64_t sfo = map->getContainingFace(goal.position, 0.4f), rfo = map->getContainingFace(robot.position, 0.4f);
    std::vector<float> pp(3 * 4096); std::vector<uint32_t> pf(4096); uint32_t np = 0;
    const float s3[3] = {goal.position.x, goal.position.y, goal.position.z}, g3[3] = {robot.position.x, robot.position.y, robot.position.z};
    CHECK(orc_cvp_backtrack(om, &cvp.getVectorMap()[0].x, s3, (uint32_t)sfo, g3, (uint32_t)rfo, 0.4, 4096, pp.data(), pf.data(), &np) == 0);
    CHECK(np == plan.size());
    for (uint32_t i = 0; i < np; ++i) CHECK(std::memcmp(&pp[3 * i], &plan[i].position.x, 12) == 0);
  }
  std::vector<float> cd(V), cdir(V); std::vector<uint32_t> cp(V); std::vector<int32_t> cc(V);
  const int64_t sf = map->getContainingFace(goal.position, 0.4f), rf = map->getContainingFace(robot.position, 0.4f);
  const float sp[3] = {goal.position.x, goal.position.y, goal.position.z};
  CHECK(orc_cvp(om, ed.data(), map->vertexCosts().data(), nullptr, (uint32_t)sf, sp, rf, 1.0, 0.3, 1, cd.data(), cp.data(), cdir.data(), cc.data(), nullptr) == 0);
  double maxrel = 0;
  for (uint32_t v = 0; v < V; ++v) {
    CHECK(std::isfinite(cd[v]) == std::isfinite(cvp.potential()[v]));
    if (std::isfinite(cd[v])) maxrel = std::fmax(maxrel, std::fabs((double)cd[v] - cvp.potential()[v]) / std::fmax((double)cd[v], 1e-30));
  }
  CHECK(maxrel <= 1e-4);                                                                  // north-star tolerance
  // INVALID_GOAL: robot far off the mesh (cvp_mesh_planner.cpp:686-690)
  PoseStamped off; off.position = {100.f, 100.f, 0.f};
  plan.clear(); CHECK(cvp.makePlan(off, goal, 0.1, plan, cost, msg) == INVALID_GOAL);
  plan.clear(); CHECK(cvp.makePlan(robot, off, 0.1, plan, cost, msg) == INVALID_START);

  // ---- InflationLayer ----
  InflationLayer infl(map);
  std::vector<uint32_t> lethals; for (uint32_t v = 0; v < V; ++v) if (std::hypot(pos[3 * v] - 4.0f, pos[3 * v + 1] - 4.0f) < 0.35f) lethals.push_back(v);
  std::vector<float> risk, dist;
  CHECK(infl.waveCostInflation(lethals, risk, dist));
  size_t inflated = 0; for (uint32_t v = 0; v < V; ++v) if (std::isfinite(dist[v]) && dist[v] > 0 && dist[v] <= 0.4f) { inflated++; CHECK(risk[v] > 0 && risk[v] <= 0.99f); }
  CHECK(inflated > 20);
  for (uint32_t v : lethals) CHECK(dist[v] == 0.0f && risk[v] == 1.0f);

  // ---- dynamic obstacle cycle: onInputChanged -> layerChanged ----
  std::vector<uint32_t> update;
  std::vector<uint32_t> moved; for (uint32_t v = 0; v < V; ++v) if (std::hypot(pos[3 * v] - 5.0f, pos[3 * v + 1] - 3.0f) < 0.3f) moved.push_back(v);
  CHECK(infl.onInputChanged(moved, update));
  // update = vertices with a riskiness entry now or after the previous wave: a superset of both lethal sets
  std::vector<uint8_t> in_update(V, 0); for (uint32_t v : update) { CHECK(v < V); in_update[v] = 1; }
  for (uint32_t v : lethals) CHECK(in_update[v]);
  for (uint32_t v : moved) CHECK(in_update[v]);
  for (size_t i = 1; i < update.size(); ++i) CHECK(update[i - 1] < update[i]);            // ascending like std::set
  std::vector<Vector> field; CHECK(infl.vectorMap(field));
  size_t with_vec = 0; for (uint32_t v = 0; v < V; ++v) if (field[v].x != 0 || field[v].y != 0 || field[v].z != 0) with_vec++;
  CHECK(with_vec > 20);
  map->edge_cost_factor = 1.0;
  CHECK(map->computeEdgeWeights()); CHECK(map->syncCosts());
  const std::vector<float> w_before = map->edgeWeights();
  CHECK(map->layerChanged(update, infl.costs(), infl.defaultValue()));
  size_t changed_w = 0; for (size_t e = 0; e < w_before.size(); ++e) if (w_before[e] != map->edgeWeights()[e]) changed_w++;
  CHECK(changed_w > 20);
  for (uint32_t v : moved) CHECK(map->vertexCosts()[v] == 1.0f);
  for (uint32_t v : lethals) CHECK(map->vertexCosts()[v] == 0.0f);                         // no entry any more -> default value
  {  // the patched tables plan exactly like a fresh install of the same arrays
    auto fresh = std::make_shared<MeshMap>(pos, faces);
    fresh->vertexCosts() = map->vertexCosts(); fresh->edgeWeights() = map->edgeWeights(); CHECK(fresh->syncCosts());
    std::vector<float> d1(V), d2(V); std::vector<uint32_t> p1(V), p2(V);
    CHECK(mnb_dijkstra(map->ctx(), seed_v, -1, 2.0, 0.3, d1.data(), p1.data()) == 0);
    CHECK(mnb_dijkstra(fresh->ctx(), seed_v, -1, 2.0, 0.3, d2.data(), p2.data()) == 0);
    CHECK(std::memcmp(d1.data(), d2.data(), sizeof(float) * V) == 0 && p1 == p2);
  }
  // ---- ObstacleLayer::processPointCloud, the shared raycaster, calcNormalClearance (f3 remainder) ----
  {
    ObstacleLayer obst(map);
    obst.config_.robot_height = 0.6; obst.config_.max_obstacle_dist = 4.0;
    std::vector<float> cloud;                                     // returns 0.3 m above the surface around (3, 3), in the frame of a sensor at (3, 3, 2)
    for (uint32_t v = 0; v < V; ++v)
      if (std::hypot(pos[3 * v] - 3.0f, pos[3 * v + 1] - 3.0f) < 0.3f) { cloud.push_back(pos[3 * v] - 3.0f); cloud.push_back(pos[3 * v + 1] - 3.0f); cloud.push_back(pos[3 * v + 2] + 0.3f - 2.0f); }
    CHECK(cloud.size() >= 3 * 20);
    const std::array<float, 12> tf = {1, 0, 0, 3.0f, 0, 1, 0, 3.0f, 0, 0, 1, 2.0f};
    std::vector<uint32_t> changed;
    CHECK(obst.processPointCloud(cloud, tf, Vector{0.0f, 0.0f, -1.0f}, changed));
    CHECK(obst.lethals().size() > 20 && obst.lethals().size() < 200 && changed == obst.lethals());
    for (uint32_t v : obst.lethals()) { CHECK(std::hypot(pos[3 * v] - 3.0f, pos[3 * v + 1] - 3.0f) < 0.7f); CHECK(std::isinf(obst.costs()[v])); }
    // the oracle's loop over all faces marks the same set
    std::vector<uint8_t> mask(V, 0), chg(V, 0);
    const float ax[3] = {0, 0, -1};
    orc_obstacle_update(om, (uint32_t)(cloud.size() / 3), cloud.data(), tf.data(), ax, 4.0, 0.6, mask.data(), chg.data());
    size_t nref = 0; for (uint32_t v = 0; v < V; ++v) nref += mask[v];
    CHECK(nref == obst.lethals().size());
    for (uint32_t v : obst.lethals()) CHECK(mask[v]);
    std::vector<float> empty;                                     // an empty cloud clears the set: everything changes back
    CHECK(obst.processPointCloud(empty, tf, Vector{0.0f, 0.0f, -1.0f}, changed) && obst.lethals().empty() && changed.size() == nref);
    RayCastResult rc;
    CHECK(castRays(*map, {3.0f, 3.0f, 5.0f, 100.0f, 100.0f, 5.0f}, {0.0f, 0.0f, -1.0f}, rc));
    CHECK(rc.hit[0] == 1 && rc.hit[1] == 0 && rc.dist[0] > 3.0f && rc.dist[0] < 7.0f && std::fabs(rc.point[0] - 3.0f) < 1e-5f);
    std::vector<float> clearance;
    CHECK(calcNormalClearance(*map, clearance));
    size_t open_sky = 0; for (uint32_t v = 0; v < V; ++v) open_sky += std::isinf(clearance[v]);
    CHECK(open_sky > V * 9 / 10);                                 // a terrain without overhangs: (almost) nothing above a vertex
  }
  orc_mesh_destroy(om);
  std::printf("cpp host mirror ok: dijkstra bit-exact, cvp max rel %.2e, %zu inflated vertices\n", maxrel, inflated);
  return 0;
}
