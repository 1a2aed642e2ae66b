// stub: result codes of mbf_msgs/action/GetPath.action as the reference planners return them
// (dijkstra_mesh_planner.h:72-85 lists the same numbers)
#pragma once
#include <cstdint>
namespace mbf_msgs { namespace action { struct GetPath { struct Result {
  static constexpr uint32_t SUCCESS = 0, FAILURE = 50, CANCELED = 51, INVALID_START = 52, INVALID_GOAL = 53, NO_PATH_FOUND = 54,
                            PAT_EXCEEDED = 55, EMPTY_PATH = 56, TF_ERROR = 57, NOT_INITIALIZED = 58, INVALID_PLUGIN = 59,
                            INTERNAL_ERROR = 60;
}; }; } }
