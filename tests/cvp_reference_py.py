"""An INDEPENDENT restatement of CVPMeshPlanner::waveFrontPropagation, written in plain Python directly from the reference
source (cvp_mesh_planner/src/cvp_mesh_planner.cpp:369-556 waveFrontUpdate, :700-886 the loop) -- not from oracle/oracle.cpp.
tests/test_oracle_golden.py runs both on the same small meshes and requires bit-identical potentials, predecessors,
directions and cutting faces: a transcription slip in either restatement shows up as a difference.  Test infrastructure only.

Conventions shared with the oracle because the reference leaves them to lvr2 (un-vendored): faces of a vertex in ascending
face id; the priority queue pops the smallest (potential, vertex id) and `insert` of a queued vertex replaces its key."""
import heapq
import math

import numpy as np

F32 = np.float32
INF = float("inf")


def wave_front_update(dist, pred, direction, cut, w_of, face_of, v1, v2, v3):            # :369-556
    u1, u2, u3 = float(dist[v1]), float(dist[v2]), float(dist[v3])
    c = float(w_of(v1, v2)); c_sq = c * c
    b = float(w_of(v1, v3)); b_sq = b * b
    a = float(w_of(v2, v3)); a_sq = a * a
    u1_sq, u2_sq = u1 * u1, u2 * u2
    with np.errstate(all="ignore"):
        sx = float(np.float64(c_sq + u1_sq - u2_sq) / np.float64(2 * c))
        sy = -math.sqrt(max(u1_sq - sx * sx, 0.0)) if not math.isnan(u1_sq - sx * sx) else float("nan")
        p = float(np.float64(b_sq + c_sq - a_sq) / np.float64(2 * c))
        hc = math.sqrt(max(b_sq - p * p, 0.0)) if not math.isnan(b_sq - p * p) else float("nan")
        dy, dx = hc - sy, p - sx
        u3tmp_sq = dx * dx + dy * dy
        u3tmp = math.sqrt(u3tmp_sq) if u3tmp_sq >= 0 else float("nan")
        if not (u3tmp < u3):
            return False
        t0a = float(np.float64(a_sq + b_sq - c_sq) / np.float64(2 * a * b))
        t1a = float(np.float64(u3tmp_sq + b_sq - u1_sq) / np.float64(2 * u3tmp * b))
        t2a = float(np.float64(a_sq + u3tmp_sq - u2_sq) / np.float64(2 * a * u3tmp))

    def edge_fallback(src, val):
        if val < u3:
            cut[v3] = face_of(v1, v2, v3); pred[v3] = src; dist[v3] = F32(val); direction[v3] = F32(0)
            return True
        return False

    if abs(t1a) > 1:                                                                       # :416
        return edge_fallback(v1, u1 + b)
    if abs(t2a) > 1:                                                                       # :436
        return edge_fallback(v2, u2 + a)
    acos = lambda x: math.acos(x) if -1.0 <= x <= 1.0 else float("nan")                    # std::acos: NaN outside [-1, 1]
    th0, th1, th2 = acos(t0a), acos(t1a), acos(t2a)
    if th1 < th0 and th2 < th0:                                                            # :489
        cut[v3] = face_of(v1, v2, v3); dist[v3] = F32(u3tmp)
        if th1 < th2:
            pred[v3] = v1; direction[v3] = F32(th1)
        else:
            pred[v3] = v2; direction[v3] = F32(-th2)
        return True
    if th1 < th2:                                                                          # :515
        return edge_fallback(v1, u1 + b)
    return edge_fallback(v2, u2 + a)                                                       # :534


def wave_front_propagation(pos, faces, edges, edge_weights, vertex_costs, seed_face, seed_pos, robot_face=-1, invalid=None,
                           cost_limit=1.0, goal_dist_offset=0.3):
    pos = np.asarray(pos, F32); faces = np.asarray(faces); V = pos.shape[0]
    wmap = {}
    for e, (x, y) in enumerate(np.asarray(edges)):
        wmap[(min(int(x), int(y)), max(int(x), int(y)))] = edge_weights[e]
    w_of = lambda x, y: wmap[(min(x, y), max(x, y))]
    faces_of = [[] for _ in range(V)]
    for f, tri in enumerate(faces):
        for v in tri:
            faces_of[int(v)].append(f)
    fmap = {tuple(sorted(int(v) for v in tri)): f for f, tri in enumerate(faces)}
    face_of = lambda x, y, z: fmap[tuple(sorted((x, y, z)))]
    inv = np.zeros(V, bool) if invalid is None else np.asarray(invalid, bool)
    dist = np.full(V, np.inf, F32); pred = np.arange(V, dtype=np.uint32)                    # :704-708
    direction = np.zeros(V, F32); cut = np.full(V, -1, np.int64)
    fixed = np.zeros(V, bool)
    heap = []
    for v in faces[seed_face]:                                                              # :711-721
        v = int(v)
        diff = (np.asarray(seed_pos, F32) - pos[v]).astype(F32)
        d = F32(np.sqrt(F32(F32(diff[0] * diff[0] + diff[1] * diff[1]) + diff[2] * diff[2])))
        dist[v] = d; cut[v] = seed_face; fixed[v] = True
        heapq.heappush(heap, (float(d), v))
    goal = [int(v) for v in faces[robot_face]] if robot_face >= 0 else []
    goal_dist = INF
    popped = np.zeros(V, bool)
    while heap:                                                                             # :747
        d, cur = heapq.heappop(heap)
        if popped[cur] or d != float(dist[cur]):
            continue                                                                        # a replaced key
        popped[cur] = True
        fixed[cur] = True
        if float(dist[cur]) > goal_dist:                                                    # :754
            continue
        if float(vertex_costs[cur]) >= cost_limit:                                          # :757
            continue
        if inv[cur]:                                                                        # :760
            continue
        if cur in goal and goal_dist == INF and all(fixed[g] for g in goal):                # :763-771
            goal_dist = float(F32(float(dist[cur]) + goal_dist_offset))             # float = float + double
        for f in faces_of[cur]:                                                             # :776-
            a, b, c = (int(x) for x in faces[f])
            if inv[a] or inv[b] or inv[c]:
                continue
            order = None
            if fixed[a] and fixed[b] and not fixed[c]: order = (a, b, c)
            elif fixed[a] and not fixed[b] and fixed[c]: order = (c, a, b)
            elif not fixed[a] and fixed[b] and fixed[c]: order = (b, c, a)
            if order is None:
                continue
            if float(vertex_costs[order[2]]) >= cost_limit:
                continue
            if wave_front_update(dist, pred, direction, cut, w_of, face_of, *order):
                heapq.heappush(heap, (float(dist[order[2]]), order[2]))
    return dict(dist=dist, pred=pred, direction=direction, cutting_face=cut)
