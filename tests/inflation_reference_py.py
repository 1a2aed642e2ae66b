"""An INDEPENDENT restatement of InflationLayer::waveCostInflation in plain Python / numpy float32, written directly from the
reference source (mesh_layers/src/inflation_layer.cpp:181-225 computeUpdateSethianMethod, :227-313 waveFrontUpdate,
:315-339 fading, :341-491 the loop) -- not from oracle/oracle.cpp.  tests/test_oracle_golden.py requires the distances and
the riskiness values of both restatements to agree bit for bit.  The repulsive vectors are left out: they depend on
lvr2::BaseVector::normalized(), which is not vendored.  Test infrastructure only.

Conventions shared with the oracle where the reference leaves the order to pmp / lvr2 (un-vendored): the neighbours of the
popped vertex in ascending edge id, the two faces of an edge in ascending face id; the queue pops the smallest
(distance, vertex id) and `insert` of a queued vertex replaces its key."""
import heapq
import math

import numpy as np

F = np.float32
INF32 = F(np.inf)
EPSILON = F(1e-9)                                                                          # inflation_layer.h:47


def sethian(d1, d2, a, b, dot, Fs):                                                        # :181-225, all float
    with np.errstate(all="ignore"):
        t = INF32
        r_cos = dot
        r_sin = F(np.sqrt(F(F(1) - dot * dot)))
        u = d2 - d1
        f2 = a * a + b * b - F(2) * a * b * r_cos
        f1 = b * u * (a * r_cos - b)
        f0 = b * b * (u * u - Fs * Fs * a * a * r_sin)
        delta = f1 * f1 - f0 * f2
        if delta >= 0:
            if abs(f2) > EPSILON:
                t = (-f1 - F(np.sqrt(delta))) / f2
                if t < u or b * (t - u) / t < a * r_cos or a / r_cos < b * (t - u) / F(2):
                    t = (-f1 + F(np.sqrt(delta))) / f2
                else:
                    t = -f0 / f1 if f1 != 0 else -INF32
        else:
            t = -INF32
        if u < t and a * r_cos < b * (t - u) / t and b * (t - u) / t < a / r_cos:
            return t + d1
        return min(b * Fs + d1, a * Fs + d2)                                               # std::min(x, y): y < x ? y : x


def fading(distance, inscribed_radius, inflation_radius, lethal_value, inscribed_value, cost_scaling_factor):   # :315-339
    d = float(distance)                                                                    # config members are double
    if d > inflation_radius:
        return F(0)
    if d > inscribed_radius:
        factor = F(math.exp(-1.0 * cost_scaling_factor * (d - inscribed_radius)))
        return F(inscribed_value * float(factor))
    if d > 0:
        return F(inscribed_value)
    return F(lethal_value)


def wave_cost_inflation(pos, faces, edges, edge_distances, lethals, invalid=None, inscribed_radius=0.25, inflation_radius=0.4,
                        lethal_value=1.0, inscribed_value=0.99, cost_scaling_factor=1.0):
    faces = np.asarray(faces); V = np.asarray(pos).shape[0]
    eid = {}
    for e, (x, y) in enumerate(np.asarray(edges)):
        eid[(min(int(x), int(y)), max(int(x), int(y)))] = e
    w = lambda x, y: F(edge_distances[eid[(min(x, y), max(x, y))]])
    edge_faces = {}
    for f, tri in enumerate(faces):
        t = [int(v) for v in tri]
        for i in range(3):
            k = (min(t[i], t[(i + 1) % 3]), max(t[i], t[(i + 1) % 3]))
            edge_faces.setdefault(eid[k], []).append(f)
    nbrs = [[] for _ in range(V)]
    for (x, y), e in eid.items():
        nbrs[x].append((e, y)); nbrs[y].append((e, x))
    for l in nbrs:
        l.sort()
    inv = np.zeros(V, bool) if invalid is None else np.asarray(invalid, bool)
    max_distance = F(inflation_radius)                                                     # `const float& max_distance` <- double config
    dist = np.full(V, np.inf, F)
    fixed = np.zeros(V, bool); queued = np.zeros(V, bool)
    key = np.full(V, np.inf)                                                               # the key a queued vertex was inserted with: a distance lowered by an
                                                                                           # update that does not re-insert (:310) leaves the queue entry alone
    heap = []
    for v in sorted(set(int(x) for x in lethals)):                                         # :397-402
        dist[v] = F(0); fixed[v] = True; queued[v] = True; key[v] = 0.0
        heapq.heappush(heap, (0.0, v))

    def update(v1, v2, v3):                                                                # :227-313 (without the vector map)
        u1, u2, u3 = dist[v1], dist[v2], dist[v3]
        if u3 == 0:
            return False
        c, b, a = w(v1, v2), w(v1, v3), w(v2, v3)
        with np.errstate(all="ignore"):
            dot = (a * a + b * b - c * c) / (F(2) * a * b)
            u3tmp = sethian(u1, u2, a, b, dot, F(1))
        if not np.isfinite(u3tmp):
            return False
        if u3tmp < u3:
            dist[v3] = F(u3tmp)
            return bool(u1 <= max_distance and u2 <= max_distance)
        return False

    while heap:                                                                            # :407
        d, cur = heapq.heappop(heap)
        if not queued[cur] or d != key[cur]:
            continue
        queued[cur] = False
        if inv[cur]:                                                                       # :417 popped, never fixed
            continue
        fixed[cur] = True
        for e, nh in nbrs[cur]:                                                            # :423
            for fh in sorted(edge_faces[e]):                                               # :427 both faces of the edge
                a, b, c = (int(x) for x in faces[fh])
                order = None
                if fixed[a] and fixed[b] and not fixed[c]: order = (a, b, c)
                elif fixed[a] and not fixed[b] and fixed[c]: order = (c, a, b)
                elif not fixed[a] and fixed[b] and fixed[c]: order = (b, c, a)
                if order is None:
                    continue
                if update(*order):
                    queued[order[2]] = True; key[order[2]] = float(dist[order[2]])
                    heapq.heappush(heap, (key[order[2]], order[2]))
    cost = np.full(V, np.nan, F)                                                           # NaN = no entry (:484-490)
    for v in range(V):
        if not np.isinf(dist[v]):
            cost[v] = fading(dist[v], inscribed_radius, inflation_radius, lethal_value, inscribed_value, cost_scaling_factor)
    return dict(dist=dist, cost=cost)
