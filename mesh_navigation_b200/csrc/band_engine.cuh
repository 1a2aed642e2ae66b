// Sliding-band pull wavefront engine (sm_100a).
//
// What it replaces: the heap loops of the reference
//   CVPMeshPlanner::waveFrontPropagation   cvp_mesh_planner.cpp:747-886
//   DijkstraMeshPlanner::dijkstra          dijkstra_mesh_planner.cpp:287-348
//   InflationLayer::waveCostInflation      inflation_layer.cpp:407-478
// which pop one vertex at a time from lvr2::Meap and push updates to the free
// vertex of every incident face (edge) that has all its other vertices fixed.
//
// B200 formulation (not a translation of the heap):
//   * every vertex carries one 16-byte word  {d : potential, a1, a2, a3 : pop time}.
//     The pop time is the moment the sequential algorithm would have popped the
//     vertex, as a monotonic stack of water levels (problems.cuh); for causal
//     updates it is simply (d, id)  ==  the oracle's canonical heap order.
//   * a *candidate* vertex is recomputed FROM SCRATCH ("pull") from its corner
//     records: the faces are visited in the order their later source vertex
//     pops, and a face only fires while the candidate itself has not popped yet.
//     That vertex-local replay reproduces the sequential result including the
//     non-causal "back-steps" of the CVP unfolding update; the
//     CPU simulator of exactly this rule is bit-identical to the oracle on the
//     10k / 1M meshes, and these very sources run on a CPU interpreter in the test-suite (tests/emu).
//   * candidates within a sliding band [lo, lo+delta) of potentials are
//     recomputed every round; everything whose tau lies strictly below the
//     smallest tau that changed in the round is a converged prefix and leaves
//     the list ("settled").  One barrier per round.
//   * one wavefront is owned by ONE thread-block cluster (CS CTAs, CS in
//     {1,2,4,8,16}); the barrier is the hardware cluster barrier (or
//     __syncthreads for CS==1), not a grid-wide sync, so a round costs ~1 us
//     instead of a kernel launch.  Batches run one wavefront per cluster on all
//     148 SMs; CS==0 selects a cooperative whole-grid group (used by the
//     multi-source inflation wave which has few, very wide rounds).
//   * the next round's candidate list is staged in shared memory per CTA and
//     flushed with one global atomic per CTA per round.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>

#include "wavefront_math.cuh"

namespace mnb {
namespace cg = cooperative_groups;

constexpr uint32_t INF_BITS = 0x7f800000u;
constexpr int STAGNATION_ROUNDS = 24;

// pop time of a vertex: monotonic stack of water levels a1 > a2 > a3 (0 = unused) + tie-break minor.
// `root` orders labels that share the bit-identical first level a1 = K: a vertex that pops at its own key carries its
// own id (the oracle's canonical heap order (key, id)); the members of a cascade carry the id of the cascade's root
// trigger t, so that they pop after every (K, id < t), right after (K, t) itself (a2 == 0 sorts first) and before every
// (K, id > t) -- float32 potentials collide millions of times on 10M-vertex meshes, and without `root` a cascade under a
// tied key was ordered after ALL plain labels of that key.
struct EvTime { float a1, a2, a3; uint32_t minor, root; };
__device__ __forceinline__ bool ev_less(const EvTime& x, const EvTime& y) {
  if (x.a1 != y.a1) return x.a1 < y.a1;
  if (x.root != y.root) return x.root < y.root;
  if (x.a2 != y.a2) return x.a2 < y.a2;
  if (x.a3 != y.a3) return x.a3 < y.a3;
  return x.minor < y.minor;
}
__device__ __forceinline__ bool ev_eq(const EvTime& x, const EvTime& y) {
  return __float_as_uint(x.a1) == __float_as_uint(y.a1) && __float_as_uint(x.a2) == __float_as_uint(y.a2) &&
         __float_as_uint(x.a3) == __float_as_uint(y.a3) && x.minor == y.minor && x.root == y.root;
}
__device__ __forceinline__ EvTime ev_normal(float key, uint32_t id) { EvTime t; t.a1 = key; t.a2 = 0.0f; t.a3 = 0.0f; t.minor = 2u * id; t.root = id; return t; }
// per-vertex label: one 16-byte word {d, a1, a2 | root flag, a3 | minor flag}; flagged roots / minors live in side arrays
struct Label { float d; EvTime t; };
__device__ __forceinline__ uint4 state_inf() { return make_uint4(INF_BITS, INF_BITS, 0u, 0u); }

// mark[] values
constexpr uint32_t MARK_NONE = 0, MARK_CAND = 1, MARK_FIXED = 2, MARK_CAND_ACT = 3;

struct GroupCtl {               // one per wavefront group, global memory
  unsigned int count[3];        // candidate list sizes (ring over rounds)
  unsigned int m_tau[3];        // float bits: min tau touched by a change in the round
  unsigned int lo[3];           // float bits: min pop time over surviving candidates
  // goal_dist (cvp:738,769 / dijkstra:279,296) and the cancel flag are read by EVERY thread at the
  // top of a round and decide whether the group leaves the loop, so they must not change while a
  // round is running: round r reads slot r&1, writers of round r only touch slot (r+1)&1.
  unsigned int goal_ring[2];    // float bits, monotonically decreasing (atomicMin)
  unsigned int stop_ring[2];
  unsigned int goal_bits;       // final goal_dist, published after the last round
  // pop time of the vertex that armed the cutoff (a1 bits, root, a2 bits, a3 bits, minor): a vertex beyond goal_dist that
  // popped BEFORE that moment did expand in the reference (cvp:754 tests the goal_dist of the moment of the pop), which
  // happens when the arming robot vertex is a cascade member that pops late with a small potential.  Written once by the
  // thread that arms, read from the next round on.
  unsigned int goal_time[5];
  unsigned int deep_labels;     // k_cvp_epilogue: finite labels with overflowed cascade levels
  int robot_left;               // robot-face vertices not yet settled
  unsigned int query;           // batch: query index owned by the group
  unsigned long long rounds, recomputes, settled;
  unsigned long long skipped;   // candidate-rounds that kept their label without a recompute (clean-candidate skip)
  unsigned int strict_armed;    // wavefronts that had to arm the strict back-step rule
  unsigned int watchdog;        // a wavefront hit the round watchdog (reported as non-convergence)
  unsigned long long t_work, t_flush, t_sync;   // clock cycles of thread 0 of the group: candidate loop / stage flush / barrier
  unsigned int barrier[2];      // grid_barrier state (whole-grid groups only)
  unsigned long long t_ph[8];   // phase cycles of thread 0 inside one candidate iteration (debug)
};

template <int CS>
__device__ __forceinline__ void group_sync(unsigned int* bar = nullptr) {
  if constexpr (CS == 1) {
    __syncthreads();
  } else if constexpr (CS == 0) {
    (void)bar;
    cg::this_grid().sync();   // a hand-rolled single-fence barrier measured slower (7.5k vs 5.1k cycles per round)
  } else {
    cg::this_cluster().sync();
  }
}

// per-CTA staging of appends to the next candidate list
struct Stage {
  static constexpr int CAP = 3072;
  static constexpr int SW_CAP = 1024;     // stage slots that take part in the in-round sweeps
  uint32_t buf[CAP];
  unsigned int n;
  unsigned int base;
  unsigned int m_tau;
  unsigned int lo;
};
// in-round sweeps (run_band_rounds_sub8<.., true>): per stage slot, the version of its vertex at the last evaluation
// and a copy of its label; the per-sweep dirty list.  Lives in the shared memory of the whole-grid kernel only.
struct SweepStage {
  uint32_t seen[Stage::SW_CAP];
  uint4 lab[Stage::SW_CAP];     // the slot vertex's own label (only its owner CTA writes it during a round)
  uint16_t dl[Stage::SW_CAP];
  unsigned int dn[2];           // dirty-list length, double-buffered over sweeps
};
constexpr uint32_t SEEN_NEVER = 0xffffffffu;

__device__ __forceinline__ unsigned int stage_push(Stage& st, uint32_t v, uint32_t* list_next, unsigned int* count_next) {
  const unsigned int p = atomicAdd(&st.n, 1u);
  if (p < Stage::CAP) st.buf[p] = v;
  else list_next[atomicAdd(count_next, 1u)] = v;  // overflow: straight to global
  return p;
}
// same, and registers the slot for the in-round sweeps with the version its vertex had at the last evaluation
__device__ __forceinline__ void stage_push_seen(Stage& st, SweepStage& ss, uint32_t v, uint32_t seen, const uint4& label_bits,
                                                uint32_t* list_next, unsigned int* count_next) {
  const unsigned int p = stage_push(st, v, list_next, count_next);
  if (p < Stage::SW_CAP) { ss.seen[p] = seen; ss.lab[p] = label_bits; }
}

// flush the CTA stage to the global list (all threads of the CTA call this)
__device__ __forceinline__ void stage_flush(Stage& st, uint32_t* list_next, unsigned int* count_next,
                                            unsigned int* g_m_tau, unsigned int* g_lo) {
  __syncthreads();
  const unsigned int n = st.n < Stage::CAP ? st.n : Stage::CAP;
  if (threadIdx.x == 0) {
    st.base = n ? atomicAdd(count_next, n) : 0u;
    if (st.m_tau != INF_BITS) atomicMin(g_m_tau, st.m_tau);
    if (st.lo != INF_BITS) atomicMin(g_lo, st.lo);
  }
  __syncthreads();
  const unsigned int base = st.base;
  for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) list_next[base + i] = st.buf[i];
  __syncthreads();
  if (threadIdx.x == 0) { st.n = 0; st.m_tau = INF_BITS; st.lo = INF_BITS; }
  // no barrier needed here: the group barrier that follows every flush orders the reset
  // before the next round's pushes
}

// -----------------------------------------------------------------------------
// The round loop.  Preconditions (established by the caller inside the kernel,
// followed by a group_sync): states of seeds written, mark[seed] = MARK_FIXED,
// list0 holds the first candidates (mark = MARK_CAND), ctl->count[0] = |list0|,
// ctl->count[1] = ctl->count[2] = 0, ctl->m_tau[0] = ctl->lo[0] = INF,
// ctl->m_tau[1] = ctl->lo[1] = INF, ctl->m_tau[2] = 0, ctl->lo[2] = bits(seed_min).
//
// Problem P provides
//   bool  recompute(c, band_end, goal, d_old, tau_old, &d_new, &tau_new)  -> true if changed
//         (writes state/aux itself)
//   void  activate(c, push)   calls push(x) for every vertex x that shares a face/edge with c
//   bool  eligible(x)         may x ever receive a label
// -----------------------------------------------------------------------------
template <int CS, class P>
__device__ void run_band_rounds(P& prob, GroupCtl* ctl, uint32_t* list0, uint32_t* list1, uint32_t* mark,
                                Stage& st, const float delta, const uint32_t gthreads, const uint32_t gtid,
                                const int has_robot, const uint32_t r0, const uint32_t r1, const uint32_t r2,
                                const double goal_dist_offset, const volatile int* cancel_flag,
                                const float band_end_init, const uint32_t max_rounds) {
  float band_end_prev = band_end_init;  // > every seed potential: seeds are available from round 0
  unsigned long long my_recomputes = 0, my_settled = 0, my_skipped = 0;
  float lo_best = -1.0f; int stagnant = 0;       // best (largest) earliest-unsettled pop time seen so far
  prob.strict = 0;
  uint32_t r = 0;
  for (;; ++r) {
    const uint32_t slot = r % 3, prev = (r + 2) % 3, next = (r + 1) % 3;
    const unsigned int n = __ldcg(&ctl->count[slot]);
    const float m_prev = __uint_as_float(__ldcg(&ctl->m_tau[prev]));
    const float lo_prev = __uint_as_float(__ldcg(&ctl->lo[prev]));
    const unsigned int goal_b = __ldcg(&ctl->goal_ring[r & 1]);
    const float goal = __uint_as_float(goal_b);
    if constexpr (P::HAS_GOAL_TIME) if (has_robot && goal_b != INF_BITS) {
      prob.goal_t.a1 = __uint_as_float(__ldcg(&ctl->goal_time[0])); prob.goal_t.root = __ldcg(&ctl->goal_time[1]);
      prob.goal_t.a2 = __uint_as_float(__ldcg(&ctl->goal_time[2])); prob.goal_t.a3 = __uint_as_float(__ldcg(&ctl->goal_time[3]));
      prob.goal_t.minor = __ldcg(&ctl->goal_time[4]);
    }
    const unsigned int stop = __ldcg(&ctl->stop_ring[r & 1]);
    if (n == 0 || stop || r > max_rounds) break;   // r is group-uniform: the watchdog cannot deadlock the barrier
    if (r > 0 && __float_as_uint(m_prev) == INF_BITS &&
        (__float_as_uint(lo_prev) == INF_BITS || lo_prev > goal)) break;
    // stagnation watch (all values are group-uniform): labels keep changing but the earliest unsettled pop
    // time does not move -> a dependency cycle between a trigger and its back-step child; arm the strict rule
    if (r > 0 && __float_as_uint(m_prev) != INF_BITS && !(lo_prev > lo_best)) { if (++stagnant >= P::STAGNATION) prob.strict = 1; }
    else { stagnant = 0; if (lo_prev > lo_best) lo_best = lo_prev; }
    float band_end = lo_prev + delta;
    if (!(band_end > band_end_prev)) band_end = band_end_prev;
    uint32_t* list_r = (r & 1) ? list1 : list0;
    uint32_t* list_n = (r & 1) ? list0 : list1;
    if (gtid == 0) {
      ctl->count[(r + 2) % 3] = 0;         // list r+2's counter (free during this round)
      ctl->m_tau[next] = INF_BITS;
      ctl->lo[next] = INF_BITS;
      atomicMin(&ctl->goal_ring[(r + 1) & 1], goal_b);                 // carry the cutoff into the next round
      ctl->stop_ring[(r + 1) & 1] = (stop || (cancel_flag && (r & 31) == 0 && *cancel_flag)) ? 1u : 0u;
    }
#ifdef MNB_EMU_ACTIVE   // round trace on the CPU interpreter of the kernels (tests/emu, MNB_EMU_TRACE=1): list length and progress per round
    if (gtid == 0 && getenv("MNB_EMU_TRACE")) fprintf(stderr, "[round %u] list %u m_prev %g lo_prev %g strict %d\n", r, n, m_prev, lo_prev, prob.strict);
#endif
    bool skip_ok = false;
    if constexpr (P::CAN_SKIP) skip_ok = prob.skip_clean && !has_robot && (P::SKIP_IN_STRICT || !prob.strict) && __float_as_uint(delta) == INF_BITS;
    float my_mtau = __uint_as_float(INF_BITS), my_lo = __uint_as_float(INF_BITS);
    for (unsigned int i = gtid; i < n; i += gthreads) {
      const uint32_t c = __ldcg(&list_r[i]);
      const Label old = prob.load_label(c);
      const float d = old.d, tau = old.t.a1;
      if (tau < m_prev && tau < band_end_prev) {
        // converged prefix: the sequential algorithm has popped c with exactly this label
        mark[c] = MARK_FIXED;
        my_settled++;
        if (has_robot && (c == r0 || c == r1 || c == r2)) {
          if (atomicSub(&ctl->robot_left, 1) == 1) {
            // c is not necessarily the last of the three in event order: take the latest (tau,minor)
            float bd = d; EvTime bt = old.t;
            const uint32_t rv[3] = {r0, r1, r2};
            for (int k = 0; k < 3; ++k) {
              const Label so = prob.load_label(rv[k]);
              if (ev_less(bt, so.t)) { bt = so.t; bd = so.d; }
            }
            ctl->goal_time[0] = __float_as_uint(bt.a1); ctl->goal_time[1] = bt.root; ctl->goal_time[2] = __float_as_uint(bt.a2);
            ctl->goal_time[3] = __float_as_uint(bt.a3); ctl->goal_time[4] = bt.minor;
            atomicMin(&ctl->goal_ring[(r + 1) & 1], __float_as_uint((float)((double)bd + goal_dist_offset)));
          }
        }
        continue;
      }
      float nd, ntau;
      if constexpr (P::CAN_SKIP) if (skip_ok) {
        // clean-candidate skip (see run_band_rounds_sub8): no source of c was re-labelled in or after the round of c's last
        // evaluation -> same inputs, same label.  (delta = inf here: no source is ever excluded by the band end.)
        const uint32_t le = __ldcg(&prob.last_eval[c]), dr = __ldcg(&prob.dirty_round[c]);
        if (le != 0u && dr < le) {
          my_lo = fminf(my_lo, tau); my_skipped++;
          stage_push(st, c, list_n, &ctl->count[next]);
          continue;
        }
      }
      my_recomputes++;
      if constexpr (P::CAN_SKIP) prob.deferred_flag = false;
      const bool changed = prob.recompute(c, band_end, goal, r, old, nd, ntau);
      if (changed) my_mtau = fminf(my_mtau, fminf(tau, ntau));
      if constexpr (P::CAN_SKIP) if (skip_ok) {
        __stcg(&prob.last_eval[c], prob.deferred_flag ? 0u : r + 1u);
        if (changed) prob.activate(c, [&](uint32_t x) { __stcg(&prob.dirty_round[x], r + 1u); });
      }
      my_lo = fminf(my_lo, ntau);      // smallest pop time still in flight: the band follows it
      stage_push(st, c, list_n, &ctl->count[next]);
      // a vertex that holds a finite label pulls its neighbours into the candidate set (once)
      if (__float_as_uint(nd) != INF_BITS && __ldcg(&mark[c]) == MARK_CAND) {
        mark[c] = MARK_CAND_ACT;
        prob.activate(c, [&](uint32_t x) {
          if (__ldcg(&mark[x]) == MARK_NONE && prob.eligible(x) && atomicCAS(&mark[x], MARK_NONE, MARK_CAND) == MARK_NONE)
            stage_push(st, x, list_n, &ctl->count[next]);
        });
      }
    }
    {
      my_mtau = fminf(my_mtau, prob.deferred_m);      // deferred back-steps are pending changes
      prob.deferred_m = __uint_as_float(INF_BITS);
      const unsigned int wm = __reduce_min_sync(0xffffffffu, __float_as_uint(my_mtau));
      const unsigned int wl = __reduce_min_sync(0xffffffffu, __float_as_uint(my_lo));
      if ((threadIdx.x & 31) == 0) {
        if (wm != INF_BITS) atomicMin(&st.m_tau, wm);
        if (wl != INF_BITS) atomicMin(&st.lo, wl);
      }
    }
    stage_flush(st, list_n, &ctl->count[next], &ctl->m_tau[slot], &ctl->lo[slot]);
    band_end_prev = band_end;
    group_sync<CS>(ctl->barrier);
  }
  // statistics
  atomicAdd(&ctl->recomputes, my_recomputes);
  atomicAdd(&ctl->settled, my_settled);
  if (my_skipped) atomicAdd(&ctl->skipped, my_skipped);
  if (gtid == 0) {
    ctl->rounds += r;
    if (prob.strict) ctl->strict_armed += 1;
    if (r > max_rounds) ctl->watchdog = 1;
    ctl->goal_bits = min(ctl->goal_ring[0], ctl->goal_ring[1]);
  }
}

// -----------------------------------------------------------------------------
// Same round loop with 8 lanes per candidate (problem provides replay_sub8 / activate via its ELL row).
// Used by the whole-grid single-plan kernel where per-round LATENCY is what matters.
// -----------------------------------------------------------------------------
template <int CS, bool SW, class P>
__device__ void run_band_rounds_sub8(P& prob, GroupCtl* ctl, uint32_t* list0, uint32_t* list1, uint32_t* mark,
                                     Stage& st, const float delta, const uint32_t gthreads, const uint32_t gtid,
                                     const int has_robot, const uint32_t r0, const uint32_t r1, const uint32_t r2,
                                     const double goal_dist_offset, const volatile int* cancel_flag,
                                     const float band_end_init, const uint32_t max_rounds, const int n_sweeps_arg,
                                     SweepStage* ss, const uint32_t n_vertices) {
  const int n_sweeps = SW ? n_sweeps_arg : 0;
  bool rescanned = false;
  float band_end_prev = band_end_init;
  unsigned long long my_recomputes = 0, my_settled = 0, my_skipped = 0;
  float lo_best = -1.0f; int stagnant = 0;       // best (largest) earliest-unsettled pop time seen so far
  prob.strict = 0;
  const uint32_t j = threadIdx.x & 7;
  const uint32_t nblk = gthreads / blockDim.x, blk = gtid / blockDim.x;   // CTAs of this group / my CTA's rank in it
  uint32_t r = 0;
  for (;; ++r) {
    const uint32_t slot = r % 3, prev = (r + 2) % 3, next = (r + 1) % 3;
    const unsigned int n = __ldcg(&ctl->count[slot]);
    const float m_prev = __uint_as_float(__ldcg(&ctl->m_tau[prev]));
    const float lo_prev = __uint_as_float(__ldcg(&ctl->lo[prev]));
    const unsigned int goal_b = __ldcg(&ctl->goal_ring[r & 1]);
    const float goal = __uint_as_float(goal_b);
    if constexpr (P::HAS_GOAL_TIME) if (has_robot && goal_b != INF_BITS) {
      prob.goal_t.a1 = __uint_as_float(__ldcg(&ctl->goal_time[0])); prob.goal_t.root = __ldcg(&ctl->goal_time[1]);
      prob.goal_t.a2 = __uint_as_float(__ldcg(&ctl->goal_time[2])); prob.goal_t.a3 = __uint_as_float(__ldcg(&ctl->goal_time[3]));
      prob.goal_t.minor = __ldcg(&ctl->goal_time[4]);
    }
    const unsigned int stop = __ldcg(&ctl->stop_ring[r & 1]);
    // the round in which the goal cutoff first becomes visible must run even if nothing else is left to do: it puts the
    // vertices that settled beyond the cutoff back into the list (see below)
    const bool need_rescan = has_robot && goal_b != INF_BITS && !rescanned;
    if (stop || r > max_rounds) break;             // r is group-uniform: the watchdog cannot deadlock the barrier
    if (!need_rescan && n == 0) break;
    if (!need_rescan && r > 0 && __float_as_uint(m_prev) == INF_BITS &&
        (__float_as_uint(lo_prev) == INF_BITS || lo_prev > goal)) break;
    // stagnation watch (all values are group-uniform): labels keep changing but the earliest unsettled pop
    // time does not move -> a dependency cycle between a trigger and its back-step child; arm the strict rule
    if (r > 0 && __float_as_uint(m_prev) != INF_BITS && !(lo_prev > lo_best)) { if (++stagnant >= STAGNATION_ROUNDS) prob.strict = 1; }
    else { stagnant = 0; if (lo_prev > lo_best) lo_best = lo_prev; }
    // the sweeps work on the first SW_CAP stage slots of a CTA: on very long fronts (tens of millions of vertices) the
    // band is narrowed so that a CTA's share of the list still fits (group-uniform, exactness does not depend on it)
#ifdef MNB_EMU_ACTIVE   // round trace on the CPU interpreter of the kernels (tests/emu, MNB_EMU_TRACE=1)
    if (gtid == 0 && getenv("MNB_EMU_TRACE")) fprintf(stderr, "[round %u] list %u m_prev %g lo_prev %g strict %d\n", r, n, m_prev, lo_prev, prob.strict);
#endif
    float delta_r = delta;
    if constexpr (SW) {
      const float fit = (float)nblk * (0.8f * (float)Stage::SW_CAP);
      if ((float)n > fit) delta_r = fmaxf(0.25f * delta, delta * fit / (float)n);
    }
    float band_end = lo_prev + delta_r;
    if (!(band_end > band_end_prev)) band_end = band_end_prev;
    uint32_t* list_r = (r & 1) ? list1 : list0;
    uint32_t* list_n = (r & 1) ? list0 : list1;
    if (gtid == 0) {
      ctl->count[(r + 2) % 3] = 0;
      ctl->m_tau[next] = INF_BITS;
      ctl->lo[next] = INF_BITS;
      atomicMin(&ctl->goal_ring[(r + 1) & 1], goal_b);                 // carry the cutoff into the next round
      ctl->stop_ring[(r + 1) & 1] = (stop || (cancel_flag && (r & 31) == 0 && *cancel_flag)) ? 1u : 0u;
    }
    // Clean-candidate skip (group-uniform switch): a label is a pure function of the source labels, the band end (only
    // through sources beyond it), the goal cutoff and -- in strict mode -- the round number.  Plans with a goal cutoff
    // and strict rounds recompute everything; otherwise a candidate none of whose sources was re-labelled in or after the
    // round of its own last evaluation, and whose relevant sources beyond the band end are still beyond it, keeps its
    // label without being recomputed.
    bool skip_ok = false;
    if constexpr (P::CAN_SKIP) skip_ok = prob.skip_clean && !has_robot && !prob.strict;
    // Goal cutoff (cvp:754 / dijkstra:299) with a band wider than goal_dist_offset: labels computed before the cutoff
    // is known may rest on sources that turn out to lie beyond it.  Once it is known, (1) vertices beyond it never
    // settle any more -- they are recomputed under the cutoff until nothing changes -- and (2) the ones that had already
    // settled are put back into the candidate list by one sweep over the vertex array (their labels only depend on
    // vertices inside the cutoff, which are unaffected, so a single recompute repairs them).
    const float settle_cap = (has_robot && goal_b != INF_BITS) ? nextafterf(goal, __uint_as_float(INF_BITS)) : __uint_as_float(INF_BITS);
    bool requeued = false;
    if (need_rescan) {
      rescanned = true;
      for (uint32_t v = gtid; v < n_vertices; v += gthreads) {
        if (__ldcg(&mark[v]) != MARK_FIXED) continue;
        const float tv = __uint_as_float(__ldcg(&prob.state[v]).y);
        if (tv > goal) {
          mark[v] = MARK_CAND_ACT;
          if constexpr (SW) stage_push_seen(st, *ss, v, SEEN_NEVER, __ldcg(&prob.state[v]), list_n, &ctl->count[next]);
          else stage_push(st, v, list_n, &ctl->count[next]);
          requeued = true;
        }
      }
    }
    const long long tp0 = clock64();
    float my_mtau = requeued ? goal : __uint_as_float(INF_BITS), my_lo = requeued ? goal : __uint_as_float(INF_BITS);   // a pending change
    // One evaluation of candidate c by its 8-lane group (every lane of the WARP calls this; idle groups pass
    // has = false so that the sub-warp shuffles can use compile-time full masks).  `fresh` = main pass (c comes from
    // the round's list and is pushed to the stage); otherwise c already sits in the stage (in-round sweep).
    auto evaluate = [&](bool has, const uint32_t c, const Label& old, const int4& ix, const float4& w, const uint32_t mk,
                        const uint32_t v0, const bool fresh, const uint32_t slot) {
      const float d = old.d, tau = old.t.a1;
      float nd; EvTime nt; int deg; uint32_t mk1 = MARK_FIXED, mk2 = MARK_FIXED; float excl = 0.0f;
      prob.replay_sub8(c, j, has, ix, w, band_end, goal, r, mark, nd, nt, deg, mk1, mk2, excl);
      const bool changed = has && (__float_as_uint(nd) != __float_as_uint(d) || !ev_eq(nt, old.t));
      if constexpr (P::CAN_SKIP) if (skip_ok) {
        // stamps of the clean-candidate skip: c was evaluated in this round; its face neighbours have a source that was
        // re-labelled in this round (plain stores: every writer of a round stores the same value, rounds are barrier-ordered)
        if (has && j == 0) { __stcg(&prob.last_eval[c], r + 1u); __stcg(&prob.excl_min[c], __float_as_uint(excl)); }
        if (changed) {
          if (ix.x != -1 && deg <= 8) { __stcg(&prob.dirty_round[ix.x], r + 1u); if constexpr (P::TWO_SOURCES) __stcg(&prob.dirty_round[ix.y], r + 1u); }
          if (j == 0 && deg > 8) prob.activate(c, [&](uint32_t x) { __stcg(&prob.dirty_round[x], r + 1u); });
        }
      }
      if (has && j == 0) {
        my_recomputes++;
        if (changed) {
          prob.store_label(c, nd, nt, __float_as_uint(d) != INF_BITS, r);
          my_mtau = fminf(my_mtau, fminf(tau, nt.a1));
        }
        my_lo = fminf(my_lo, nt.a1);
        if constexpr (SW) {
          if (fresh) stage_push_seen(st, *ss, c, v0, prob.pack_label(c, nd, nt), list_n, &ctl->count[next]);
          else if (changed) ss->lab[slot] = prob.pack_label(c, nd, nt);
        } else {
          stage_push(st, c, list_n, &ctl->count[next]);
        }
      }
      if constexpr (SW) if (n_sweeps > 0) {
        // tell the vertices that read c's label (its face neighbours) that it changed by bumping their version.
        // No fence / release-acquire pairing on purpose: the sweeps are opportunistic -- a neighbour that polls the
        // bump but still reads the old label merely misses one in-round update; the next round's main pass
        // recomputes every candidate and the change is accounted in m_tau, so exactness never depends on it.
        __syncwarp();
        if (changed) {
          if (ix.x != -1 && deg <= 8) { atomicAdd(&prob.ver[ix.x], 1u); if constexpr (P::TWO_SOURCES) atomicAdd(&prob.ver[ix.y], 1u); }
          if (j == 0 && deg > 8) prob.activate(c, [&](uint32_t x) { atomicAdd(&prob.ver[x], 1u); });
        }
      }
      if (has && __float_as_uint(nd) != INF_BITS && mk == MARK_CAND) {
        // every lane pulls the two source vertices of its own corner into the candidate set; their marks
        // were fetched together with their labels, so only genuinely new vertices cost an atomic
        if (ix.x != -1) {
          if (!prob.prefetch_marks) { mk1 = __ldcg(&mark[ix.x]); if (P::TWO_SOURCES) mk2 = __ldcg(&mark[ix.y]); }
          if (mk1 == MARK_NONE && prob.eligible((uint32_t)ix.x) && atomicCAS(&mark[ix.x], MARK_NONE, MARK_CAND) == MARK_NONE)
            { if constexpr (SW) stage_push_seen(st, *ss, (uint32_t)ix.x, SEEN_NEVER, state_inf(), list_n, &ctl->count[next]); else stage_push(st, (uint32_t)ix.x, list_n, &ctl->count[next]); }
          if (P::TWO_SOURCES && mk2 == MARK_NONE && prob.eligible((uint32_t)ix.y) && atomicCAS(&mark[ix.y], MARK_NONE, MARK_CAND) == MARK_NONE)
            { if constexpr (SW) stage_push_seen(st, *ss, (uint32_t)ix.y, SEEN_NEVER, state_inf(), list_n, &ctl->count[next]); else stage_push(st, (uint32_t)ix.y, list_n, &ctl->count[next]); }
        }
        if (j == 0 && deg > 8)
          prob.activate(c, [&](uint32_t x) {
            if (__ldcg(&mark[x]) == MARK_NONE && prob.eligible(x) && atomicCAS(&mark[x], MARK_NONE, MARK_CAND) == MARK_NONE)
              { if constexpr (SW) stage_push_seen(st, *ss, x, SEEN_NEVER, state_inf(), list_n, &ctl->count[next]); else stage_push(st, x, list_n, &ctl->count[next]); }
          });
        if (j == 0) mark[c] = MARK_CAND_ACT;
      }
    };
    // warp-uniform trip count: every lane of the warp runs every iteration (idle groups carry has = false) so
    // that the sub-warp shuffles below can use compile-time full masks (no MATCH.ANY / WARPSYNC sequences)
    // the candidates are dealt to the CTAs of the group in equal contiguous chunks and packed into the
    // lowest warps of each CTA: a round's cost is the instruction stream of its busiest SM, so an even
    // spread (instead of filling the first CTAs completely) is what shortens the round
    // ... and dealt round-robin (candidate i -> CTA i mod nblk): the list is ordered by flush time, i.e. by how busy the
    // producing CTA was, so contiguous chunks would hand all the "hot" candidates (the ones whose labels are still
    // moving) to a few CTAs while the rest idle at the barrier
    const unsigned int cnt = n > blk ? (n - blk + nblk - 1) / nblk : 0u;
    for (unsigned int qb = (threadIdx.x >> 5) * 4u; qb < cnt; qb += (blockDim.x >> 3)) {
      const unsigned int q = qb + ((threadIdx.x & 31) >> 3);
      bool has = q < cnt;
      uint32_t c = 0;
      if (has) c = __ldcg(&list_r[(size_t)q * nblk + blk]);
      // issue the independent loads of the candidate together: its label, its ELL row, its mark (and version)
      Label old = prob.load_label(has ? c : 0u);
      int4 ix = prob.load_row_idx(has ? c : 0u, j);
      float4 w = prob.load_row_w(has ? c : 0u, j);
      const uint32_t mk = __ldcg(&mark[has ? c : 0u]);
      uint32_t v0 = 0;
      if constexpr (SW) if (n_sweeps > 0) v0 = __ldcg(&prob.ver[has ? c : 0u]);
      const float d = old.d, tau = old.t.a1;
      if (has && tau < m_prev && tau < band_end_prev && tau < settle_cap) {
        if (j == 0) {
          mark[c] = MARK_FIXED;
          my_settled++;
          if (has_robot && (c == r0 || c == r1 || c == r2)) {
            if (atomicSub(&ctl->robot_left, 1) == 1) {
              float bd = d; EvTime bt = old.t;
              const uint32_t rv[3] = {r0, r1, r2};
              for (int k = 0; k < 3; ++k) {
                const Label so = prob.load_label(rv[k]);
                if (ev_less(bt, so.t)) { bt = so.t; bd = so.d; }
              }
              ctl->goal_time[0] = __float_as_uint(bt.a1); ctl->goal_time[1] = bt.root; ctl->goal_time[2] = __float_as_uint(bt.a2);
            ctl->goal_time[3] = __float_as_uint(bt.a3); ctl->goal_time[4] = bt.minor;
            atomicMin(&ctl->goal_ring[(r + 1) & 1], __float_as_uint((float)((double)bd + goal_dist_offset)));
            }
          }
        }
        has = false;
      }
      if constexpr (P::CAN_SKIP) if (skip_ok && has) {
        const uint32_t le = __ldcg(&prob.last_eval[c]), dr = __ldcg(&prob.dirty_round[c]);
        const float em = __uint_as_float(__ldcg(&prob.excl_min[c]));
        if (le != 0u && dr < le && !(band_end > em)) {
          // clean: survives with its label as it is (the 8 lanes of the group agree: same loads)
          if (j == 0) {
            my_lo = fminf(my_lo, tau);
            my_skipped++;
            if constexpr (SW) stage_push_seen(st, *ss, c, v0, prob.pack_label(c, old.d, old.t), list_n, &ctl->count[next]);
            else stage_push(st, c, list_n, &ctl->count[next]);
          }
          has = false;
        }
      }
      evaluate(has, c, old, ix, w, mk, v0, true, 0u);
    }
    // ---- in-round sweeps: the CTA keeps relaxing the candidates it staged (survivors + newly activated) whose
    // inputs changed since their last evaluation, so a dependency chain advances several hops per barrier ----
    const long long tps = clock64();
    if (gtid == 0) { ctl->t_ph[0] += (unsigned long long)(tps - tp0); ctl->t_ph[2] += cnt; }
    if constexpr (SW) for (int sw = 0; sw < n_sweeps; ++sw) {
      const long long tq0 = clock64();
      __syncthreads();                       // stage pushes / label-cache writes of the previous phase are visible
      const unsigned int ns = min(st.n, (unsigned int)Stage::SW_CAP);
      unsigned int* dcur = &ss->dn[sw & 1];
      if (threadIdx.x == 0) ss->dn[(sw + 1) & 1] = 0;
      for (unsigned int i = threadIdx.x; i < ns; i += blockDim.x) {
        const uint32_t c = st.buf[i];
        const uint32_t v0 = __ldcg(&prob.ver[c]);
        if (v0 != ss->seen[i]) { ss->seen[i] = v0; ss->dl[atomicAdd(dcur, 1u)] = (uint16_t)i; }
      }
      __syncthreads();
      const unsigned int dn = *dcur;
      const long long tq1 = clock64();
      if (gtid == 0) { ctl->t_ph[3] += dn; ctl->t_ph[4] += ns; ctl->t_ph[5] += (unsigned long long)(tq1 - tq0); }
      for (unsigned int ib = (threadIdx.x >> 5) * 4u; ib < dn; ib += (blockDim.x >> 3)) {
        const unsigned int i = ib + ((threadIdx.x & 31) >> 3);
        const bool has = i < dn;
        const uint32_t slot = has ? ss->dl[i] : 0u;
        const uint32_t c = has ? st.buf[slot] : 0u;
        // own label from the CTA's cache, ELL row through L1: the source labels are the only L2 trip of the chain
        const Label old = has ? prob.unpack_label(c, ss->lab[slot]) : prob.unpack_label(0u, state_inf());
        int4 ix = prob.load_row_idx(c, j);
        float4 w = prob.load_row_w(c, j);
        const uint32_t mk = __ldcg(&mark[c]);
        evaluate(has, c, old, ix, w, mk, 0u, false, slot);
      }
      if (gtid == 0) ctl->t_ph[6] += (unsigned long long)(clock64() - tq1);
    }
    {
      my_mtau = fminf(my_mtau, prob.deferred_m);      // deferred back-steps are pending changes
      prob.deferred_m = __uint_as_float(INF_BITS);
      const unsigned int wm = __reduce_min_sync(0xffffffffu, __float_as_uint(my_mtau));
      const unsigned int wl = __reduce_min_sync(0xffffffffu, __float_as_uint(my_lo));
      if ((threadIdx.x & 31) == 0) {
        if (wm != INF_BITS) atomicMin(&st.m_tau, wm);
        if (wl != INF_BITS) atomicMin(&st.lo, wl);
      }
    }
    __syncthreads();
    const long long tp1 = clock64();
    stage_flush(st, list_n, &ctl->count[next], &ctl->m_tau[slot], &ctl->lo[slot]);
    const long long tp2 = clock64();
    band_end_prev = band_end;
    group_sync<CS>(ctl->barrier);
    if (gtid == 0) { const long long tp3 = clock64(); ctl->t_work += (unsigned long long)(tp1 - tp0); ctl->t_flush += (unsigned long long)(tp2 - tp1); ctl->t_sync += (unsigned long long)(tp3 - tp2); }
  }
  atomicAdd(&ctl->recomputes, my_recomputes);
  atomicAdd(&ctl->settled, my_settled);
  if (my_skipped) atomicAdd(&ctl->skipped, my_skipped);
  if (gtid == 0) {
    ctl->rounds += r;
    if (prob.strict) ctl->strict_armed += 1;
    if (r > max_rounds) ctl->watchdog = 1;
    ctl->goal_bits = min(ctl->goal_ring[0], ctl->goal_ring[1]);
  }
}

}  // namespace mnb
