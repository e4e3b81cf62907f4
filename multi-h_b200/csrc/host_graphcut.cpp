// ============================================================================
// host_graphcut.cpp — host-side combinatorial steps that consume the GPU-built
// data costs: 4-D neighbourhood and alpha-expansion.
//
// Own implementation (nothing is taken from the reference's vendored GCO /
// maxflow sources, whose licence forbids redistribution).  It plays the role of
// GCoptimizationGeneralGraph in MultiH::LabelingStep (MultiH/MultiH/MultiH.cpp:
// 520-543) and of FlannBasedMatcher::radiusMatch in ClusterMergingAndLabeling
// (MultiH.cpp:231-253).
//
// Equivalence with the reference's optimiser (checked in tests against the
// reference's GCO compiled in place): GCO's expansion(iter, max) sweeps the
// labels 0..L-1 in order, accepts a move iff it strictly lowers the energy and
// stops when a full cycle leaves the energy unchanged (GCoptimization.cpp:
// 1032-1049, 1212-1289).  Inside a move it labels a site alpha unless the site
// belongs to the sink tree of the BK search, i.e. unless it can still reach the
// sink in the residual graph (energy.h get_var -> what_segment(default SOURCE)).
// That set is a property of the binary energy alone (it is the minimiser with
// the most alpha labels), so any exact max-flow followed by a reverse
// reachability pass from the sink reproduces GCO's moves bit for bit.  Totals
// are int64 (the reference's EnergyType is a 32-bit int, GCoptimization.h:165-170).
// ============================================================================
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/multih_b200.h"
#ifdef MH_GC_PROFILE
#include <x86intrin.h>
#endif
#if defined(MH_GC_PROFILE) || defined(MH_GC_CHECK_TALLY)
#include <cstdio>
#endif

namespace mh {

#ifdef MH_GC_PROFILE   // make EXTRA=-DMH_GC_PROFILE: cycle counters per phase of the expansion, printed at exit (single-threaded runs)
static unsigned long long g_prof[16];
static unsigned long long g_prof_n[16];
struct GcProfPrinter {
  ~GcProfPrinter() {
    static const char* names[] = {"candidate scan", "bounds init", "reduction", "network build", "max-flow", "tally", "memo checks",
                                  "setup", "total", "commit", "", "", "", "", "", ""};
    std::fprintf(stderr, "[gc profile] moves with candidates %llu: mean C0 %.1f, decided at once %.1f, network sites %.1f (non-empty networks %llu), sure switchers %.1f\n",
                 g_prof_n[1], g_prof[10] / (double)g_prof_n[1], g_prof[11] / (double)g_prof_n[1], g_prof[12] / (double)g_prof_n[1], g_prof_n[12],
                 g_prof[13] / (double)g_prof_n[1]);
    std::fprintf(stderr, "[gc profile] accepted moves %llu, sites switched %llu\n", g_prof_n[14], g_prof[14]);
    for (int i = 0; i < 10; ++i)
      std::fprintf(stderr, "[gc profile] %-16s %10.1f Mcycles  (%llu)  %5.1f %%\n", names[i], g_prof[i] / 1e6, g_prof_n[i],
                   100.0 * g_prof[i] / (double)(g_prof[8] ? g_prof[8] : 1));
  }
};
static GcProfPrinter g_prof_printer;
#define GC_TICK(var) const unsigned long long var = __rdtsc()
#define GC_ACC(slot, from) do { g_prof[slot] += __rdtsc() - (from); ++g_prof_n[slot]; } while (0)
#else
#define GC_TICK(var) do { } while (0)
#define GC_ACC(slot, from) do { } while (0)
#endif

// ---------------------------------------------------------------------------
// Dinic max-flow on a static arc array (paired arcs e, e^1).
// ---------------------------------------------------------------------------
class MaxFlow {
 public:
  void reset(int n_nodes, size_t arc_hint) {
    n_ = n_nodes + 2; s_ = n_nodes; t_ = n_nodes + 1;
    flow_ = 0;
    head_.assign(n_, -1);
    arc_.clear();
    arc_.reserve(arc_hint);
  }
  void add_edge(int u, int v, int64_t c_uv, int64_t c_vu) {
    arc_.push_back({c_uv, v, head_[u]}); head_[u] = (int)arc_.size() - 1;
    arc_.push_back({c_vu, u, head_[v]}); head_[v] = (int)arc_.size() - 1;
  }
  // terminal capacities: only the difference needs an arc
  void add_terminal(int x, int64_t from_source, int64_t to_sink) {
    if (from_source > to_sink) add_edge(s_, x, from_source - to_sink, 0);
    else if (to_sink > from_source) add_edge(x, t_, to_sink - from_source, 0);
  }
  void solve() {
    std::vector<int>&level = level_, &it = it_, &queue = queue_, &path = path_;   // arcs on the current DFS path
    level.resize(n_); it.resize(n_); queue.resize(n_);
    Arc* const arc = arc_.data();
    for (;;) {
      std::fill(level.begin(), level.end(), -1);
      int qh = 0, qt = 0;
      queue[qt++] = s_; level[s_] = 0;
      while (qh < qt && level[t_] < 0) {   // nodes beyond the sink's level cannot lie on a shortest augmenting path
        const int u = queue[qh++];
        for (int e = head_[u]; e >= 0; e = arc[e].next)
          if (arc[e].cap > 0 && level[arc[e].to] < 0) { level[arc[e].to] = level[u] + 1; queue[qt++] = arc[e].to; }
      }
      if (level[t_] < 0) break;
      for (int i = 0; i < n_; ++i) it[i] = head_[i];
      // iterative blocking flow
      path.clear();
      int u = s_;
      for (;;) {
        if (u == t_) {
          int64_t f = INT64_MAX;
          for (int e : path) f = std::min(f, arc[e].cap);
          for (int e : path) { arc[e].cap -= f; arc[e ^ 1].cap += f; }
          flow_ += f;
          // restart from the first saturated arc
          size_t k = 0;
          while (k < path.size() && arc[path[k]].cap > 0) ++k;
          path.resize(k);
          u = path.empty() ? s_ : arc[path.back()].to;
          continue;
        }
        int& e = it[u];
        while (e >= 0 && !(arc[e].cap > 0 && level[arc[e].to] == level[u] + 1)) e = arc[e].next;
        if (e >= 0) { path.push_back(e); u = arc[e].to; }
        else {
          level[u] = -1;  // dead end
          if (path.empty()) break;
          const int back = path.back();
          path.pop_back();
          u = arc[back ^ 1].to;
        }
      }
    }
    // sink side = nodes that can still reach t in the residual graph
    reach_t_.assign(n_, 0);
    std::vector<int>& st = queue;
    int top = 0;
    st[top++] = t_; reach_t_[t_] = 1;
    while (top) {
      const int v = st[--top];
      for (int e = head_[v]; e >= 0; e = arc[e].next) {
        const int u = arc[e].to;  // arc v->u is e, arc u->v is e^1
        if (!reach_t_[u] && arc[e ^ 1].cap > 0) { reach_t_[u] = 1; st[top++] = u; }
      }
    }
  }
  bool sink_side(int x) const { return reach_t_[x] != 0; }
  int64_t flow_value() const { return flow_; }   // value of the maximum flow (over the terminal-capacity differences)

 private:
  struct Arc { int64_t cap; int to, next; };
  int64_t flow_ = 0;
  int n_ = 0, s_ = 0, t_ = 0;
  std::vector<int> head_, level_, it_, queue_, path_;
  std::vector<Arc> arc_;
  std::vector<char> reach_t_;
};

// ---------------------------------------------------------------------------
// Push-relabel (FIFO, gap heuristic, periodic global relabelling), first phase only: a maximum PREFLOW is enough, because
// all the caller needs is the set of nodes that can still reach the sink in the residual graph, and returning the excess
// that is stranded on the source side to the source only changes arcs between nodes outside that set.  Same interface as
// MaxFlow.  For the flow networks of large moves (thousands of sites, ~30 arcs each, pair capacities far below the terminal
// capacities) Dinic needs ~15 phases and ~10^4 augmentations; used from PR_MIN_NODES network nodes on.
// ---------------------------------------------------------------------------
class MaxFlowPR {
 public:
  void reset(int n_nodes, size_t arc_hint) {
    n_ = n_nodes;
    flow_ = 0;
    first_.assign((size_t)n_, -1);
    excess_.assign((size_t)n_, 0);
    tcap_.assign((size_t)n_, 0);
    arc_.clear();
    arc_.reserve(arc_hint);
  }
  void add_edge(int u, int v, int64_t c_uv, int64_t c_vu) {
    arc_.push_back({c_uv, v, first_[u]}); first_[u] = (int)arc_.size() - 1;
    arc_.push_back({c_vu, u, first_[v]}); first_[v] = (int)arc_.size() - 1;
  }
  // terminal capacities: only the difference matters; the source arc is saturated at once (it becomes the node's excess)
  void add_terminal(int x, int64_t from_source, int64_t to_sink) {
    if (from_source > to_sink) excess_[x] = from_source - to_sink;
    else tcap_[x] = to_sink - from_source;
  }
  bool sink_side(int x) const { return d_[x] < dead_; }
  int64_t flow_value() const { return flow_; }   // what reached the sink: the value of a maximum preflow = of a maximum flow

  void solve() {
    dead_ = n_ + 1;
    d_.assign((size_t)n_, dead_);
    cur_.assign((size_t)n_, -1);
    inq_.assign((size_t)n_, 0);
    count_.assign((size_t)n_ + 2, 0);
    queue_.clear(); qhead_ = 0;
    global_relabel();
    Arc* const arc = arc_.data();
    long relabels = 0;
    while (qhead_ < queue_.size()) {
      const int i = queue_[qhead_++];
      inq_[i] = 0;
      if (qhead_ > (size_t)n_ && qhead_ * 2 > queue_.size()) {   // keep the FIFO compact
        queue_.erase(queue_.begin(), queue_.begin() + (long)qhead_);
        qhead_ = 0;
      }
      // discharge
      while (excess_[i] > 0 && d_[i] < dead_) {
        if (tcap_[i] > 0 && d_[i] == 1) {
          const int64_t f = std::min(excess_[i], tcap_[i]);
          tcap_[i] -= f; excess_[i] -= f;
          flow_ += f;
          continue;
        }
        int e = cur_[i];
        for (; e >= 0; e = arc[e].next) {
          const int j = arc[e].to;
          if (arc[e].cap > 0 && d_[i] == d_[j] + 1) {
            const int64_t f = std::min(excess_[i], arc[e].cap);
            arc[e].cap -= f; arc[e ^ 1].cap += f;
            excess_[i] -= f; excess_[j] += f;
            if (!inq_[j]) { inq_[j] = 1; queue_.push_back(j); }
            if (excess_[i] == 0) break;
          }
        }
        cur_[i] = e;
        if (excess_[i] == 0) break;
        // relabel: one above the lowest neighbour reachable on a residual arc (the sink counts as height 0)
        int nd = dead_;
        if (tcap_[i] > 0) nd = 1;
        for (int a = first_[i]; a >= 0; a = arc[a].next)
          if (arc[a].cap > 0 && d_[arc[a].to] + 1 < nd) nd = d_[arc[a].to] + 1;
        const int old = d_[i];
        --count_[old];
        d_[i] = nd;
        ++count_[nd];
        cur_[i] = first_[i];
        if (count_[old] == 0 && old < dead_) {   // gap: nothing left at height `old`, so everything above it is cut off from the sink
          for (int k = 0; k < n_; ++k)
            if (d_[k] > old && d_[k] < dead_) { --count_[d_[k]]; d_[k] = dead_; ++count_[dead_]; }
        }
        if (++relabels >= (long)n_) {
          relabels = 0;
          global_relabel();
          break;   // the queue was rebuilt
        }
      }
    }
    global_relabel();   // exact distances: d < dead  <=>  the node can still reach the sink
  }

 private:
  struct Arc { int64_t cap; int to, next; };
  // exact heights by a reverse breadth-first search from the sink; rebuilds the queue of active nodes
  void global_relabel() {
    const Arc* arc = arc_.data();
    std::fill(d_.begin(), d_.end(), dead_);
    std::fill(count_.begin(), count_.end(), 0);
    bfs_.clear();
    for (int i = 0; i < n_; ++i)
      if (tcap_[i] > 0) { d_[i] = 1; bfs_.push_back(i); }
    for (size_t h = 0; h < bfs_.size(); ++h) {
      const int v = bfs_[h];
      for (int e = first_[v]; e >= 0; e = arc[e].next) {
        const int u = arc[e].to;   // arc u -> v is e ^ 1
        if (d_[u] == dead_ && arc[e ^ 1].cap > 0) { d_[u] = d_[v] + 1; bfs_.push_back(u); }
      }
    }
    queue_.clear(); qhead_ = 0;
    for (int i = 0; i < n_; ++i) {
      ++count_[d_[i]];
      cur_[i] = first_[i];
      inq_[i] = 0;
      if (excess_[i] > 0 && d_[i] < dead_) { inq_[i] = 1; queue_.push_back(i); }
    }
  }
  int64_t flow_ = 0;
  int n_ = 0, dead_ = 1;
  std::vector<int> first_, d_, cur_, count_, queue_, bfs_;
  std::vector<char> inq_;
  std::vector<int64_t> excess_, tcap_;
  std::vector<Arc> arc_;
  size_t qhead_ = 0;
};
constexpr int PR_MIN_NODES = 128;   // measured: equal on ~70-node networks, 10 % faster at ~200-600, 2x at ~6600

// symmetric weighted adjacency from the directed CSR the caller hands in: each directed entry (i -> j) stands for one
// setNeighbors(i, j) call of the reference (MultiH.cpp:532-540), which inserts the pair into BOTH lists.
struct SymGraph {
  std::vector<int64_t> off;
  std::vector<int32_t> nbr;
  std::vector<int32_t> w;
};

static void symmetrise(int N, const int64_t* offsets, const int32_t* adj, SymGraph& g) {
  std::vector<int64_t> deg(N + 1, 0);
  for (int i = 0; i < N; ++i)
    for (int64_t e = offsets[i]; e < offsets[i + 1]; ++e) {
      const int j = adj[e];
      if (j == i || j < 0 || j >= N) continue;
      ++deg[i]; ++deg[j];
    }
  std::vector<int64_t> start(N + 1, 0);
  for (int i = 0; i < N; ++i) start[i + 1] = start[i] + deg[i];
  std::vector<int32_t> tmp(start[N]);
  std::vector<int64_t> cur(start.begin(), start.end() - 1);
  for (int i = 0; i < N; ++i)
    for (int64_t e = offsets[i]; e < offsets[i + 1]; ++e) {
      const int j = adj[e];
      if (j == i || j < 0 || j >= N) continue;
      tmp[cur[i]++] = j; tmp[cur[j]++] = i;
    }
  g.off.assign(N + 1, 0);
  g.nbr.clear(); g.w.clear();
  g.nbr.reserve(tmp.size()); g.w.reserve(tmp.size());
  for (int i = 0; i < N; ++i) {
    std::sort(tmp.begin() + start[i], tmp.begin() + start[i + 1]);
    for (int64_t k = start[i]; k < start[i + 1];) {
      int64_t m = k;
      while (m < start[i + 1] && tmp[m] == tmp[k]) ++m;
      g.nbr.push_back(tmp[k]); g.w.push_back((int32_t)(m - k));
      k = m;
    }
    g.off[i + 1] = (int64_t)g.nbr.size();
  }
}

static int64_t total_energy(const int32_t* cost, int N, int L, int potts, const SymGraph& g, const int32_t* lab) {
  int64_t e = 0;
  for (int i = 0; i < N; ++i) {
    e += cost[(size_t)i * L + lab[i]];
    for (int64_t k = g.off[i]; k < g.off[i + 1]; ++k) {
      const int j = g.nbr[k];
      if (j < i && lab[j] != lab[i]) e += (int64_t)potts * g.w[k];
    }
  }
  return e;
}

// ---------------------------------------------------------------------------
// One expansion move, evaluated on a read-only labelling: which sites would switch to alpha, and by how much the energy
// would drop.  Everything it reads — labels of the initial candidates C0 (sites whose data cost alone does not forbid
// alpha) and of their neighbours — is recorded in `touched`, so a caller that evaluated the move on an older labelling
// can tell whether the result still holds.
// ---------------------------------------------------------------------------
struct MoveWS {   // per-thread scratch
  std::vector<int> var, cand, work, fsw;
  std::vector<char> sw_mark, state;
  std::vector<int32_t> trial;
  std::vector<int64_t> src, snk, Wcur, Ucur, Dcur;
  MaxFlow mf;
  MaxFlowPR mfpr;
  void prepare(int N) { if ((int)var.size() != N) { var.assign(N, -1); trial.assign(N, 0); sw_mark.assign(N, 0); } }
};
struct MoveResult {
  int alpha = -1;
  int64_t delta = 0;               // energy change if applied (< 0, or 0 with an empty switch list)
  std::vector<int> sw;             // sites that switch to alpha
  std::vector<uint64_t> touched;   // bitset over sites: the initial candidates C0 (what the move read = C0 and its neighbours)
  bool any_c0 = false;
};
struct MoveProblem {
  const int32_t* cost; int N, L, potts;
  const int32_t* costT;            // label-major copy [L][N]: a move's scan reads one contiguous row
  const int32_t* cur;              // cost of every site at its current label (kept up to date by the sweep)
  const SymGraph* g;
  const int64_t* Wall;             // SUM_j w_ij
  int solver;                      // 0 = Dinic below PR_MIN_NODES network nodes, push-relabel above; 1 / 2 force one (MH_GC_SOLVER)
};

static void eval_move(const MoveProblem& P, const int32_t* lab, int alpha, MoveWS& ws, MoveResult& r, bool want_touched) {
  const int N = P.N, L = P.L, potts = P.potts;
  const int32_t* cost = P.cost;
  const SymGraph& g = *P.g;
  std::vector<int>&var = ws.var, &cand = ws.cand, &work = ws.work, &fsw = ws.fsw;
  std::vector<char>& swm = ws.sw_mark;
  std::vector<int64_t>&src = ws.src, &snk = ws.snk, &Wcur = ws.Wcur, &Ucur = ws.Ucur, &Dcur = ws.Dcur;
  r.alpha = alpha; r.delta = 0; r.sw.clear(); r.any_c0 = false;
  if (want_touched) r.touched.assign(((size_t)N + 63) / 64, 0);
  // Exact two-sided reduction of the move's binary problem.  Let S be the set of sites that may still switch to alpha and
  // D_i the rise of site i's data cost when it does.
  //  (keep)   Switching i lowers its pairwise terms by at most
  //               W_i = SUM_{j : alpha already, or j in S} w_ij  -  SUM_{j keeps for sure, l_j = l_i} w_ij
  //           (a sure keeper with another label costs w either way).  If D_i > W_i, dropping i from any switch set inside S
  //           strictly lowers the energy: no minimiser switches i; i leaves S.
  //  (switch) Switching i raises its pairwise terms by at most
  //               U_i = SUM_{j not alpha, l_j = l_i} w_ij  -  SUM_{j : alpha already} w_ij
  //           If D_i + U_i < 0, adding i to any switch set strictly lowers the energy: every minimiser switches i; i counts
  //           as "alpha already" for its neighbours from then on.
  // Both rules only ever tighten the other sites' bounds, so the work list below reaches the unique fixed point.  What is
  // left for the flow network are the genuinely ambiguous sites — often none: cold sweeps are mostly sure switchers (the
  // support of hypothesis alpha leaving the outlier label), later sweeps mostly sure keepers.  Every minimiser contains the
  // sure switchers and lies inside S, so the reduced problem has the same minimisers and the same maximal one (the labelling
  // GCO returns).
  cand.clear();
  GC_TICK(t_scan);
  {
    const int32_t* ca = P.costT + (size_t)alpha * N;
    const int32_t* cc = P.cur;
    const int64_t* wa = P.Wall;
    for (int i = 0; i < N; ++i)
      if (((int64_t)ca[i] - cc[i] <= wa[i]) & (lab[i] != alpha)) cand.push_back(i);
    for (size_t a = 0; a < cand.size(); ++a) var[cand[a]] = (int)a;
  }
  GC_ACC(0, t_scan);
  if (cand.empty()) return;
  r.any_c0 = true;
  GC_TICK(t_init);
  const size_t n0 = cand.size();
  Wcur.resize(n0); Ucur.resize(n0); Dcur.resize(n0);
  std::vector<char>& st = ws.state;   // per initial candidate: 0 = undecided, 1 = queued/decided keep, 2 = queued/decided switch
  st.assign(n0, 0);
  work.clear();
  for (size_t a = 0; a < n0; ++a) {
    const int i = cand[a];
    if (want_touched) r.touched[(size_t)i >> 6] |= 1ull << (i & 63);
    const int li = lab[i];
    int32_t wsum = 0, usum = 0;   // branch-free: the tests are data-dependent coin flips
    for (int64_t k = g.off[i]; k < g.off[i + 1]; ++k) {
      const int j = g.nbr[k];
      const int lj = lab[j];
      const int isA = lj == alpha, may = isA | (var[j] >= 0), same = lj == li;
      wsum += g.w[k] * (may - ((may ^ 1) & same));
      usum += g.w[k] * (((isA ^ 1) & same) - isA);
    }
    Wcur[a] = (int64_t)potts * wsum;
    Ucur[a] = (int64_t)potts * usum;
    Dcur[a] = (int64_t)cost[(size_t)i * L + alpha] - cost[(size_t)i * L + li];
  }
  for (size_t a = 0; a < n0; ++a) {
    if (Dcur[a] > Wcur[a]) { st[a] = 1; work.push_back((int)a); }
    else if (Dcur[a] + Ucur[a] < 0) { st[a] = 2; work.push_back((int)a); }
  }
  GC_ACC(1, t_init);
#ifdef MH_GC_PROFILE
  g_prof[10] += n0; g_prof[11] += work.size();
#endif
  GC_TICK(t_red);
  fsw.clear();
  for (size_t q = 0; q < work.size(); ++q) {
    const int a = work[q], i = cand[a];
    const bool keeps = st[a] == 1;
    if (!keeps) { swm[i] = 1; fsw.push_back(i); }
    for (int64_t k = g.off[i]; k < g.off[i + 1]; ++k) {
      const int j = g.nbr[k];
      const int b2 = var[j];
      if (b2 < 0 || st[b2] != 0) continue;
      const int64_t d = (int64_t)potts * g.w[k] * (lab[j] == lab[i] ? 2 : 1);
      if (keeps) {   // i can no longer turn alpha (-w), and now certainly disagrees with an alpha-j if l_j = l_i (-w)
        Wcur[b2] -= d;
        if (Dcur[b2] > Wcur[b2]) { st[b2] = 1; work.push_back(b2); }
      } else {       // i is alpha from now on: its pair term can only fall when j switches (-w), and no longer rises (-w if l_j = l_i)
        Ucur[b2] -= d;
        if (Dcur[b2] + Ucur[b2] < 0) { st[b2] = 2; work.push_back(b2); }
      }
    }
  }
  // survivors = the flow network's nodes; decided sites leave `var`
  {
    size_t keep = 0;
    for (size_t a = 0; a < n0; ++a) {
      if (st[a] == 0) cand[keep++] = cand[a];
      else var[cand[a]] = -1;
    }
    cand.resize(keep);
  }
  GC_ACC(2, t_red);
#ifdef MH_GC_PROFILE
  g_prof[12] += cand.size(); g_prof[13] += fsw.size(); g_prof_n[12] += !cand.empty();
#endif
  if (cand.empty() && fsw.empty()) return;
  for (size_t a = 0; a < cand.size(); ++a) var[cand[a]] = (int)a;
  std::vector<int32_t>& trial = ws.trial;
  bool any = !fsw.empty();
  // energy of the network's optimum minus the energy with every network site keeping its label (both with the sure switchers
  // switched): the network represents the energy exactly up to a constant, so this is (flow + SUM min(src, snk)) - SUM src
  int64_t delta_net = 0;
  if (!cand.empty()) {
    size_t arcs = 0;
    for (int i : cand) arcs += (size_t)(g.off[i + 1] - g.off[i]);
    auto run_network = [&](auto& mf) {
      GC_TICK(t_build);
      mf.reset((int)cand.size(), 2 * arcs + 2 * cand.size());
      src.assign(cand.size(), 0);
      snk.assign(cand.size(), 0);
      for (size_t a = 0; a < cand.size(); ++a) {
        const int i = cand[a], li = lab[i];
        // x = 0 (source side) takes alpha and pays E0 on the arc to the sink; x = 1 keeps its label.  Per neighbour j:
        //   alpha already (or a sure switcher): pay w iff i keeps its label                      -> source arc += w
        //   j keeps for sure (l_j != alpha)    : alpha pays w, keeping pays w [l_i != l_j]        -> sink += w, source += w [..]
        //   j in the network, j < i            : E00 = 0, E01 = E10 = w, E11 = w [l_i != l_j]: pay E11 on i's source arc, the
        //                                        remaining table [0, w; w - E11, 0] becomes the arc pair
        // (sums kept branch-free; only the arc insertion branches)
        int32_t s_w = 0, t_w = 0;
        for (int64_t k = g.off[i]; k < g.off[i + 1]; ++k) {
          const int j = g.nbr[k], lj = lab[j], vj = var[j], w = g.w[k];
          const int isA = (lj == alpha) | swm[j], inS = vj >= 0, diff = li != lj;
          const int fixed = (isA | inS) ^ 1, both = inS & (j < i);
          s_w += w * (isA | (diff & (fixed | both)));
          t_w += w * fixed;
          if (both) mf.add_edge((int)a, vj, (int64_t)potts * w, (int64_t)potts * w * (diff ^ 1));
        }
        snk[a] = cost[(size_t)i * L + alpha] + (int64_t)potts * t_w;
        src[a] = cost[(size_t)i * L + li] + (int64_t)potts * s_w;
      }
      int64_t sum_src = 0, sum_min = 0;
      for (size_t a = 0; a < cand.size(); ++a) {
        mf.add_terminal((int)a, src[a], snk[a]);
        sum_src += src[a];
        sum_min += std::min(src[a], snk[a]);
      }
      GC_ACC(3, t_build);
      GC_TICK(t_flow);
      mf.solve();
      GC_ACC(4, t_flow);
      delta_net = mf.flow_value() + sum_min - sum_src;
      for (size_t a = 0; a < cand.size(); ++a) {
        const bool sw = !mf.sink_side((int)a);
        trial[cand[a]] = sw ? alpha : lab[cand[a]];
        any |= sw;
      }
    };
    const int solver = P.solver;   // 0 = by size, 1 = Dinic, 2 = push-relabel
    if (solver == 2 || (solver == 0 && (int)cand.size() >= PR_MIN_NODES)) run_network(ws.mfpr);
    else run_network(ws.mf);
  }
  GC_TICK(t_tally);
  if (any) {
    // Energy change of the move = delta_net (above) + the change from switching the sure switchers alone, network sites keeping:
    // every term that involves a sure switcher, before and after (each pair among them once).  GCO accepts a move only if it
    // strictly lowers the energy.
    int64_t before = 0, after = 0;
    for (int i : fsw) {
      const int li = lab[i];
      before += cost[(size_t)i * L + li];
      after += cost[(size_t)i * L + alpha];
      int32_t cb = 0, ca = 0;
      for (int64_t k = g.off[i]; k < g.off[i + 1]; ++k) {
        const int j = g.nbr[k], lj = lab[j];
        const int sj = swm[j];
        const int once = (sj ^ 1) | (j < i);
        cb += g.w[k] * (once & (li != lj));
        ca += g.w[k] * (once & ((sj | (lj == alpha)) ^ 1));   // i is alpha afterwards: the pair costs w unless j is (or becomes) alpha
      }
      before += (int64_t)potts * cb;
      after += (int64_t)potts * ca;
    }
    const int64_t delta = delta_net + (after - before);
#ifdef MH_GC_CHECK_TALLY   // the direct tally over every network site and sure switcher (the previous implementation)
    {
      int64_t b2 = 0, a2 = 0;
      auto tally = [&](int i, int ti) {
        const int li = lab[i];
        b2 += cost[(size_t)i * L + li];
        a2 += cost[(size_t)i * L + ti];
        int32_t cb = 0, ca = 0;
        for (int64_t k = g.off[i]; k < g.off[i + 1]; ++k) {
          const int j = g.nbr[k], vj = var[j], lj = lab[j];
          const int inU = (vj >= 0) | swm[j];
          const int tj = swm[j] ? alpha : (vj >= 0 ? trial[j] : lj);
          const int once = (inU ^ 1) | (j < i);
          cb += g.w[k] * (once & (li != lj));
          ca += g.w[k] * (once & (ti != tj));
        }
        b2 += (int64_t)potts * cb;
        a2 += (int64_t)potts * ca;
      };
      for (int i : cand) tally(i, trial[i]);
      for (int i : fsw) tally(i, alpha);
      if (a2 - b2 != delta) { std::fprintf(stderr, "[gc check] delta %lld vs tally %lld\n", (long long)delta, (long long)(a2 - b2)); std::abort(); }
    }
#endif
    if (delta < 0) {
      r.delta = delta;
      for (int i : fsw) r.sw.push_back(i);
      for (int i : cand)
        if (trial[i] == alpha) r.sw.push_back(i);
    }
  }
  for (int i : cand) var[i] = -1;
  for (int i : fsw) swm[i] = 0;
  GC_ACC(5, t_tally);
}

// ---------------------------------------------------------------------------
// Worker threads for speculative move evaluation (below).  One pool per process, created on first use and re-created
// after a fork; a caller that finds it busy (another host thread is inside mh_alpha_expansion) simply runs sequentially.
// Workers sleep between sessions and spin inside one (a move takes ~50 us: a futex round trip per move would eat it).
// ---------------------------------------------------------------------------
class MovePool {
 public:
  static std::mutex& global_mutex() { static std::mutex gm; return gm; }
  static MovePool*& instance() { static MovePool* pool = nullptr; return pool; }
  // joins the workers; called when the library is unloaded (dlclose / process exit), see mh_movepool_shutdown below
  static void shutdown() {
    std::lock_guard<std::mutex> lk(global_mutex());
    MovePool*& pool = instance();
    if (!pool || pool->owner_ != getpid()) return;   // (a forked child never owned the threads)
    { std::lock_guard<std::mutex> l2(pool->m_); pool->quit_ = true; pool->session_ = true; }
    pool->cv_.notify_all();
    pool->epoch_.fetch_add(2);
    for (std::thread& t : pool->th_) t.join();
    delete pool;
    pool = nullptr;
  }
  static MovePool* acquire() {   // nullptr: no threads wanted / pool busy
    std::mutex& gm = global_mutex();
    MovePool*& pool = instance();
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(gm);
    int want = (int)std::thread::hardware_concurrency();
    if (const char* e = std::getenv("MH_GC_THREADS")) want = std::atoi(e);
    want = std::min(want, 8);   // measured on the 16-core B200 host: 8 -> 16 threads gains 2 %
    if (want < 2) return nullptr;
    if (!pool || owner != getpid()) { pool = new MovePool(want - 1); owner = getpid(); pool->owner_ = owner; }   // (a forked child leaks the parent's)
    if (pool->busy_) return nullptr;
    pool->busy_ = true;
    pool->begin_session();
    return pool;
  }
  void release() {
    end_session();
    busy_ = false;   // only the owner of busy_ writes it back; acquire() reads it under the global mutex
  }
  int threads() const { return (int)th_.size() + 1; }
  // fn(job, thread) for job in [0, n): the caller takes part as thread 0.  The epoch is a sequence lock: odd while the
  // caller rewrites the job state, which it does only after every worker has left work() (active_ == 0); a worker registers
  // in active_ first and re-checks the epoch, so it can never run on half-written state or on a finished run's counters.
  template <typename F> void run(int n, F&& fn) {
    epoch_.fetch_add(1);                        // odd: no new worker may enter
    while (active_.load() != 0) cpu_relax();    // stragglers of the previous run leave
    fn_ = [&](int j, int t) { fn(j, t); };
    njobs_ = n;
    done_.store(0);
    next_.store(0);
    epoch_.fetch_add(1);                        // even: go
    work(0);
    while (done_.load() < n) cpu_relax();
  }

 private:
  explicit MovePool(int workers) {
    for (int i = 0; i < workers; ++i) th_.emplace_back([this, i] { loop(i + 1); });
  }
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  void work(int t) {
    for (;;) {
      const int j = next_.fetch_add(1, std::memory_order_relaxed);
      if (j >= njobs_) break;
      fn_(j, t);
      done_.fetch_add(1, std::memory_order_release);
    }
  }
  void begin_session() {
    { std::lock_guard<std::mutex> lk(m_); session_ = true; }
    cv_.notify_all();
  }
  void end_session() {
    { std::lock_guard<std::mutex> lk(m_); session_ = false; }
    epoch_.fetch_add(2);   // (stays even) spinners re-check the session flag
  }
  void loop(int t) {
    uint64_t seen = epoch_.load();
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return session_; });
        if (quit_) return;
      }
      for (;;) {   // inside a session: spin on the epoch
        if (quit_) return;
        const uint64_t e = epoch_.load();
        if (e != seen && (e & 1) == 0) {
          active_.fetch_add(1);
          if (epoch_.load() != e) { active_.fetch_sub(1); continue; }   // the caller moved on: look again
          seen = e;
          bool in_session;
          { std::lock_guard<std::mutex> lk(m_); in_session = session_; }
          if (in_session) work(t);
          active_.fetch_sub(1);
          if (!in_session) break;
        } else {
          cpu_relax();
        }
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_;
  bool session_ = false;
  std::atomic<bool> quit_{false};
  pid_t owner_ = 0;
  std::atomic<bool> busy_{false};
  std::function<void(int, int)> fn_;
  int njobs_ = 0;
  std::atomic<int> next_{0}, done_{0}, active_{0};
  std::atomic<uint64_t> epoch_{0};
};

// the worker threads must not outlive the library's code: join them when it is unloaded
__attribute__((destructor)) static void mh_movepool_shutdown() { MovePool::shutdown(); }

// The neighbourhood is the same for every labelling step of a pair (MultiH.cpp:231-253 builds it once), so the symmetrised
// graph is kept between calls: one entry per thread, keyed on the content of the caller's CSR.
struct PreparedGraph {
  uint64_t key = 0;
  int N = -1;
  int64_t nnz = -1;
  SymGraph g;
  std::vector<int64_t> Wall;
};
static uint64_t csr_key(int N, const int64_t* offsets, const int32_t* adj) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
  for (int i = 0; i <= N; ++i) mix((uint64_t)offsets[i]);
  for (int64_t e = 0; e < offsets[N]; ++e) mix((uint32_t)adj[e]);
  return h;
}

// GCO's expansion(): sweep the labels 0..L-1 in order, apply a move iff it strictly lowers the energy, stop when a full sweep
// changes nothing.  The sweep is sequential by definition (move alpha+1 sees the result of move alpha), but a move reads only
// the labels of its candidate set and their neighbours, and moves of different hypotheses mostly live in different parts of the
// image.  So a window of upcoming moves is evaluated SPECULATIVELY in parallel on the current labelling, and the results are
// committed in label order: a result is used iff no site that changed in the meantime belongs to what the move read (its
// `touched` set) or newly qualifies as a candidate; otherwise the move is evaluated again on the current labelling.  The
// sequence of applied moves — hence labels and energy — is exactly the sequential one.
mh_status alpha_expansion(const int32_t* cost, int N, int L, int potts, const int64_t* offsets, const int32_t* adj,
                          const int32_t* init, int max_cycles, int32_t* lab, int64_t* energy_out) {
  if (N <= 0 || L < 1) return MH_EINVAL;
  GC_TICK(t_total);
  for (int i = 0; i < N; ++i) {
    lab[i] = init ? init[i] : 0;
    if (lab[i] < 0 || lab[i] >= L) return MH_EINVAL;
  }
  static thread_local PreparedGraph prepared;
  static thread_local SymGraph no_edges;
  const SymGraph* gp = &no_edges;
  if (offsets && adj) {
    const uint64_t key = csr_key(N, offsets, adj);
    if (prepared.N != N || prepared.nnz != offsets[N] || prepared.key != key) {
      symmetrise(N, offsets, adj, prepared.g);
      prepared.N = N; prepared.nnz = offsets[N]; prepared.key = key;
      prepared.Wall.assign(N, 0);   // SUM_j w_ij (times potts below: potts may differ between calls)
      for (int i = 0; i < N; ++i)
        for (int64_t k = prepared.g.off[i]; k < prepared.g.off[i + 1]; ++k) prepared.Wall[i] += prepared.g.w[k];
    }
    gp = &prepared.g;
  } else {
    no_edges.off.assign(N + 1, 0);
  }
  const SymGraph& g = *gp;
  const bool have_edges = !g.nbr.empty();

  if (!have_edges) {  // GCO solveSpecialCases: data costs only -> per-site argmin, first label wins ties
    for (int i = 0; i < N; ++i) {
      int best = 0;
      for (int l = 1; l < L; ++l)
        if (cost[(size_t)i * L + l] < cost[(size_t)i * L + best]) best = l;
      lab[i] = best;
    }
    if (energy_out) *energy_out = total_energy(cost, N, L, potts, g, lab);
    return MH_OK;
  }

  std::vector<int64_t> Wall(N);
  for (int i = 0; i < N; ++i) Wall[i] = (int64_t)potts * prepared.Wall[i];
  // label-major copy of the costs and the cost of every site at its current label: a move's candidate scan becomes a
  // contiguous pass (the site-major matrix would be read with stride L)
  std::vector<int32_t> costT((size_t)N * L), cur(N);
  for (int i = 0; i < N; ++i) {
    const int32_t* row = cost + (size_t)i * L;
    for (int l = 0; l < L; ++l) costT[(size_t)l * N + i] = row[l];
    cur[i] = row[lab[i]];
  }
  int solver = 0;
  if (const char* e = std::getenv("MH_GC_SOLVER")) solver = !std::strcmp(e, "dinic") ? 1 : !std::strcmp(e, "pr") ? 2 : 0;
  const MoveProblem P{cost, N, L, potts, costT.data(), cur.data(), &g, Wall.data(), solver};

  MovePool* pool = (L >= 4 && (int64_t)N * L >= 4096) ? MovePool::acquire() : nullptr;
  const int nthreads = pool ? pool->threads() : 1;
  static thread_local std::vector<MoveWS> ws;
  static thread_local std::vector<MoveResult> res;
  if ((int)ws.size() < nthreads) ws.resize(nthreads);
  for (int t = 0; t < nthreads; ++t) ws[t].prepare(N);
  const int window = pool ? std::min(L, 2 * nthreads) : 1;
  if ((int)res.size() < window) res.resize(window);
  MoveWS* const wsp = ws.data();          // (thread_local objects: workers must go through the caller's pointers)
  MoveResult* const resp = res.data();

  // Move memo.  Right after move alpha has been processed the labelling admits no improving alpha-expansion, and what a
  // re-evaluation would read lies inside the `touched` set of that evaluation (the candidates only shrink by the sites that
  // switched).  So when alpha comes round again and none of the sites that changed in between is in that set or newly
  // qualifies as a candidate, the move is known to change nothing and is not evaluated at all — most moves of the later
  // cycles.  `log` lists the changed sites in commit order; memo[alpha].pos = its length when alpha was last processed.
  struct Memo { bool valid = false, any_c0 = false; size_t pos = 0; std::vector<uint64_t> touched; };
  std::vector<Memo> memo(L);
  std::vector<int> log;
  auto clean_since = [&](int alpha, const std::vector<uint64_t>& touched, bool any_c0, size_t from) {
    for (size_t k = from; k < log.size(); ++k) {
      const int d = log[k];
      if (any_c0) {   // d in C0, or a neighbour of a C0 site (the graph is symmetric)
        if ((touched[(size_t)d >> 6] >> (d & 63)) & 1ull) return false;
        for (int64_t q = g.off[d]; q < g.off[d + 1]; ++q)
          if ((touched[(size_t)g.nbr[q] >> 6] >> (g.nbr[q] & 63)) & 1ull) return false;
      }
      if (lab[d] != alpha && (int64_t)cost[(size_t)d * L + alpha] - cost[(size_t)d * L + lab[d]] <= Wall[d]) return false;
    }
    return true;
  };
  auto known_idle = [&](int alpha) {
    Memo& m = memo[alpha];
    if (!m.valid || log.size() - m.pos > 256) return false;   // (many changes: evaluating is cheaper than checking, and rarely idle)
    if (!clean_since(alpha, m.touched, m.any_c0, m.pos)) return false;
    m.pos = log.size();
    return true;
  };

  GC_ACC(7, t_total);
  if (max_cycles < 0) max_cycles = 1 << 30;
  const int64_t max_moves = (int64_t)std::min<int64_t>(max_cycles, (1LL << 40) / L) * L;
  int idle_moves = 0;   // consecutive moves that changed nothing: L of them = a full sweep over an unchanged labelling
  int64_t E_delta = 0;  // energy change within the current cycle (GCO stops after a cycle without change)
  int64_t move = 0;
  bool stop = false;
  std::vector<char> speculated(window);
  while (!stop && move < max_moves && idle_moves < L) {
    const int nw = (int)std::min<int64_t>(window, max_moves - move);
    const int a0 = (int)(move % L);
    const size_t snapshot = log.size();   // speculative results below are evaluated on the labelling at this point
    if (pool && nw > 1) {
      int todo = 0;
      for (int j = 0; j < nw; ++j) todo += (speculated[j] = !known_idle((a0 + j) % L));
      if (todo > 1)
        pool->run(nw, [&, wsp, resp](int j, int t) {
          if (speculated[j]) eval_move(P, lab, (a0 + j) % L, wsp[t], resp[j], true);
        });
      else
        std::fill(speculated.begin(), speculated.end(), 0);
    }
    for (int j = 0; j < nw && !stop; ++j, ++move) {
      const int alpha = (a0 + j) % L;
      ++idle_moves;
      GC_TICK(t_memo);
      const bool idle_known = known_idle(alpha);
      GC_ACC(6, t_memo);
      if (!idle_known) {
        const bool usable = pool && nw > 1 && speculated[j] && clean_since(alpha, res[j].touched, res[j].any_c0, snapshot);
        if (!usable) eval_move(P, lab, alpha, ws[0], res[j], true);
        GC_TICK(t_commit);
#ifdef MH_GC_PROFILE
        g_prof[14] += res[j].sw.size(); g_prof_n[14] += !res[j].sw.empty();
#endif
        if (!res[j].sw.empty()) {
          for (int i : res[j].sw) { lab[i] = alpha; cur[i] = cost[(size_t)i * L + alpha]; log.push_back(i); }
          E_delta += res[j].delta;
          idle_moves = 0;
        }
        Memo& m = memo[alpha];
        m.valid = true; m.any_c0 = res[j].any_c0; m.pos = log.size();
        m.touched.swap(res[j].touched);
        GC_ACC(9, t_commit);
      }
      if (alpha == L - 1) {   // end of a cycle: GCO stops when the cycle left the energy unchanged
        if (E_delta == 0) stop = true;
        E_delta = 0;
      }
      if (idle_moves >= L) stop = true;
    }
  }
  if (pool) pool->release();
  if (energy_out) *energy_out = total_energy(cost, N, L, potts, g, lab);
  GC_ACC(8, t_total);
  return MH_OK;
}

// ---------------------------------------------------------------------------
// Neighbourhood on float (x1,y1,x2,y2): the `max_neighbours` nearest sites (ties by index) among those with
// d^2 <= radius^2 (OpenCV's FlannBasedMatcher squares maxDistance for the L2 index); max_neighbours <= 0 = the full
// ball.  This restates what the reference's radiusMatch call returns with FLANN's default parameters (4 randomised
// KD-trees, checks = 32, radius result set always "full" => the search stops after 32 examined points, the query
// among them) as an exactly defined set — see oracle/multih_oracle.cpp orc_radius_neighbours.
// Uniform grid over (x1,y1), ring search with early termination.  Neighbours are listed in ascending index order;
// the site itself is excluded (MultiH.cpp:537).
// ---------------------------------------------------------------------------
int64_t radius_neighbourhood(const double* pts, int N, double radius, int max_neighbours, int64_t* offsets,
                             int32_t* adj) {
  if (N <= 0) { if (offsets) offsets[0] = 0; return 0; }
  const float r = (float)radius, r2 = r * r;
  float minx = 1e30f, miny = 1e30f, maxx = -1e30f, maxy = -1e30f;
  std::vector<float> p(4 * (size_t)N);
  for (int i = 0; i < N; ++i) {
    for (int k = 0; k < 4; ++k) p[4 * (size_t)i + k] = (float)pts[4 * (size_t)i + k];
    minx = std::min(minx, p[4 * (size_t)i]); maxx = std::max(maxx, p[4 * (size_t)i]);
    miny = std::min(miny, p[4 * (size_t)i + 1]); maxy = std::max(maxy, p[4 * (size_t)i + 1]);
  }
  const bool knn = max_neighbours > 0;
  // cell size: the radius for ball queries; ~4 sites per cell for k-nearest queries (never above the radius)
  float cell = std::max(r, 1e-6f);
  if (knn) {
    const double area = std::max(1e-12, (double)(maxx - minx) * (double)(maxy - miny));
    cell = std::min(cell, (float)std::sqrt(area * 4.0 / N));
    cell = std::max(cell, 1e-6f);
  }
  const int gx = std::max(1, std::min(4096, (int)((maxx - minx) / cell) + 1));
  const int gy = std::max(1, std::min(4096, (int)((maxy - miny) / cell) + 1));
  const float cw = std::max(cell, (maxx - minx) / gx + 1e-6f), ch = std::max(cell, (maxy - miny) / gy + 1e-6f);
  auto cx = [&](float x) { return std::min(gx - 1, std::max(0, (int)((x - minx) / cw))); };
  auto cy = [&](float y) { return std::min(gy - 1, std::max(0, (int)((y - miny) / ch))); };
  const float cmin = std::min(cw, ch);
  std::vector<int> cstart((size_t)gx * gy + 1, 0), order(N);
  for (int i = 0; i < N; ++i) ++cstart[(size_t)cy(p[4 * (size_t)i + 1]) * gx + cx(p[4 * (size_t)i]) + 1];
  for (size_t c = 0; c < (size_t)gx * gy; ++c) cstart[c + 1] += cstart[c];
  {
    std::vector<int> cur(cstart.begin(), cstart.end() - 1);
    for (int i = 0; i < N; ++i) order[cur[(size_t)cy(p[4 * (size_t)i + 1]) * gx + cx(p[4 * (size_t)i])]++] = i;
  }
  int64_t total = 0;
  typedef std::pair<float, int> Cand;
  std::vector<Cand> heap;  // max-heap on (d2, index): top = worst kept candidate
  std::vector<int32_t> row;
  const int max_ring = std::max(gx, gy);
  for (int i = 0; i < N; ++i) {
    if (offsets) offsets[i] = total;
    const float* a = &p[4 * (size_t)i];
    const int ix = cx(a[0]), iy = cy(a[1]);
    heap.clear();
    for (int ring = 0; ring <= max_ring; ++ring) {
      if (ring > 0) {
        const float reach = (ring - 1) * cmin;  // every site of this ring is at least this far away in (x1,y1)
        if (reach * reach > r2) break;
        if (knn && (int)heap.size() == max_neighbours && reach * reach > heap.front().first) break;
      }
      if (ring > std::max(std::max(ix, gx - 1 - ix), std::max(iy, gy - 1 - iy))) break;  // past the grid
      const int y0 = iy - ring, y1 = iy + ring, x0 = ix - ring, x1 = ix + ring;
      auto visit = [&](int xx, int yy) {
        if (xx < 0 || yy < 0 || xx >= gx || yy >= gy) return;
        const size_t c = (size_t)yy * gx + xx;
        for (int k = cstart[c]; k < cstart[c + 1]; ++k) {
          const int j = order[k];
          if (j == i) continue;
          const float* b = &p[4 * (size_t)j];
          const float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2], d3 = a[3] - b[3];
          const float d = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          if (d > r2) continue;
          const Cand cand(d, j);
          if (!knn || (int)heap.size() < max_neighbours) {
            heap.push_back(cand);
            if (knn) std::push_heap(heap.begin(), heap.end());
          } else if (cand < heap.front()) {
            std::pop_heap(heap.begin(), heap.end());
            heap.back() = cand;
            std::push_heap(heap.begin(), heap.end());
          }
        }
      };
      if (ring == 0) visit(ix, iy);
      else {
        for (int xx = x0; xx <= x1; ++xx) { visit(xx, y0); visit(xx, y1); }
        for (int yy = y0 + 1; yy <= y1 - 1; ++yy) { visit(x0, yy); visit(x1, yy); }
      }
    }
    if (adj) {
      row.clear();
      for (const Cand& c : heap) row.push_back(c.second);
      std::sort(row.begin(), row.end());
      std::memcpy(adj + total, row.data(), sizeof(int32_t) * row.size());
    }
    total += (int64_t)heap.size();
  }
  if (offsets) offsets[N] = total;
  return total;
}

}  // namespace mh
