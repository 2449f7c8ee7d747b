// Exact linear minimisation oracle of GWD-B for rectangular plans: the transportation problem
//     min <cost, G>   s.t.  G 1 = 1/n,  G^T 1 = 1/m,  G >= 0
// that POT's ot.emd solves inside ot.gromov.gromov_wasserstein (representations/representation_search/
// gromov_wasserstein.py:62-69 pairs n events with m != n pixels, :85-184).  Host code: the solve is a sequential
// combinatorial algorithm on an (n + m)-node network; the dense work of a conditional-gradient step (gradient, contraction,
// line search) stays on the GPU (gw_kl.cu).
//
// Successive shortest paths with node potentials, dense ("Hungarian for transportation"): supplies and demands are
// the integers m / g and n / g (g = gcd(n, m); one unit = g / (n m)), so flows are exact; one source at a time ships
// its supply along shortest residual paths found by a Dijkstra over the reduced costs (array-based: the graph is
// complete bipartite), and every reached node's potential moves by (D - dist) so that reduced costs stay >= 0 and are
// 0 on arcs that carry flow.  The result is an optimal vertex (at most n + m - 1 positive entries) in CSR form.
#include <float.h>
#include <limits.h>

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "evrep_common.cuh"

namespace evrep {

struct FlowArc {
  int node;
  long long x;
};

static void arc_add(std::vector<FlowArc>& lst, int node, long long d) {
  for (size_t k = 0; k < lst.size(); ++k)
    if (lst[k].node == node) {
      lst[k].x += d;
      if (lst[k].x == 0) {
        lst[k] = lst.back();
        lst.pop_back();
      }
      return;
    }
  lst.push_back({node, d});
}

// cost: n x m row major (host).  row_ptr (n + 1), col / weight (capacity cap): the plan in CSR, weights sum to 1.
int transport_plan_host(const float* cost, int n, int m, int cap, int* row_ptr, int* col, double* weight, int* nnz_out) {
  if (n < 1 || m < 1) {
    set_error("transport: need n, m >= 1");
    return EVREP_EINVAL;
  }
  for (size_t e = 0; e < (size_t)n * m; ++e)
    if (!std::isfinite(cost[e])) {
      set_error("transport: non-finite cost at entry %zu", e);
      return EVREP_EINVAL;
    }
  const long long g = std::gcd((long long)n, (long long)m);
  const long long sup0 = m / g, dem0 = n / g;
  std::vector<long long> dem(m, dem0);
  std::vector<double> u(n, 0.0), v(m), minv(m), ds(n);
  std::vector<int> way(m), prev_sink(n), reached, scanned;
  std::vector<char> used(m), seen(n);
  std::vector<std::vector<FlowArc>> by_sink(m), by_src(n);
  for (int j = 0; j < m; ++j) {  // column reduction: reduced costs start >= 0
    double lo = DBL_MAX;
    for (int i = 0; i < n; ++i) lo = std::min(lo, (double)cost[(size_t)i * m + j]);
    v[j] = lo;
  }
  reached.reserve(n);
  scanned.reserve(m);
  for (int s = 0; s < n; ++s) {
    long long rem = sup0;
    while (rem > 0) {
      // Dijkstra from s over the residual graph
      std::fill(used.begin(), used.end(), 0);
      for (int i : reached) seen[i] = 0;
      reached.clear();
      scanned.clear();
      seen[s] = 1;
      ds[s] = 0.0;
      reached.push_back(s);
      {
        const float* crow = cost + (size_t)s * m;
        for (int j = 0; j < m; ++j) {
          minv[j] = (double)crow[j] - u[s] - v[j];
          way[j] = s;
        }
      }
      int t = -1;
      double D = 0.0;
      while (true) {
        int j1 = -1;
        double best = DBL_MAX;
        for (int j = 0; j < m; ++j)
          if (!used[j] && minv[j] < best) {
            best = minv[j];
            j1 = j;
          }
        if (j1 < 0) {
          set_error("transport: internal error (no augmenting path)");
          return EVREP_EINVAL;
        }
        used[j1] = 1;
        scanned.push_back(j1);
        D = best;
        if (dem[j1] > 0) {
          t = j1;
          break;
        }
        for (const FlowArc& a : by_sink[j1]) {  // backward arcs: reduced cost 0
          const int i = a.node;
          if (seen[i]) continue;
          seen[i] = 1;
          ds[i] = D;
          prev_sink[i] = j1;
          reached.push_back(i);
          const float* crow = cost + (size_t)i * m;
          const double base = D - u[i];
          for (int j = 0; j < m; ++j) {
            if (used[j]) continue;
            const double cand = base + (double)crow[j] - v[j];
            if (cand < minv[j]) {
              minv[j] = cand;
              way[j] = i;
            }
          }
        }
      }
      // bottleneck along the path t <- way[t] <- prev_sink[..] <- ... <- s
      long long delta = std::min(rem, dem[t]);
      for (int j = t;;) {
        const int i = way[j];
        if (i == s) break;
        const int jb = prev_sink[i];
        for (const FlowArc& a : by_sink[jb])
          if (a.node == i) {
            delta = std::min(delta, a.x);
            break;
          }
        j = jb;
      }
      for (int j = t;;) {
        const int i = way[j];
        arc_add(by_sink[j], i, delta);
        arc_add(by_src[i], j, delta);
        if (i == s) break;
        const int jb = prev_sink[i];
        arc_add(by_sink[jb], i, -delta);
        arc_add(by_src[i], jb, -delta);
        j = jb;
      }
      rem -= delta;
      dem[t] -= delta;
      for (int i : reached) u[i] += D - ds[i];
      for (int j : scanned) v[j] -= D - minv[j];
    }
  }
  const double unit = (double)g / ((double)n * (double)m);
  int nnz = 0;
  for (int i = 0; i < n; ++i) {
    row_ptr[i] = nnz;
    std::sort(by_src[i].begin(), by_src[i].end(), [](const FlowArc& a, const FlowArc& b) { return a.node < b.node; });
    for (const FlowArc& a : by_src[i]) {
      if (nnz >= cap) {  // a vertex has at most n + m - 1 entries; tied costs can leave a few more
        set_error("transport: the plan has more than cap = %d entries", cap);
        return EVREP_EWORKSPACE;
      }
      col[nnz] = a.node;
      weight[nnz] = (double)a.x * unit;
      ++nnz;
    }
  }
  row_ptr[n] = nnz;
  if (nnz_out) *nnz_out = nnz;
  return EVREP_OK;
}

}  // namespace evrep
