// Pins the oracle's restatement of std::sort-with-a-non-strict-weak-comparator (LongestEdgeRefinement::coarsening,
// longestedgerefinement.hh:92-112) against the real std::sort of this toolchain's libstdc++.
// stdin: n, then per case n x (level len constrained); stdout: the sorted index order per case.
#include <algorithm>
#include <array>
#include <cstdio>
#include <utility>
#include <vector>

struct V { int idx; long level; bool constrained; };

int main() {
  int n;
  if (std::scanf("%d", &n) != 1 || n < 1 || n > 8) return 1;
  std::vector<std::pair<V, double>> v(n);
  for (;;) {
    for (int k = 0; k < n; ++k) {
      long lev; double len; int con;
      if (std::scanf("%ld %lf %d", &lev, &len, &con) != 3) return 0;
      v[k] = { V{ k, lev, con != 0 }, len };
    }
    std::sort(v.begin(), v.end(), [&](auto a, auto b) {
      if (a.first.level < b.first.level) return true;
      if (a.second < b.second - 1e-14) return true;
      auto constrained = [&](V x) { return x.constrained; };
      if (!constrained(a.first) and constrained(b.first)) return true;
      return false;
    });
    for (int k = 0; k < n; ++k) std::printf("%d%c", v[k].first.idx, k + 1 == n ? '\n' : ' ');
  }
}
