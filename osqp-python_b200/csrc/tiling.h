// tiling.h -- host side of tile_spmv.cuh: cut a CSR pattern into tiles once at setup and hand every CTA of
// the persistent grid a contiguous, cost-balanced run of tiles.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

namespace b200qp {

// Whole-row tiles hold <= tile_nnz nonzeros and <= max_rows rows; a row longer than tile_nnz becomes a run of
// chunk tiles (y = -1, last one y = -2).
inline void make_tiles(int rows, const int* ptr, int tile_nnz, int max_rows, std::vector<int4>& tiles) {
  tiles.clear();
  int r = 0;
  while (r < rows) {
    const int len = ptr[r + 1] - ptr[r];
    if (len > tile_nnz) {
      for (int k = ptr[r]; k < ptr[r + 1]; k += tile_nnz) {
        const int e = std::min(k + tile_nnz, ptr[r + 1]);
        tiles.push_back(make_int4(r, e == ptr[r + 1] ? -2 : -1, k, e));
      }
      ++r;
      continue;
    }
    int e = r;
    while (e < rows && e - r < max_rows && ptr[e + 1] - ptr[r] <= tile_nnz) ++e;
    tiles.push_back(make_int4(r, e, ptr[r], ptr[e]));
    r = e;
  }
}

// cta_tile[c] .. cta_tile[c+1] = tiles of CTA c; the chunk tiles of one row stay in one CTA.
// part[c] .. part[c+1] = the rows those tiles cover (used for the element-wise phases).
inline void partition_tiles(const std::vector<int4>& tiles, int rows, int G, std::vector<int>& cta_tile,
                            std::vector<int>& part) {
  const int T = (int)tiles.size();
  std::vector<long long> pre(T + 1, 0);
  for (int t = 0; t < T; ++t) {
    const int nr = tiles[t].y >= 0 ? tiles[t].y - tiles[t].x : 1;
    pre[t + 1] = pre[t] + (tiles[t].w - tiles[t].z) + 2LL * nr + 64;
  }
  cta_tile.assign(G + 1, T);
  cta_tile[0] = 0;
  int t = 0;
  for (int g = 1; g < G; ++g) {
    const long long target = pre[T] * g / G;
    t = (int)(std::lower_bound(pre.begin() + t, pre.end(), target) - pre.begin());
    while (t < T && t > 0 && tiles[t - 1].y == -1) ++t;   // do not cut a long row
    t = std::min(t, T);
    cta_tile[g] = t;
  }
  part.resize(G + 1);
  for (int g = 0; g <= G; ++g) part[g] = cta_tile[g] < T ? tiles[cta_tile[g]].x : rows;
  part[0] = 0;
}

}  // namespace b200qp
