// Host-logic check of csrc/csl_host.h (no GPU): build compacted-slice tiles from a random ragged CSR pair, decode the
// blobs back and compare; check the value maps and the tile partitions (whole range and per-rank sub-ranges).
// Built and run by tests/test_csl_host.py:  g++ -O1 -std=c++17 -I osqp-python_b200/csrc tests/host/csl_host_check.cpp
#include <cstdio>
#include <cstdlib>
#include <random>

#include "csl_host.h"

using namespace b200qp;

struct Csr { int rows, cols; std::vector<int> ptr, idx; std::vector<double> val; };
static Csr random_csr(int rows, int cols, int maxlen, unsigned seed) {
  std::mt19937 rng(seed);
  Csr M{rows, cols, {0}, {}, {}};
  for (int r = 0; r < rows; ++r) {
    const int len = (int)(rng() % (unsigned)(maxlen + 1));          // empty rows included
    std::vector<int> c;
    for (int k = 0; k < len; ++k) c.push_back((int)(rng() % (unsigned)cols));
    std::sort(c.begin(), c.end());
    c.erase(std::unique(c.begin(), c.end()), c.end());
    for (int j : c) { M.idx.push_back(j); M.val.push_back(1.0 + (double)(rng() % 1000) / 7.0); }
    M.ptr.push_back((int)M.idx.size());
  }
  return M;
}
#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int check(int rows, const Csr* A, const Csr* B) {
  CslHost H;
  csl_build(rows, A ? A->ptr.data() : nullptr, A ? A->idx.data() : nullptr, B ? B->ptr.data() : nullptr, B ? B->idx.data() : nullptr, H);
  REQUIRE(H.ok && H.ntiles == (rows + 31) / 32 && (int)H.toff.size() == H.ntiles + 1);
  double* bd = reinterpret_cast<double*>(H.blob.data());
  if (A) for (size_t e = 0; e < H.src1.size(); ++e) bd[H.dst1[e]] = A->val[H.src1[e]];
  if (B) for (size_t e = 0; e < H.src2.size(); ++e) bd[H.dst2[e]] = B->val[H.src2[e]];
  REQUIRE(!A || H.src1.size() == A->idx.size());
  REQUIRE(!B || H.src2.size() == B->idx.size());
  for (int t = 0; t < H.ntiles; ++t) {
    const unsigned char* tp = H.blob.data() + (size_t)H.toff[t] * 16;
    REQUIRE(((size_t)H.toff[t] * 16) % 128 == 0);                      // tiles start on 128-byte lines
    const int* hdr = reinterpret_cast<const int*>(tp);
    const int c1 = hdr[0], c2 = hdr[1], m1 = hdr[2], m2 = hdr[3];
    REQUIRE(c1 % 4 == 0 && c2 % 4 == 0);
    REQUIRE(80 + 12 * (size_t)(c1 + c2) <= ((size_t)H.toff[t + 1] - H.toff[t]) * 16);
    const double* v1 = reinterpret_cast<const double*>(tp + 80);
    const double* v2 = v1 + c1;
    const int* i1 = reinterpret_cast<const int*>(v2 + c2);
    const int* i2 = i1 + c1;
    for (int part = 0; part < 2; ++part) {
      const Csr* M = part ? B : A;
      if (!M) continue;
      const unsigned char* len = tp + 16 + 32 * part;
      const double* v = part ? v2 : v1;
      const int* ix = part ? i2 : i1;
      const int mx = part ? m2 : m1;
      int k = 0, seen_max = 0;
      for (int j = 0; j < mx; ++j)
        for (int l = 0; l < 32; ++l) {
          const int r = t * 32 + l;
          const int rl = r < rows ? M->ptr[r + 1] - M->ptr[r] : 0;
          if (j == 0) { REQUIRE(len[l] == rl); seen_max = std::max(seen_max, rl); }
          if (rl > j) {                                                // column-major, compacted, row order
            REQUIRE(ix[k] == M->idx[M->ptr[r] + j]);
            REQUIRE(v[k] == M->val[M->ptr[r] + j]);
            ++k;
          }
        }
      REQUIRE(mx == seen_max || mx == 0);
    }
  }
  // partitions: monotone, cover everything, sub-ranges nest
  for (int G : {1, 3, 7, 64}) {
    std::vector<int> p;
    csl_partition(H, G, p);
    REQUIRE((int)p.size() == G + 1 && p[0] == 0 && p[G] == H.ntiles);
    for (int g = 0; g < G; ++g) REQUIRE(p[g] <= p[g + 1]);
  }
  for (int world : {1, 2, 3, 8}) {
    std::vector<int> rk;
    csl_partition_range(H.toff, 0, H.ntiles, world, rk);
    REQUIRE(rk[0] == 0 && rk[world] == H.ntiles);
    int covered = 0;
    for (int h = 0; h < world; ++h) {
      REQUIRE(rk[h] <= rk[h + 1]);
      std::vector<int> pc;
      csl_partition_range(H.toff, rk[h], rk[h + 1], 5, pc);
      REQUIRE(pc[0] == rk[h] && pc[5] == rk[h + 1]);
      for (int g = 0; g < 5; ++g) { REQUIRE(pc[g] <= pc[g + 1]); covered += pc[g + 1] - pc[g]; }
    }
    REQUIRE(covered == H.ntiles);
  }
  return 0;
}

int main() {
  for (int rows : {1, 31, 32, 33, 1000, 4097}) {
    Csr A = random_csr(rows, 777, 12, 11u + rows), B = random_csr(rows, 1234, 30, 5u + rows);
    if (check(rows, &A, &B)) return 1;          // n-side tile: two parts
    if (check(rows, &A, nullptr)) return 1;     // m-side tile: one part
  }
  {   // a row longer than 255 entries cannot be tiled: the builder must say so (the engine then keeps the CSR kernel)
    Csr L = random_csr(40, 5000, 3, 3);
    Csr W{40, 5000, {0}, {}, {}};
    for (int r = 0; r < 40; ++r) { const int len = r == 17 ? 300 : 2; for (int k = 0; k < len; ++k) { W.idx.push_back(k * 3); W.val.push_back(1.0); } W.ptr.push_back((int)W.idx.size()); }
    CslHost H;
    csl_build(40, L.ptr.data(), L.idx.data(), W.ptr.data(), W.idx.data(), H);
    REQUIRE(!H.ok);
  }
  std::printf("csl_host OK\n");
  return 0;
}
