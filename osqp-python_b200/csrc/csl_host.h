// csl_host.h -- host side of admm_kernel_csl.cuh: cut CSR patterns into compacted-slice tiles (once, at setup).
//
// A stream is a sequence of 32-row tiles over one row set; every tile carries the rows of up to two matrices
// (n-side: P and A'; m-side: A).  Tile blob (16-byte aligned, see admm_kernel_csl.cuh):
//   int4 {cnt1, cnt2, maxlen1, maxlen2} | u8 len1[32] | u8 len2[32] | f64 val1[cnt1] | f64 val2[cnt2] | i32 idx1[cnt1] | i32 idx2[cnt2]
// with cnt = entries rounded up to a multiple of 4.  Column indices are written here; values are (re)packed on the
// device from the scaled CSR arrays through the (src, dst) maps, after equilibration and after every value update.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace b200qp {

struct CslHost {
  int rows = 0, ntiles = 0;
  std::vector<unsigned> toff;          // ntiles + 1, 16-byte units
  std::vector<unsigned char> blob;
  std::vector<int> src1, src2;         // CSR position of every entry of part 1 / part 2
  std::vector<unsigned> dst1, dst2;    // its value slot in the blob, in 8-byte units
  unsigned max_tile_bytes = 0;
  bool ok = true;                      // false: a row is longer than 255 entries (u8 length)
  long long waste4 = 0, waste5 = 0;    // padded gather slots of part 2 (part 1 when there is no part 2) for U = 4 / 5
};

// ptr1 / idx1 may be null (no part 1).  Parts share the row set [0, rows).
inline void csl_build(int rows, const int* ptr1, const int* idx1, const int* ptr2, const int* idx2, CslHost& H) {
  H = CslHost();
  H.rows = rows;
  H.ntiles = (rows + 31) / 32;
  H.toff.assign(H.ntiles + 1, 0);
  // pass 1: sizes
  size_t bytes = 0;
  for (int t = 0; t < H.ntiles; ++t) {
    H.toff[t] = (unsigned)(bytes / 16);
    long long c1 = 0, c2 = 0;
    int m1 = 0, m2 = 0;
    for (int l = 0; l < 32; ++l) {
      const int r = t * 32 + l;
      if (r >= rows) break;
      if (ptr1) { const int len = ptr1[r + 1] - ptr1[r]; c1 += len; m1 = std::max(m1, len); }
      if (ptr2) { const int len = ptr2[r + 1] - ptr2[r]; c2 += len; m2 = std::max(m2, len); }
    }
    if (m1 > 255 || m2 > 255) H.ok = false;
    const int mm = ptr2 ? m2 : m1;
    H.waste4 += (mm + 3) / 4 * 4;
    H.waste5 += (mm + 4) / 5 * 5;
    c1 = (c1 + 3) & ~3LL; c2 = (c2 + 3) & ~3LL;
    // tiles start on 128-byte lines: a bulk copy that begins or ends inside a line re-fetches the shared line and
    // runs at sector granularity
    const size_t tb = (80 + 12 * (size_t)(c1 + c2) + 127) & ~(size_t)127;
    H.max_tile_bytes = std::max<unsigned>(H.max_tile_bytes, (unsigned)tb);
    bytes += tb;
  }
  H.toff[H.ntiles] = (unsigned)(bytes / 16);
  if (!H.ok) return;
  H.blob.assign(bytes + 16, 0);
  if (ptr1) { H.src1.resize(ptr1[rows]); H.dst1.resize(ptr1[rows]); }
  if (ptr2) { H.src2.resize(ptr2[rows]); H.dst2.resize(ptr2[rows]); }
  // pass 2: fill
  size_t e1 = 0, e2 = 0;
  for (int t = 0; t < H.ntiles; ++t) {
    unsigned char* tp = H.blob.data() + (size_t)H.toff[t] * 16;
    const int nr = std::min(32, rows - t * 32);
    int len1[32] = {0}, len2[32] = {0}, m1 = 0, m2 = 0;
    long long c1 = 0, c2 = 0;
    for (int l = 0; l < nr; ++l) {
      const int r = t * 32 + l;
      if (ptr1) { len1[l] = ptr1[r + 1] - ptr1[r]; c1 += len1[l]; m1 = std::max(m1, len1[l]); }
      if (ptr2) { len2[l] = ptr2[r + 1] - ptr2[r]; c2 += len2[l]; m2 = std::max(m2, len2[l]); }
    }
    const int p1 = (int)((c1 + 3) & ~3LL), p2 = (int)((c2 + 3) & ~3LL);
    int hdr[4] = {p1, p2, m1, m2};
    std::memcpy(tp, hdr, 16);
    for (int l = 0; l < 32; ++l) { tp[16 + l] = (unsigned char)len1[l]; tp[48 + l] = (unsigned char)len2[l]; }
    const size_t v1off = (size_t)H.toff[t] * 16 + 80;                 // bytes
    const size_t v2off = v1off + 8 * (size_t)p1;
    int* i1 = reinterpret_cast<int*>(H.blob.data() + v2off + 8 * (size_t)p2);
    int* i2 = i1 + p1;
    int k = 0;
    for (int j = 0; j < m1; ++j)
      for (int l = 0; l < nr; ++l)
        if (len1[l] > j) {
          const int pos = ptr1[t * 32 + l] + j;
          i1[k] = idx1[pos];
          H.src1[e1] = pos; H.dst1[e1] = (unsigned)(v1off / 8 + k); ++e1; ++k;
        }
    k = 0;
    for (int j = 0; j < m2; ++j)
      for (int l = 0; l < nr; ++l)
        if (len2[l] > j) {
          const int pos = ptr2[t * 32 + l] + j;
          i2[k] = idx2[pos];
          H.src2[e2] = pos; H.dst2[e2] = (unsigned)(v2off / 8 + k); ++e2; ++k;
        }
  }
}

// contiguous tile ranges per CTA, balanced by tile bytes (+ a constant per tile for the epilogue)
inline void csl_partition(const CslHost& H, int G, std::vector<int>& part) {
  part.assign(G + 1, H.ntiles);
  part[0] = 0;
  if (H.ntiles == 0) { std::fill(part.begin(), part.end(), 0); return; }
  std::vector<long long> pre(H.ntiles + 1, 0);
  for (int t = 0; t < H.ntiles; ++t) pre[t + 1] = pre[t] + 16LL * (H.toff[t + 1] - H.toff[t]) + 512;
  int t = 0;
  for (int g = 1; g < G; ++g) {
    const long long target = pre[H.ntiles] * g / G;
    t = (int)(std::lower_bound(pre.begin() + t, pre.end(), target) - pre.begin());
    part[g] = std::min(t, H.ntiles);
  }
  part[G] = H.ntiles;
}

// the same over the tile sub-range [t_lo, t_hi) (multi-GPU: rank ranges first, then the CTAs of a rank); toff = tile
// offsets in 16-byte units (CslHost::toff); part receives G + 1 absolute tile numbers
inline void csl_partition_range(const std::vector<unsigned>& toff, int t_lo, int t_hi, int G, std::vector<int>& part) {
  part.assign(G + 1, t_hi);
  part[0] = t_lo;
  const int T = t_hi - t_lo;
  if (T <= 0) { std::fill(part.begin(), part.end(), t_lo); return; }
  std::vector<long long> pre(T + 1, 0);
  for (int t = 0; t < T; ++t) pre[t + 1] = pre[t] + 16LL * (toff[t_lo + t + 1] - toff[t_lo + t]) + 512;
  int t = 0;
  for (int g = 1; g < G; ++g) {
    const long long target = pre[T] * g / G;
    t = (int)(std::lower_bound(pre.begin() + t, pre.end(), target) - pre.begin());
    part[g] = t_lo + std::min(t, T);
  }
  part[G] = t_hi;
}

}  // namespace b200qp
