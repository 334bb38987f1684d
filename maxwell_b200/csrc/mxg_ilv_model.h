// Host-only cost model that picks the thread -> row assignment of the dictionary SpMM kernel (mxg_spmv.cu).
//
// ncu shows that kernel bound by L1 data-pipe wavefronts (profiles/README_r01.md), so the model counts, for sampled
// 96-row tiles, the distinct 128-byte lines each warp-wide load touches under the two assignments:
//   plain      : warp w of the tile takes rows tile + 32 w + lane
//   interleaved: warp c of the tile takes rows tile + 3 lane + c   (32 cells of ONE field component when the DOFs
//                come in component triples, GID = comp + 3 cell)
// Per pattern entry a warp issues one pattern-table load (cost: distinct pattern lines, paid once per group of up
// to 4 columns) and one x gather per column (cost: distinct x lines); per row it reads the pattern id and stores y.
// Measured on B200, pillbox-256: curl-curl (three components, different offsets per component) 0.297 -> 0.262 ms
// with the interleave; the 7-point vector Laplacian (same offsets for every component, already contiguous) loses
// with it -- which is what the model predicts.
#pragma once
#include <algorithm>
#include <cstdint>

namespace mxg {

struct IlvCost {
  double pat[2] = {0, 0};   // [0] plain, [1] interleaved: lines per tile, paid once per column group
  double col[2] = {0, 0};   // lines per tile and column
  int tiles = 0;
};

inline int distinctCount(int64_t* v, int n) {
  if (n == 0) return 0;
  std::sort(v, v + n);
  return int(std::unique(v, v + n) - v);
}

// rowPat[r] < 0: not a dictionary row (skipped). delta(q) = column offset of pattern entry q relative to the row.
template <class Delta>
IlvCost ilvCostModel(const int32_t* rowPat, const int32_t* patOff, Delta delta, int64_t rowBegin, int64_t rowEnd, int xBytes,
                     int patEntryBytes, int maxTiles = 64) {
  IlvCost c;
  const int64_t span = rowEnd - rowBegin;
  if (span < 96) return c;
  const int64_t nTiles = span / 96;
  const int64_t step = std::max<int64_t>(1, nTiles / maxTiles);
  for (int64_t t = 0; t < nTiles && c.tiles < maxTiles; t += step) {
    const int64_t base = rowBegin + t * 96;
    for (int m = 0; m < 2; ++m)
      for (int w = 0; w < 3; ++w) {
        int64_t rows[32];
        int n = 0, maxLen = 0;
        for (int l = 0; l < 32; ++l) {
          const int64_t r = m == 0 ? base + 32 * w + l : base + 3 * l + w;
          if (rowPat[r] < 0) continue;
          rows[n++] = r;
          maxLen = std::max(maxLen, int(patOff[rowPat[r] + 1] - patOff[rowPat[r]]));
        }
        if (n == 0) continue;
        int64_t a[32];
        for (int i = 0; i < n; ++i) a[i] = rows[i] * 4 / 128;            // pattern-id loads
        c.pat[m] += distinctCount(a, n);
        for (int i = 0; i < n; ++i) a[i] = rows[i] * xBytes / 128;       // y stores
        c.col[m] += distinctCount(a, n);
        for (int k = 0; k < maxLen; ++k) {
          int64_t xl[32], pl[32];
          int cnt = 0;
          for (int i = 0; i < n; ++i) {
            const int32_t p = rowPat[rows[i]];
            const int32_t q = patOff[p] + k;
            if (q >= patOff[p + 1]) continue;
            const int64_t xAddr = (rows[i] + int64_t(delta(q))) * xBytes;
            xl[cnt] = xAddr >= 0 ? xAddr / 128 : -((-xAddr + 127) / 128);
            pl[cnt] = int64_t(q) * patEntryBytes / 128;
            ++cnt;
          }
          c.col[m] += distinctCount(xl, cnt);
          c.pat[m] += distinctCount(pl, cnt);
        }
      }
    ++c.tiles;
  }
  return c;
}

// true: use the interleaved assignment for this operator. The criterion is the single-column line count with a
// 2 % margin. B200 measurements behind it (pillbox-256): curl-curl, model ratio 0.94 -> interleaved is faster at 1, 4
// and 10 columns (0.262 vs 0.297 ms, 0.706 vs 0.757 ms, 1.74 vs 2.06 ms); vector Laplacian, model ratio 1.8 -> the
// multigrid-preconditioned eigensolve is 35 % slower with it (5.06 s vs 3.75 s).
inline double ilvRatio(const IlvCost& c) { return (c.pat[1] + c.col[1]) / (c.pat[0] + c.col[0]); }
inline bool ilvWins(const IlvCost& c) {
  if (c.tiles == 0 || c.pat[0] + c.col[0] <= 0) return false;
  return ilvRatio(c) < 0.98;
}

}  // namespace mxg
