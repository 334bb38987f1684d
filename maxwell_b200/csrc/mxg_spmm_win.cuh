// Windowed dictionary SpMM: the x-vector neighbourhood of a row tile is staged in shared memory with 1-D TMA
// (cp.async.bulk + mbarrier) and every gather of the stencil is served from there.
//
// Why: ncu on the round-1 kernel (one thread per row gathering x through L1, profiles/README_r01.md) showed the apply
// bound by l1tex data-pipe wavefronts (93 %), DRAM at 13 %: a warp-wide 8-byte gather with stride 24 B touches 6-7
// 128-byte lines. On a Yee grid the columns a tile of consecutive rows touches fall into (at most) three compact index
// ranges -- the tile's own neighbourhood (same x-plane: +-z, +-y lines) and the two neighbouring x-planes. Each range is
// ONE contiguous piece of the x column (DOFs are stored in ascending GID order, z fastest), so three bulk copies bring
// everything the tile needs into shared memory, and the stencil gathers become conflict-free shared loads
// (stride-3 doubles: 2 wavefronts per warp instead of 6-7). Reference call: MxCrsMatrix::apply (src/MxCrsMatrix.cpp:347-353).
//
// The arithmetic is untouched: per row, ascending column order, separately rounded multiply and add (Epetra order), so the
// result is bit-identical to the gather kernels and to the oracle.
//
// Tiles are planned on the host (planWinTiles): for each tile of R consecutive rows the distinct column offsets of its
// patterns are clustered into <= 3 groups (split at the two largest gaps), each group's [min column, max column] range
// becomes a window; overlapping windows share storage. A tile whose windows do not fit the shared-memory budget, or that
// touches ghost columns, is flagged and served by the gather path inside the same kernel.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <vector>

namespace mxg {

constexpr int kWinThreads = 384;   // 12 warps = 4 groups of 3 (component-interleaved assignment)

struct alignas(16) WinTile {
  int32_t segLo[3];    // first column (local index) of copy segment s, aligned to 16 bytes
  int32_t segLen[3];   // elements, multiple of 16 bytes; 0 = unused
  int32_t shift[3];    // shared index of column e in class k = e + shift[k]
  int32_t dLo, dHi;    // class of an entry with offset d: d < dLo -> 0, d > dHi -> 2, else 1
  int32_t valid;       // 1: windows usable, 0: gather path
  int32_t total;       // shared elements of all segments
  int32_t pad[3];
};
static_assert(sizeof(WinTile) == 64, "WinTile layout");

// rowPat[r] < 0: not a dictionary row. delta(q): column offset of pattern entry q relative to its row.
// align: elements per 16 bytes (2 for double, 1 for complex). Columns outside [0, nLoc) are ghosts -> tile not windowed.
template <class Delta>
inline std::vector<WinTile> planWinTiles(const int32_t* rowPat, const int32_t* patOff, int64_t numPats, Delta delta, int64_t nRows,
                                         int64_t nLoc, int R, int align, int64_t budgetElems, int64_t* maxTotal, int64_t* validTiles) {
  const int64_t nTiles = (nRows + R - 1) / R;
  std::vector<WinTile> tiles(static_cast<size_t>(nTiles));
  std::vector<int64_t> minRow(static_cast<size_t>(numPats), 0), maxRow(static_cast<size_t>(numPats), 0);
  std::vector<int64_t> stamp(static_cast<size_t>(numPats), -1);
  std::vector<int32_t> touched;
  struct Ent { int64_t d, lo, hi; };
  std::vector<Ent> ents;
  *maxTotal = 0;
  *validTiles = 0;
  for (int64_t t = 0; t < nTiles; ++t) {
    WinTile& W = tiles[size_t(t)];
    W = WinTile{};
    W.dLo = INT_MIN;
    W.dHi = INT_MAX;
    const int64_t r0 = t * R, r1 = std::min<int64_t>(nRows, r0 + R);
    touched.clear();
    for (int64_t r = r0; r < r1; ++r) {
      const int32_t p = rowPat[r];
      if (p < 0) continue;
      if (stamp[p] != t) { stamp[p] = t; minRow[p] = maxRow[p] = r; touched.push_back(p); }
      else maxRow[p] = r;   // rows ascend
    }
    if (touched.empty()) continue;
    ents.clear();
    for (int32_t p : touched)
      for (int32_t q = patOff[p]; q < patOff[p + 1]; ++q) {
        const int64_t d = delta(q);
        ents.push_back({d, minRow[p] + d, maxRow[p] + d + 1});
      }
    if (ents.empty()) continue;
    std::sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.d < b.d; });
    // two largest gaps between consecutive offsets -> three clusters
    int64_t g1 = 0, g2 = 0;
    size_t s1 = 0, s2 = 0;   // split AFTER index s (0 = none)
    for (size_t i = 0; i + 1 < ents.size(); ++i) {
      const int64_t g = ents[i + 1].d - ents[i].d;
      if (g > g1) { g2 = g1; s2 = s1; g1 = g; s1 = i + 1; }
      else if (g > g2) { g2 = g; s2 = i + 1; }
    }
    // a split is only worth a window of its own when the gap is large
    const int64_t minGap = 64;
    if (g2 < minGap) s2 = 0;
    if (g1 < minGap) { s1 = s2; s2 = 0; }
    size_t a = s1, b = s2;
    if (a && b && a > b) std::swap(a, b);
    if (!a) { a = b; b = 0; }
    // cluster index ranges: [0,a) [a,b) [b,n) with empty ones dropped; classes are assigned so that the kernel's test
    // (d < dLo -> 0, d > dHi -> 2, else 1) reproduces them
    const size_t n = ents.size();
    size_t cb[4];
    int ncl;
    if (a && b) { cb[0] = 0; cb[1] = a; cb[2] = b; cb[3] = n; ncl = 3; }
    else if (a) { cb[0] = 0; cb[1] = a; cb[2] = n; ncl = 2; }
    else { cb[0] = 0; cb[1] = n; ncl = 1; }
    int cls[3];
    if (ncl == 3) { cls[0] = 0; cls[1] = 1; cls[2] = 2; W.dLo = int32_t(ents[a].d); W.dHi = int32_t(ents[b - 1].d); }
    else if (ncl == 2) { cls[0] = 1; cls[1] = 2; W.dHi = int32_t(ents[a - 1].d); }
    else cls[0] = 1;
    // column range of every cluster, then the union of the ranges as copy segments (ranges of different clusters may
    // overlap or even be out of order when different patterns use them on different rows; overlapping ranges share storage)
    bool ok = true;
    int64_t lo[3], hi[3];
    for (int c = 0; c < ncl; ++c) {
      lo[c] = INT64_MAX;
      hi[c] = INT64_MIN;
      for (size_t i = cb[c]; i < cb[c + 1]; ++i) { lo[c] = std::min(lo[c], ents[i].lo); hi[c] = std::max(hi[c], ents[i].hi); }
      if (lo[c] < 0 || hi[c] > nLoc) ok = false;
      lo[c] = lo[c] / align * align;
      hi[c] = (hi[c] + align - 1) / align * align;
    }
    int order[3] = {0, 1, 2};
    std::sort(order, order + ncl, [&](int x, int y) { return lo[x] < lo[y]; });
    int64_t total = 0;
    int nseg = 0, segOf[3] = {0, 0, 0};
    int64_t segLo[3] = {0, 0, 0}, segHi[3] = {0, 0, 0}, segOff[3] = {0, 0, 0};
    for (int k = 0; k < ncl && ok; ++k) {
      const int c = order[k];
      if (nseg > 0 && lo[c] <= segHi[nseg - 1]) segHi[nseg - 1] = std::max(segHi[nseg - 1], hi[c]);
      else { segLo[nseg] = lo[c]; segHi[nseg] = hi[c]; ++nseg; }
      segOf[c] = nseg - 1;
    }
    for (int sgi = 0; sgi < nseg && ok; ++sgi) {
      segOff[sgi] = total;
      total += segHi[sgi] - segLo[sgi];
      W.segLo[sgi] = int32_t(segLo[sgi]);
      W.segLen[sgi] = int32_t(segHi[sgi] - segLo[sgi]);
    }
    for (int c = 0; c < ncl && ok; ++c) W.shift[cls[c]] = int32_t(segOff[segOf[c]] - segLo[segOf[c]]);
    if (!ok || total > budgetElems || total >= (int64_t(1) << 30)) {
      W = WinTile{};
      W.dLo = INT_MIN;
      W.dHi = INT_MAX;
      continue;
    }
    W.valid = 1;
    W.total = int32_t(total);
    *maxTotal = std::max<int64_t>(*maxTotal, total);
    ++*validTiles;
  }
  return tiles;
}

#ifdef __CUDACC__
// ---- PTX helpers: mbarrier + 1-D bulk (TMA) copies ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemU32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemU32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarFenceInit() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemU32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemU32(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smemU32(bar))
               : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smemU32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
#endif

}  // namespace mxg
