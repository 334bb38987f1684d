// CPU check of the build-time cost model that chooses the dictionary kernel's thread -> row assignment
// (maxwell_b200/csrc/mxg_ilv_model.h): synthetic pattern tables shaped like the operators on the path.
#include <cstdio>
#include <vector>

#include "mxg_ilv_model.h"

static int failures = 0;
#define CHECK(cond, msg)                                                   \
  do {                                                                     \
    if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, msg); ++failures; } \
  } while (0)

struct Table {
  std::vector<int32_t> rowPat, patOff, d;
};

// rows = 3 components x cells, GID = comp + 3 cell, cell z-fastest with plane strides sy, sx (in cells)
static Table make(int64_t cells, int sy, int sx, bool curlLike) {
  Table t;
  t.patOff.push_back(0);
  for (int c = 0; c < 3; ++c) {
    std::vector<int32_t> off;
    if (curlLike) {
      // curl-curl row of component c: 5 own-component entries and 4 + 4 entries of the two other components,
      // whose cell offsets depend on c (13 entries)
      const int cellStride[3] = {sx, sy, 1};
      const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
      for (int s : {-cellStride[c1], -cellStride[c2], 0, cellStride[c2], cellStride[c1]}) off.push_back(3 * s);
      for (int oc : {c1, c2})
        for (int a : {0, -1})
          for (int b : {0, 1}) off.push_back(3 * (a * cellStride[oc] + b * cellStride[c]) + (oc - c));
      std::sort(off.begin(), off.end());
    } else {
      for (int s : {-sx, -sy, -1, 0, 1, sy, sx}) off.push_back(3 * s);   // 7-point Laplacian, same for every component
    }
    for (int32_t o : off) t.d.push_back(o);
    t.patOff.push_back(int32_t(t.d.size()));
  }
  t.rowPat.resize(3 * cells);
  for (int64_t r = 0; r < 3 * cells; ++r) t.rowPat[r] = int32_t(r % 3);
  return t;
}

int main() {
  using namespace mxg;
  const int sy = 257, sx = 257 * 257;
  const int64_t cells = 40000;
  {
    Table t = make(cells, sy, sx, true);
    const int32_t* d = t.d.data();
    IlvCost c = ilvCostModel(t.rowPat.data(), t.patOff.data(), [d](int32_t q) { return d[q]; }, 0, 3 * cells, 8, 16);
    std::printf("curl-like : pat %.0f / %.0f, col %.0f / %.0f over %d tiles\n", c.pat[0], c.pat[1], c.col[0], c.col[1], c.tiles);
    CHECK(c.tiles > 0, "no tiles sampled");
    CHECK(c.pat[1] < 0.5 * c.pat[0], "interleaving must cut the pattern-table lines of a 3-component operator");
    CHECK(ilvWins(c), "curl-curl: interleaved assignment expected");
  }
  {
    Table t = make(cells, sy, sx, false);
    const int32_t* d = t.d.data();
    IlvCost c = ilvCostModel(t.rowPat.data(), t.patOff.data(), [d](int32_t q) { return d[q]; }, 0, 3 * cells, 8, 16);
    std::printf("laplacian : pat %.0f / %.0f, col %.0f / %.0f over %d tiles\n", c.pat[0], c.pat[1], c.col[0], c.col[1], c.tiles);
    CHECK(c.col[1] > 1.5 * c.col[0], "same-offset stencil: contiguous gathers become strided when interleaved");
    CHECK(!ilvWins(c) && ilvRatio(c) > 1.3, "vector Laplacian: plain assignment expected");
  }
  {
    // one-component operator (scalar Laplacian): all rows share a pattern, interleaving only strides the accesses
    Table t;
    t.patOff = {0, 7};
    for (int s : {-sx, -sy, -1, 0, 1, sy, sx}) t.d.push_back(s);
    t.rowPat.assign(cells, 0);
    const int32_t* d = t.d.data();
    IlvCost c = ilvCostModel(t.rowPat.data(), t.patOff.data(), [d](int32_t q) { return d[q]; }, 0, cells, 8, 16);
    CHECK(!ilvWins(c), "scalar operator: plain assignment expected");
  }
  {
    // degenerate inputs: fewer than one tile, and tiles without dictionary rows
    std::vector<int32_t> rp(50, 0), po = {0, 1};
    IlvCost c = ilvCostModel(rp.data(), po.data(), [](int32_t) { return 0; }, 0, 50, 8, 16);
    CHECK(c.tiles == 0 && !ilvWins(c), "short range must fall back to the plain assignment");
    std::vector<int32_t> none(960, -1);
    c = ilvCostModel(none.data(), po.data(), [](int32_t) { return 0; }, 0, 960, 8, 16);
    CHECK(!ilvWins(c), "no dictionary rows: plain assignment");
    // negative addresses (ghost columns below the first owned row)
    std::vector<int32_t> rp2(960, 0);
    c = ilvCostModel(rp2.data(), po.data(), [](int32_t) { return -5000; }, 0, 960, 16, 24);
    CHECK(c.tiles > 0, "ghost offsets");
  }
  if (!failures) std::printf("PASSED\n");
  return failures ? 1 : 0;
}
