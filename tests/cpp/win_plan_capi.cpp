// extern "C" wrapper around the host-side tile planner of the windowed SpMM (maxwell_b200/csrc/mxg_spmm_win.cuh) that ALSO
// replays the kernel's shared-memory addressing on the CPU: every tile's segments are "copied" from an x vector holding
// its own indices, and each (row, pattern entry) must find column row + d at shared index row + d + shift[class].
#include "mxg_spmm_win.cuh"

extern "C" int win_plan_eval(const int32_t* rowPat, const int32_t* patOff, int64_t numPats, const int32_t* delta, int64_t nRows,
                             int64_t nLoc, int R, int align, int64_t budgetElems, int64_t out[6]) {
  int64_t maxTotal = 0, valid = 0;
  std::vector<mxg::WinTile> tiles =
      mxg::planWinTiles(rowPat, patOff, numPats, [delta](int32_t q) { return int64_t(delta[q]); }, nRows, nLoc, R, align, budgetElems, &maxTotal, &valid);
  int64_t bad = 0, rowsWindowed = 0, copied = 0, misaligned = 0;
  std::vector<int64_t> buf;
  for (size_t t = 0; t < tiles.size(); ++t) {
    const mxg::WinTile& W = tiles[t];
    if (!W.valid) continue;
    buf.assign(size_t(W.total), -1);
    int64_t off = 0;
    for (int s = 0; s < 3; ++s) {
      if (W.segLen[s] % align || W.segLo[s] % align) ++misaligned;
      if (W.segLo[s] < 0 || W.segLo[s] + W.segLen[s] > (nLoc + align - 1) / align * align) ++bad;
      for (int32_t i = 0; i < W.segLen[s]; ++i) buf[size_t(off + i)] = W.segLo[s] + i;
      off += W.segLen[s];
      copied += W.segLen[s];
    }
    if (off != W.total) ++bad;
    const int64_t r0 = int64_t(t) * R, r1 = std::min<int64_t>(nRows, r0 + R);
    for (int64_t r = r0; r < r1; ++r) {
      const int32_t p = rowPat[r];
      if (p < 0) continue;
      ++rowsWindowed;
      for (int32_t q = patOff[p]; q < patOff[p + 1]; ++q) {
        const int32_t d = delta[q];
        const int32_t sh = d < W.dLo ? W.shift[0] : (d > W.dHi ? W.shift[2] : W.shift[1]);
        const int64_t idx = r + d + sh;
        if (idx < 0 || idx >= W.total || buf[size_t(idx)] != r + d) ++bad;
      }
    }
  }
  out[0] = int64_t(tiles.size()); out[1] = valid; out[2] = maxTotal; out[3] = bad + misaligned; out[4] = rowsWindowed; out[5] = copied;
  return 0;
}
