// Host-only building block for a component-major device ordering of multivectors.
// NOT yet used by libmxgpu: the line-count model (mxg_ilv_model.h) predicts 0.54x the L1 lines of today's order for the
// curl-curl SpMM when consecutive lanes read consecutive cells of ONE field component, which needs the vectors themselves
// stored component-major on the device. These routines are the host half of that change, validated on the CPU
// (tests/test_ilv_model.py): the permutation, its inverse, and the CSR re-indexing that keeps every row's entry
// order -- ascending REFERENCE local column, Epetra's summation order -- so results stay bit-identical.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace mxg {

// perm[devicePosition] = reference local index. Groups the DOFs by component (GID mod ncomp, the reference's
// globCompIndx = comp + ncomp * cell, MxGridField.hpp:142-145), ascending GID (= ascending cell) inside a group.
inline std::vector<int32_t> componentMajorOrder(const int64_t* gids, int64_t n, int ncomp) {
  if (ncomp < 1) throw std::invalid_argument("componentMajorOrder: ncomp must be >= 1");
  std::vector<int32_t> perm(static_cast<size_t>(n));
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return gids[a] % ncomp < gids[b] % ncomp; });
  return perm;
}

inline std::vector<int32_t> inversePermutation(const std::vector<int32_t>& perm) {
  std::vector<int32_t> inv(perm.size());
  for (size_t i = 0; i < perm.size(); ++i) inv[static_cast<size_t>(perm[i])] = int32_t(i);
  return inv;
}

// Rows re-ordered by rowPerm (device row i = reference row rowPerm[i]); owned columns (0 <= c < nLoc) translated with
// colInv (reference local column -> device position); ghost columns (c < 0 or c >= nLoc) are left as they are.
template <class T>
void permuteCsr(const std::vector<int64_t>& rowptr, const std::vector<int32_t>& col, const std::vector<T>& val,
                const std::vector<int32_t>& rowPerm, const std::vector<int32_t>& colInv, int64_t nLoc,
                std::vector<int64_t>& outRowptr, std::vector<int32_t>& outCol, std::vector<T>& outVal) {
  const int64_t nRows = int64_t(rowPerm.size());
  if (int64_t(rowptr.size()) != nRows + 1) throw std::invalid_argument("permuteCsr: rowptr / permutation size mismatch");
  outRowptr.assign(size_t(nRows) + 1, 0);
  for (int64_t i = 0; i < nRows; ++i) {
    const int64_t r = rowPerm[size_t(i)];
    outRowptr[size_t(i) + 1] = outRowptr[size_t(i)] + (rowptr[size_t(r) + 1] - rowptr[size_t(r)]);
  }
  outCol.resize(size_t(outRowptr[size_t(nRows)]));
  outVal.resize(outCol.size());
  for (int64_t i = 0; i < nRows; ++i) {
    const int64_t r = rowPerm[size_t(i)];
    int64_t o = outRowptr[size_t(i)];
    for (int64_t k = rowptr[size_t(r)]; k < rowptr[size_t(r) + 1]; ++k, ++o) {
      const int32_t c = col[size_t(k)];
      outCol[size_t(o)] = (c >= 0 && c < nLoc) ? colInv[size_t(c)] : c;
      outVal[size_t(o)] = val[size_t(k)];
    }
  }
}

}  // namespace mxg
