// Compiled and run by tests/test_cpp_shims.py. Exercises the reference-named C++ shims
// (include/mx/*.hpp) the way test/AnasaziInterface.cpp of the reference does: an MxMap of length 5, an
// MxAnasaziMV with 2 vectors, then the MultiVecTraits-style calls. Without a GPU the very first call must
// throw (no CPU fallback); with a GPU it runs the checks and prints PASSED.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mx/MxSolver.hpp"

template <class Scalar>
static int run() {
  std::shared_ptr<MxComm> comm(new MxComm(0));
  std::shared_ptr<MxMap> map(new MxMap(5, comm));                       // test/AnasaziInterface.cpp:44
  MxAnasaziMV<Scalar> ivec(map, 2);                                     // :47
  ivec.MvRandom();
  if (ivec.GetVecLength() != 5 || ivec.GetNumberVecs() != 2) return 1;
  std::vector<double> nrm;
  ivec.MvNorm(nrm);
  if (!(nrm[0] > 0 && nrm[1] > 0)) return 2;
  std::unique_ptr<mx::MultiVec<Scalar>> copy(ivec.CloneCopy());
  std::unique_ptr<mx::MultiVec<Scalar>> view(ivec.CloneViewNonConst(std::vector<int>{1}));
  view->MvInit(Scalar(3.0));
  ivec.MvNorm(nrm);
  if (std::fabs(nrm[1] - 3.0 * std::sqrt(5.0)) > 1e-12) return 3;       // writes through the view reach the parent
  mx::SerialDenseMatrix<int, Scalar> G(2, 2);
  ivec.MvTransMv(Scalar(1.0), *copy, G);                                // copy^H * ivec
  std::vector<Scalar> d;
  ivec.MvDot(*copy, d);
  if (std::abs(G(0, 0) - d[0]) > 1e-12 * std::abs(d[0])) return 4;
  mx::SerialDenseMatrix<int, Scalar> B(2, 2);
  B(0, 0) = Scalar(1.0); B(1, 1) = Scalar(1.0);
  std::unique_ptr<mx::MultiVec<Scalar>> z(ivec.Clone(2));
  z->MvTimesMatAddMv(Scalar(1.0), ivec, B, Scalar(0.0));                // z = ivec * I
  z->MvAddMv(Scalar(1.0), *z, Scalar(-1.0), ivec);
  z->MvNorm(nrm);
  if (nrm[0] != 0.0 || nrm[1] != 0.0) return 5;
  // MxCrsMatrix assembly through insertRowValues / fillComplete, then apply
  MxCrsMatrix<Scalar> A(map);
  for (MxIndex g = 0; g < 5; ++g) {
    std::vector<MxIndex> cols{g, (g + 1) % 5};
    std::vector<Scalar> vals{Scalar(2.0), Scalar(-1.0)};
    A.insertRowValues(g, cols, vals);
  }
  A.fillComplete(map, map);
  MxMultiVector<Scalar> ones(map, 1), y(map, 1);
  ones.set(Scalar(1.0));
  A.apply(ones, y);
  y.norm2(nrm);
  if (std::fabs(nrm[0] - std::sqrt(5.0)) > 1e-14) return 6;
  // successive MvRandom calls -- also on a Clone()d block -- draw independent numbers (Epetra's Random() advances its state)
  MxAnasaziMV<Scalar> r1(map, 2);
  r1.MvRandom();
  std::unique_ptr<mx::MultiVec<Scalar>> r2(r1.Clone(2));
  r2->MvRandom();
  std::unique_ptr<mx::MultiVec<Scalar>> r1copy(r1.CloneCopy());
  r1.MvRandom();
  for (mx::MultiVec<Scalar>* other : {r2.get(), r1copy.get()}) {
    std::unique_ptr<mx::MultiVec<Scalar>> diff(r1.CloneCopy());
    diff->MvAddMv(Scalar(1.0), r1, Scalar(-1.0), *other);
    diff->MvNorm(nrm);
    if (!(nrm[0] > 1e-3 && nrm[1] > 1e-3)) return 7;
  }
  return 0;
}

int main() {
  try {
    const int r = run<double>();
    const int c = run<MxComplex>();
    if (r || c) { std::printf("FAILED real=%d complex=%d\n", r, c); return 1; }
    std::printf("PASSED\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("EXCEPTION: %s\n", e.what());
    return 3;
  }
}
