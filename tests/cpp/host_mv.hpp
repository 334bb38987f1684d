// Plain host implementation of the mx::MultiVec surface (what MxAnasaziMV forwards to the GPU), used to run the
// templated eigensolver driver of include/mx/MxSolver.hpp in the CPU test-suite. Test infrastructure only.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <vector>

#include "mx/MxTypes.hpp"

namespace hostmv {

struct Comm {
  void sync() const {}
};
struct Map {
  explicit Map(int64_t n) : n(n), comm(new Comm) {}
  std::shared_ptr<Comm> getComm() const { return comm; }
  int64_t n;
  std::shared_ptr<Comm> comm;
};

template <class S>
class HostMVT : public mx::MultiVec<S> {
  typedef mx::ScalarTraits<S> ST;

 public:
  HostMVT(std::shared_ptr<Map> map, size_t numVecs)
      : map_(map), data_(new std::vector<S>(size_t(map->n) * numVecs, S(0.0))), cols_(numVecs) {
    for (size_t j = 0; j < numVecs; ++j) cols_[j] = int(j);
  }
  // view
  HostMVT(const HostMVT& parent, const std::vector<int>& index, bool) : map_(parent.map_), data_(parent.data_) {
    for (int j : index) {
      if (j < 0 || j >= int(parent.cols_.size())) throw std::runtime_error("HostMV: view index out of range");
      cols_.push_back(parent.cols_[j]);
    }
  }
  // deep copy
  HostMVT(const HostMVT& o) : map_(o.map_), data_(new std::vector<S>(size_t(o.map_->n) * o.cols_.size())), cols_(o.cols_.size()) {
    for (size_t j = 0; j < cols_.size(); ++j) {
      cols_[j] = int(j);
      std::copy(o.col(j), o.col(j) + n(), col(j));
    }
  }
  HostMVT& operator=(const HostMVT& o) {   // value assignment into existing storage (views write through)
    if (o.cols_.size() != cols_.size() || o.n() != n()) throw std::runtime_error("HostMV: assignment shape mismatch");
    if (&o == this) return *this;
    std::vector<S> tmp(size_t(n()) * cols_.size());
    for (size_t j = 0; j < cols_.size(); ++j) std::copy(o.col(j), o.col(j) + n(), tmp.begin() + j * n());
    for (size_t j = 0; j < cols_.size(); ++j) std::copy(tmp.begin() + j * n(), tmp.begin() + (j + 1) * n(), col(j));
    return *this;
  }
  std::shared_ptr<Map> getMap() const { return map_; }
  static constexpr bool kTimesMatInPlace = false;
  void setSeed(uint64_t s) { seed_ = s; }
  void swap(HostMVT& o) {
    std::swap(map_, o.map_);
    std::swap(data_, o.data_);
    std::swap(cols_, o.cols_);
  }
  int64_t n() const { return map_->n; }
  S* col(size_t j) { return data_->data() + size_t(cols_[j]) * n(); }
  const S* col(size_t j) const { return data_->data() + size_t(cols_[j]) * n(); }

  mx::MultiVec<S>* Clone(const int numVecs) const override { return new HostMVT(map_, size_t(numVecs)); }
  mx::MultiVec<S>* CloneCopy() const override { return new HostMVT(*this); }
  mx::MultiVec<S>* CloneCopy(const std::vector<int>& index) const override {
    HostMVT v(*this, index, false);
    return new HostMVT(v);
  }
  const mx::MultiVec<S>* CloneView(const std::vector<int>& index) const override { return new HostMVT(*this, index, false); }
  mx::MultiVec<S>* CloneViewNonConst(const std::vector<int>& index) override { return new HostMVT(*this, index, false); }
  int GetVecLength() const override { return int(n()); }
  int GetNumberVecs() const override { return int(cols_.size()); }
  void MvTimesMatAddMv(S alpha, const mx::MultiVec<S>& A_, const mx::SerialDenseMatrix<int, S>& B, S beta) override {
    const HostMVT& A = dynamic_cast<const HostMVT&>(A_);
    if (B.numRows() != A.GetNumberVecs() || B.numCols() != GetNumberVecs()) throw std::runtime_error("HostMV: MvTimesMatAddMv shapes");
    std::vector<S> out(size_t(n()) * cols_.size(), S(0.0));   // A may alias this
    for (int j = 0; j < GetNumberVecs(); ++j)
      for (int k = 0; k < A.GetNumberVecs(); ++k) {
        const S b = alpha * B(k, j);
        if (b == S(0.0)) continue;
        const S* a = A.col(k);
        S* o = out.data() + size_t(j) * n();
        for (int64_t i = 0; i < n(); ++i) o[i] += a[i] * b;
      }
    for (int j = 0; j < GetNumberVecs(); ++j) {
      S* y = col(j);
      const S* o = out.data() + size_t(j) * n();
      for (int64_t i = 0; i < n(); ++i) y[i] = o[i] + (beta == S(0.0) ? S(0.0) : beta * y[i]);
    }
  }
  void MvAddMv(S alpha, const mx::MultiVec<S>& A_, S beta, const mx::MultiVec<S>& B_) override {
    const HostMVT& A = dynamic_cast<const HostMVT&>(A_);
    const HostMVT& B = dynamic_cast<const HostMVT&>(B_);
    for (int j = 0; j < GetNumberVecs(); ++j) {
      const S *a = A.col(j), *b = B.col(j);
      S* y = col(j);
      for (int64_t i = 0; i < n(); ++i) y[i] = alpha * a[i] + beta * b[i];
    }
  }
  void MvTransMv(S alpha, const mx::MultiVec<S>& A_, mx::SerialDenseMatrix<int, S>& B) const override {   // B = alpha A^H this
    const HostMVT& A = dynamic_cast<const HostMVT&>(A_);
    if (B.numRows() != A.GetNumberVecs() || B.numCols() != GetNumberVecs()) throw std::runtime_error("HostMV: MvTransMv shapes");
    for (int j = 0; j < GetNumberVecs(); ++j)
      for (int k = 0; k < A.GetNumberVecs(); ++k) {
        const S *a = A.col(k), *x = col(j);
        S s = S(0.0);
        for (int64_t i = 0; i < n(); ++i) s += ST::conj(a[i]) * x[i];
        B(k, j) = alpha * s;
      }
  }
  void MvDot(const mx::MultiVec<S>& A_, std::vector<S>& b) const override {
    const HostMVT& A = dynamic_cast<const HostMVT&>(A_);
    b.resize(cols_.size());
    for (size_t j = 0; j < cols_.size(); ++j) {
      S s = S(0.0);
      for (int64_t i = 0; i < n(); ++i) s += ST::conj(A.col(j)[i]) * col(j)[i];
      b[j] = s;
    }
  }
  void MvNorm(std::vector<double>& normvec) const override {
    normvec.resize(cols_.size());
    for (size_t j = 0; j < cols_.size(); ++j) {
      double s = 0.0;
      for (int64_t i = 0; i < n(); ++i) s += std::norm(col(j)[i]);
      normvec[j] = std::sqrt(s);
    }
  }
  void SetBlock(const mx::MultiVec<S>& A_, const std::vector<int>& index) override {
    const HostMVT& A = dynamic_cast<const HostMVT&>(A_);
    for (size_t k = 0; k < index.size(); ++k) std::copy(A.col(k), A.col(k) + n(), col(size_t(index[k])));
  }
  void MvScale(S alpha) override {
    for (size_t j = 0; j < cols_.size(); ++j)
      for (int64_t i = 0; i < n(); ++i) col(j)[i] *= alpha;
  }
  void MvScale(const std::vector<S>& alpha) override {
    for (size_t j = 0; j < cols_.size(); ++j)
      for (int64_t i = 0; i < n(); ++i) col(j)[i] *= alpha[j];
  }
  void MvRandom() override {   // splitmix64 keyed by (seed, row, column): reproducible
    for (size_t j = 0; j < cols_.size(); ++j)
      for (int64_t i = 0; i < n(); ++i) {
        uint64_t z = seed_ + 0x9E3779B97F4A7C15ull * (uint64_t(i) * 1315423911ull + j + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const double re = double(z >> 11) / double(1ull << 52) - 1.0;
        z = (z ^ (z >> 29)) * 0xD6E8FEB86659FD93ull;
        const double im = double(z >> 11) / double(1ull << 52) - 1.0;
        double parts[2] = {re, ST::isComplex ? im : 0.0};
        col(j)[i] = ST::unpack(parts);
      }
  }
  void MvInit(S alpha) override {
    for (size_t j = 0; j < cols_.size(); ++j) std::fill(col(j), col(j) + n(), alpha);
  }
  void MvPrint(std::ostream& os) const override { os << "HostMV " << n() << " x " << cols_.size() << "\n"; }

 private:
  std::shared_ptr<Map> map_;
  std::shared_ptr<std::vector<S>> data_;
  std::vector<int> cols_;
  uint64_t seed_ = 1;
};

typedef HostMVT<double> HostMV;
typedef HostMVT<std::complex<double>> HostMVC;

}  // namespace hostmv
