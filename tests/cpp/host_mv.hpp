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

class HostMV : public mx::MultiVec<double> {
 public:
  HostMV(std::shared_ptr<Map> map, size_t numVecs)
      : map_(map), data_(new std::vector<double>(size_t(map->n) * numVecs, 0.0)), cols_(numVecs) {
    for (size_t j = 0; j < numVecs; ++j) cols_[j] = int(j);
  }
  // view
  HostMV(const HostMV& parent, const std::vector<int>& index, bool) : map_(parent.map_), data_(parent.data_) {
    for (int j : index) {
      if (j < 0 || j >= int(parent.cols_.size())) throw std::runtime_error("HostMV: view index out of range");
      cols_.push_back(parent.cols_[j]);
    }
  }
  // deep copy
  HostMV(const HostMV& o) : map_(o.map_), data_(new std::vector<double>(size_t(o.map_->n) * o.cols_.size())), cols_(o.cols_.size()) {
    for (size_t j = 0; j < cols_.size(); ++j) {
      cols_[j] = int(j);
      std::copy(o.col(j), o.col(j) + n(), col(j));
    }
  }
  HostMV& operator=(const HostMV& o) {   // value assignment into existing storage (views write through)
    if (o.cols_.size() != cols_.size() || o.n() != n()) throw std::runtime_error("HostMV: assignment shape mismatch");
    if (&o == this) return *this;
    std::vector<double> tmp(size_t(n()) * cols_.size());
    for (size_t j = 0; j < cols_.size(); ++j) std::copy(o.col(j), o.col(j) + n(), tmp.begin() + j * n());
    for (size_t j = 0; j < cols_.size(); ++j) std::copy(tmp.begin() + j * n(), tmp.begin() + (j + 1) * n(), col(j));
    return *this;
  }
  std::shared_ptr<Map> getMap() const { return map_; }
  void setSeed(uint64_t s) { seed_ = s; }
  void swap(HostMV& o) {
    std::swap(map_, o.map_);
    std::swap(data_, o.data_);
    std::swap(cols_, o.cols_);
  }
  int64_t n() const { return map_->n; }
  double* col(size_t j) { return data_->data() + size_t(cols_[j]) * n(); }
  const double* col(size_t j) const { return data_->data() + size_t(cols_[j]) * n(); }

  mx::MultiVec<double>* Clone(const int numVecs) const override { return new HostMV(map_, size_t(numVecs)); }
  mx::MultiVec<double>* CloneCopy() const override { return new HostMV(*this); }
  mx::MultiVec<double>* CloneCopy(const std::vector<int>& index) const override {
    HostMV v(*this, index, false);
    return new HostMV(v);
  }
  const mx::MultiVec<double>* CloneView(const std::vector<int>& index) const override { return new HostMV(*this, index, false); }
  mx::MultiVec<double>* CloneViewNonConst(const std::vector<int>& index) override { return new HostMV(*this, index, false); }
  int GetVecLength() const override { return int(n()); }
  int GetNumberVecs() const override { return int(cols_.size()); }
  void MvTimesMatAddMv(double alpha, const mx::MultiVec<double>& A_, const mx::SerialDenseMatrix<int, double>& B, double beta) override {
    const HostMV& A = dynamic_cast<const HostMV&>(A_);
    if (B.numRows() != A.GetNumberVecs() || B.numCols() != GetNumberVecs()) throw std::runtime_error("HostMV: MvTimesMatAddMv shapes");
    std::vector<double> out(size_t(n()) * cols_.size(), 0.0);   // A may alias this
    for (int j = 0; j < GetNumberVecs(); ++j)
      for (int k = 0; k < A.GetNumberVecs(); ++k) {
        const double b = alpha * B(k, j);
        if (b == 0.0) continue;
        const double* a = A.col(k);
        double* o = out.data() + size_t(j) * n();
        for (int64_t i = 0; i < n(); ++i) o[i] += a[i] * b;
      }
    for (int j = 0; j < GetNumberVecs(); ++j) {
      double* y = col(j);
      const double* o = out.data() + size_t(j) * n();
      for (int64_t i = 0; i < n(); ++i) y[i] = o[i] + (beta == 0.0 ? 0.0 : beta * y[i]);
    }
  }
  void MvAddMv(double alpha, const mx::MultiVec<double>& A_, double beta, const mx::MultiVec<double>& B_) override {
    const HostMV& A = dynamic_cast<const HostMV&>(A_);
    const HostMV& B = dynamic_cast<const HostMV&>(B_);
    for (int j = 0; j < GetNumberVecs(); ++j) {
      const double *a = A.col(j), *b = B.col(j);
      double* y = col(j);
      for (int64_t i = 0; i < n(); ++i) y[i] = alpha * a[i] + beta * b[i];
    }
  }
  void MvTransMv(double alpha, const mx::MultiVec<double>& A_, mx::SerialDenseMatrix<int, double>& B) const override {
    const HostMV& A = dynamic_cast<const HostMV&>(A_);
    if (B.numRows() != A.GetNumberVecs() || B.numCols() != GetNumberVecs()) throw std::runtime_error("HostMV: MvTransMv shapes");
    for (int j = 0; j < GetNumberVecs(); ++j)
      for (int k = 0; k < A.GetNumberVecs(); ++k) {
        const double *a = A.col(k), *x = col(j);
        double s = 0.0;
        for (int64_t i = 0; i < n(); ++i) s += a[i] * x[i];
        B(k, j) = alpha * s;
      }
  }
  void MvDot(const mx::MultiVec<double>& A_, std::vector<double>& b) const override {
    const HostMV& A = dynamic_cast<const HostMV&>(A_);
    b.resize(cols_.size());
    for (size_t j = 0; j < cols_.size(); ++j) {
      double s = 0.0;
      for (int64_t i = 0; i < n(); ++i) s += A.col(j)[i] * col(j)[i];
      b[j] = s;
    }
  }
  void MvNorm(std::vector<double>& normvec) const override {
    normvec.resize(cols_.size());
    for (size_t j = 0; j < cols_.size(); ++j) {
      double s = 0.0;
      for (int64_t i = 0; i < n(); ++i) s += col(j)[i] * col(j)[i];
      normvec[j] = std::sqrt(s);
    }
  }
  void SetBlock(const mx::MultiVec<double>& A_, const std::vector<int>& index) override {
    const HostMV& A = dynamic_cast<const HostMV&>(A_);
    for (size_t k = 0; k < index.size(); ++k) std::copy(A.col(k), A.col(k) + n(), col(size_t(index[k])));
  }
  void MvScale(double alpha) override {
    for (size_t j = 0; j < cols_.size(); ++j)
      for (int64_t i = 0; i < n(); ++i) col(j)[i] *= alpha;
  }
  void MvScale(const std::vector<double>& alpha) override {
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
        col(j)[i] = double(z >> 11) / double(1ull << 52) - 1.0;
      }
  }
  void MvInit(double alpha) override {
    for (size_t j = 0; j < cols_.size(); ++j) std::fill(col(j), col(j) + n(), alpha);
  }
  void MvPrint(std::ostream& os) const override { os << "HostMV " << n() << " x " << cols_.size() << "\n"; }

 private:
  std::shared_ptr<Map> map_;
  std::shared_ptr<std::vector<double>> data_;
  std::vector<int> cols_;
  uint64_t seed_ = 1;
};

}  // namespace hostmv
