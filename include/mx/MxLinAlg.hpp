// MxComm / MxMap / MxMultiVector / MxVector / MxAnasaziMV / MxCrsMatrix / MxOperator with the
// reference's method names (src/MxComm.hpp, MxMap.hpp, MxMultiVector.hpp, MxVector.hpp,
// MxAnasaziMV.hpp, MxCrsMatrix.hpp, MxOperator.hpp), forwarding to libmxgpu.
#pragma once
#include <algorithm>
#include <map>
#include <ostream>

#include "MxTypes.hpp"

// ---- MxComm (src/MxComm.hpp:14-37): one GPU + NCCL communicator instead of Epetra_MpiComm ----
class MxComm {
 public:
  explicit MxComm(int device = 0) { mx::check(mxg_ctx_create(device, &ctx_)); }
  // non-owning wrapper around an existing context
  MxComm(mxg_ctx* ctx, bool own) : ctx_(ctx), own_(own) {}
  ~MxComm() { if (own_ && ctx_) mxg_ctx_destroy(ctx_); }
  MxComm(const MxComm&) = delete;
  MxComm& operator=(const MxComm&) = delete;
  int myPID() const { return mxg_ctx_rank(ctx_); }
  size_t numProc() const { return size_t(mxg_ctx_num_ranks(ctx_)); }
  void commInit(int rank, int nranks, const void* uniqueId) { mx::check(mxg_ctx_comm_init(ctx_, rank, nranks, uniqueId)); }
  void sync() const { mx::check(mxg_ctx_sync(ctx_)); }
  mxg_ctx* raw() const { return ctx_; }

 private:
  mxg_ctx* ctx_ = nullptr;
  bool own_ = true;
};

// ---- MxMap (src/MxMap.hpp:22-103) --------------------------------------------------------------
class MxMap {
 public:
  // MxMap(numGlobalIndices, myGlobalIndices, comm) (MxMap.hpp:33-34)
  MxMap(size_t numGlobalIndices, const std::vector<MxIndex>& myGlobalIndices, std::shared_ptr<MxComm> comm)
      : comm_(comm), gids_(myGlobalIndices) {
    mx::check(mxg_map_create(comm->raw(), int64_t(numGlobalIndices), gids_.data(), int64_t(gids_.size()), &map_));
  }
  // linear map (MxMap.hpp:25): this rank owns a contiguous block
  MxMap(size_t numGlobalIndices, std::shared_ptr<MxComm> comm) : comm_(comm) {
    const int64_t P = int64_t(comm->numProc()), r = comm->myPID(), n = int64_t(numGlobalIndices);
    const int64_t lo = n * r / P, hi = n * (r + 1) / P;
    for (int64_t g = lo; g < hi; ++g) gids_.push_back(g);
    mx::check(mxg_map_create(comm->raw(), n, gids_.data(), int64_t(gids_.size()), &map_));
  }
  MxMap(mxg_map* map, std::shared_ptr<MxComm> comm, bool own) : comm_(comm), map_(map), own_(own) {}
  ~MxMap() { if (own_ && map_) mxg_map_destroy(map_); }
  MxMap(const MxMap&) = delete;
  MxMap& operator=(const MxMap&) = delete;
  size_t getNodeNumIndices() const { return size_t(mxg_map_local_size(map_)); }
  MxIndex getGlobalNumIndices() const { return mxg_map_global_size(map_); }
  const MxIndex* getNodeIndexList() const { return gids_.data(); }
  MxIndex getGlobalIndex(MxIndex localIndex) const { return gids_[size_t(localIndex)]; }
  MxIndex getLocalIndex(MxIndex globIndex) const {
    auto it = std::lower_bound(gids_.begin(), gids_.end(), globIndex);
    return (it != gids_.end() && *it == globIndex) ? MxIndex(it - gids_.begin()) : MxIndex(-1);
  }
  bool isNodeGlobalIndex(MxIndex g) const { return getLocalIndex(g) >= 0; }
  std::shared_ptr<MxComm> getComm() const { return comm_; }
  bool operator==(const MxMap& o) const { return getGlobalNumIndices() == o.getGlobalNumIndices() && gids_ == o.gids_; }
  bool operator!=(const MxMap& o) const { return !(*this == o); }
  mxg_map* raw() const { return map_; }

 private:
  std::shared_ptr<MxComm> comm_;
  std::vector<MxIndex> gids_;
  mxg_map* map_ = nullptr;
  bool own_ = true;
};

template <class Scalar> class MxCrsMatrix;
template <class Scalar> class MxVector;

// ---- MxMultiVector (src/MxMultiVector.hpp:16-89) ------------------------------------------------
template <class Scalar>
class MxMultiVector {
  typedef mx::ScalarTraits<Scalar> ST;

 public:
  MxMultiVector(std::shared_ptr<MxMap> map, size_t numVecs) : map_(map) {
    mx::check(mxg_mv_create(map->raw(), int(numVecs), ST::isComplex, &mv_));
  }
  // copy constructor = deep copy (MxMultiVector.cpp:48-58)
  MxMultiVector(const MxMultiVector<Scalar>& mv) : map_(mv.map_) { mx::check(mxg_mv_clone_copy(mv.mv_, nullptr, 0, &mv_)); }
  // view (deepcopy == false) or copy of selected columns (MxMultiVector.cpp:29-44)
  MxMultiVector(const MxMultiVector<Scalar>& mv, const std::vector<size_t>& vecInds, bool deepcopy) : map_(mv.map_) {
    std::vector<int> idx(vecInds.begin(), vecInds.end());
    if (deepcopy) mx::check(mxg_mv_clone_copy(mv.mv_, idx.data(), int(idx.size()), &mv_));
    else mx::check(mxg_mv_view(mv.mv_, idx.data(), int(idx.size()), &mv_));
  }
  // adopt a raw handle
  MxMultiVector(mxg_mv* raw, std::shared_ptr<MxMap> map, bool own) : map_(map), mv_(raw), own_(own) {}
  virtual ~MxMultiVector() { if (own_ && mv_) mxg_mv_destroy(mv_); }

  size_t getNumVecs() const { return size_t(mxg_mv_num_cols(mv_)); }
  MxIndex getLocalLength() const { return mxg_mv_local_length(mv_); }
  std::shared_ptr<MxMap> getMap() const { return map_; }

  void conj() { mx::check(mxg_mv_conj(mv_)); }
  void scale(Scalar val) { double a[2]; ST::pack(val, a); mx::check(mxg_mv_scale(mv_, a)); }
  void scale(const std::vector<Scalar>& vals) {
    if (vals.size() != getNumVecs()) throw std::runtime_error("MxMultiVector::scale: one scalar per column expected");
    mx::check(mxg_mv_scale_cols(mv_, reinterpret_cast<const double*>(vals.data())));
  }
  // uniform (-1,1) keyed by (seed, call epoch of the context, global DOF id, column): every call draws new numbers
  // (Epetra's Random() advances its state), identical on every rank count
  virtual void random() {
    const uint64_t ep = mxg_ctx_random_epoch(map_->getComm()->raw());
    uint64_t z = seed_ + 0x9E3779B97F4A7C15ull * (ep + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    mx::check(mxg_mv_random(mv_, ep == 0 ? seed_ : (z ^ (z >> 31))));
  }
  void setSeed(uint64_t seed) { seed_ = seed; }
  void norm2(std::vector<double>& norms) const { norms.resize(getNumVecs()); mx::check(mxg_mv_norm2(mv_, norms.data())); }
  // divides by the norm (the reference multiplies -- MxMultiVector.cpp:157-172)
  void normalize() { mx::check(mxg_mv_normalize(mv_)); }
  void set(Scalar val) { double a[2]; ST::pack(val, a); mx::check(mxg_mv_fill(mv_, a)); }
  // res = mv^dagger . this, column by column (true complex inner product)
  void dot(const MxMultiVector<Scalar>& mv, std::vector<Scalar>& res) const {
    res.resize(getNumVecs());
    mx::check(mxg_mv_dot(mv.mv_, mv_, reinterpret_cast<double*>(res.data())));
  }
  // this = scalarA * mvA + scalarThis * this (MxMultiVector.cpp:205-227)
  void update(Scalar scalarA, const MxMultiVector<Scalar>& mvA, Scalar scalarThis) {
    double a[2], s[2]; ST::pack(scalarA, a); ST::pack(scalarThis, s);
    mx::check(mxg_mv_update(mv_, a, mvA.mv_, s));
  }
  MxMultiVector<Scalar>& operator=(const MxMultiVector<Scalar>& mv) { mx::check(mxg_mv_assign(mv_, mv.mv_)); return *this; }

  // element access goes through the host: fine for setup and tests, not for hot loops
  Scalar operator()(MxIndex localIndex, size_t vec) const {
    std::vector<Scalar> h(size_t(getLocalLength()) * getNumVecs());
    mx::check(mxg_mv_download(mv_, reinterpret_cast<double*>(h.data()), getLocalLength()));
    return h[size_t(vec) * getLocalLength() + localIndex];
  }
  void replaceLocalValue(MxIndex row, size_t vec, Scalar value) {
    std::vector<Scalar> h(size_t(getLocalLength()) * getNumVecs());
    mx::check(mxg_mv_download(mv_, reinterpret_cast<double*>(h.data()), getLocalLength()));
    h[size_t(vec) * getLocalLength() + row] = value;
    mx::check(mxg_mv_upload(mv_, reinterpret_cast<const double*>(h.data()), getLocalLength()));
  }
  void replaceGlobalValue(MxIndex row, size_t vec, Scalar value) {
    const MxIndex l = map_->getLocalIndex(row);
    if (l >= 0) replaceLocalValue(l, vec, value);
  }
  void fromHost(const Scalar* data, MxIndex ld) { mx::check(mxg_mv_upload(mv_, reinterpret_cast<const double*>(data), ld)); }
  void toHost(Scalar* data, MxIndex ld) const { mx::check(mxg_mv_download(mv_, reinterpret_cast<double*>(data), ld)); }

  std::shared_ptr<MxVector<Scalar>> getVectorNonConst(size_t vecIndex, bool copy) {
    return std::make_shared<MxVector<Scalar>>(*this, vecIndex, copy);
  }
  std::shared_ptr<const MxVector<Scalar>> getVector(size_t vecIndex, bool copy) const {
    return std::make_shared<const MxVector<Scalar>>(*this, vecIndex, copy);
  }

  // exchange the underlying storage of two multivectors of the same shape (no data movement)
  void swap(MxMultiVector<Scalar>& other) {
    std::swap(map_, other.map_);
    std::swap(mv_, other.mv_);
    std::swap(own_, other.own_);
  }

  mxg_mv* getRawMV() const { return mv_; }   // the reference exposes the Epetra object the same way (MxMultiVector.hpp:85-89)

 protected:
  std::shared_ptr<MxMap> map_;
  mxg_mv* mv_ = nullptr;
  bool own_ = true;
  uint64_t seed_ = 12345;
};

// ---- MxVector (src/MxVector.hpp) -------------------------------------------------------------------
template <class Scalar>
class MxVector : public MxMultiVector<Scalar> {
 public:
  explicit MxVector(std::shared_ptr<MxMap> map) : MxMultiVector<Scalar>(map, 1) {}
  MxVector(const MxMultiVector<Scalar>& mv, size_t vecIndex, bool deepcopy)
      : MxMultiVector<Scalar>(mv, std::vector<size_t>(1, vecIndex), deepcopy) {}
  void replaceGlobalValue(MxIndex row, Scalar value) { MxMultiVector<Scalar>::replaceGlobalValue(row, 0, value); }
  void replaceLocalValue(MxIndex row, Scalar value) { MxMultiVector<Scalar>::replaceLocalValue(row, 0, value); }
  Scalar getValue(MxIndex localIndex) const { return (*this)(localIndex, 0); }
};

// ---- MxAnasaziMV (src/MxAnasaziMV.hpp:23-136, .cpp) -------------------------------------------------
template <class Scalar>
class MxAnasaziMV : public mx::MultiVec<Scalar>, public MxMultiVector<Scalar> {
  typedef mx::ScalarTraits<Scalar> ST;
  static const MxAnasaziMV<Scalar>& cast(const mx::MultiVec<Scalar>& A) {
    const MxAnasaziMV<Scalar>* p = dynamic_cast<const MxAnasaziMV<Scalar>*>(&A);
    if (!p) throw std::runtime_error("MxAnasaziMV: operand is not an MxAnasaziMV");   // reference: exit(EXIT_FAILURE)
    return *p;
  }
  static std::vector<size_t> toSize(const std::vector<int>& index) { return std::vector<size_t>(index.begin(), index.end()); }

 public:
  // MvTimesMatAddMv accepts a result that shares columns with A (real scalars, <= 64 source / <= 48 result columns):
  // the eigensolver right-multiplies basis blocks in place instead of through a temporary block + copy
  static constexpr bool kTimesMatInPlace = !ST::isComplex;
  MxAnasaziMV(std::shared_ptr<MxMap> map, size_t numVecs) : MxMultiVector<Scalar>(map, numVecs) {}
  MxAnasaziMV(const MxMultiVector<Scalar>& mv) : MxMultiVector<Scalar>(mv) {}
  MxAnasaziMV(const MxMultiVector<Scalar>& mv, const std::vector<size_t>& vecInds, bool deepcopy)
      : MxMultiVector<Scalar>(mv, vecInds, deepcopy) {}
  MxAnasaziMV(mxg_mv* raw, std::shared_ptr<MxMap> map, bool own) : MxMultiVector<Scalar>(raw, map, own) {}

  mx::MultiVec<Scalar>* Clone(const int numVecs) const override { return new MxAnasaziMV<Scalar>(this->getMap(), numVecs); }
  mx::MultiVec<Scalar>* CloneCopy() const override { return new MxAnasaziMV<Scalar>(*this); }
  mx::MultiVec<Scalar>* CloneCopy(const std::vector<int>& index) const override {
    return new MxAnasaziMV<Scalar>(*this, toSize(index), true);
  }
  const mx::MultiVec<Scalar>* CloneView(const std::vector<int>& index) const override {
    return new MxAnasaziMV<Scalar>(*this, toSize(index), false);
  }
  mx::MultiVec<Scalar>* CloneViewNonConst(const std::vector<int>& index) override {
    return new MxAnasaziMV<Scalar>(*this, toSize(index), false);
  }
  int GetVecLength() const override { return int(this->getMap()->getGlobalNumIndices()); }
  int GetNumberVecs() const override { return int(this->getNumVecs()); }

  // this = alpha*A*B + beta*this (MxAnasaziMV.cpp:8-86)
  void MvTimesMatAddMv(Scalar alpha, const mx::MultiVec<Scalar>& A, const mx::SerialDenseMatrix<int, Scalar>& B, Scalar beta) override {
    double a[2], b[2]; ST::pack(alpha, a); ST::pack(beta, b);
    mx::check(mxg_mv_times_mat_add_mv(a, cast(A).getRawMV(), reinterpret_cast<const double*>(B.values()), B.stride(), b, this->mv_));
  }
  // this = alpha*A + beta*B (MxAnasaziMV.cpp:89-111)
  void MvAddMv(Scalar alpha, const mx::MultiVec<Scalar>& A, Scalar beta, const mx::MultiVec<Scalar>& B) override {
    double a[2], b[2]; ST::pack(alpha, a); ST::pack(beta, b);
    mx::check(mxg_mv_add_mv(this->mv_, a, cast(A).getRawMV(), b, cast(B).getRawMV()));
  }
  // B = alpha * A^H * this (MxAnasaziMV.cpp:114-197)
  void MvTransMv(Scalar alpha, const mx::MultiVec<Scalar>& A, mx::SerialDenseMatrix<int, Scalar>& B) const override {
    double a[2]; ST::pack(alpha, a);
    mx::check(mxg_mv_trans_mv(a, cast(A).getRawMV(), this->mv_, reinterpret_cast<double*>(B.values()), B.stride()));
  }
  void MvDot(const mx::MultiVec<Scalar>& A, std::vector<Scalar>& b) const override { this->dot(cast(A), b); }
  void MvNorm(std::vector<double>& normvec) const override { this->norm2(normvec); }
  // column index[j] of this = column j of A (MxAnasaziMV.cpp:201-213)
  void SetBlock(const mx::MultiVec<Scalar>& A, const std::vector<int>& index) override {
    mx::check(mxg_mv_set_block(this->mv_, cast(A).getRawMV(), index.data(), int(index.size())));
  }
  void MvScale(Scalar alpha) override { this->scale(alpha); }
  void MvScale(const std::vector<Scalar>& alpha) override { this->scale(alpha); }
  void MvRandom() override { this->random(); }
  void MvInit(Scalar alpha) override { this->set(alpha); }
  void MvPrint(std::ostream& os) const override { os << "MxMultiVector\n"; }
};

// ---- MxCrsMatrix (src/MxCrsMatrix.hpp:14-82): host assembly, device apply ----------------------------
template <class Scalar>
class MxCrsMatrix {
  typedef mx::ScalarTraits<Scalar> ST;

 public:
  explicit MxCrsMatrix(std::shared_ptr<MxMap> rowMap) : rowMap_(rowMap) {}
  MxCrsMatrix(std::shared_ptr<MxMap> rowMap, std::shared_ptr<MxMap> colMap) : rowMap_(rowMap), colMap_(colMap) {}
  // diagonal from a scalar (MxCrsMatrix.cpp:31-45)
  MxCrsMatrix(std::shared_ptr<MxMap> rowMap, Scalar diag) : rowMap_(rowMap), colMap_(rowMap) {
    for (size_t i = 0; i < rowMap->getNodeNumIndices(); ++i) {
      const MxIndex g = rowMap->getGlobalIndex(MxIndex(i));
      insertRowValues(g, 1, &g, &diag);
    }
    fillComplete(rowMap, rowMap);
  }
  MxCrsMatrix(mxg_crs* raw, std::shared_ptr<MxMap> rowMap, std::shared_ptr<MxMap> colMap, bool own)
      : rowMap_(rowMap), colMap_(colMap), crs_(raw), own_(own), filled_(true) {}
  ~MxCrsMatrix() { if (own_ && crs_) mxg_crs_destroy(crs_); }
  MxCrsMatrix(const MxCrsMatrix&) = delete;
  MxCrsMatrix& operator=(const MxCrsMatrix&) = delete;

  // indices are always global here (MxCrsMatrix.hpp:31-36); duplicates are summed at fillComplete
  void insertRowValues(MxIndex row, size_t numEntries, const MxIndex* cols, const Scalar* vals) {
    if (filled_) throw std::runtime_error("MxCrsMatrix::insertRowValues: matrix is already fill-completed");
    auto& r = rows_[row];
    for (size_t i = 0; i < numEntries; ++i) r.emplace_back(cols[i], vals[i]);
  }
  void insertRowValues(MxIndex row, std::vector<MxIndex> cols, std::vector<Scalar> vals) {
    insertRowValues(row, cols.size(), cols.data(), vals.data());
  }
  void scale(Scalar val) {
    if (filled_) throw std::runtime_error("MxCrsMatrix::scale: scale before fillComplete (device layout is immutable)");
    for (auto& kv : rows_) for (auto& e : kv.second) e.second *= val;
  }
  void fillComplete(std::shared_ptr<MxMap> domainMap, std::shared_ptr<MxMap> rangeMap) {
    (void)rangeMap;
    colMap_ = domainMap;
    const size_t n = rowMap_->getNodeNumIndices();
    std::vector<int64_t> rowptr(n + 1, 0), cols;
    std::vector<Scalar> vals;
    for (size_t i = 0; i < n; ++i) {
      auto it = rows_.find(rowMap_->getGlobalIndex(MxIndex(i)));
      if (it != rows_.end())
        for (auto& e : it->second) { cols.push_back(e.first); vals.push_back(e.second); }
      rowptr[i + 1] = int64_t(cols.size());
    }
    mx::check(mxg_crs_create(rowMap_->raw(), colMap_->raw(), rowptr.data(), cols.data(),
                             reinterpret_cast<const double*>(vals.data()), ST::isComplex, &crs_));
    rows_.clear();
    filled_ = true;
  }
  bool isFilled() const { return filled_; }
  std::shared_ptr<MxMap> getDomainMap() const { return colMap_; }
  std::shared_ptr<MxMap> getRangeMap() const { return rowMap_; }
  // y = A x (MxCrsMatrix.cpp:347-353)
  void apply(const MxMultiVector<Scalar>& x, MxMultiVector<Scalar>& y) const { mx::check(mxg_crs_apply(crs_, x.getRawMV(), y.getRawMV())); }
  mxg_crs* getRawMatrix() const { return crs_; }

 private:
  std::shared_ptr<MxMap> rowMap_, colMap_;
  std::map<MxIndex, std::vector<std::pair<MxIndex, Scalar>>> rows_;
  mxg_crs* crs_ = nullptr;
  bool own_ = true;
  bool filled_ = false;
};

// ---- MxOperator (src/MxOperator.hpp): abstract tag; grid/field accessors are host-side setup and
// are reduced here to the maps the hot path needs ------------------------------------------------------
class MxOperatorBase {
 public:
  virtual ~MxOperatorBase() {}
  virtual std::shared_ptr<MxMap> getDomainMap() const = 0;
  virtual std::shared_ptr<MxMap> getRangeMap() const = 0;
  virtual void setImaginary(bool imag) { imaginary = imag; }
  virtual bool isImaginary() { return imaginary; }

 protected:
  bool imaginary = false;
};
