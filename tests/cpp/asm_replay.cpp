// CPU replay of the operator assembly: the same row functions, chains and C entry points as maxwell_b200/csrc/mxg_asm.cu,
// run by plain loops instead of CUDA kernels (symbols mxr_* instead of mxg_*). Test infrastructure: it lets the CPU suite
// compare the assembly logic with the oracle without a GPU. Built by tests/test_asm_replay.py with -ffp-contract=off.
#include <cstdlib>
#include <cstring>
#include <string>

#include "mxg_asm_impl.h"
#include "mxg_shape.h"

namespace {

struct HostExec {
  template <class T>
  T* alloc(int64_t n) { return static_cast<T*>(std::malloc(size_t(n > 0 ? n : 1) * sizeof(T))); }
  void free(void* p) { std::free(p); }
  template <class F>
  void forEach(int64_t n, const F& f) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) f(i);
  }
  int64_t scan(const int32_t* in, int64_t* out, int64_t n) {
    int64_t s = 0;
    for (int64_t i = 0; i < n; ++i) { out[i] = s; s += in[i]; }
    out[n] = s;
    return s;
  }
  int maxOf(const int32_t* in, int64_t n) {
    int m = 0;
    for (int64_t i = 0; i < n; ++i) m = in[i] > m ? in[i] : m;
    return m;
  }
  void toHost(void* dst, const void* src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); }
  void toExec(void* dst, const void* src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); }
  void zero(void* p, size_t bytes) { if (bytes) std::memset(p, 0, bytes); }
  void sync() {}
};

thread_local std::string gError;

}  // namespace

struct mxr_sim;
struct mxr_dcsr;
struct mxr_shape;

extern "C" const char* mxr_last_error() { return gError.c_str(); }

#define MXA_FN(name) mxr_##name
#define MXA_EXEC HostExec
#define MXA_CTX void
#define MXA_SIM_T mxr_sim
#define MXA_DCSR_T mxr_dcsr
#define MXA_SHAPE_T mxr_shape
#define MXA_NEW_EXEC(ctx) (new HostExec())
#define MXA_FAIL(code, msg) \
  do {                      \
    gError = (msg);         \
    return (code);          \
  } while (0)

#include "mxg_asm_api.inc"
