/* The drop-in boundary is a C ABI: the public headers must compile as plain C99. */
#include "mxgpu.h"
#include "mxsolver.h"
#include "mxasm.h"

int main(void) {
  mxs_params p;
  mxg_shape* s = 0;
  const double zero[3] = {0.0, 0.0, 0.0};
  (void)p;
  if (mxg_shape_sphere(1.0, zero, &s) != 0 || mxg_shape_destroy(s) != 0) return 2;
  return mxg_version() >= 100 ? 0 : 1;
}
