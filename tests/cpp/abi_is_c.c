/* The drop-in boundary is a C ABI: both public headers must compile as plain C99. */
#include "mxgpu.h"
#include "mxsolver.h"

int main(void) {
  mxs_params p;
  (void)p;
  return mxg_version() >= 100 ? 0 : 1;
}
