// TEST INFRASTRUCTURE ONLY: instantiates the __host__ __device__ numerical cores of the
// device D&C (zquatev_b200/csrc/dc_core.cuh) with the 1-lane policy so they can be checked
// on the CPU against LAPACK (tests/test_dc_core_host.py).  Never linked into the product.
#include "dc_core.cuh"
using namespace zq;
extern "C" {
int zqh_secular(int k, const double* dl, const double* z2, double rho, int* org, double* mu, int* iters) {
  int worst = 0;
  for (int j = 0; j < k; ++j) {
    iters[j] = secular_root<OneLane>(j, k, dl, z2, rho, &org[j], &mu[j]);
    if (iters[j] > worst) worst = iters[j];
  }
  return worst;
}
int zqh_leaf(int m, double* d, double* e, double* Z, int ldz) { return leaf_ql<OneLane>(m, d, e, Z, ldz); }
void zqh_deflate(int nm, int n1, double rho, const double* ds, const double* zs, const int* col, double* dlam, double* wz,
                 int* ndcol, int* ndtype, double* dfval, int* dfcol, int* rc1, int* rc2, double* rcc, double* rss, int* out3) {
  DeflateOut o = deflate_scan<OneLane>(nm, n1, rho, ds, zs, col, dlam, wz, ndcol, ndtype, dfval, dfcol, rc1, rc2, rcc, rss);
  out3[0] = o.k; out3[1] = o.ndefl; out3[2] = o.nrot;
}
}
