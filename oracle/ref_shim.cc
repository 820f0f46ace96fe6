// Test infrastructure only. A C-ABI door onto the UNMODIFIED reference
// ts::zquatev (zquatev.h:54 / zquatev.cc:42), so Python (ctypes) can call the
// reference compiled in place from /root/reference. Also exposes LAPACK zheev
// on the full 2n x 2n matrix the way test.cc:84-95 uses it.
#include <complex>
#include <memory>
#include "zquatev.h"
extern "C" {
void zheev_(const char*, const char*, const int*, std::complex<double>*, const int*, double*,
            std::complex<double>*, const int*, double*, int*);
void scipy_openblas_set_num_threads(int);
int  scipy_openblas_get_num_threads(void);

int zq_ref_zquatev(int n2, void* D, int ld2, double* eig) {
  return ts::zquatev(n2, static_cast<std::complex<double>*>(D), ld2, eig);
}
int zq_ref_zheev(int n2, void* C, int ld, double* eig) {
  std::complex<double>* c = static_cast<std::complex<double>*>(C);
  int lwork = -1, info = 0;
  std::unique_ptr<double[]> rwork(new double[3 * n2 > 2 ? 3 * n2 - 2 : 1]);
  std::complex<double> q;
  zheev_("V", "U", &n2, c, &ld, eig, &q, &lwork, rwork.get(), &info);
  lwork = static_cast<int>(q.real());
  std::unique_ptr<std::complex<double>[]> work(new std::complex<double>[lwork]);
  zheev_("V", "U", &n2, c, &ld, eig, work.get(), &lwork, rwork.get(), &info);
  return info;
}
void zq_ref_set_threads(int t) { scipy_openblas_set_num_threads(t); }
int  zq_ref_get_threads(void) { return scipy_openblas_get_num_threads(); }
}
