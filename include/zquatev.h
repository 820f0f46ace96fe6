// zquatev.h -- C++ drop-in header of the B200-native solver.
//
// Declares the ONE public function of the reference with the reference's exact signature
// (reference zquatev.h:54, defined at zquatev.cc:42), so code written against
// qsimulate-open/zquatev -- including its own test.cc:100 -- compiles and links against
// libzquatev_b200.so without modification.  The reference's `ts::impl::*` prototypes
// (zquatev.h:57-76) are internal to its BLAS implementation and are intentionally absent:
// here the work is done by CUDA kernels behind the C ABI in zquatev_b200.h.
//
// Diagonalises the quaternionic (time-reversal symmetric) Hermitian matrix
//     (  A  -B* )        A Hermitian, B antisymmetric (B^T = -B),
//     (  B   A* )
// given through the LEFT half of the column-major 2n x 2n array D (leading dimension nld2).
// On return eig[0..n) holds the n distinct (each doubly degenerate) eigenvalues in ascending
// order and D the symmetry-adapted eigenvectors ( U -V* ; V U* ).  Return value: 0 on success
// (see zquatev_b200.h for the other codes).
#ifndef ZQUATEV_B200_CXX_H
#define ZQUATEV_B200_CXX_H

#include <complex>
#include <memory>

namespace ts {
extern int zquatev(const int n2, std::complex<double>* const D, const int nld2, double* const eig);
}

#endif
