/* Test infrastructure only (see oracle/README.md).
 * Maps the Fortran BLAS/LAPACK names declared in the reference's f77.h:40-70
 * onto the `scipy_`-prefixed symbols exported by the OpenBLAS that ships in the
 * scipy wheel of this image (LP64, self-contained). Force-included when the
 * reference sources are compiled in place from /root/reference. */
#ifndef ZQ_ORACLE_BLAS_RENAME_H
#define ZQ_ORACLE_BLAS_RENAME_H
#define zscal_   scipy_zscal_
#define zdotc_   scipy_zdotc_
#define zaxpy_   scipy_zaxpy_
#define zgemv_   scipy_zgemv_
#define ztrmv_   scipy_ztrmv_
#define zgemm3m_ scipy_zgemm3m_
#define zgemm_   scipy_zgemm_
#define zrot_    scipy_zrot_
#define zgerc_   scipy_zgerc_
#define zgeru_   scipy_zgeru_
#define zheev_   scipy_zheev_
#define zhbev_   scipy_zhbev_
#define zlartg_  scipy_zlartg_
#define zlarfg_  scipy_zlarfg_
#endif
