// CBLAS prototypes (include/caffe/util/mkl_alternate.hpp:11 includes <cblas.h>); the symbols come from the
// OpenBLAS that ships inside the image's Python wheels (OpenBLAS is one of the reference's three BLAS choices,
// Makefile:361-363), resolved at load time by oracle/build_ref.py.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void cblas_sgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K, float alpha, const float* A, int lda,
                 const float* B, int ldb, float beta, float* C, int ldc);
void cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K, double alpha, const double* A, int lda,
                 const double* B, int ldb, double beta, double* C, int ldc);
void cblas_sgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int M, int N, float alpha, const float* A, int lda, const float* X, int incX,
                 float beta, float* Y, int incY);
void cblas_dgemv(CBLAS_ORDER, CBLAS_TRANSPOSE, int M, int N, double alpha, const double* A, int lda, const double* X, int incX,
                 double beta, double* Y, int incY);
void cblas_saxpy(int N, float alpha, const float* X, int incX, float* Y, int incY);
void cblas_daxpy(int N, double alpha, const double* X, int incX, double* Y, int incY);
void cblas_sscal(int N, float alpha, float* X, int incX);
void cblas_dscal(int N, double alpha, double* X, int incX);
void cblas_scopy(int N, const float* X, int incX, float* Y, int incY);
void cblas_dcopy(int N, const double* X, int incX, double* Y, int incY);
float cblas_sdot(int N, const float* X, int incX, const float* Y, int incY);
double cblas_ddot(int N, const double* X, int incX, const double* Y, int incY);
float cblas_sasum(int N, const float* X, int incX);
double cblas_dasum(int N, const double* X, int incX);
#ifdef __cplusplus
}
#endif
