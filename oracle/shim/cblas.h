/* oracle/shim/cblas.h — enum-only stand-in for <cblas.h> so that the reference's
 * dnamatrix_contiguous.hpp (which declares the BLAS/LAPACK prototypes itself,
 * MATH:165-190) compiles without system BLAS headers.  Values are the standard
 * CBLAS enumerators. */
#ifndef GADJ_ORACLE_SHIM_CBLAS_H_
#define GADJ_ORACLE_SHIM_CBLAS_H_
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113, CblasConjNoTrans = 114 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
#endif
