// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" doorway onto the *unmodified* reference matrix class
// (dynadjust/include/math/dnamatrix_contiguous.cpp, compiled in place from
// /root/reference by oracle/Makefile into oracle/_ref/libref_matrix.so).
// Nothing here restates reference arithmetic: every call lands in the
// reference's own matrix_2d methods, so the Cholesky inverse the oracle uses
// is the reference's dpotrf + dpotri call sequence (MATC:952-1020) and its
// dspmv product (MATC:1471-1510).
#include <cstdint>
#include <cstring>
#include <include/math/dnamatrix_contiguous.hpp>

using dynadjust::math::matrix_2d;

extern "C" {

// optional: only resolved when linked against OpenBLAS
void openblas_set_num_threads(int) __attribute__((weak));

void ref_set_threads(int n)
{
    if (openblas_set_num_threads)
        openblas_set_num_threads(n);
}

// packed lower, column-major, n(n+1)/2 doubles, inverted in place.
// returns 0 on success, 1 on MatrixInversionFailure, 2 on any other exception.
int ref_cholesky_inverse_packed(double* ap, uint32_t n)
{
    try {
        matrix_2d m;
        m.redim_packed(n);
        std::memcpy(m.getbuffer(), ap, matrix_2d::packed_size(n) * sizeof(double));
        m.cholesky_inverse(false, false);
        std::memcpy(ap, m.getbuffer(), matrix_2d::packed_size(n) * sizeof(double));
        return 0;
    } catch (const dynadjust::math::MatrixInversionFailure&) {
        return 1;
    } catch (...) {
        return 2;
    }
}

// dense n x n column-major; on entry either the upper (lower_is_cleared=1) or the lower triangle is valid;
// on exit the full symmetric inverse.
int ref_cholesky_inverse_full(double* a, uint32_t n, int lower_is_cleared)
{
    try {
        matrix_2d m(n, n);
        for (uint32_t j = 0; j < n; ++j)
            for (uint32_t i = 0; i < n; ++i)
                m.put(i, j, a[(size_t)j * n + i]);
        m.cholesky_inverse(lower_is_cleared != 0, false);
        for (uint32_t j = 0; j < n; ++j)
            for (uint32_t i = 0; i < n; ++i)
                a[(size_t)j * n + i] = m.get(i, j);
        return 0;
    } catch (const dynadjust::math::MatrixInversionFailure&) {
        return 1;
    } catch (...) {
        return 2;
    }
}

// y = A x with A packed-lower symmetric (the reference's dspmv path, MATC:1489-1497)
int ref_multiply_sym_packed(const double* ap, uint32_t n, const double* x, double* y)
{
    try {
        matrix_2d A;
        A.redim_packed(n);
        std::memcpy(A.getbuffer(), ap, matrix_2d::packed_size(n) * sizeof(double));
        matrix_2d X(n, 1), Y(n, 1);
        std::memcpy(X.getbuffer(), x, n * sizeof(double));
        Y.multiply_sym(A, X);
        std::memcpy(y, Y.getbuffer(), n * sizeof(double));
        return 0;
    } catch (...) {
        return 2;
    }
}

// A <- S A S on packed-lower storage (MATC:1145-1152)
int ref_scale_symmetric_diagonal_packed(double* ap, uint32_t n, const double* d)
{
    try {
        matrix_2d A;
        A.redim_packed(n);
        std::memcpy(A.getbuffer(), ap, matrix_2d::packed_size(n) * sizeof(double));
        A.scale_symmetric_diagonal(d);
        std::memcpy(ap, A.getbuffer(), matrix_2d::packed_size(n) * sizeof(double));
        return 0;
    } catch (...) {
        return 2;
    }
}

}  // extern "C"
