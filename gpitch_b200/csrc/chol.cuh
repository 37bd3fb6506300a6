// Batched fp64 Cholesky factorisation + triangular inverse (blocked; trailing/panel work on the DMMA GEMM).
// Replaces tf.cholesky / tf.matrix_triangular_solve of sgpr_ss.py:44,48,51,53 and GPflow conditional().
#pragma once
#include "common.cuh"

namespace gpx {
// A [batch, M, lda]: on entry the symmetric matrix (lower triangle read), on exit L (upper triangle zeroed).
// Linv [batch, M, ldi] (may be null): L^-1, lower, upper triangle zeroed.
// work [batch, 64, M] doubles (only needed when Linv != null).
// info [batch] int: 0 ok, k>0 = leading minor of order k not positive definite (LAPACK convention).
int potrf_trinv(double* A, long long sA, int lda, double* Linv, long long sI, int ldi, double* work, int* info,
                int M, int batch, cudaStream_t st);
}
