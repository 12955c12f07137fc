// TEST INFRASTRUCTURE (oracle) -- stands in for <mkl.h>.  The reference's
// element sources include it transitively; every BLAS/LAPACK/PARDISO call
// lives in the reference's Matrix.cpp / SparseMatrix.cpp, which the oracle
// replaces by ref_shims/matrix_twin.cpp and ref_shims/sparse_twin.cpp.
#pragma once
inline void MKL_Set_Num_Threads(int) {}
