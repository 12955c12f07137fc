// TEST INFRASTRUCTURE (oracle) -- members of the reference's SparseMatrix
// class (declared in the UNMODIFIED /root/reference/src/SparseMatrix.h)
// restated without MKL PARDISO / ARPACK.
//
// Follows reference SparseMatrix.cpp:16-71 (constructors, Clear, setValue,
// Mount) and :169-194 (CSR x vector).  The linear solve and the eigen-solver
// (SparseMatrix.cpp:74-166, 270-451) are outside the assembly path and abort.
#include "SparseMatrix.h"
#include "Matrix.h"
#include <stdio.h>
#include <stdlib.h>

SparseMatrix::SparseMatrix()
{
	mounted = false;
	rows = 1;
	cols = 1;
	non_null_estimative = 1;
	m_matrix.resize(1, 1);
	tripletList.reserve(1);
}
SparseMatrix::SparseMatrix(int e_rows, int e_cols, int e_non_null_estimative)
{
	rows = e_rows;
	cols = e_cols;
	non_null_estimative = e_non_null_estimative;
	m_matrix.resize(e_rows, e_cols);
	m_matrix.reserve(e_non_null_estimative);
	// the reference reserves twice the estimate (SparseMatrix.cpp:32)
	tripletList.reserve(2 * (size_t)(e_non_null_estimative > 0 ? e_non_null_estimative : 0));
	mounted = false;
}
SparseMatrix::SparseMatrix(SparseMatrix &src)
{
	rows = src.rows;
	cols = src.cols;
	m_matrix = src.m_matrix;
	non_null_estimative = src.non_null_estimative;
	mounted = src.mounted;
	tripletList = src.tripletList;
}
SparseMatrix::~SparseMatrix() {}

SparseMatrix &SparseMatrix::operator = (SparseMatrix const &src)
{
	rows = src.rows;
	cols = src.cols;
	m_matrix = src.m_matrix;
	non_null_estimative = src.non_null_estimative;
	mounted = src.mounted;
	// as in the reference (SparseMatrix.cpp:257-267): vector copy-assignment,
	// so the constructor's reserve does NOT survive SetGlobalSize; capacity is
	// grown by push_back on the first iteration and then kept by Clear().
	tripletList = src.tripletList;
	return *this;
}

void SparseMatrix::Clear()
{
	tripletList.erase(tripletList.begin(), tripletList.end());
	mounted = false;
}
void SparseMatrix::setValue(int i, int j, double v)
{
	if (i + 1 > m_matrix.rows() || j + 1 > m_matrix.cols())
		printf("Error assigning sparse matrix value\n");
	else
		tripletList.push_back(T(i, j, v));
	if (mounted)
		mounted = false;
}
void SparseMatrix::Mount()
{
	m_matrix.setFromTriplets(tripletList.begin(), tripletList.end());
	mounted = true;
}

Matrix operator * (SparseMatrix &m, Matrix &x)
{
	if (!m.mounted)
		m.Mount();
	if (m.m_matrix.cols() != x.getLines())
	{
		printf("Impossible to multiply matrices. Dimensions are not compatible\n");
		return Matrix(0L);
	}
	Matrix y((long)m.m_matrix.rows(), 1);
	const double* a = m.m_matrix.valuePtr();
	const int* ia = m.m_matrix.outerIndexPtr();
	const int* ja = m.m_matrix.innerIndexPtr();
	for (long i = 0; i < m.m_matrix.rows(); i++)
		for (int p = ia[i]; p < ia[i + 1]; p++)
			y(i, 0) += a[p] * x(ja[p], 0);
	return y;
}

Matrix sparsesystem(SparseMatrix &, Matrix &, int *, int, int)
{
	fprintf(stderr, "oracle: sparsesystem (PARDISO) is outside the assembly path\n");
	abort();
	return Matrix(0L);
}
