// TEST INFRASTRUCTURE (oracle) -- implementation of the dense-matrix twin.
// Restates the arithmetic semantics of the reference's Matrix.cpp
// (/root/reference/src/Matrix.cpp:7-365 construction/operators,
//  :1793-2032 dyadic/skew/norm/transp/V/d_V) with plain loops instead of MKL.
// See matrix_twin.h for the list of quirks that are kept on purpose.
#include "Matrix.h"
#include <stdlib.h>
#include <string.h>

static const double kPi = 3.1415926535897932384626433832795;

// ---- storage ------------------------------------------------------------
static void init_shape(Matrix* m, long r, long c)
{
	m->m_lines_deleted = true;
	m->m_alloced_lines = 0;
	m->m_lines = r;
	m->m_columns = c;
	m->m_matrix = NULL;
	if (!m->alloc())
		printf("Nao foi possivel alocar matriz! \n");
}
Matrix::Matrix(void) { init_shape(this, 1, 1); }
Matrix::Matrix(long lines) { init_shape(this, lines, 1); }
Matrix::Matrix(long lines, long columns) { init_shape(this, lines, columns); }
Matrix::Matrix(const Matrix &src)
{
	init_shape(this, src.m_lines, src.m_columns);
	for (long i = 0; i < src.m_alloced_lines; i++)
		m_matrix[i] = src.m_matrix[i];
}
Matrix::~Matrix(void) { flush(); }

bool Matrix::alloc()
{
	flush();
	long n = m_lines * m_columns;
	m_matrix = new double[n > 0 ? n : 0];
	m_alloced_lines = n;
	m_lines_deleted = false;
	for (long i = 0; i < n; i++)
		m_matrix[i] = 0.0;
	return true;
}
bool Matrix::flush()
{
	if (!m_lines_deleted)
	{
		delete[] m_matrix;
		m_matrix = NULL;
		m_lines_deleted = true;
		m_alloced_lines = 0;
	}
	return true;
}
void Matrix::clear()
{
	for (long i = 0; i < m_alloced_lines; i++)
		m_matrix[i] = 0.0;
}

void Matrix::print()
{
	printf("\n");
	for (long i = 0; i < m_lines; i++)
	{
		printf("|");
		for (long j = 0; j < m_columns; j++)
			printf(" %.4e ", m_matrix[i + j * m_lines]);
		printf("|\n");
	}
	printf("\n");
}
void Matrix::fprint(char* s)
{
	FILE* f = fopen(s, "w");
	if (!f) return;
	fprintf(f, "\n");
	for (long i = 0; i < m_lines; i++)
	{
		for (long j = 0; j < m_columns; j++)
			fprintf(f, " %.14e\t", m_matrix[i + j * m_lines]);
		fprintf(f, "\n");
	}
	fprintf(f, "\n");
	fclose(f);
}

void Matrix::MatrixToPtr(double** ptr, int order)
{
	for (int i = 0; i < order; i++)
		for (int j = 0; j < order; j++)
			ptr[i][j] = m_matrix[i + j * order];
}
void Matrix::PtrToMatrix(double** ptr, int order)
{
	for (int i = 0; i < order; i++)
		for (int j = 0; j < order; j++)
			m_matrix[i + j * order] = ptr[i][j];
}
void Matrix::PtrToMatrix(double** ptr, int lines, int columns)
{
	for (int i = 0; i < lines; i++)
		for (int j = 0; j < columns; j++)
			m_matrix[i + j * lines] = ptr[i][j];
}

// Out-of-range access: message + a throw-away cell (kept from the reference).
double &Matrix::operator() (long line, long column) const
{
	if (line > m_lines - 1 || column > m_columns - 1 || line < 0 || column < 0)
	{
		printf("Not valid position accessed in matrix! (%ld,%ld)\n", line, column);
		double* scratch = new double[1];
		scratch[0] = 0;
		return scratch[0];
	}
	return m_matrix[line + column * m_lines];
}

// Re-shape only when the ELEMENT COUNT differs (kept from the reference).
Matrix &Matrix::operator = (Matrix const &src)
{
	if (src.m_alloced_lines != m_alloced_lines)
	{
		flush();
		m_lines = src.m_lines;
		m_columns = src.m_columns;
		m_matrix = NULL;
		if (!alloc())
			printf("Nao foi possivel alocar matriz! \n");
	}
	for (long i = 0; i < src.m_alloced_lines; i++)
		m_matrix[i] = src.m_matrix[i];
	return *this;
}

// ---- elementwise --------------------------------------------------------
static bool same_shape(const Matrix &a, const Matrix &b)
{
	return a.m_lines == b.m_lines && a.m_columns == b.m_columns;
}
Matrix operator + (const Matrix &a, const Matrix &b)
{
	if (!same_shape(a, b))
	{
		printf("Matrizes devem possuir a mesma dimensao! \n");
		return Matrix(0L);
	}
	Matrix r(a.m_lines, a.m_columns);
	for (long i = 0; i < a.m_alloced_lines; i++)
		r.m_matrix[i] = a.m_matrix[i] + b.m_matrix[i];
	return r;
}
Matrix operator - (const Matrix &a, const Matrix &b)
{
	if (!same_shape(a, b))
	{
		printf("Matrizes devem possuir a mesma dimensao! \n");
		return Matrix(0L);
	}
	Matrix r(a.m_lines, a.m_columns);
	for (long i = 0; i < a.m_alloced_lines; i++)
		r.m_matrix[i] = a.m_matrix[i] - b.m_matrix[i];
	return r;
}
Matrix operator * (double s, const Matrix &a)
{
	Matrix r(a.m_lines, a.m_columns);
	for (long i = 0; i < r.m_alloced_lines; i++)
		r.m_matrix[i] = a.m_matrix[i] * s;
	return r;
}
Matrix operator * (const Matrix &a, double s) { return s * a; }

// ---- products -----------------------------------------------------------
// Column-major C(m x n) += A(m x k) B(k x n) on a zero-filled C; the inner
// index runs fastest, one fused-free multiply-add at a time (no FMA
// contraction is requested; the compiler flags decide).
Matrix operator * (const Matrix &a, const Matrix &b)
{
	if (a.m_columns != b.m_lines)
	{
		if (a.m_lines == b.m_lines)
		{
			Matrix r(1, 1);
			for (long i = 0; i < a.m_lines; i++)
				r(0, 0) += a(i, 0) * b(i, 0);
			return r;
		}
		printf("Nao e possivel multiplicar as matrizes. Dimensoes incompativeis!");
		return Matrix(0L);
	}
	const long m = a.m_lines, kk = a.m_columns, n = b.m_columns;
	Matrix r(m, n);
	const double* A = a.m_matrix;
	const double* B = b.m_matrix;
	double* C = r.m_matrix;
	for (long j = 0; j < n; j++)
		for (long i = 0; i < m; i++)
		{
			double acc = 0.0;
			for (long k = 0; k < kk; k++)
				acc += A[i + k * m] * B[k + j * kk];
			C[i + j * m] += acc;
		}
	return r;
}
Matrix Dot(const Matrix &a, const Matrix &b) { return a * b; }

bool operator == (const Matrix &a, const Matrix &b)
{
	if (!same_shape(a, b)) return false;
	for (long i = 0; i < a.m_alloced_lines; i++)
		if (a.m_matrix[i] != b.m_matrix[i]) return false;
	return true;
}
bool operator != (const Matrix &a, const Matrix &b) { return !(a == b); }

double dot(const Matrix &a, const Matrix &b)
{
	if (a.m_lines != b.m_lines)
	{
		printf("Nao e possivel calcular o produto escalar. Dimensoes incompativeis!");
		return 0;
	}
	double s = 0.0;
	for (long i = 0; i < a.m_alloced_lines; i++)
		s += a.m_matrix[i] * b.m_matrix[i];
	return s;
}
Matrix cross(const Matrix &a, const Matrix &b)
{
	if (a.m_columns != 1 || b.m_columns != 1 || a.m_lines != 3 || b.m_lines != 3)
	{
		printf("Nao e possivel calcular o produto vetorial. Dimensoes incompativeis!");
		return Matrix(0L);
	}
	Matrix r(3);
	r(0, 0) = a(1, 0) * b(2, 0) - a(2, 0) * b(1, 0);
	r(1, 0) = a(2, 0) * b(0, 0) - a(0, 0) * b(2, 0);
	r(2, 0) = a(0, 0) * b(1, 0) - a(1, 0) * b(0, 0);
	return r;
}
Matrix dyadic(const Matrix &a, const Matrix &b)
{
	if (a.m_columns != 1 || b.m_columns != 1 || a.m_lines != b.m_lines)
	{
		printf("Nao e possivel calcular o produto tensorial. Dimensoes incompativeis!");
		return Matrix(0L);
	}
	const int n = (int)a.m_lines;
	Matrix r(n, n);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++)
			r(i, j) = a(i, 0) * b(j, 0);
	return r;
}
Matrix skew(const Matrix &a)
{
	if (a.m_columns != 1 || a.m_lines != 3)
	{
		printf("Nao e possivel calcular o produto escalar. Dimensoes incompativeis!");
		return Matrix(0L);
	}
	Matrix r(3, 3);
	r(0, 1) = -a(2, 0);
	r(0, 2) = +a(1, 0);
	r(1, 2) = -a(0, 0);
	r(1, 0) = +a(2, 0);
	r(2, 0) = -a(1, 0);
	r(2, 1) = +a(0, 0);
	return r;
}
Matrix axial(const Matrix &a)
{
	if (a.m_columns != 3 || a.m_lines != 3)
	{
		printf("Nao e possivel calcular o produto escalar. Dimensoes incompativeis!");
		return Matrix(0L);
	}
	Matrix r(3);
	r(0, 0) = -a(1, 2);
	r(1, 0) = +a(0, 2);
	r(2, 0) = -a(0, 1);
	return r;
}
Matrix transp(const Matrix &a)
{
	Matrix r(a.m_columns, a.m_lines);
	for (long j = 0; j < a.m_columns; j++)
		for (long i = 0; i < a.m_lines; i++)
			r(j, i) = a(i, j);
	return r;
}
void zeros(Matrix* a)
{
	for (long j = 0; j < a->m_columns; j++)
		for (long i = 0; i < a->m_lines; i++)
			(*a)(i, j) = 0.0;
}

// Euclidean norm for 2/3/4/6-vectors, otherwise the infinity norm with the
// reference's NaN/Inf sentinel of 1e100 (reference Matrix.cpp:1888-1956).
double norm(const Matrix &a)
{
	if (a.m_columns != 1)
	{
		printf("Dimensao nao consistente para calculo da norma");
		return 0;
	}
	const long n = a.m_lines;
	if (n == 2 || n == 3 || n == 4 || n == 6)
	{
		double s = a(0, 0) * a(0, 0);
		for (long i = 1; i < n; i++)
			s = s + a(i, 0) * a(i, 0);
		return sqrt(s);
	}
	double mx = 0;
	for (long i = 0; i < n; i++)
	{
		double v = a(i, 0);
		if (v != v) return 1e100;
		if (v >= 1e300 || v <= -1e300) return 1e100;
		if (mx < fabs(v)) mx = fabs(v);
	}
	return mx;
}
double norm4(const Matrix &a)
{
	if (a.m_lines < 4)
	{
		printf("Error. Function norm4\n");
		return 0;
	}
	return sqrt(a(0, 0) * a(0, 0) + a(1, 0) * a(1, 0) + a(2, 0) * a(2, 0) + a(3, 0) * a(3, 0));
}

// ---- rotation-tangent helpers (reference Matrix.cpp:1999-2032) ----------
// The expressions keep the reference's term structure, including the terms
// multiplied by the zero coefficients h3/h5/h7, so that NaN/Inf propagation
// and association order are the same.
Matrix V(Matrix x, Matrix t, double alpha_escalar)
{
	double h = 4.0 / (4.0 + alpha_escalar * alpha_escalar);
	double h2 = 0.5 * h, h3 = 0, h4 = -0.25 * h * h, h5 = 0, h8 = -0.5 * h * h;
	Matrix left = dyadic(h8 * t - h4 * (skew(x) * t) + h5 * (skew(x) * skew(x)) * t, x);
	return left + h2 * skew(t) - h3 * (2 * skew(x) * skew(t) - skew(t) * skew(x));
}
Matrix d_V(Matrix x, Matrix d_x, Matrix t, double alpha_escalar)
{
	double h = 4.0 / (4.0 + alpha_escalar * alpha_escalar);
	double h3 = 0, h4 = -0.25 * h * h, h5 = 0, h6 = 0.25 * h * h * h, h7 = 0;
	double h8 = -0.5 * h * h, h9 = 0.5 * h * h * h;
	return dot(x, d_x) * (dyadic(h9 * t - h6 * (skew(x) * t) + h7 * (skew(x) * skew(x)) * t, x)) +
		dyadic(h8 * t - h4 * skew(x) * t + h5 * (skew(x) * skew(x)) * t, d_x) +
		dyadic(h5 * (skew(x) * skew(d_x) + skew(d_x) * skew(x)) * t - h4 * skew(d_x) * t, x) +
		h4 * dot(x, d_x) * skew(t) - h5 * dot(x, d_x) * (2 * skew(x) * skew(t) - skew(t) * skew(x)) -
		h3 * (2 * skew(d_x) * skew(t) - skew(t) * skew(d_x));
}

// ---- small dense solves (off the hot path; Gauss-Jordan, partial pivot) --
static Matrix gauss_jordan_inverse(const Matrix &a, int n)
{
	if (a.m_lines != n || a.m_columns != n)
	{
		printf("Matrix inversion: unexpected dimensions\n");
		return Matrix(0L);
	}
	Matrix w(a), inv(n, n);
	for (int i = 0; i < n; i++) inv(i, i) = 1.0;
	for (int c = 0; c < n; c++)
	{
		int p = c;
		for (int r = c + 1; r < n; r++)
			if (fabs(w(r, c)) > fabs(w(p, c))) p = r;
		if (p != c)
			for (int j = 0; j < n; j++)
			{
				double t1 = w(c, j); w(c, j) = w(p, j); w(p, j) = t1;
				double t2 = inv(c, j); inv(c, j) = inv(p, j); inv(p, j) = t2;
			}
		double d = 1.0 / w(c, c);
		for (int j = 0; j < n; j++) { w(c, j) *= d; inv(c, j) *= d; }
		for (int r = 0; r < n; r++)
		{
			if (r == c) continue;
			double f = w(r, c);
			if (f == 0.0) continue;
			for (int j = 0; j < n; j++) { w(r, j) -= f * w(c, j); inv(r, j) -= f * inv(c, j); }
		}
	}
	return inv;
}
Matrix invert2x2(const Matrix &a) { return gauss_jordan_inverse(a, 2); }
Matrix invert3x3(const Matrix &a) { return gauss_jordan_inverse(a, 3); }
Matrix invert4x4(const Matrix &a) { return gauss_jordan_inverse(a, 4); }
Matrix invert5x5(const Matrix &a) { return gauss_jordan_inverse(a, 5); }
Matrix invert6x6(const Matrix &a) { return gauss_jordan_inverse(a, 6); }
Matrix invert(const Matrix &a) { return gauss_jordan_inverse(a, (int)a.m_lines); }

Matrix fullsystem(Matrix &A, Matrix &b, int *flag_error)
{
	const int n = (int)A.m_lines;
	*flag_error = 0;
	for (int c = 0; c < n; c++)
	{
		int p = c;
		for (int r = c + 1; r < n; r++)
			if (fabs(A(r, c)) > fabs(A(p, c))) p = r;
		if (A(p, c) == 0.0) { *flag_error = 1; return b; }
		if (p != c)
		{
			for (int j = 0; j < n; j++) { double t = A(c, j); A(c, j) = A(p, j); A(p, j) = t; }
			double t = b(c, 0); b(c, 0) = b(p, 0); b(p, 0) = t;
		}
		for (int r = c + 1; r < n; r++)
		{
			double f = A(r, c) / A(c, c);
			for (int j = c; j < n; j++) A(r, j) -= f * A(c, j);
			b(r, 0) -= f * b(c, 0);
		}
	}
	for (int r = n - 1; r >= 0; r--)
	{
		double s = b(r, 0);
		for (int j = r + 1; j < n; j++) s -= A(r, j) * b(j, 0);
		b(r, 0) = s / A(r, r);
	}
	return b;
}

// Symmetric eigen-solvers are LAPACK calls in the reference and are never
// reached from the assembly path; the oracle refuses to guess.
static int no_lapack(const char* what)
{
	fprintf(stderr, "oracle: %s needs LAPACK and is outside the assembly path\n", what);
	abort();
	return 1;
}
int fulleigen1(Matrix &, Matrix &, Matrix &, double) { return no_lapack("fulleigen1"); }
int fulleigen2(Matrix &, Matrix &, Matrix &) { return no_lapack("fulleigen2"); }
double mineigen(Matrix &, Matrix &, Matrix &, double) { return (double)no_lapack("mineigen"); }

// ---- scalar helpers -----------------------------------------------------
static double wrap_quadrants(double arc, bool positive_range)
{
	double c = cos(arc), s = sin(arc);
	if (c > 1.0) c = 1.0;
	if (c < -1.0) c = -1.0;
	if (s > 1.0) s = 1.0;
	if (s < -1.0) s = -1.0;
	double r = 0.0;
	if (s >= 0 && c >= 0) r = asin(s);
	if (s >= 0 && c < 0) r = -asin(s) + kPi;
	if (s < 0 && c >= 0) r = asin(s) + (positive_range ? 2 * kPi : 0.0);
	if (s < 0 && c < 0) r = -asin(s) - kPi + (positive_range ? 2 * kPi : 0.0);
	return r;
}
double ArcReduction(double arc) { return wrap_quadrants(arc, false); }
double ArcReduction2p(double arc) { return wrap_quadrants(arc, true); }

Matrix List(double a, double b, double c)
{
	Matrix r(3);
	r(0, 0) = a; r(1, 0) = b; r(2, 0) = c;
	return r;
}
double Power(double a, double b) { return pow(a, b); }
double Power(Matrix a, double b)
{
	if (a.m_lines != 1 || a.m_columns != 1)
	{
		if (a.m_lines == 3 && a.m_columns == 1 && b == 2)
			return dot(a, a);
		printf("Error in Power function. Supposed to receive a 1x1 or 3x1 matrix!\n");
	}
	return pow(a(0, 0), b);
}
double Sin(double a) { return sin(a); }
double Cos(double a) { return cos(a); }
double operator + (double a, const Matrix &b)
{
	if (b.m_lines != 1 || b.m_columns != 1)
	{
		printf("Matriz deve ser unitaria! operator+(double,matrix)\n");
		return 0;
	}
	return b(0, 0) + a;
}
