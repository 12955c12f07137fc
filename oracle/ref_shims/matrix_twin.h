// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product path.
//
// Const-correct twin of the reference's dense matrix interface, installed as
// "Matrix.h" into the oracle/_ref build farm so that the UNMODIFIED reference
// element sources (Beam_1.cpp, Shell_1.cpp, Node.cpp, Solution.cpp, ...)
// compile with g++.  The reference header (/root/reference/src/Matrix.h:5-85)
// declares its free functions on `Matrix&`, which only MSVC accepts for
// temporaries; this twin declares the same names on `const Matrix&`.
//
// Behaviours deliberately kept (SURVEY.md section 8c "hazards"):
//   * operator* degrades to a dot product when inner sizes mismatch but row
//     counts agree                       (reference Matrix.cpp:221-230)
//   * GEMM accumulates into a zero-filled result (beta = 1, Matrix.cpp:246)
//   * operator= keeps the destination SHAPE when element counts are equal
//                                          (reference Matrix.cpp:293-314)
//   * operator() on an out-of-range index prints and hands back a scratch
//     cell instead of failing              (reference Matrix.cpp:316-330)
// MKL is replaced by plain triple loops (column-major, k innermost), which is
// a restatement: last-bit rounding may differ from the MSVC+MKL build.
#pragma once
#include <stdio.h>
#include <cmath>

class Matrix
{
public:
	Matrix(void);
	Matrix(long lines);
	Matrix(long lines, long columns);
	Matrix(const Matrix &copied);
	~Matrix(void);

	long getLines() const { return m_lines; }
	long getColumns() const { return m_columns; }
	void setLines(long value) { m_lines = value; }
	void setColumns(long value) { m_columns = value; }
	double* getMatrix() const { return m_matrix; }

	void print();
	void fprint(char* s);
	bool alloc();
	bool flush();
	void clear();

	void MatrixToPtr(double** ptr, int order);
	void PtrToMatrix(double** ptr, int order);
	void PtrToMatrix(double** ptr, int lines, int columns);
	double &operator() (long line, long column) const;
	Matrix &operator = (Matrix const &matrix1);

	double*  m_matrix;
	long     m_lines;
	long     m_columns;
	long	 m_alloced_lines;
	bool	 m_lines_deleted;
};

Matrix operator + (const Matrix &a, const Matrix &b);
Matrix operator - (const Matrix &a, const Matrix &b);
Matrix operator * (const Matrix &a, const Matrix &b);
Matrix operator * (double s, const Matrix &a);
Matrix operator * (const Matrix &a, double s);
bool operator == (const Matrix &a, const Matrix &b);
bool operator != (const Matrix &a, const Matrix &b);
double dot(const Matrix &a, const Matrix &b);
Matrix cross(const Matrix &a, const Matrix &b);
Matrix dyadic(const Matrix &a, const Matrix &b);
Matrix skew(const Matrix &a);
Matrix axial(const Matrix &a);
Matrix fullsystem(Matrix &A, Matrix &b, int *flag_error);
double norm(const Matrix &a);
double norm4(const Matrix &a);
Matrix transp(const Matrix &a);
void zeros(Matrix* a);
Matrix invert2x2(const Matrix &a);
Matrix invert3x3(const Matrix &a);
Matrix invert4x4(const Matrix &a);
Matrix invert5x5(const Matrix &a);
Matrix invert6x6(const Matrix &a);
Matrix invert(const Matrix &a);

int fulleigen1(Matrix &A, Matrix &P, Matrix &D, double abstol);
int fulleigen2(Matrix &A, Matrix &P, Matrix &D);
double mineigen(Matrix &A, Matrix &P, Matrix &D, double abstol);

Matrix V(Matrix x, Matrix t, double alpha_escalar);
Matrix d_V(Matrix x, Matrix d_x, Matrix t, double alpha_escalar);

double ArcReduction(double arc);
double ArcReduction2p(double arc);
Matrix List(double a, double b, double c);
double Power(double a, double b);
double Power(Matrix a, double b);
double Sin(double a);
double Cos(double a);
Matrix Dot(const Matrix &a, const Matrix &b);
double operator + (double a, const Matrix &b);
