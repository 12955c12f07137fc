// TEST INFRASTRUCTURE (oracle) -- CPU restatement of GIRAFFE's per-Newton-
// iteration element assembly, written densely "as the reference writes it"
// (no structure exploitation) so that it is an independent check of the
// structure-exploiting CUDA kernels.  Interface and parity status: see
// gfa_oracle.h.  Every block cites the reference lines it restates
// (paths relative to /root/reference/src).
#include "gfa_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <omp.h>

namespace {

// ------------------------------------------------------------------------
// small fixed-size dense algebra (row-major storage, value semantics)
// ------------------------------------------------------------------------
template <int R, int C>
struct Mx
{
	double a[R * C];
	Mx() { for (int i = 0; i < R * C; i++) a[i] = 0.0; }
	double& operator()(int i, int j) { return a[i * C + j]; }
	double operator()(int i, int j) const { return a[i * C + j]; }
	double& operator[](int i) { return a[i]; }
	double operator[](int i) const { return a[i]; }
};
typedef Mx<3, 1> V3;
typedef Mx<3, 3> M3;

template <int R, int K, int C>
Mx<R, C> operator*(const Mx<R, K>& A, const Mx<K, C>& B)
{
	Mx<R, C> o;
	for (int i = 0; i < R; i++)
		for (int j = 0; j < C; j++)
		{
			double s = 0.0;
			for (int k = 0; k < K; k++) s += A(i, k) * B(k, j);
			o(i, j) = s;
		}
	return o;
}
template <int R, int C> Mx<R, C> operator+(const Mx<R, C>& A, const Mx<R, C>& B)
{ Mx<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] + B.a[i]; return o; }
template <int R, int C> Mx<R, C> operator-(const Mx<R, C>& A, const Mx<R, C>& B)
{ Mx<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] - B.a[i]; return o; }
template <int R, int C> Mx<R, C> operator*(double s, const Mx<R, C>& A)
{ Mx<R, C> o; for (int i = 0; i < R * C; i++) o.a[i] = A.a[i] * s; return o; }
template <int R, int C> Mx<C, R> tr(const Mx<R, C>& A)
{ Mx<C, R> o; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) o(j, i) = A(i, j); return o; }
template <int R, int C, int r, int c> void put(Mx<R, C>& dst, int i0, int j0, const Mx<r, c>& src)
{ for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) dst(i0 + i, j0 + j) = src(i, j); }

V3 vec(double x, double y, double z) { V3 v; v[0] = x; v[1] = y; v[2] = z; return v; }
double dot(const V3& a, const V3& b) { double s = 0.0; for (int i = 0; i < 3; i++) s += a[i] * b[i]; return s; }
double norm(const V3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
V3 cross(const V3& a, const V3& b)
{ return vec(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]); }
M3 skew(const V3& v)
{
	M3 s;
	s(0, 1) = -v[2]; s(0, 2) = +v[1]; s(1, 2) = -v[0];
	s(1, 0) = +v[2]; s(2, 0) = -v[1]; s(2, 1) = +v[0];
	return s;
}
M3 dyad(const V3& a, const V3& b) { M3 o; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o(i, j) = a[i] * b[j]; return o; }
M3 eye3() { M3 o; o(0, 0) = o(1, 1) = o(2, 2) = 1.0; return o; }

// Matrix.cpp:1999-2012 (terms with the zero coefficients h3,h5 dropped)
M3 Vop(const V3& x, const V3& t, double alpha)
{
	double h = 4.0 / (4.0 + alpha * alpha);
	double h2 = 0.5 * h, h4 = -0.25 * h * h, h8 = -0.5 * h * h;
	return dyad(h8 * t - h4 * (skew(x) * t), x) + h2 * skew(t);
}
// Matrix.cpp:2014-2032 (terms with h3,h5,h7 = 0 dropped)
M3 dVop(const V3& x, const V3& dx, const V3& t, double alpha)
{
	double h = 4.0 / (4.0 + alpha * alpha);
	double h4 = -0.25 * h * h, h6 = 0.25 * h * h * h, h8 = -0.5 * h * h, h9 = 0.5 * h * h * h;
	double xd = dot(x, dx);
	return xd * dyad(h9 * t - h6 * (skew(x) * t), x) + dyad(h8 * t - h4 * (skew(x) * t), dx)
		+ dyad((-h4) * (skew(dx) * t), x) + (h4 * xd) * skew(t);
}

// ------------------------------------------------------------------------
// model tables
// ------------------------------------------------------------------------
enum { T_BEAM = 1, T_PIPE = 2, T_SHELL = 3, T_SOLID = 7 };    // Pipe_1 is held as a BeamEl: same Mount (Pipe_1.cpp:836-974)

struct ShellEl
{
	double area, alpha1, lambda, mu, stiff_drill, thick, rho;
	M3 T3;                                   // rows e1r,e2r,e3r (Shell_1.cpp:1470-1511)
	double Nu[3][6], Nu1[3][6], Nu2[3][6];    // translation shape fns and x1/x2 derivatives at the 3 points
	double Na[3][3], Na1[3][3], Na2[3][3];    // rotation shape fns (mid nodes)
	double grav[6];                           // sum_g w4[g]*N_a4[g] per node (one application)
	// committed state per Gauss point (Shell_1.h:117-127)
	M3 Q_i[3]; V3 zx1_i[3], zx2_i[3], k1_i[3], k2_i[3];
	// trial values kept by Mount for SaveLagrange (Shell_1.cpp:1650-1664)
	M3 Q_d[3], Xi_d[3]; V3 a_x1[3], a_x2[3], u_x1[3], u_x2[3];
	double K[27 * 27], Fint[27], P[27], energy;
	// Newmark dynamics (Shell_1.cpp:2406-2547): committed Rodrigues vector, trial increments of Mount,
	// 6-point rule data of MountMassModal, inertia constants (:2022-2024), stored Rayleigh matrix
	V3 alpha_i[3], a_d[3], u_d[3];
	double N4u[6][6], N4a[6][3], w4[6];
	double coef1, coef2, coef3;
	std::vector<double> CR;
	double res[3][24];                        // eta_r1 eta_r2 kappa_r1 kappa_r2 n_r1 n_r2 m_r1 m_r2 per point (Shell_1.h:153-160)
};
struct BeamEl
{
	Mx<6, 6> D; M3 T3; V3 e3r; double jac, length, rhoA;
	double N[2][3], dN[2][3];
	M3 Q_i[2]; V3 dz_i[2], k_i[2];
	M3 Q_d[2]; V3 dz[2], kr[2];
	M3 Xi_d[2], dXi_d[2]; double g_last, Aint;  // left by Mount for MountPipeSpecialLoads (g: the LAST point's, Pipe_1.cpp:887)
	double K[18 * 18], Fint[18], P[18], energy;
	// Newmark dynamics (Beam_1.cpp:1564-1673): committed Rodrigues vector, trial increments of Mount,
	// section inertia (:582-596), stored Rayleigh matrix
	V3 alpha_i[2], a_d[2], u_d[2];
	M3 Mr, Jr; V3 br;
	std::vector<double> CR;
	bool energy_on;                           // Pipe_1::Mount never adds to strain_energy
	double res[2][12];                        // epsilon_r(6) sigma_r(6) per point (Beam_1.h:80-81)
};
struct SolidEl
{
	double lambda, mu, rho;
	double K[24 * 24], Fint[24], P[24], energy;
};

struct World
{
	int n_nodes = 0;
	std::vector<double> ref, copy;            // [n][3], [n][6]
	std::vector<double> hooke, sec, thick, cs, pipe;
	int n_el = 0;
	std::vector<int> type, mat, secid, csid, nptr, nodes;
	std::vector<char> is_pipe;
	std::vector<double> pret;
	std::vector<int> pipe_load_ptr, pipe_load_el; std::vector<double> pipe_load_p;   // PipeLoad objects (gfo_set_pipe_loads)
	int g_on = 0; double g[3] = { 0, 0, 0 };
	std::vector<int> cmask, gls;
	int n_free = 0, n_fixed = 0;
	std::vector<int> slot;                    // element -> index into its type table
	std::vector<ShellEl> shells; std::vector<BeamEl> beams; std::vector<SolidEl> solids;
	std::vector<double> disp;
	// global system
	struct Trip { int r, c; double v; };
	std::vector<Trip> trip[4], extra[4];
	std::vector<int> outer[4], inner[4]; std::vector<double> val[4];
	int rows[4] = { 0, 0, 0, 0 }, cols[4] = { 0, 0, 0, 0 };
	std::vector<double> PA, IA, PB;
	// Newmark dynamics: Node::vel/accel/copy_vel/copy_accel, Dynamic::a1..a6, alpha, beta
	std::vector<double> vel, accel, copy_vel, copy_accel;
	double nm[6] = { 0, 0, 0, 0, 0, 0 }, ray_alpha = 0.0, ray_beta = 0.0;
} W;

const double* X(int node1) { return &W.ref[3 * (size_t)(node1 - 1)]; }
V3 Xv(int node1) { const double* p = X(node1); return vec(p[0], p[1], p[2]); }

// ------------------------------------------------------------------------
// Shell_1
// ------------------------------------------------------------------------
// Shell_1.cpp:1666-2365 (homogeneous section branch)
void shell_precalc(ShellEl& s, const int* nd, double E, double nu, double rho, double t)
{
	V3 x[6];
	for (int a = 0; a < 6; a++) x[a] = Xv(nd[a]);
	double A = 0.5 * norm(cross(x[1] - x[0], x[2] - x[0]));
	s.area = A; s.thick = t; s.rho = rho;
	V3 n = cross(x[1] - x[0], x[2] - x[0]);
	V3 e3 = (1.0 / norm(n)) * n;
	V3 eg = vec(1.0, 0.0, 0.0);
	if (std::fabs(dot(eg, e3)) >= 1.0 - 1e-4) eg = vec(0.0, 1.0, 0.0);   // :1990
	V3 e1 = eg - dot(eg, e3) * e3;
	e1 = (1.0 / norm(e1)) * e1;
	V3 e2 = cross(e3, e1);
	for (int j = 0; j < 3; j++) { s.T3(0, j) = e1[j]; s.T3(1, j) = e2[j]; s.T3(2, j) = e3[j]; }
	s.mu = E / (2.0 * (1 + nu));
	s.lambda = 2.0 * s.mu * nu / (1 - 2.0 * nu);
	s.stiff_drill = E * t * t * t;

	double b1 = dot(x[1] - x[2], e2), b2 = dot(x[2] - x[0], e2), b3 = dot(x[0] - x[1], e2);
	double c1 = dot(x[2] - x[1], e1), c2 = dot(x[0] - x[2], e1), c3 = dot(x[1] - x[0], e1);
	double L1x = 0.5 * b1 / A, L2x = 0.5 * b2 / A, L3x = 0.5 * b3 / A;
	double L1y = 0.5 * c1 / A, L2y = 0.5 * c2 / A, L3y = 0.5 * c3 / A;
	for (int g = 0; g < 3; g++)
	{
		const V3& xp = x[3 + g];                                            // :2031-2039
		double A1 = 0.5 * norm(cross(x[1] - xp, x[2] - xp));
		double A2 = 0.5 * norm(cross(x[2] - xp, x[0] - xp));
		double A3 = 0.5 * norm(cross(x[0] - xp, x[1] - xp));
		double L1 = A1 / A, L2 = A2 / A, L3 = A3 / A;
		double* N = s.Nu[g]; double* N1 = s.Nu1[g]; double* N2 = s.Nu2[g];
		N[0] = (2 * L1 - 1) * L1; N[1] = (2 * L2 - 1) * L2; N[2] = (2 * L3 - 1) * L3;
		N[3] = 4 * L1 * L2; N[4] = 4 * L2 * L3; N[5] = 4 * L3 * L1;
		s.Na[g][0] = 1 - 2 * L3; s.Na[g][1] = 1 - 2 * L1; s.Na[g][2] = 1 - 2 * L2;
		N1[0] = 4 * L1x * L1 - L1x; N1[1] = 4 * L2x * L2 - L2x; N1[2] = 4 * L3x * L3 - L3x;
		N1[3] = 4 * L1x * L2 + 4 * L1 * L2x; N1[4] = 4 * L2x * L3 + 4 * L2 * L3x; N1[5] = 4 * L3x * L1 + 4 * L3 * L1x;
		s.Na1[g][0] = -2 * L3x; s.Na1[g][1] = -2 * L1x; s.Na1[g][2] = -2 * L2x;
		N2[0] = 4 * L1y * L1 - L1y; N2[1] = 4 * L2y * L2 - L2y; N2[2] = 4 * L3y * L3 - L3y;
		N2[3] = 4 * L1y * L2 + 4 * L1 * L2y; N2[4] = 4 * L2y * L3 + 4 * L2 * L3y; N2[5] = 4 * L3y * L1 + 4 * L3 * L1y;
		s.Na2[g][0] = -2 * L3y; s.Na2[g][1] = -2 * L1y; s.Na2[g][2] = -2 * L2y;
		s.Q_i[g] = eye3();                                                   // :2357-2362
		s.zx1_i[g] = vec(1, 0, 0); s.zx2_i[g] = vec(0, 1, 0);
		s.k1_i[g] = V3(); s.k2_i[g] = V3();
	}
	s.alpha1 = A / 3.0;                                                      // :2364
	// 6-point Cowper rule, area coordinates are the constants themselves (:2185-2246)
	static const double cw[6][4] = {
		{ 0.816847572980459, 0.091576213509771, 0.091576213509771, 0.109951743655322 },
		{ 0.091576213509771, 0.816847572980459, 0.091576213509771, 0.109951743655322 },
		{ 0.091576213509771, 0.091576213509771, 0.816847572980459, 0.109951743655322 },
		{ 0.108103018168070, 0.445948490915965, 0.445948490915965, 0.223381589678011 },
		{ 0.445948490915965, 0.108103018168070, 0.445948490915965, 0.223381589678011 },
		{ 0.445948490915965, 0.445948490915965, 0.108103018168070, 0.223381589678011 } };
	for (int a = 0; a < 6; a++) s.grav[a] = 0.0;
	for (int g = 0; g < 6; g++)
	{
		double L1 = cw[g][0], L2 = cw[g][1], L3 = cw[g][2], w = A * cw[g][3];
		double N4[6] = { (2 * L1 - 1) * L1, (2 * L2 - 1) * L2, (2 * L3 - 1) * L3, 4 * L1 * L2, 4 * L2 * L3, 4 * L3 * L1 };
		for (int a = 0; a < 6; a++) s.grav[a] += w * N4[a];
		for (int a = 0; a < 6; a++) s.N4u[g][a] = N4[a];                    // :2261-2269
		s.N4a[g][0] = 1 - 2 * L3; s.N4a[g][1] = 1 - 2 * L1; s.N4a[g][2] = 1 - 2 * L2;
		s.w4[g] = w;
	}
	s.coef1 = t * rho;                                                       // :2022-2024
	s.coef2 = (1.0 / 12.0) * t * t * t * rho;
	s.coef3 = rho * t * A / (3 * 3.1415926535897932384626433832795);
	for (int g = 0; g < 3; g++) s.alpha_i[g] = V3();
	s.CR.assign(27 * 27, 0.0);
}

// Shell_1.cpp:899-1332
void shell_mount(ShellEl& s, const int* nd)
{
	Mx<27, 27> K; Mx<27, 1> F;
	s.energy = 0.0;
	const V3 e1l = vec(1, 0, 0), e2l = vec(0, 1, 0), e3l = vec(0, 0, 1);
	const M3 I3 = eye3();
	Mx<27, 27> T;
	for (int b = 0; b < 9; b++) put(T, 3 * b, 3 * b, s.T3);
	for (int g = 0; g < 3; g++)
	{
		V3 u_d1, u_d2, a_d, a_d1, a_d2;
		for (int k = 0; k < 3; k++)
		{
			for (int a = 0; a < 6; a++)
			{
				double d = W.disp[6 * (size_t)(nd[a] - 1) + k];
				u_d1[k] += d * s.Nu1[g][a];
				u_d2[k] += d * s.Nu2[g][a];
			}
			for (int a = 0; a < 3; a++)
			{
				double r = W.disp[6 * (size_t)(nd[3 + a] - 1) + 3 + k];
				a_d[k] += r * s.Na[g][a];
				a_d1[k] += r * s.Na1[g][a];
				a_d2[k] += r * s.Na2[g][a];
			}
		}
		u_d1 = s.T3 * u_d1; u_d2 = s.T3 * u_d2;                              // :993-998
		a_d = s.T3 * a_d; a_d1 = s.T3 * a_d1; a_d2 = s.T3 * a_d2;

		double alpha = norm(a_d);                                           // :1001-1014
		M3 A = skew(a_d);
		double gg = 4.0 / (4.0 + alpha * alpha);
		M3 Qd = I3 + gg * (A + 0.5 * (A * A));
		M3 Xi = gg * (I3 + 0.5 * A);
		M3 Xi1 = (-0.5 * gg) * (dot(a_d, a_d1) * Xi - skew(a_d1));
		M3 Xi2 = (-0.5 * gg) * (dot(a_d, a_d2) * Xi - skew(a_d2));
		V3 z1 = u_d1 + s.zx1_i[g], z2 = u_d2 + s.zx2_i[g];
		M3 Z1 = skew(z1), Z2 = skew(z2);
		M3 Q = Qd * s.Q_i[g];
		M3 Qt = tr(Q);
		V3 eta1 = Qt * z1 - e1l, eta2 = Qt * z2 - e2l;                      // :1017-1020
		V3 kap1 = tr(s.Q_i[g]) * (tr(Xi) * a_d1) + s.k1_i[g];
		V3 kap2 = tr(s.Q_i[g]) * (tr(Xi) * a_d2) + s.k2_i[g];
		s.Q_d[g] = Qd; s.Xi_d[g] = Xi; s.a_x1[g] = a_d1; s.a_x2[g] = a_d2; s.u_x1[g] = u_d1; s.u_x2[g] = u_d2;
		V3 u_d;                                                              // :906-923, 993
		for (int k = 0; k < 3; k++)
			for (int a = 0; a < 6; a++) u_d[k] += W.disp[6 * (size_t)(nd[a] - 1) + k] * s.Nu[g][a];
		s.u_d[g] = s.T3 * u_d; s.a_d[g] = a_d;

		// thickness integration, Simo-Ciarlet plane-stress neo-Hookean (:1056-1163)
		V3 n1, n2, m1, m2;
		M3 dn1e1, dn1e2, dn1k1, dn1k2, dn2e1, dn2e2, dn2k1, dn2k2;
		M3 dm1e1, dm1e2, dm1k1, dm1k2, dm2e1, dm2e2, dm2k1, dm2k2;
		double psi_t = 0.0;
		const double lam = s.lambda, mu = s.mu, jac = s.thick / 2.0;
		const M3 E3 = skew(e3l);
		static const double gp[3] = { -0.77459666924148337703585307995648, 0.0, +0.77459666924148337703585307995648 };
		static const double gw[3] = { 0.55555555555555555555555555555556, 0.88888888888888888888888888888889, 0.55555555555555555555555555555556 };
		for (int q = 0; q < 3; q++)
		{
			double zeta = s.thick * gp[q] / 2.0, w = gw[q];
			V3 ga1 = eta1 + zeta * cross(kap1, e3l), ga2 = eta2 + zeta * cross(kap2, e3l);
			double g11 = ga1[0], g12 = ga1[1], g13 = ga1[2], g21 = ga2[0], g22 = ga2[1], g23 = ga2[2];
			double jb = (1.0 + g11) * (1.0 + g22) - g12 * g21;
			double v = (lam * (jb * jb * jb - 1.0) + 2.0 * mu * (jb - 1.0)) / (lam * jb * jb * jb + 2.0 * mu * jb);
			double dv = ((lam + 2.0 * mu) * (3.0 * lam * jb * jb + 2.0 * mu)) / (jb * jb * (lam * jb * jb + 2.0 * mu) * (lam * jb * jb + 2.0 * mu));
			V3 t1 = vec(mu * v * (1.0 + g22) + mu * (g11 - g22), mu * v * (-g21) + mu * (g12 + g21), 0 + mu * g13);
			V3 t2 = vec(mu * v * (-g12) + mu * (g12 + g21), mu * v * (1.0 + g11) + mu * (g22 - g11), 0 + mu * g23);
			M3 C11, C22, C12;
			C11(0, 0) = mu * ((1.0 + g22) * (1.0 + g22) * dv + 1.0);
			C11(0, 1) = -mu * (1.0 + g22) * g21 * dv; C11(1, 0) = C11(0, 1);
			C11(1, 1) = mu * (g21 * g21 * dv + 1.0); C11(2, 2) = mu;
			C22(0, 0) = mu * (g12 * g12 * dv + 1.0);
			C22(0, 1) = -mu * ((1.0 + g11) * g12 * dv); C22(1, 0) = C22(0, 1);
			C22(1, 1) = mu * ((1.0 + g11) * (1.0 + g11) * dv + 1.0); C22(2, 2) = mu;
			C12(0, 0) = -mu * ((1.0 + g22) * g12 * dv);
			C12(0, 1) = mu * (v - 1.0 + (1.0 + g11) * (1.0 + g22) * dv);
			C12(1, 0) = mu * (1.0 - v + g12 * g21 * dv);
			C12(1, 1) = -mu * ((1.0 + g11) * g21 * dv);
			M3 C21 = tr(C12);
			double wj = w * jac;
			dn1e1 = dn1e1 + wj * C11; dn1e2 = dn1e2 + wj * C12;
			dn1k1 = dn1k1 - (wj * zeta) * (C11 * E3); dn1k2 = dn1k2 - (wj * zeta) * (C12 * E3);
			dn2e1 = dn2e1 + wj * C21; dn2e2 = dn2e2 + wj * C22;
			dn2k1 = dn2k1 - (wj * zeta) * (C21 * E3); dn2k2 = dn2k2 - (wj * zeta) * (C22 * E3);
			dm1e1 = dm1e1 + (wj * zeta) * (E3 * C11); dm1e2 = dm1e2 + (wj * zeta) * (E3 * C12);
			dm1k1 = dm1k1 - (wj * zeta * zeta) * (E3 * C11 * E3); dm1k2 = dm1k2 - (wj * zeta * zeta) * (E3 * C12 * E3);
			dm2e1 = dm2e1 + (wj * zeta) * (E3 * C21); dm2e2 = dm2e2 + (wj * zeta) * (E3 * C22);
			dm2k1 = dm2k1 - (wj * zeta * zeta) * (E3 * C21 * E3); dm2k2 = dm2k2 - (wj * zeta * zeta) * (E3 * C22 * E3);
			n1 = n1 + wj * t1; n2 = n2 + wj * t2;
			m1 = m1 + (wj * zeta) * cross(e3l, t1); m2 = m2 + (wj * zeta) * cross(e3l, t2);
			double g33 = std::sqrt((lam + 2.0 * mu) / (lam * jb * jb + 2.0 * mu)) - 1.0;
			double jF = jb * (1 + g33);
			double I1 = (1.0 + g11) * (1.0 + g11) + g12 * g12 + g13 * g13 + g21 * g21 + (1.0 + g22) * (1.0 + g22) + g23 * g23 + (1.0 + g33) * (1.0 + g33);
			double psi = 0.5 * lam * (0.5 * (jF * jF - 1.0) - std::log(jF)) + 0.5 * mu * (I1 - 3.0 - 2.0 * std::log(jF));
			psi_t += wj * psi;
		}
		Mx<12, 12> D;                                                        // :1171-1226
		put(D, 0, 0, dn1e1); put(D, 0, 3, dn1k1); put(D, 0, 6, dn1e2); put(D, 0, 9, dn1k2);
		put(D, 3, 0, dm1e1); put(D, 3, 3, dm1k1); put(D, 3, 6, dm1e2); put(D, 3, 9, dm1k2);
		put(D, 6, 0, dn2e1); put(D, 6, 3, dn2k1); put(D, 6, 6, dn2e2); put(D, 6, 9, dn2k2);
		put(D, 9, 0, dm2e1); put(D, 9, 3, dm2k1); put(D, 9, 6, dm2e2); put(D, 9, 9, dm2k2);
		D(5, 5) = s.stiff_drill; D(11, 11) = s.stiff_drill;
		m1[2] = s.stiff_drill * kap1[2]; m2[2] = s.stiff_drill * kap2[2];
		{
			const V3* keep[8] = { &eta1, &eta2, &kap1, &kap2, &n1, &n2, &m1, &m2 };
			for (int k = 0; k < 8; k++) for (int i = 0; i < 3; i++) s.res[g][3 * k + i] = (*keep[k])[i];
		}
		Mx<12, 1> sig;
		for (int i = 0; i < 3; i++) { sig[i] = n1[i]; sig[3 + i] = m1[i]; sig[6 + i] = n2[i]; sig[9 + i] = m2[i]; }

		Mx<12, 15> Psi;                                                      // :1239-1270
		put(Psi, 0, 0, Qt); put(Psi, 6, 6, Qt);
		M3 QtXi = Qt * Xi;
		put(Psi, 3, 3, QtXi); put(Psi, 9, 9, QtXi);
		put(Psi, 0, 12, Qt * Z1 * Xi); put(Psi, 3, 12, Qt * Xi1);
		put(Psi, 6, 12, Qt * Z2 * Xi); put(Psi, 9, 12, Qt * Xi2);

		Mx<15, 27> dN;                                                       // :2118-2180
		for (int a = 0; a < 6; a++)
			for (int k = 0; k < 3; k++) { dN(k, 3 * a + k) = s.Nu1[g][a]; dN(6 + k, 3 * a + k) = s.Nu2[g][a]; }
		for (int a = 0; a < 3; a++)
			for (int k = 0; k < 3; k++)
			{
				dN(3 + k, 18 + 3 * a + k) = s.Na1[g][a];
				dN(9 + k, 18 + 3 * a + k) = s.Na2[g][a];
				dN(12 + k, 18 + 3 * a + k) = s.Na[g][a];
			}
		K = K + s.alpha1 * (tr(dN) * ((((tr(Psi)) * D) * Psi) * dN));         // :1273

		V3 sn1 = Q * n1, sn2 = Q * n2, sm1 = Q * m1, sm2 = Q * m2;           // :1277-1302
		M3 VZ1n1 = Vop(a_d, Z1 * sn1, alpha), VZ2n2 = Vop(a_d, Z2 * sn2, alpha);
		M3 Vm1 = Vop(a_d, sm1, alpha), Vm2 = Vop(a_d, sm2, alpha);
		M3 dV1 = dVop(a_d, a_d1, sm1, alpha), dV2 = dVop(a_d, a_d2, sm2, alpha);
		M3 Gua1 = (-1.0) * (skew(sn1) * Xi), Gua2 = (-1.0) * (skew(sn2) * Xi);
		M3 Gaa1 = tr(Xi) * (Z1 * skew(sn1)) * Xi - VZ1n1 + dV1 - tr(Xi1) * (skew(sm1) * Xi);
		M3 Gaa2 = tr(Xi) * (Z2 * skew(sn2)) * Xi - VZ2n2 + dV2 - tr(Xi2) * (skew(sm2) * Xi);
		Mx<15, 15> G;                                                        // :1305-1320
		put(G, 0, 12, Gua1); put(G, 3, 12, tr(Vm1)); put(G, 6, 12, Gua2); put(G, 9, 12, tr(Vm2));
		put(G, 12, 0, tr(Gua1)); put(G, 12, 3, Vm1); put(G, 12, 6, tr(Gua2)); put(G, 12, 9, Vm2);
		put(G, 12, 12, Gaa1 + Gaa2);
		K = K + s.alpha1 * (tr(dN) * (G * dN));                              // :1322
		F = F + s.alpha1 * (tr(dN) * (tr(Psi) * sig));                       // :1324
		s.energy += s.alpha1 * psi_t;                                        // :1326
	}
	K = (tr(T) * K) * T;                                                     // :1330-1331
	F = tr(T) * F;
	for (int i = 0; i < 27; i++) { s.Fint[i] = F[i]; for (int j = 0; j < 27; j++) s.K[i * 27 + j] = K(i, j); }
}

// Shell_1.cpp:1335-1389 -- self-weight is applied TWICE (both blocks execute)
void shell_loads(ShellEl& s, double lfac)
{
	double e[27];
	for (int i = 0; i < 27; i++) e[i] = 0.0;
	if (W.g_on)
		for (int rep = 0; rep < 2; rep++)
		{
			double mult = lfac * 1.0 * s.rho * s.thick;
			for (int a = 0; a < 6; a++)
				for (int k = 0; k < 3; k++) e[3 * a + k] += s.grav[a] * (mult * W.g[k]);
		}
	for (int i = 0; i < 27; i++) s.P[i] = s.Fint[i] - e[i];
}

// Shell_1.cpp:1650-1664
void shell_commit(ShellEl& s)
{
	for (int g = 0; g < 3; g++)
	{
		s.k1_i[g] = tr(s.Q_i[g]) * (tr(s.Xi_d[g]) * s.a_x1[g]) + s.k1_i[g];
		s.k2_i[g] = tr(s.Q_i[g]) * (tr(s.Xi_d[g]) * s.a_x2[g]) + s.k2_i[g];
		s.Q_i[g] = s.Q_d[g] * s.Q_i[g];
		s.zx1_i[g] = s.u_x1[g] + s.zx1_i[g];
		s.zx2_i[g] = s.u_x2[g] + s.zx2_i[g];
		s.alpha_i[g] = (4.0 / (4.0 - dot(s.a_d[g], s.alpha_i[g]))) * (s.a_d[g] + s.alpha_i[g] + 0.5 * cross(s.a_d[g], s.alpha_i[g]));   // :1659-1660
	}
}

// ------------------------------------------------------------------------
// Beam_1
// ------------------------------------------------------------------------
// Beam_1.cpp:501-692 (Hooke + plain section branch), :1385-1427
void beam_precalc(BeamEl& b, const int* nd, const double* hk, const double* sc, const double* cs, double T0)
{
	double E = hk[0], nu = hk[1], rho = hk[2];
	double G = E / (2 * (1 + nu)), sf = 1.0;
	double A = sc[0], I1 = sc[1], I2 = sc[2], I12 = sc[3], It = sc[5];
	b.D = Mx<6, 6>();
	b.D(0, 0) = sf * G * A; b.D(1, 1) = sf * G * A; b.D(2, 2) = E * A;
	b.D(3, 3) = E * I1; b.D(4, 4) = E * I2; b.D(3, 4) = E * I12; b.D(4, 3) = E * I12; b.D(5, 5) = G * It;
	b.rhoA = rho * A;
	b.energy_on = true;
	b.Mr = M3(); b.Jr = M3(); b.br = V3();                                   // :582-596
	b.Mr(0, 0) = rho * A; b.Mr(1, 1) = rho * A; b.Mr(2, 2) = rho * A;
	b.Jr(0, 0) = rho * I1; b.Jr(1, 1) = rho * I2; b.Jr(2, 2) = rho * sc[4]; b.Jr(0, 1) = rho * I12; b.Jr(1, 0) = rho * I12;
	b.alpha_i[0] = V3(); b.alpha_i[1] = V3();
	b.CR.assign(18 * 18, 0.0);
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) b.T3(i, j) = cs[3 * i + j];
	V3 e3 = Xv(nd[2]) - Xv(nd[0]);
	e3 = (1.0 / norm(e3)) * e3;
	b.e3r = b.T3 * e3;                                                       // :609-614
	double du0 = T0 / b.D(2, 2);                                             // :616-621
	V3 d = Xv(nd[2]) - Xv(nd[0]);
	double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	b.length = len / (1.0 + du0);
	b.jac = b.length / 2.0;
	for (int g = 0; g < 2; g++)
	{
		double xi = g == 0 ? -0.577350269189626 : +0.577350269189626;
		b.N[g][0] = 0.5 * xi * (xi - 1.0); b.N[g][1] = 1.0 - xi * xi; b.N[g][2] = 0.5 * xi * (1.0 + xi);
		b.dN[g][0] = (1.0 / b.jac) * (xi - 0.5); b.dN[g][1] = (1.0 / b.jac) * (-2.0 * xi); b.dN[g][2] = (1.0 / b.jac) * (0.5 + xi);
		b.Q_i[g] = eye3(); b.dz_i[g] = vec(0, 0, 1.0 + du0); b.k_i[g] = V3();
	}
}

// Pipe_1::PreCalc (Pipe_1.cpp:1106-1180): D = diag(GA, GA, EA, EI, EI, GJ) from the PipeSection, mass per
// unit length Rho, plain chord length, no pre-tension; ps = EA EI GJ GA Rho CDt CDn CAt CAn De Di
void pipe_precalc(BeamEl& b, const int* nd, const double* ps, const double* cs)
{
	const double unit_hooke[3] = { 1.0, 0.0, 0.0 }, no_section[6] = { 1.0, 0, 0, 0, 0, 0 };
	beam_precalc(b, nd, unit_hooke, no_section, cs, 0.0);                    // frame, length, shape functions, state
	b.D = Mx<6, 6>();
	b.D(0, 0) = ps[3]; b.D(1, 1) = ps[3]; b.D(2, 2) = ps[0]; b.D(3, 3) = ps[1]; b.D(4, 4) = ps[1]; b.D(5, 5) = ps[2];
	b.rhoA = ps[4];
	b.energy_on = false;
	b.Aint = 3.1415926535897932384626433832795 * ps[10] * ps[10] / 4.0;      // Pipe_1.cpp:1151 (PI as in the reference's Matrix.h)
	// inertia per unit length (Pipe_1.cpp:1131-1144, as written: radii squared are SUBTRACTED), no ocean data:
	// Mr = Rho I in MountMass / MountMassModal (Pipe_1.cpp:1580-1583, 1799-1803)
	const double rr = (ps[9] / 2.0) * (ps[9] / 2.0) - (ps[10] / 2.0) * (ps[10] / 2.0);
	b.Mr = M3(); b.Jr = M3(); b.br = V3();
	b.Mr(0, 0) = ps[4]; b.Mr(1, 1) = ps[4]; b.Mr(2, 2) = ps[4];
	b.Jr(0, 0) = (ps[4] * rr / 4.0); b.Jr(1, 1) = (ps[4] * rr / 4.0); b.Jr(2, 2) = (ps[4] * rr / 2.0);
}

// Beam_1.cpp:695-835
void beam_mount(BeamEl& b, const int* nd)
{
	Mx<18, 18> K; Mx<18, 1> F;
	b.energy = 0.0;
	const M3 I3 = eye3();
	Mx<18, 18> T;
	for (int k = 0; k < 6; k++) put(T, 3 * k, 3 * k, b.T3);
	for (int g = 0; g < 2; g++)
	{
		V3 a_d, da_d, du_d;
		for (int k = 0; k < 3; k++)
			for (int a = 0; a < 3; a++)
			{
				const double* d = &W.disp[6 * (size_t)(nd[a] - 1)];
				a_d[k] += d[3 + k] * b.N[g][a];
				da_d[k] += d[3 + k] * b.dN[g][a];
				du_d[k] += d[k] * b.dN[g][a];
			}
		a_d = b.T3 * a_d; da_d = b.T3 * da_d; du_d = b.T3 * du_d;            // :743-746
		double alpha = norm(a_d);
		M3 A = skew(a_d);
		double gg = 4.0 / (4.0 + alpha * alpha);
		M3 Qd = I3 + gg * (A + 0.5 * (A * A));
		M3 Xi = gg * (I3 + 0.5 * A);
		M3 dXi = (-0.5 * gg) * (dot(a_d, da_d) * Xi - skew(da_d));
		V3 dz = du_d + b.dz_i[g];
		M3 dZ = skew(dz);
		M3 Q = Qd * b.Q_i[g];
		M3 Qt = tr(Q);
		Mx<6, 6> B1; put(B1, 0, 0, Qt); put(B1, 3, 3, Qt);                  // :758-774
		Mx<6, 9> B2; put(B2, 0, 0, I3); put(B2, 0, 6, dZ * Xi); put(B2, 3, 3, Xi); put(B2, 3, 6, dXi);
		Mx<6, 9> B = B1 * B2;
		Mx<9, 18> dN;                                                        // :641-669
		for (int a = 0; a < 3; a++)
			for (int k = 0; k < 3; k++)
			{
				dN(k, 6 * a + k) = b.dN[g][a];
				dN(3 + k, 6 * a + 3 + k) = b.dN[g][a];
				dN(6 + k, 6 * a + 3 + k) = b.N[g][a];
			}
		Mx<18, 18> Kc = tr(dN) * ((((tr(B)) * b.D) * B) * dN);               // :776
		V3 eta = Qt * dz - b.e3r;                                            // :778-797
		V3 kap = tr(b.Q_i[g]) * (tr(Xi) * da_d) + b.k_i[g];
		Mx<6, 1> eps; for (int i = 0; i < 3; i++) { eps[i] = eta[i]; eps[3 + i] = kap[i]; }
		Mx<6, 1> sig = b.D * eps;
		V3 nr = vec(sig[0], sig[1], sig[2]), mr = vec(sig[3], sig[4], sig[5]);
		V3 n = Q * nr, m = Q * mr;
		M3 Vdzn = Vop(a_d, dZ * n, alpha), Vm = Vop(a_d, m, alpha), dVm = dVop(a_d, da_d, m, alpha);   // :799-822
		M3 Gua = (-1.0) * (skew(n) * Xi);
		M3 Gau = tr(Xi) * skew(n);
		M3 Gaa = tr(Xi) * (dZ * skew(n)) * Xi - Vdzn + dVm - tr(dXi) * (skew(m) * Xi);
		M3 Gada = Vm - tr(Xi) * (skew(m) * Xi);
		Mx<9, 9> G;
		put(G, 0, 6, Gua); put(G, 6, 0, Gau); put(G, 6, 6, Gaa); put(G, 3, 6, Gada); put(G, 6, 3, Vm);
		Mx<18, 18> Kg = tr(dN) * (G * dN);                                   // :824-828
		K = K + (1.0 * b.jac) * (Kc + Kg);
		F = F + (1.0 * b.jac) * ((tr(dN) * tr(B)) * sig);
		if (b.energy_on) b.energy += 0.5 * (1.0 * b.jac) * (tr(sig) * eps)[0];
		for (int i = 0; i < 6; i++) { b.res[g][i] = eps[i]; b.res[g][6 + i] = sig[i]; }
		b.Q_d[g] = Qd; b.dz[g] = dz; b.kr[g] = kap;
		b.Xi_d[g] = Xi; b.dXi_d[g] = dXi; b.g_last = gg;
		V3 u_d;                                                              // :722-731, 743
		for (int k = 0; k < 3; k++)
			for (int a = 0; a < 3; a++) u_d[k] += W.disp[6 * (size_t)(nd[a] - 1) + k] * b.N[g][a];
		b.u_d[g] = b.T3 * u_d; b.a_d[g] = a_d;
	}
	K = (tr(T) * K) * T;                                                     // :833-834
	F = tr(T) * F;
	for (int i = 0; i < 18; i++) { b.Fint[i] = F[i]; for (int j = 0; j < 18; j++) b.K[i * 18 + j] = K(i, j); }
}

// Beam_1.cpp:838-904 (gravity only; wind/BEM branch out of scope)
void beam_loads(BeamEl& b, double lfac)
{
	double e[18];
	for (int i = 0; i < 18; i++) e[i] = 0.0;
	if (W.g_on)
	{
		double mult = lfac * 1.0 * 1.0 * b.jac * b.rhoA;
		for (int g = 0; g < 2; g++)
			for (int a = 0; a < 3; a++)
				for (int k = 0; k < 3; k++) e[6 * a + k] += mult * b.N[g][a] * W.g[k];
	}
	for (int i = 0; i < 18; i++) b.P[i] = b.Fint[i] - e[i];
}

// Pipe_1::MountPipeSpecialLoads (Pipe_1.cpp:1443-1494), called by PipeLoad::Mount (PipeLoad.cpp:117-133) during
// MountLoads: internal pressure p0i on the current configuration of the pipe axis.  Uses what Mount left in the
// element -- including the scalar g of the LAST Gauss point for both points, as the reference does.
void pipe_pressure(BeamEl& b, double p0i)
{
	Mx<18, 18> T;
	for (int k = 0; k < 6; k++) put(T, 3 * k, 3 * k, b.T3);
	Mx<18, 18> K; Mx<18, 1> P;
	for (int i = 0; i < 18; i++) { P[i] = b.P[i]; for (int j = 0; j < 18; j++) K(i, j) = b.K[i * 18 + j]; }
	const double mult = 1.0 * b.jac;
	for (int g = 0; g < 2; g++)
	{
		const M3 Q = b.Q_d[g] * b.Q_i[g];
		const M3& Xi = b.Xi_d[g];
		const V3 kip = Q * b.kr[g], e3ip = Q * b.e3r;
		const V3 tf = (-p0i * b.Aint) * cross(kip, e3ip);
		const V3 tm = (-p0i * b.Aint) * (tr(Xi) * cross(b.dz[g], e3ip));
		Mx<6, 1> tl; for (int i = 0; i < 3; i++) { tl[i] = tf[i]; tl[3 + i] = tm[i]; }
		const M3 E3 = skew(e3ip), Kip = skew(kip), dZ = skew(b.dz[g]);
		const V3 c = cross(e3ip, b.dz[g]);
		const V3 Xtc = tr(Xi) * c;
		M3 outer;
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) outer(i, j) = Xtc[i] * b.a_d[g][j];
		const M3 O1 = (-0.5 * b.g_last) * (outer - skew(c));
		const M3 K1ua = Kip * E3 * Xi + E3 * b.dXi_d[g] - 1.0 * (E3 * skew(kip) * Xi);
		const M3 K1aa = O1 + tr(Xi) * dZ * E3 * Xi;
		const M3 K2ua = E3 * Xi;
		const M3 K2au = tr(Xi) * E3;
		Mx<6, 12> Kext;
		put(Kext, 0, 3, K1ua); put(Kext, 3, 3, K1aa); put(Kext, 0, 9, K2ua); put(Kext, 3, 6, K2au);
		Mx<6, 18> N; Mx<12, 18> UN;                                            // Pipe_1.cpp:1213-1273
		for (int a = 0; a < 3; a++)
			for (int k = 0; k < 6; k++) { N(k, 6 * a + k) = b.N[g][a]; UN(k, 6 * a + k) = b.N[g][a]; UN(6 + k, 6 * a + k) = b.dN[g][a]; }
		const Mx<18, 1> fl = mult * (tr(T) * (tr(N) * tl));
		P = P - fl;
		K = K - (mult * p0i * b.Aint) * (tr(T) * ((tr(N) * Kext) * UN) * T);
	}
	for (int i = 0; i < 18; i++) { b.P[i] = P[i]; for (int j = 0; j < 18; j++) b.K[i * 18 + j] = K(i, j); }
}

// Beam_1.cpp:1494-1506
void beam_commit(BeamEl& b)
{
	for (int g = 0; g < 2; g++)
	{
		b.Q_i[g] = b.Q_d[g] * b.Q_i[g];
		b.k_i[g] = b.kr[g];
		b.dz_i[g] = b.dz[g];
		b.alpha_i[g] = (4.0 / (4.0 - dot(b.a_d[g], b.alpha_i[g]))) * (b.a_d[g] + b.alpha_i[g] + 0.5 * cross(b.a_d[g], b.alpha_i[g]));   // :1502-1503
	}
}


// ------------------------------------------------------------------------
// Newmark dynamics: MountMass / MountDamping / MountDyn of Beam_1 and Shell_1
// (Beam_1.cpp:1537-1673, Shell_1.cpp:2367-2547).  The reference evaluates the
// Gauss-point inertial pseudo-forces dT and their tangent DdT with AceGen-
// generated code (Beam_1.cpp:1781-2361, Shell_1.cpp:2568-2890).  Restated here
// from the formulation that code implements (Newmark in the tangent space of
// the incremental rotation, Dynamic.cpp:480-556):
//   Q = Q(alpha_d) Q(alpha_i),  Xi = Xi(alpha_d)
//   omega  = Q(alpha_d) (a4 alpha_d + a5 omega_i + a6 domega_i)
//   domega = Q(alpha_d) (a1 alpha_d - a2 omega_i - a3 domega_i)
//   ddu    = a1 u_d - a2 du_i - a3 ddu_i
//   Beam_1 : M = Q Mr Q^T, J = Q Jr Q^T, b = Q br
//            f  = M (ddu + domega x b + omega x (omega x b))
//            mu = M (b x ddu) + J domega + omega x (J omega)
//   Shell_1: e3 = Q e3r,  f = coef1 ddu,  mu = coef2 e3 x (domega x e3 + omega x (omega x e3))
//   dT = [f ; Xi^T mu],   DdT = d dT / d(u_d, alpha_d)
// The tangent is obtained by forward-mode differentiation of dT (what AceGen
// does symbolically), so DdT is the exact derivative of the same function.
// ------------------------------------------------------------------------
struct Du
{
	double v, d[6];
	Du(double x = 0.0) : v(x) { for (int i = 0; i < 6; i++) d[i] = 0.0; }
};
Du operator+(Du a, const Du& b) { a.v += b.v; for (int i = 0; i < 6; i++) a.d[i] += b.d[i]; return a; }
Du operator-(Du a, const Du& b) { a.v -= b.v; for (int i = 0; i < 6; i++) a.d[i] -= b.d[i]; return a; }
Du operator-(Du a) { a.v = -a.v; for (int i = 0; i < 6; i++) a.d[i] = -a.d[i]; return a; }
Du operator*(const Du& a, const Du& b) { Du o(a.v * b.v); for (int i = 0; i < 6; i++) o.d[i] = a.d[i] * b.v + a.v * b.d[i]; return o; }
Du operator/(const Du& a, const Du& b) { Du o(a.v / b.v); for (int i = 0; i < 6; i++) o.d[i] = (a.d[i] - o.v * b.d[i]) / b.v; return o; }

template <class T> void rot_Q(const T* a, T Q[3][3])             // I + g (A + A A / 2), g = 4 / (4 + |a|^2)
{
	T g = T(4.0) / (T(4.0) + a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
	T A[3][3] = { { T(0.0), -a[2], a[1] }, { a[2], T(0.0), -a[0] }, { -a[1], a[0], T(0.0) } };
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++)
		{
			T s2(0.0);
			for (int k = 0; k < 3; k++) s2 = s2 + A[i][k] * A[k][j];
			Q[i][j] = T(i == j ? 1.0 : 0.0) + g * (A[i][j] + T(0.5) * s2);
		}
}
template <class T> void rot_Xi(const T* a, T X[3][3])            // g (I + A / 2)
{
	T g = T(4.0) / (T(4.0) + a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
	T A[3][3] = { { T(0.0), -a[2], a[1] }, { a[2], T(0.0), -a[0] }, { -a[1], a[0], T(0.0) } };
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) X[i][j] = g * (T(i == j ? 1.0 : 0.0) + T(0.5) * A[i][j]);
}
template <class T> void crossT(const T* a, const T* b, T* o)
{ o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
template <class T> void mulT(const T A[3][3], const T* x, T* o)
{ for (int i = 0; i < 3; i++) o[i] = A[i][0] * x[0] + A[i][1] * x[1] + A[i][2] * x[2]; }

// omega, domega, ddu, Q, Xi shared by both elements
template <class T>
void newmark_kinematics(const V3& alpha_i, const T* ad, const T* ud, const V3& om_i, const V3& dom_i, const V3& du_i,
	const V3& ddu_i, T Q[3][3], T Xi[3][3], T* om, T* dom, T* ddu)
{
	const double a1 = W.nm[0], a2 = W.nm[1], a3 = W.nm[2], a4 = W.nm[3], a5 = W.nm[4], a6 = W.nm[5];
	T Qd[3][3], Qi[3][3], ai[3] = { T(alpha_i[0]), T(alpha_i[1]), T(alpha_i[2]) };
	rot_Q(ad, Qd); rot_Q(ai, Qi); rot_Xi(ad, Xi);
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) Q[i][j] = Qd[i][0] * Qi[0][j] + Qd[i][1] * Qi[1][j] + Qd[i][2] * Qi[2][j];
	T wl[3], dwl[3];
	for (int k = 0; k < 3; k++)
	{
		wl[k] = ad[k] * T(a4) + T(om_i[k] * a5) + T(dom_i[k] * a6);
		dwl[k] = ad[k] * T(a1) - T(om_i[k] * a2) - T(dom_i[k] * a3);
		ddu[k] = ud[k] * T(a1) - T(du_i[k] * a2) - T(ddu_i[k] * a3);
	}
	mulT(Qd, wl, om); mulT(Qd, dwl, dom);
}

template <class T>
void beam_inertia(const V3& alpha_i, const T* ad, const T* ud, const V3& om_i, const V3& dom_i, const V3& du_i,
	const V3& ddu_i, const M3& Jr, const M3& Mr, const V3& br, T* dT)
{
	T Q[3][3], Xi[3][3], w[3], dw[3], ddu[3];
	newmark_kinematics(alpha_i, ad, ud, om_i, dom_i, du_i, ddu_i, Q, Xi, w, dw, ddu);
	T M[3][3], J[3][3], b[3];
	for (int i = 0; i < 3; i++)
	{
		b[i] = Q[i][0] * T(br[0]) + Q[i][1] * T(br[1]) + Q[i][2] * T(br[2]);
		for (int j = 0; j < 3; j++)
		{
			T m(0.0), jj(0.0);
			for (int k = 0; k < 3; k++)
				for (int l = 0; l < 3; l++) { m = m + Q[i][k] * T(Mr(k, l)) * Q[j][l]; jj = jj + Q[i][k] * T(Jr(k, l)) * Q[j][l]; }
			M[i][j] = m; J[i][j] = jj;
		}
	}
	T wb[3], wwb[3], dwb[3], acc[3], f[3], bu[3], Mbu[3], Jdw[3], Jw[3], wJw[3], mu[3];
	crossT(w, b, wb); crossT(w, wb, wwb); crossT(dw, b, dwb);
	for (int k = 0; k < 3; k++) acc[k] = ddu[k] + dwb[k] + wwb[k];
	mulT(M, acc, f);
	crossT(b, ddu, bu); mulT(M, bu, Mbu); mulT(J, dw, Jdw); mulT(J, w, Jw); crossT(w, Jw, wJw);
	for (int k = 0; k < 3; k++) mu[k] = Mbu[k] + Jdw[k] + wJw[k];
	for (int k = 0; k < 3; k++)
	{
		dT[k] = f[k];
		dT[3 + k] = Xi[0][k] * mu[0] + Xi[1][k] * mu[1] + Xi[2][k] * mu[2];
	}
}

template <class T>
void shell_inertia(const V3& alpha_i, const T* ad, const T* ud, const V3& om_i, const V3& dom_i, const V3& du_i,
	const V3& ddu_i, const V3& e3r, double coef1, double coef2, T* dT)
{
	T Q[3][3], Xi[3][3], w[3], dw[3], ddu[3];
	newmark_kinematics(alpha_i, ad, ud, om_i, dom_i, du_i, ddu_i, Q, Xi, w, dw, ddu);
	T e3[3], we[3], wwe[3], dwe[3], acc[3], mu[3];
	for (int i = 0; i < 3; i++) e3[i] = Q[i][0] * T(e3r[0]) + Q[i][1] * T(e3r[1]) + Q[i][2] * T(e3r[2]);
	crossT(w, e3, we); crossT(w, we, wwe); crossT(dw, e3, dwe);
	for (int k = 0; k < 3; k++) acc[k] = dwe[k] + wwe[k];
	crossT(e3, acc, mu);
	for (int k = 0; k < 3; k++)
	{
		dT[k] = T(coef1) * ddu[k];
		dT[3 + k] = T(coef2) * (Xi[0][k] * mu[0] + Xi[1][k] * mu[1] + Xi[2][k] * mu[2]);
	}
}

// seeds u_d (directions 0-2) and alpha_d (3-5), returns dT values and DdT
template <class F>
void with_tangent(const V3& a_d, const V3& u_d, F&& eval, Mx<6, 1>& dT, Mx<6, 6>& DdT)
{
	Du ad[3], ud[3], out[6];
	for (int k = 0; k < 3; k++) { ud[k] = Du(u_d[k]); ud[k].d[k] = 1.0; ad[k] = Du(a_d[k]); ad[k].d[3 + k] = 1.0; }
	eval(ad, ud, out);
	for (int i = 0; i < 6; i++) { dT[i] = out[i].v; for (int j = 0; j < 6; j++) DdT(i, j) = out[i].d[j]; }
}

V3 nodal3(const std::vector<double>& arr, int node1, int off)
{ const double* p = &arr[6 * (size_t)(node1 - 1) + off]; return vec(p[0], p[1], p[2]); }

// Beam_1::MountMass (Beam_1.cpp:1564-1636), MountDamping (:1639-1664), MountDyn (:1667-1672)
void beam_dynamics(BeamEl& b, const int* nd, bool update_rayleigh)
{
	Mx<18, 18> T; for (int k = 0; k < 6; k++) put(T, 3 * k, 3 * k, b.T3);
	Mx<18, 18> mass; Mx<18, 1> inertial;
	for (int g = 0; g < 2; g++)
	{
		V3 om, dom, du, ddu;
		for (int a = 0; a < 3; a++)
		{
			om = om + b.N[g][a] * nodal3(W.copy_vel, nd[a], 3); dom = dom + b.N[g][a] * nodal3(W.copy_accel, nd[a], 3);
			du = du + b.N[g][a] * nodal3(W.copy_vel, nd[a], 0); ddu = ddu + b.N[g][a] * nodal3(W.copy_accel, nd[a], 0);
		}
		om = b.T3 * om; dom = b.T3 * dom; du = b.T3 * du; ddu = b.T3 * ddu;     // :1618-1622
		Mx<6, 1> dT; Mx<6, 6> DdT;
		with_tangent(b.a_d[g], b.u_d[g], [&](const Du* ad, const Du* ud, Du* out)
			{ beam_inertia(b.alpha_i[g], ad, ud, om, dom, du, ddu, b.Jr, b.Mr, b.br, out); }, dT, DdT);
		Mx<6, 18> Nm;
		for (int a = 0; a < 3; a++) for (int k = 0; k < 6; k++) Nm(k, 6 * a + k) = b.N[g][a];
		inertial = inertial + (1.0 * b.jac) * (tr(Nm) * dT);
		mass = mass + (1.0 * b.jac) * ((tr(Nm) * DdT) * Nm);
	}
	inertial = tr(T) * inertial;
	mass = (tr(T) * mass) * T;
	if (update_rayleigh)                                                         // MountMassModal :1537-1552
	{
		Mx<18, 18> mm;
		for (int g = 0; g < 2; g++)
		{
			double ai[3] = { b.alpha_i[g][0], b.alpha_i[g][1], b.alpha_i[g][2] }, Qa[3][3];
			rot_Q(ai, Qa);
			M3 Q; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q(i, j) = Qa[i][j];
			M3 M = (Q * b.Mr) * tr(Q), J = (Q * b.Jr) * tr(Q);
			V3 bb = Q * b.br;
			Mx<6, 6> MM;                                                         // EvaluateMassModal :1681-1777
			put(MM, 0, 0, M); put(MM, 3, 3, J);
			for (int i = 0; i < 3; i++)
			{
				V3 L = cross(vec(M(i, 0), M(i, 1), M(i, 2)), bb);
				for (int j = 0; j < 3; j++) { MM(3 + i, j) = L[j]; MM(i, 3 + j) = -L[j]; }
			}
			Mx<6, 18> Nm;
			for (int a = 0; a < 3; a++) for (int k = 0; k < 6; k++) Nm(k, 6 * a + k) = b.N[g][a];
			mm = mm + (1.0 * b.jac) * ((tr(Nm) * MM) * Nm);
		}
		mm = (tr(T) * mm) * T;
		for (int i = 0; i < 18 * 18; i++) b.CR[i] = W.ray_alpha * mm.a[i] + W.ray_beta * b.K[i];
	}
	double v[18], dl[18];
	for (int a = 0; a < 3; a++) for (int k = 0; k < 6; k++) v[6 * a + k] = W.vel[6 * (size_t)(nd[a] - 1) + k];
	for (int i = 0; i < 18; i++) { double sacc = 0.0; for (int j = 0; j < 18; j++) sacc += b.CR[i * 18 + j] * v[j]; dl[i] = sacc; }
	for (int i = 0; i < 18; i++)
	{
		b.P[i] = b.P[i] + inertial[i] + dl[i];
		for (int j = 0; j < 18; j++) b.K[i * 18 + j] = b.K[i * 18 + j] + mass(i, j) + W.nm[3] * b.CR[i * 18 + j];
	}
}

// Shell_1::MountMass (Shell_1.cpp:2406-2498), MountDamping (:2501-2533), MountDyn (:2536-2541)
void shell_dynamics(ShellEl& s, const int* nd, bool update_rayleigh)
{
	Mx<27, 27> T; for (int k = 0; k < 9; k++) put(T, 3 * k, 3 * k, s.T3);
	Mx<27, 27> mass; Mx<27, 1> inertial;
	const V3 e3r = vec(0, 0, 1);                                                 // e3rlocal :2006-2008
	for (int g = 0; g < 3; g++)
	{
		V3 om, dom, du, ddu;
		for (int a = 0; a < 3; a++)
		{ om = om + s.Na[g][a] * nodal3(W.copy_vel, nd[3 + a], 3); dom = dom + s.Na[g][a] * nodal3(W.copy_accel, nd[3 + a], 3); }
		for (int a = 0; a < 6; a++)
		{ du = du + s.Nu[g][a] * nodal3(W.copy_vel, nd[a], 0); ddu = ddu + s.Nu[g][a] * nodal3(W.copy_accel, nd[a], 0); }
		om = s.T3 * om; dom = s.T3 * dom; du = s.T3 * du; ddu = s.T3 * ddu;     // :2471-2474
		Mx<6, 1> dT; Mx<6, 6> DdT;
		with_tangent(s.a_d[g], s.u_d[g], [&](const Du* ad, const Du* ud, Du* out)
			{ shell_inertia(s.alpha_i[g], ad, ud, om, dom, du, ddu, e3r, s.coef1, s.coef2, out); }, dT, DdT);
		Mx<6, 27> Nm;
		for (int k = 0; k < 3; k++)
		{
			for (int a = 0; a < 6; a++) Nm(k, 3 * a + k) = s.Nu[g][a];
			for (int a = 0; a < 3; a++) Nm(3 + k, 18 + 3 * a + k) = s.Na[g][a];
		}
		inertial = inertial + s.alpha1 * (tr(Nm) * dT);
		mass = mass + s.alpha1 * ((tr(Nm) * DdT) * Nm);
	}
	inertial = tr(T) * inertial;
	mass = (tr(T) * mass) * T;
	if (update_rayleigh)                                                         // MountMassModal :2367-2396
	{
		Mx<27, 27> mm;
		for (int g = 0; g < 6; g++)
		{
			V3 ai;
			for (int a = 0; a < 3; a++) ai = ai + s.N4a[g][a] * nodal3(W.copy, nd[3 + a], 3);
			ai = s.T3 * ai;
			double aa[3] = { ai[0], ai[1], ai[2] }, Qa[3][3];
			rot_Q(aa, Qa);
			double e3[3];
			for (int i = 0; i < 3; i++) e3[i] = Qa[i][0] * e3r[0] + Qa[i][1] * e3r[1] + Qa[i][2] * e3r[2];
			Mx<6, 6> MM;                                                         // EvaluateMassModal :3171-3221
			for (int k = 0; k < 3; k++) MM(k, k) = s.coef1;
			const double q0 = e3[0] * e3[0], q1 = e3[1] * e3[1], q2 = e3[2] * e3[2], dc = -s.coef2 + s.coef3;
			MM(3, 3) = s.coef3 * q0 + s.coef2 * (q1 + q2); MM(4, 4) = s.coef3 * q1 + s.coef2 * (q0 + q2); MM(5, 5) = s.coef2 * (q1 + q0) + s.coef3 * q2;
			MM(3, 4) = e3[0] * (e3[1] * dc); MM(3, 5) = e3[0] * e3[2] * dc; MM(4, 5) = e3[2] * (e3[1] * dc);
			MM(4, 3) = MM(3, 4); MM(5, 3) = MM(3, 5); MM(5, 4) = MM(4, 5);
			Mx<6, 27> Nm;
			for (int k = 0; k < 3; k++)
			{
				for (int a = 0; a < 6; a++) Nm(k, 3 * a + k) = s.N4u[g][a];
				for (int a = 0; a < 3; a++) Nm(3 + k, 18 + 3 * a + k) = s.N4a[g][a];
			}
			mm = mm + s.w4[g] * ((tr(Nm) * MM) * Nm);
		}
		mm = (tr(T) * mm) * T;
		for (int i = 0; i < 27 * 27; i++) s.CR[i] = W.ray_alpha * mm.a[i] + W.ray_beta * s.K[i];
	}
	double v[27], dl[27];
	for (int a = 0; a < 6; a++) for (int k = 0; k < 3; k++) v[3 * a + k] = W.vel[6 * (size_t)(nd[a] - 1) + k];
	for (int a = 0; a < 3; a++) for (int k = 0; k < 3; k++) v[18 + 3 * a + k] = W.vel[6 * (size_t)(nd[3 + a] - 1) + 3 + k];
	for (int i = 0; i < 27; i++) { double sacc = 0.0; for (int j = 0; j < 27; j++) sacc += s.CR[i * 27 + j] * v[j]; dl[i] = sacc; }
	for (int i = 0; i < 27; i++)
	{
		s.P[i] = s.P[i] + inertial[i] + dl[i];
		for (int j = 0; j < 27; j++) s.K[i * 27 + j] = s.K[i * 27 + j] + mass(i, j) + W.nm[3] * s.CR[i * 27 + j];
	}
}

// Dynamic::UpdateDyn (Dynamic.cpp:480-556), node DOFs.  vel_aux / ace_aux live outside the node loop in
// the reference: a rotational DOF that is not free keeps the value left by the previous node.
void update_dyn()
{
	const double a1 = W.nm[0], a2 = W.nm[1], a3 = W.nm[2], a4 = W.nm[3], a5 = W.nm[4], a6 = W.nm[5];
	V3 vel_aux, ace_aux;
	for (int i = 0; i < W.n_nodes; i++)
	{
		const double* d = &W.disp[6 * (size_t)i]; const int* gl = &W.gls[6 * (size_t)i];
		const double* cv = &W.copy_vel[6 * (size_t)i]; const double* ca = &W.copy_accel[6 * (size_t)i];
		double* v = &W.vel[6 * (size_t)i]; double* ac = &W.accel[6 * (size_t)i];
		for (int j = 0; j < 3; j++)
			if (gl[j] > 0)
			{
				v[j] = d[j] * a4 + cv[j] * a5 + ca[j] * a6;
				ac[j] = d[j] * a1 - cv[j] * a2 - ca[j] * a3;
			}
		V3 ad = vec(d[3], d[4], d[5]);
		double al = norm(ad);
		M3 A = skew(ad);
		double g = 4.0 / (4.0 + al * al);
		M3 Qd = eye3() + g * (A + 0.5 * (A * A));
		for (int j = 3; j < 6; j++)
			if (gl[j] > 0)
			{
				vel_aux[j - 3] = d[j] * a4 + cv[j] * a5 + ca[j] * a6;
				ace_aux[j - 3] = d[j] * a1 - cv[j] * a2 - ca[j] * a3;
			}
		vel_aux = Qd * vel_aux; ace_aux = Qd * ace_aux;
		for (int j = 3; j < 6; j++)
			if (gl[j] > 0) { v[j] = vel_aux[j - 3]; ac[j] = ace_aux[j - 3]; }
	}
}

// ------------------------------------------------------------------------
// Solid_1 -- BUILDER-DEFINED (reference bodies are empty, Solid_1.cpp:148-176)
// 8-node trilinear hexahedron, total Lagrangian, St.Venant-Kirchhoff:
//   F = I + sum_a u_a (x) dN_a/dX,  E = (F^T F - I)/2,  S = lambda tr(E) I + 2 mu E
//   Fint_a = sum_gp w |J| B_a^T S,  Kt_ab = sum_gp w |J| (B_a^T C B_b + (dN_a . S dN_b) I3)
// with u_a = (copy - ref) + displacements, 2x2x2 Gauss points, Voigt order
// (11,22,33,12,23,13).  Gravity: consistent nodal load rho*g*l_factor.
// ------------------------------------------------------------------------
void solid_mount(SolidEl& s, const int* nd, double lfac)
{
	static const double sg[8][3] = { {-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1} };
	const double gp = 0.57735026918962576451;
	double Xn[8][3], un[8][3];
	for (int a = 0; a < 8; a++)
		for (int k = 0; k < 3; k++)
		{
			size_t n0 = (size_t)(nd[a] - 1);
			Xn[a][k] = W.ref[3 * n0 + k];
			un[a][k] = (W.copy[6 * n0 + k] - W.ref[3 * n0 + k]) + W.disp[6 * n0 + k];
		}
	Mx<24, 24> K; double F[24], Fe[24];
	for (int i = 0; i < 24; i++) { F[i] = 0.0; Fe[i] = 0.0; }
	s.energy = 0.0;
	const double lam = s.lambda, mu = s.mu;
	for (int q = 0; q < 8; q++)
	{
		double xi = gp * sg[q][0], et = gp * sg[q][1], ze = gp * sg[q][2];
		double N[8], dNl[8][3];
		for (int a = 0; a < 8; a++)
		{
			double sx = sg[a][0], sy = sg[a][1], sz = sg[a][2];
			N[a] = 0.125 * (1 + sx * xi) * (1 + sy * et) * (1 + sz * ze);
			dNl[a][0] = 0.125 * sx * (1 + sy * et) * (1 + sz * ze);
			dNl[a][1] = 0.125 * sy * (1 + sx * xi) * (1 + sz * ze);
			dNl[a][2] = 0.125 * sz * (1 + sx * xi) * (1 + sy * et);
		}
		M3 J;   // J(i,j) = dX_i / dxi_j
		for (int a = 0; a < 8; a++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) J(i, j) += Xn[a][i] * dNl[a][j];
		double det = J(0, 0) * (J(1, 1) * J(2, 2) - J(1, 2) * J(2, 1)) - J(0, 1) * (J(1, 0) * J(2, 2) - J(1, 2) * J(2, 0)) + J(0, 2) * (J(1, 0) * J(2, 1) - J(1, 1) * J(2, 0));
		M3 Ji;
		Ji(0, 0) = (J(1, 1) * J(2, 2) - J(1, 2) * J(2, 1)) / det; Ji(0, 1) = (J(0, 2) * J(2, 1) - J(0, 1) * J(2, 2)) / det; Ji(0, 2) = (J(0, 1) * J(1, 2) - J(0, 2) * J(1, 1)) / det;
		Ji(1, 0) = (J(1, 2) * J(2, 0) - J(1, 0) * J(2, 2)) / det; Ji(1, 1) = (J(0, 0) * J(2, 2) - J(0, 2) * J(2, 0)) / det; Ji(1, 2) = (J(0, 2) * J(1, 0) - J(0, 0) * J(1, 2)) / det;
		Ji(2, 0) = (J(1, 0) * J(2, 1) - J(1, 1) * J(2, 0)) / det; Ji(2, 1) = (J(0, 1) * J(2, 0) - J(0, 0) * J(2, 1)) / det; Ji(2, 2) = (J(0, 0) * J(1, 1) - J(0, 1) * J(1, 0)) / det;
		double dN[8][3];   // dN_a/dX_j = sum_k dNl[a][k] * Ji(k,j)
		for (int a = 0; a < 8; a++) for (int j = 0; j < 3; j++) { double v = 0.0; for (int k = 0; k < 3; k++) v += dNl[a][k] * Ji(k, j); dN[a][j] = v; }
		M3 Fd = eye3();
		for (int a = 0; a < 8; a++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Fd(i, j) += un[a][i] * dN[a][j];
		M3 Cg = tr(Fd) * Fd;
		double Ev[6] = { 0.5 * (Cg(0, 0) - 1.0), 0.5 * (Cg(1, 1) - 1.0), 0.5 * (Cg(2, 2) - 1.0), Cg(0, 1), Cg(1, 2), Cg(0, 2) };   // engineering shears
		double trE = Ev[0] + Ev[1] + Ev[2];
		double Sv[6] = { lam * trE + 2 * mu * Ev[0], lam * trE + 2 * mu * Ev[1], lam * trE + 2 * mu * Ev[2], mu * Ev[3], mu * Ev[4], mu * Ev[5] };
		M3 S; S(0, 0) = Sv[0]; S(1, 1) = Sv[1]; S(2, 2) = Sv[2]; S(0, 1) = S(1, 0) = Sv[3]; S(1, 2) = S(2, 1) = Sv[4]; S(0, 2) = S(2, 0) = Sv[5];
		double Cm[6][6];
		for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Cm[i][j] = 0.0;
		for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Cm[i][j] = lam; Cm[i][i] = lam + 2 * mu; Cm[3 + i][3 + i] = mu; }
		double B[8][6][3];
		for (int a = 0; a < 8; a++)
			for (int k = 0; k < 3; k++)
			{
				B[a][0][k] = Fd(k, 0) * dN[a][0];
				B[a][1][k] = Fd(k, 1) * dN[a][1];
				B[a][2][k] = Fd(k, 2) * dN[a][2];
				B[a][3][k] = Fd(k, 0) * dN[a][1] + Fd(k, 1) * dN[a][0];
				B[a][4][k] = Fd(k, 1) * dN[a][2] + Fd(k, 2) * dN[a][1];
				B[a][5][k] = Fd(k, 0) * dN[a][2] + Fd(k, 2) * dN[a][0];
			}
		double wdet = det;   // Gauss weights are 1
		for (int a = 0; a < 8; a++)
		{
			for (int k = 0; k < 3; k++)
			{
				double f = 0.0;
				for (int r = 0; r < 6; r++) f += B[a][r][k] * Sv[r];
				F[3 * a + k] += wdet * f;
				if (W.g_on) Fe[3 * a + k] += wdet * (s.rho * lfac) * N[a] * W.g[k];
			}
			for (int b = 0; b < 8; b++)
			{
				double CB[6][3];
				for (int r = 0; r < 6; r++) for (int k = 0; k < 3; k++) { double v = 0.0; for (int c = 0; c < 6; c++) v += Cm[r][c] * B[b][c][k]; CB[r][k] = v; }
				double geo = 0.0;
				for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) geo += dN[a][i] * S(i, j) * dN[b][j];
				for (int i = 0; i < 3; i++)
					for (int k = 0; k < 3; k++)
					{
						double v = 0.0;
						for (int r = 0; r < 6; r++) v += B[a][r][i] * CB[r][k];
						if (i == k) v += geo;
						K(3 * a + i, 3 * b + k) += wdet * v;
					}
			}
		}
		double en = 0.0;
		for (int r = 0; r < 6; r++) en += 0.5 * Sv[r] * Ev[r];
		s.energy += wdet * en;
	}
	for (int i = 0; i < 24; i++) { s.Fint[i] = F[i]; s.P[i] = F[i] - Fe[i]; for (int j = 0; j < 24; j++) s.K[i * 24 + j] = K(i, j); }
}

// ------------------------------------------------------------------------
// global system
// ------------------------------------------------------------------------
// local DOF -> (local node, nodal DOF): Shell_1.cpp:1523-1557, Beam_1.cpp:1439-1444
inline void local_dof(int type, int i, int& a, int& k)
{
	if (type == T_SHELL) { if (i < 18) { a = i / 3; k = i % 3; } else { a = 3 + (i - 18) / 3; k = 3 + (i - 18) % 3; } }
	else if (type == T_BEAM) { a = i / 6; k = i % 6; }
	else { a = i / 3; k = i % 3; }
}
inline int ndof_of(int type) { return type == T_SHELL ? 27 : type == T_BEAM ? 18 : 24; }

// Eigen setFromTriplets semantics (see oracle/ref_shims/eigen_sparsecore_stub.h),
// restated with a stable sort: columns ascending inside a row, duplicates
// summed in insertion order, explicit zeros kept.
void build_csr(int w)
{
	std::vector<World::Trip>& t = W.trip[w];
	const int nr = W.rows[w];
	std::vector<size_t> start((size_t)nr + 1, 0);
	for (size_t i = 0; i < t.size(); i++) start[(size_t)t[i].r + 1]++;
	for (int r = 0; r < nr; r++) start[r + 1] += start[r];
	std::vector<std::pair<int, double> > byrow(t.size());
	{
		std::vector<size_t> fill(start.begin(), start.end() - 1);
		for (size_t i = 0; i < t.size(); i++) byrow[fill[t[i].r]++] = std::make_pair(t[i].c, t[i].v);
	}
	W.outer[w].assign((size_t)nr + 1, 0);
	W.inner[w].clear(); W.val[w].clear();
	for (int r = 0; r < nr; r++)
	{
		std::stable_sort(byrow.begin() + start[r], byrow.begin() + start[r + 1],
			[](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
		for (size_t p = start[r]; p < start[r + 1]; p++)
		{
			if (p > start[r] && byrow[p].first == byrow[p - 1].first) W.val[w].back() += byrow[p].second;
			else { W.inner[w].push_back(byrow[p].first); W.val[w].push_back(byrow[p].second); }
		}
		W.outer[w][r + 1] = (int)W.inner[w].size();
	}
}

double now_s()
{
	using namespace std::chrono;
	return duration_cast<duration<double> >(high_resolution_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" {

int gfo_reset(void) { W = World(); return 0; }
int gfo_set_threads(int n) { omp_set_num_threads(n); return omp_get_max_threads(); }
int gfo_set_nodes(int n, const double* xyz)
{
	W.n_nodes = n;
	W.ref.assign(xyz, xyz + 3 * (size_t)n);
	W.copy.assign(6 * (size_t)n, 0.0);
	for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) W.copy[6 * (size_t)i + k] = xyz[3 * (size_t)i + k];
	W.cmask.assign(n, 0);
	return 0;
}
int gfo_set_materials(int n, const double* h) { W.hooke.assign(h, h + 3 * (size_t)n); return 0; }
int gfo_set_sections(int n, const double* s) { W.sec.assign(s, s + 6 * (size_t)n); return 0; }
int gfo_set_shell_sections(int n, const double* t) { W.thick.assign(t, t + n); return 0; }
int gfo_set_cs(int n, const double* e) { W.cs.assign(e, e + 9 * (size_t)n); return 0; }
int gfo_set_pipe_sections(int n, const double* p) { W.pipe.assign(p, p + 11 * (size_t)n); return 0; }
int gfo_set_elements(int n, const int* type, const int* mat, const int* sec, const int* cs,
	const int* node_ptr, const int* nodes, const double* pretension)
{
	W.n_el = n;
	W.is_pipe.assign(n, 0);
	W.type.assign(type, type + n); W.mat.assign(mat, mat + n); W.secid.assign(sec, sec + n); W.csid.assign(cs, cs + n);
	W.nptr.assign(node_ptr, node_ptr + n + 1);
	W.nodes.assign(nodes, nodes + node_ptr[n]);
	W.pret.assign(n, 0.0);
	if (pretension) W.pret.assign(pretension, pretension + n);
	for (int e = 0; e < n; e++) if (W.type[e] == T_PIPE) { W.is_pipe[e] = 1; W.type[e] = T_BEAM; W.pret[e] = 0.0; }
	return 0;
}
// PipeLoad objects in load-number order: element lists (0-based) and the pressure P0I each one has at the
// evaluation time (the caller interpolates the table: Load::GetValueAt(last_converged_time + current_time_step, 0))
int gfo_set_pipe_loads(int n_loads, const int* ptr, const int* elements, const double* p0i)
{
	W.pipe_load_ptr.assign(ptr, ptr + n_loads + 1);
	W.pipe_load_el.assign(elements, elements + ptr[n_loads]);
	W.pipe_load_p.assign(p0i, p0i + n_loads);
	for (int e : W.pipe_load_el) if (e < 0 || e >= W.n_el || !W.is_pipe[e]) return -1;    // PipeLoad::Check (PipeLoad.cpp:91-106)
	return 0;
}
int gfo_set_gravity(int on, double gx, double gy, double gz) { W.g_on = on; W.g[0] = gx; W.g[1] = gy; W.g[2] = gz; return 0; }
int gfo_set_constraint_mask(const int* m) { W.cmask.assign(m, m + W.n_nodes); return 0; }

// Database::PreCalc element loop (Database.cpp:713-714) followed by
// Solution::DOFsActive (Solution.cpp:121-224) and SetGlobalDOFs (:40-118).
int gfo_precalc(void)
{
	W.slot.assign(W.n_el, 0);
	W.shells.clear(); W.beams.clear(); W.solids.clear();
	for (int e = 0; e < W.n_el; e++)
	{
		if (W.type[e] == T_SHELL) { W.slot[e] = (int)W.shells.size(); W.shells.push_back(ShellEl()); }
		else if (W.type[e] == T_BEAM) { W.slot[e] = (int)W.beams.size(); W.beams.push_back(BeamEl()); }
		else if (W.type[e] == T_SOLID) { W.slot[e] = (int)W.solids.size(); W.solids.push_back(SolidEl()); }
		else return -1;
	}
#pragma omp parallel for
	for (int e = 0; e < W.n_el; e++)
	{
		const int* nd = &W.nodes[W.nptr[e]];
		const double* hk = W.is_pipe[e] ? nullptr : &W.hooke[3 * (size_t)(W.mat[e] - 1)];
		if (W.is_pipe[e]) pipe_precalc(W.beams[W.slot[e]], nd, &W.pipe[11 * (size_t)(W.secid[e] - 1)], &W.cs[9 * (size_t)(W.csid[e] - 1)]);
		else if (W.type[e] == T_SHELL) shell_precalc(W.shells[W.slot[e]], nd, hk[0], hk[1], hk[2], W.thick[W.secid[e] - 1]);
		else if (W.type[e] == T_BEAM) beam_precalc(W.beams[W.slot[e]], nd, hk, &W.sec[6 * (size_t)(W.secid[e] - 1)], &W.cs[9 * (size_t)(W.csid[e] - 1)], W.pret[e]);
		else
		{
			SolidEl& s = W.solids[W.slot[e]];
			s.mu = hk[0] / (2.0 * (1 + hk[1]));
			s.lambda = hk[0] * hk[1] / ((1 + hk[1]) * (1 - 2.0 * hk[1]));
			s.rho = hk[2];
		}
	}
	std::vector<int> active(6 * (size_t)W.n_nodes, 0);
	for (int e = 0; e < W.n_el; e++)
	{
		int nn = W.nptr[e + 1] - W.nptr[e];
		for (int a = 0; a < nn; a++)
		{
			size_t n0 = (size_t)(W.nodes[W.nptr[e] + a] - 1);
			for (int k = 0; k < 3; k++) active[6 * n0 + k] = 1;
			bool rot = W.type[e] == T_BEAM || (W.type[e] == T_SHELL && a > 2);   // Shell_1.cpp:51-68, Beam_1.cpp:28-57
			if (rot) for (int k = 3; k < 6; k++) active[6 * n0 + k] = 1;
		}
	}
	W.gls.assign(6 * (size_t)W.n_nodes, 0);
	int nf = 0, nx = 0;
	for (int i = 0; i < W.n_nodes; i++)
		for (int k = 0; k < 6; k++)
		{
			if (!active[6 * (size_t)i + k]) continue;
			if ((W.cmask[i] >> k) & 1) W.gls[6 * (size_t)i + k] = -(++nx);
			else W.gls[6 * (size_t)i + k] = ++nf;
		}
	W.n_free = nf; W.n_fixed = nx;
	W.rows[0] = nf; W.cols[0] = nf; W.rows[1] = nf; W.cols[1] = nx;
	W.rows[2] = nx; W.cols[2] = nf; W.rows[3] = nx; W.cols[3] = nx;
	W.disp.assign(6 * (size_t)W.n_nodes, 0.0);
	return 0;
}
int gfo_n_free(void) { return W.n_free; }
int gfo_n_fixed(void) { return W.n_fixed; }
int gfo_get_gls(int* g) { std::memcpy(g, W.gls.data(), sizeof(int) * W.gls.size()); return 0; }

int gfo_set_extra_triplets(int which, long n, const int* r, const int* c, const double* v)
{
	W.extra[which].resize((size_t)n);
	for (long i = 0; i < n; i++) { W.extra[which][i].r = r[i]; W.extra[which][i].c = c[i]; W.extra[which][i].v = v[i]; }
	return 0;
}

static int assemble(const double* disp6, double lfac, double* seconds, int dynamic, int update_rayleigh)
{
	W.disp.assign(disp6, disp6 + 6 * (size_t)W.n_nodes);
	if (dynamic)
		for (int e = 0; e < W.n_el; e++)
			if (W.type[e] == T_SOLID) return -7;   // no arithmetic in the reference (Pipe_1: restated without ocean data)
	double t0 = now_s();
	for (int w = 0; w < 4; w++) W.trip[w] = W.extra[w];                      // Clear + MountLoads
	W.PA.assign(W.n_free, 0.0); W.IA.assign(W.n_free, 0.0); W.PB.assign(W.n_fixed, 0.0);
	double t1 = now_s();
#pragma omp parallel for schedule(static)
	for (int e = 0; e < W.n_el; e++)                                         // MountLocal (Solution.cpp:227-248)
	{
		const int* nd = &W.nodes[W.nptr[e]];
		if (W.type[e] == T_SHELL) shell_mount(W.shells[W.slot[e]], nd);
		else if (W.type[e] == T_BEAM) beam_mount(W.beams[W.slot[e]], nd);
		else solid_mount(W.solids[W.slot[e]], nd, lfac);
	}
	double t2 = now_s();
#pragma omp parallel for schedule(static)
	for (int e = 0; e < W.n_el; e++)                                         // MountElementLoads (:251-265)
	{
		if (W.type[e] == T_SHELL) shell_loads(W.shells[W.slot[e]], lfac);
		else if (W.type[e] == T_BEAM) beam_loads(W.beams[W.slot[e]], lfac);
	}
	for (size_t l = 0; l + 1 < W.pipe_load_ptr.size(); l++)                  // MountLoads: PipeLoad::Mount, load by load
		for (int k = W.pipe_load_ptr[l]; k < W.pipe_load_ptr[l + 1]; k++)
			pipe_pressure(W.beams[W.slot[W.pipe_load_el[k]]], W.pipe_load_p[l]);
	if (dynamic)
	{
#pragma omp parallel for schedule(static)
		for (int e = 0; e < W.n_el; e++)                                     // MountMass, MountDamping, MountDyn (Solution.cpp:711-759)
		{
			const int* nd = &W.nodes[W.nptr[e]];
			if (W.type[e] == T_SHELL) shell_dynamics(W.shells[W.slot[e]], nd, update_rayleigh != 0);
			else beam_dynamics(W.beams[W.slot[e]], nd, update_rayleigh != 0);
		}
	}
	double t3 = now_s();
	for (int e = 0; e < W.n_el; e++)                                         // MountGlobal, serial (:322-349)
	{
		const int* nd = &W.nodes[W.nptr[e]];
		const int ty = W.type[e], n = ndof_of(ty);
		const double* K = ty == T_SHELL ? W.shells[W.slot[e]].K : ty == T_BEAM ? W.beams[W.slot[e]].K : W.solids[W.slot[e]].K;
		const double* P = ty == T_SHELL ? W.shells[W.slot[e]].P : ty == T_BEAM ? W.beams[W.slot[e]].P : W.solids[W.slot[e]].P;
		int gl[27];
		for (int i = 0; i < n; i++) { int a, k; local_dof(ty, i, a, k); gl[i] = W.gls[6 * (size_t)(nd[a] - 1) + k]; }
		for (int i = 0; i < n; i++)
		{
			const int g1 = gl[i];
			if (g1 > 0) { W.PA[g1 - 1] += P[i]; W.IA[g1 - 1] += P[i]; }
			else if (g1 < 0) W.PB[-g1 - 1] += P[i];
			for (int j = 0; j < n; j++)
			{
				const int g2 = gl[j];
				World::Trip t; t.v = K[i * n + j];
				if (g1 > 0 && g2 > 0) { t.r = g1 - 1; t.c = g2 - 1; W.trip[0].push_back(t); }
				if (g1 < 0 && g2 < 0) { t.r = -g1 - 1; t.c = -g2 - 1; W.trip[3].push_back(t); }
				if (g1 > 0 && g2 < 0) { t.r = g1 - 1; t.c = -g2 - 1; W.trip[1].push_back(t); }
				if (g1 < 0 && g2 > 0) { t.r = -g1 - 1; t.c = g2 - 1; W.trip[2].push_back(t); }
			}
		}
	}
	double t4 = now_s();
	for (int w = 0; w < 4; w++) build_csr(w);                                // MountSparse (:851-863)
	double t5 = now_s();
	if (seconds) { seconds[0] = t2 - t1; seconds[1] = t3 - t2; seconds[2] = t4 - t3; seconds[3] = t5 - t4; seconds[4] = t1 - t0; }
	return 0;
}

int gfo_assemble(const double* disp6, double lfac, double* seconds) { return assemble(disp6, lfac, seconds, 0, 0); }

// ---- Newmark dynamics (Dynamic.cpp:303-340) ----
int gfo_set_dynamic(const double* newmark6, double rayleigh_alpha, double rayleigh_beta)
{
	for (int i = 0; i < 6; i++) W.nm[i] = newmark6[i];
	W.ray_alpha = rayleigh_alpha; W.ray_beta = rayleigh_beta;
	return 0;
}
static void ensure_kinematics()
{
	const size_t n = 6 * (size_t)W.n_nodes;
	if (W.vel.size() != n) { W.vel.assign(n, 0.0); W.accel.assign(n, 0.0); W.copy_vel.assign(n, 0.0); W.copy_accel.assign(n, 0.0); }
}
int gfo_set_kinematics(const double* vel, const double* accel, const double* copy_vel, const double* copy_accel)
{
	ensure_kinematics();
	const size_t n = 6 * (size_t)W.n_nodes;
	if (vel) W.vel.assign(vel, vel + n);
	if (accel) W.accel.assign(accel, accel + n);
	if (copy_vel) W.copy_vel.assign(copy_vel, copy_vel + n);
	if (copy_accel) W.copy_accel.assign(copy_accel, copy_accel + n);
	return 0;
}
int gfo_get_kinematics(double* vel, double* accel, double* copy_vel, double* copy_accel)
{
	ensure_kinematics();
	const size_t n = sizeof(double) * 6 * (size_t)W.n_nodes;
	if (vel) std::memcpy(vel, W.vel.data(), n);
	if (accel) std::memcpy(accel, W.accel.data(), n);
	if (copy_vel) std::memcpy(copy_vel, W.copy_vel.data(), n);
	if (copy_accel) std::memcpy(copy_accel, W.copy_accel.data(), n);
	return 0;
}
int gfo_update_dyn(const double* disp6)
{
	ensure_kinematics();
	W.disp.assign(disp6, disp6 + 6 * (size_t)W.n_nodes);
	update_dyn();
	return 0;
}
int gfo_assemble_dynamic(const double* disp6, double lfac, int update_rayleigh)
{
	ensure_kinematics();
	return assemble(disp6, lfac, NULL, 1, update_rayleigh);
}
int gfo_get_alpha_i(int e, double* out)
{
	int w = 0;
	if (W.type[e] == T_SHELL) { const ShellEl& s = W.shells[W.slot[e]]; for (int g = 0; g < 3; g++) for (int i = 0; i < 3; i++) out[w++] = s.alpha_i[g][i]; }
	else if (W.type[e] == T_BEAM) { const BeamEl& b = W.beams[W.slot[e]]; for (int g = 0; g < 2; g++) for (int i = 0; i < 3; i++) out[w++] = b.alpha_i[g][i]; }
	return w;
}

long gfo_triplet_count(int w) { return (long)W.trip[w].size(); }
int gfo_csr_rows(int w) { return W.rows[w]; }
int gfo_csr_cols(int w) { return W.cols[w]; }
long gfo_csr_nnz(int w) { return (long)W.val[w].size(); }
int gfo_csr_get(int w, int* outer, int* inner, double* val)
{
	if (outer) std::memcpy(outer, W.outer[w].data(), sizeof(int) * W.outer[w].size());
	if (inner) std::memcpy(inner, W.inner[w].data(), sizeof(int) * W.inner[w].size());
	if (val) std::memcpy(val, W.val[w].data(), sizeof(double) * W.val[w].size());
	return 0;
}
int gfo_get_vectors(double* PA, double* IA, double* PB)
{
	if (PA) std::memcpy(PA, W.PA.data(), sizeof(double) * W.PA.size());
	if (IA) std::memcpy(IA, W.IA.data(), sizeof(double) * W.IA.size());
	if (PB) std::memcpy(PB, W.PB.data(), sizeof(double) * W.PB.size());
	return 0;
}
int gfo_get_element(int e, double* K, double* P, double* energy)
{
	const int ty = W.type[e], n = ndof_of(ty);
	const double* k = ty == T_SHELL ? W.shells[W.slot[e]].K : ty == T_BEAM ? W.beams[W.slot[e]].K : W.solids[W.slot[e]].K;
	const double* p = ty == T_SHELL ? W.shells[W.slot[e]].P : ty == T_BEAM ? W.beams[W.slot[e]].P : W.solids[W.slot[e]].P;
	double en = ty == T_SHELL ? W.shells[W.slot[e]].energy : ty == T_BEAM ? W.beams[W.slot[e]].energy : W.solids[W.slot[e]].energy;
	if (K) std::memcpy(K, k, sizeof(double) * n * n);
	if (P) std::memcpy(P, p, sizeof(double) * n);
	if (energy) *energy = en;
	return n;
}
// same layout as ref_get_state (oracle/ref_shims/ref_driver.cpp)
int gfo_get_state(int e, double* out)
{
	int w = 0;
	if (W.type[e] == T_SHELL)
	{
		const ShellEl& s = W.shells[W.slot[e]];
		for (int g = 0; g < 3; g++)
		{
			for (int i = 0; i < 9; i++) out[w++] = s.Q_i[g].a[i];
			for (int i = 0; i < 3; i++) out[w++] = s.zx1_i[g][i];
			for (int i = 0; i < 3; i++) out[w++] = s.zx2_i[g][i];
			for (int i = 0; i < 3; i++) out[w++] = s.k1_i[g][i];
			for (int i = 0; i < 3; i++) out[w++] = s.k2_i[g][i];
		}
	}
	else if (W.type[e] == T_BEAM)
	{
		const BeamEl& b = W.beams[W.slot[e]];
		for (int g = 0; g < 2; g++)
		{
			for (int i = 0; i < 9; i++) out[w++] = b.Q_i[g].a[i];
			for (int i = 0; i < 3; i++) out[w++] = b.dz_i[g][i];
			for (int i = 0; i < 3; i++) out[w++] = b.k_i[g][i];
		}
	}
	return w;
}

int gfo_get_results(int e, double* out)
{
	if (e < 0 || e >= (int)W.type.size()) return -1;
	int w = 0;
	const int ty = W.type[e];
	if (ty == T_SHELL)
	{
		const ShellEl& s = W.shells[W.slot[e]];
		out[w++] = s.energy;
		for (int g = 0; g < 3; g++) for (int i = 0; i < 24; i++) out[w++] = s.res[g][i];
	}
	else if (ty == T_BEAM)
	{
		const BeamEl& b = W.beams[W.slot[e]];
		out[w++] = b.energy;
		for (int g = 0; g < 2; g++) for (int i = 0; i < 12; i++) out[w++] = b.res[g][i];
	}
	return w;
}

// Node::SaveConfiguration (Node.cpp:325-349, spatial description) +
// Element::SaveLagrange, then the next increment's Zeros().
int gfo_commit(void)
{
	for (int i = 0; i < W.n_nodes; i++)
	{
		double* c = &W.copy[6 * (size_t)i];
		const double* d = &W.disp[6 * (size_t)i];
		for (int k = 0; k < 3; k++) c[k] += d[k];
		V3 a1 = vec(c[3], c[4], c[5]), a2 = vec(d[3], d[4], d[5]);
		V3 a3 = (4.0 / (4.0 - dot(a2, a1))) * (a2 + a1 + 0.5 * cross(a2, a1));
		c[3] = a3[0]; c[4] = a3[1]; c[5] = a3[2];
	}
#pragma omp parallel for
	for (int e = 0; e < W.n_el; e++)
	{
		if (W.type[e] == T_SHELL) shell_commit(W.shells[W.slot[e]]);
		else if (W.type[e] == T_BEAM) beam_commit(W.beams[W.slot[e]]);
	}
	std::fill(W.disp.begin(), W.disp.end(), 0.0);
	if (!W.vel.empty()) { W.copy_vel = W.vel; W.copy_accel = W.accel; }       // Node.cpp:375-380
	return 0;
}
int gfo_get_copy_coordinates(double* c) { std::memcpy(c, W.copy.data(), sizeof(double) * W.copy.size()); return 0; }

} // extern "C"
