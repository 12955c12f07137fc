// Hand-written sm_100a kernels for GIRAFFE's per-Newton-iteration assembly.
//
// Element evaluation (Element::Mount + MountElementLoads):
//   one warp owns a batch of elements.  Phase A: one lane per Gauss point
//   evaluates the kinematics, the constitutive/geometric tangent in the
//   15x15 (shell) / 9x9 (beam, solid) "gradient space" C' and the generalised
//   force f, already rotated to global axes, and stages them in shared
//   memory.  Phase B: the lanes of the warp sweep (element, K-column) work
//   items and apply the shape-function congruence
//        K = sum_gp  (S (x) I3)^T C' (S (x) I3),   F = sum_gp (S (x) I3)^T f
//   exploiting that every block of the reference's deltaN matrices is a
//   scalar times I3 (Shell_1.cpp:2118-2180, Beam_1.cpp:641-669).  The element
//   block goes to the arena as contiguous 3x3 blocks in the reference's local
//   DOF order (Shell_1: the 48 stored blocks of gfa_device.h only).
//
// Scatter (Element::MountGlobal + SparseMatrix::Mount): one thread per row of a
//   CSR patch (= one (group-node, neighbour) pair) sums the contributing element
//   blocks in ascending element order -- the order the reference pushes its
//   triplets (Solution.cpp:327-328) -- and writes each CSR row once, in whole
//   sectors.  No atomics; results are bitwise reproducible.
//
// Also here: the on-demand Gauss-point result kernels, the state commit kernels and
// the Newton-loop vector steps (sign flip, K_AB X_B, max-norms, UpdateDisps).
#include <cuda_runtime.h>
#include <cstdlib>
#include "gfa_device.h"
#include "gfa_math.cuh"

namespace gfa {

// =========================================================================
// Shell_1
// =========================================================================
namespace shell {

constexpr int NGP = 3;                  // EPW (elements per warp batch) is a template parameter of eval_kernel
// Shared-memory record of one Gauss point (REC doubles):
//   [0, 24)    the four symmetric diagonal blocks C'(R,R), R < 4, 6 values each (00 01 02 11 12 22)
//   [24, 78)   the off-diagonal blocks (0,1) (0,2) (0,3) (1,2) (1,3) (2,3), row-major 3x3
//   [78, 123)  column 4: (0,4) (1,4) (2,4) (3,4) and the full, non-symmetric (4,4)
//   [123, 138) f = Psi'^T sigma (15)
//   [138, 159) S: N,1[6] N,2[6] Na,1[3] Na,2[3] Na[3]
// While phase A runs, [0, 72) parks eight intermediate 3x3 blocks (Y0..Y3, G0..G3).
constexpr int F_OFF = 123;
constexpr int S_OFF = 138;
constexpr int REC = 159;                // odd: the per-lane stores of phase A fall into 16 distinct bank pairs
__host__ __device__ constexpr int smem_bytes(int epw) { return epw * (NGP * REC + 27) * 8; }      // records, then the batch's P staging

// offset of block (p, q), p <= q, of the 5x5 block matrix C' inside the record
__host__ __device__ constexpr int boff(int p, int q) {
    return q == 4 ? 78 + 9 * p : p == q ? 6 * p : 24 + 9 * (p == 0 ? q - 1 : p == 1 ? q + 1 : 5);
}
__host__ __device__ constexpr int park(int s) { return 9 * s; }      // parking slot s = 0..7

struct DBlk { double d00, d01, d10, d11, d22; };

// D(r,s) of the 4x4 block constitutive matrix in the order [eta1,kappa1,eta2,kappa2]
// from the thickness moments X[m][pair][00,01,10,11] (Shell_1.cpp:1125-1143,
// 1171-1215); pairs: 0 = C11, 1 = C12, 2 = C22, C21 = C12^T.
template <int R, int S>
GFA_DI DBlk getD(const double (&X)[3][3][4], double smu, double drill) {
    constexpr int a = R / 2, b = S / 2, kr = R % 2, ks = S % 2;
    constexpr int pair = (a == 0 && b == 0) ? 0 : (a == 1 && b == 1) ? 2 : 1;
    constexpr bool tr = (a == 1 && b == 0);
    constexpr int m = kr + ks;
    const double e00 = X[m][pair][0], e01 = X[m][pair][tr ? 2 : 1], e10 = X[m][pair][tr ? 1 : 2], e11 = X[m][pair][3];
    DBlk d;
    if (kr == 0 && ks == 0) { d.d00 = e00; d.d01 = e01; d.d10 = e10; d.d11 = e11; d.d22 = (a == b) ? smu : 0.0; }
    else if (kr == 0 && ks == 1) { d.d00 = -e01; d.d01 = e00; d.d10 = -e11; d.d11 = e10; d.d22 = 0.0; }
    else if (kr == 1 && ks == 0) { d.d00 = -e10; d.d01 = -e11; d.d10 = e00; d.d11 = e01; d.d22 = 0.0; }
    else { d.d00 = e11; d.d01 = -e10; d.d10 = -e01; d.d11 = e00; d.d22 = (a == b) ? drill : 0.0; }
    return d;
}
GFA_DI void dmul(double* o, const DBlk& D, const double* P) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        o[j] = D.d00 * P[j] + D.d01 * P[3 + j];
        o[3 + j] = D.d10 * P[j] + D.d11 * P[3 + j];
        o[6 + j] = D.d22 * P[6 + j];
    }
}
GFA_DI void dmul_acc(double* o, const DBlk& D, const double* P) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        o[j] += D.d00 * P[j] + D.d01 * P[3 + j];
        o[3 + j] += D.d10 * P[j] + D.d11 * P[3 + j];
        o[6 + j] += D.d22 * P[6 + j];
    }
}
// rec[block (P,Q)] = w * M; a diagonal block of the first four groups is symmetric and keeps its
// upper triangle only
template <int P, int Q>
GFA_DI void put_block(double* rec, double w, const double* M) {
    if (P == Q && P < 4) {
        rec[boff(P, Q) + 0] = w * M[0]; rec[boff(P, Q) + 1] = w * M[1]; rec[boff(P, Q) + 2] = w * M[2];
        rec[boff(P, Q) + 3] = w * M[4]; rec[boff(P, Q) + 4] = w * M[5]; rec[boff(P, Q) + 5] = w * M[8];
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) rec[boff(P, Q) + i] = w * M[i];
    }
}

// Element frame and area-coordinate gradients (Shell_1::PreCalc, :1985-2060)
struct Frame {
    double R[9];            // rows e1r, e2r, e3r = transform3 (:1470-1495)
    double area;
    double Lx[3], Ly[3];    // dL_a/dx1, dL_a/dx2
};
GFA_DI void frame_of(const double (&x)[6][3], Frame& fr) {
    // strict arithmetic in the reference's order: the shape-function values feed
    // the strain chain (see gfa_math.cuh, "Strict" arithmetic)
    double d21[3], d31[3], n[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { d21[k] = s_sub(x[1][k], x[0][k]); d31[k] = s_sub(x[2][k], x[0][k]); }
    s_cross3(n, d21, d31);
    const double nn = s_norm3(n);
    const double A = s_mul(0.5, nn);
    double e3[3], e1[3], e2[3];
    const double inv = 1.0 / nn;
#pragma unroll
    for (int k = 0; k < 3; k++) e3[k] = s_mul(n[k], inv);
    double eg[3] = { 1.0, 0.0, 0.0 };
    if (fabs(e3[0]) >= 1.0 - 1e-4) { eg[0] = 0.0; eg[1] = 1.0; }     // Shell_1.cpp:1990
    const double ege3 = s_dot3(eg, e3);
#pragma unroll
    for (int k = 0; k < 3; k++) e1[k] = s_sub(eg[k], s_mul(e3[k], ege3));
    const double inv1 = 1.0 / s_norm3(e1);
#pragma unroll
    for (int k = 0; k < 3; k++) e1[k] = s_mul(e1[k], inv1);
    s_cross3(e2, e3, e1);
#pragma unroll
    for (int k = 0; k < 3; k++) { fr.R[k] = e1[k]; fr.R[3 + k] = e2[k]; fr.R[6 + k] = e3[k]; }
    fr.area = A;
    double d23[3], d12[3], d32[3], d13[3], d31b[3], d21b[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        d23[k] = s_sub(x[1][k], x[2][k]); d31b[k] = s_sub(x[2][k], x[0][k]); d12[k] = s_sub(x[0][k], x[1][k]);
        d32[k] = s_sub(x[2][k], x[1][k]); d13[k] = s_sub(x[0][k], x[2][k]); d21b[k] = s_sub(x[1][k], x[0][k]);
    }
    fr.Lx[0] = s_mul(0.5, s_dot3(d23, e2)) / A; fr.Lx[1] = s_mul(0.5, s_dot3(d31b, e2)) / A; fr.Lx[2] = s_mul(0.5, s_dot3(d12, e2)) / A;
    fr.Ly[0] = s_mul(0.5, s_dot3(d32, e1)) / A; fr.Ly[1] = s_mul(0.5, s_dot3(d13, e1)) / A; fr.Ly[2] = s_mul(0.5, s_dot3(d21b, e1)) / A;
}
// Shape functions at in-plane point g (located at mid-side node 4+g), :2029-2088
struct Shape { double N1[6], N2[6], A0[3], A1[3], A2[3]; };
GFA_DI void shape_of(const double (&x)[6][3], const Frame& fr, int g, Shape& s) {
    const double* xp = x[3 + g];
    double a[3], b[3], c[3], t[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { a[k] = s_sub(x[0][k], xp[k]); b[k] = s_sub(x[1][k], xp[k]); c[k] = s_sub(x[2][k], xp[k]); }
    s_cross3(t, b, c); const double L1 = s_mul(0.5, s_norm3(t)) / fr.area;
    s_cross3(t, c, a); const double L2 = s_mul(0.5, s_norm3(t)) / fr.area;
    s_cross3(t, a, b); const double L3 = s_mul(0.5, s_norm3(t)) / fr.area;
    const double L[3] = { L1, L2, L3 };
    const double* Lx = fr.Lx; const double* Ly = fr.Ly;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        s.N1[k] = s_sub(s_mul(s_mul(4.0, Lx[k]), L[k]), Lx[k]);
        s.N2[k] = s_sub(s_mul(s_mul(4.0, Ly[k]), L[k]), Ly[k]);
    }
#define GFA_MIX(D_, i_, j_) s_add(s_mul(s_mul(4.0, D_[i_]), L[j_]), s_mul(s_mul(4.0, L[i_]), D_[j_]))
    s.N1[3] = GFA_MIX(Lx, 0, 1); s.N1[4] = GFA_MIX(Lx, 1, 2); s.N1[5] = GFA_MIX(Lx, 2, 0);
    s.N2[3] = GFA_MIX(Ly, 0, 1); s.N2[4] = GFA_MIX(Ly, 1, 2); s.N2[5] = GFA_MIX(Ly, 2, 0);
#undef GFA_MIX
    s.A0[0] = s_sub(1.0, s_mul(2.0, L3)); s.A0[1] = s_sub(1.0, s_mul(2.0, L1)); s.A0[2] = s_sub(1.0, s_mul(2.0, L2));
    s.A1[0] = -2 * Lx[2]; s.A1[1] = -2 * Lx[0]; s.A1[2] = -2 * Lx[1];
    s.A2[0] = -2 * Ly[2]; s.A2[1] = -2 * Ly[0]; s.A2[2] = -2 * Ly[1];
}

// Increment kinematics at one point, in the element frame (:906-1020)
struct Kin {
    double a[3], a1[3], a2[3], u1[3], u2[3];   // alpha_delta, its x1/x2 derivatives, u_delta,1 u_delta,2 (element frame)
    double ga[3], ga1[3], ga2[3];               // the same rotation quantities in global axes
};
GFA_DI void interpolate(const EvalArgs& A, const int* nd, const Frame& fr, const Shape& s, Kin& k) {
    double gu1[3], gu2[3], ga[3], ga1[3], ga2[3];
#pragma unroll
    for (int n = 0; n < 6; n++) {
        // a node's six increments are one 48-byte record, 16-byte aligned: three (corner nodes: two) vector loads
        const double2* d = reinterpret_cast<const double2*>(A.disp + 6 * (size_t)nd[n]);
        const double2 u01 = __ldg(d), u2r0 = __ldg(d + 1);
        const double u[3] = { u01.x, u01.y, u2r0.x };
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double v = u[c];
            gu1[c] = n == 0 ? s_mul(v, s.N1[0]) : s_add(gu1[c], s_mul(v, s.N1[n]));
            gu2[c] = n == 0 ? s_mul(v, s.N2[0]) : s_add(gu2[c], s_mul(v, s.N2[n]));
        }
        if (n >= 3) {
            const double2 r12 = __ldg(d + 2);
            const double rr[3] = { u2r0.y, r12.x, r12.y };
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double r = rr[c];
                ga[c] = n == 3 ? s_mul(r, s.A0[0]) : s_add(ga[c], s_mul(r, s.A0[n - 3]));
                ga1[c] = n == 3 ? s_mul(r, s.A1[0]) : s_add(ga1[c], s_mul(r, s.A1[n - 3]));
                ga2[c] = n == 3 ? s_mul(r, s.A2[0]) : s_add(ga2[c], s_mul(r, s.A2[n - 3]));
            }
        }
    }
    s_mv(k.u1, fr.R, gu1); s_mv(k.u2, fr.R, gu2);
    s_mv(k.a, fr.R, ga); s_mv(k.a1, fr.R, ga1); s_mv(k.a2, fr.R, ga2);
#pragma unroll
    for (int c = 0; c < 3; c++) { k.ga[c] = ga[c]; k.ga1[c] = ga1[c]; k.ga2[c] = ga2[c]; }
}

GFA_DI void load_nodes(const EvalArgs& A, int e, int* nd, double (&x)[6][3]) {
#pragma unroll
    for (int n = 0; n < 6; n++) {
        nd[n] = __ldg(A.conn + 6 * (size_t)e + n);
        const double* p = A.xyz + 3 * (size_t)nd[n];
#pragma unroll
        for (int c = 0; c < 3; c++) x[n][c] = __ldg(p + c);
    }
}

// Per-element / per-point constants produced once by precalc_kernel:
//   geo[k * n_el + e], k = 0..8 : R (rows e1r,e2r,e3r), k = 9 : area
//   shp[k * n_gp + gp], k = 0..20: N,1[6] N,2[6] Na,1[3] Na,2[3] Na[3]
GFA_DI void load_precalc(const EvalArgs& A, int e, int g, Frame& fr, Shape& sh) {
    const size_t ne = (size_t)A.n_el, n_gp = ne * NGP, gp = (size_t)e * NGP + g;
#pragma unroll
    for (int k = 0; k < 9; k++) fr.R[k] = __ldg(A.geo + k * ne + e);
    fr.area = __ldg(A.geo + 9 * ne + e);
#pragma unroll
    for (int k = 0; k < 6; k++) { sh.N1[k] = __ldg(A.shp + k * n_gp + gp); sh.N2[k] = __ldg(A.shp + (6 + k) * n_gp + gp); }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        sh.A1[k] = __ldg(A.shp + (12 + k) * n_gp + gp);
        sh.A2[k] = __ldg(A.shp + (15 + k) * n_gp + gp);
        sh.A0[k] = __ldg(A.shp + (18 + k) * n_gp + gp);
    }
}
__global__ void precalc_kernel(EvalArgs A, double* geo, double* shp) {
    const size_t ne = (size_t)A.n_el, n_gp = ne * NGP;
    const size_t gp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= n_gp) return;
    const int e = (int)(gp / NGP), g = (int)(gp % NGP);
    int nd[6];
    double x[6][3];
    load_nodes(A, e, nd, x);
    Frame fr; frame_of(x, fr);
    Shape sh; shape_of(x, fr, g, sh);
    if (g == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) geo[k * ne + e] = fr.R[k];
        geo[9 * ne + e] = fr.area;
    }
#pragma unroll
    for (int k = 0; k < 6; k++) { shp[k * n_gp + gp] = sh.N1[k]; shp[(6 + k) * n_gp + gp] = sh.N2[k]; }
#pragma unroll
    for (int k = 0; k < 3; k++) { shp[(12 + k) * n_gp + gp] = sh.A1[k]; shp[(15 + k) * n_gp + gp] = sh.A2[k]; shp[(18 + k) * n_gp + gp] = sh.A0[k]; }
}

// Thickness integration at one in-plane point (Shell_1.cpp:1056-1163).
// RESULTANTS: strict arithmetic, yields n_beta, m_beta (they set the internal
// force).  Otherwise: the moments X[m][pair][..] of the tangent blocks.
struct Strains { double eta1[3], eta2[3], kap1[3], kap2[3]; };
template <bool RESULTANTS>
GFA_DI void thickness(const Strains& st, double lam, double mu, double thick,
                      double (&X)[3][3][4], double& smu, double* n1, double* n2, double* m1, double* m2) {
    const double jac = thick / 2.0;
    const double mu2 = s_mul(2.0, mu);
    if (RESULTANTS) {
#pragma unroll
        for (int i = 0; i < 3; i++) { n1[i] = 0.0; n2[i] = 0.0; m1[i] = 0.0; m2[i] = 0.0; }
    } else {
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
            for (int p = 0; p < 3; p++)
#pragma unroll
                for (int c = 0; c < 4; c++) X[m][p][c] = 0.0;
        smu = 0.0;
    }
    // rolled on purpose: keeps the live register set of the three points apart
#pragma unroll 1
    for (int q = 0; q < 3; q++) {
        const double csi = (q == 0) ? -0.77459666924148337703585307995648 : (q == 1) ? 0.0 : 0.77459666924148337703585307995648;
        const double al2 = (q == 1) ? 0.88888888888888888888888888888889 : 0.55555555555555555555555555555556;
        const double zeta = s_mul(thick, csi) / 2.0;
        // gamma = eta + zeta * kappa x e3
        const double g11 = s_add(st.eta1[0], s_mul(zeta, st.kap1[1])), g12 = s_add(st.eta1[1], s_mul(zeta, -st.kap1[0])), g13 = st.eta1[2];
        const double g21 = s_add(st.eta2[0], s_mul(zeta, st.kap2[1])), g22 = s_add(st.eta2[1], s_mul(zeta, -st.kap2[0])), g23 = st.eta2[2];
        const double p11 = s_add(1.0, g11), p22 = s_add(1.0, g22);
        const double jb = s_sub(s_mul(p11, p22), s_mul(g12, g21));
        const double jb3 = s_mul(s_mul(jb, jb), jb);
        const double ljj = s_mul(s_mul(lam, jb), jb);                      // lambda*jb*jb
        const double v = s_add(s_mul(lam, s_sub(jb3, 1.0)), s_mul(mu2, s_sub(jb, 1.0))) / s_add(s_mul(ljj, jb), s_mul(mu2, jb));
        const double wj = s_mul(al2, jac), wz = s_mul(wj, zeta);
        if (RESULTANTS) {
            const double muv = s_mul(mu, v);
            const double t1[3] = { s_add(s_mul(muv, p22), s_mul(mu, s_sub(g11, g22))), s_add(s_mul(muv, -g21), s_mul(mu, s_add(g12, g21))), s_mul(mu, g13) };
            const double t2[3] = { s_add(s_mul(muv, -g12), s_mul(mu, s_add(g12, g21))), s_add(s_mul(muv, p11), s_mul(mu, s_sub(g22, g11))), s_mul(mu, g23) };
#pragma unroll
            for (int i = 0; i < 3; i++) {
                n1[i] = s_add(n1[i], s_mul(wj, t1[i]));
                n2[i] = s_add(n2[i], s_mul(wj, t2[i]));
            }
            // m += wz * e3 x tau
            m1[0] = s_add(m1[0], s_mul(wz, -t1[1])); m1[1] = s_add(m1[1], s_mul(wz, t1[0]));
            m2[0] = s_add(m2[0], s_mul(wz, -t2[1])); m2[1] = s_add(m2[1], s_mul(wz, t2[0]));
        } else {
            const double cden = s_add(ljj, mu2);
            const double dv = s_mul(s_add(lam, mu2), s_add(s_mul(s_mul(s_mul(3.0, lam), jb), jb), mu2)) / s_mul(s_mul(s_mul(jb, jb), cden), cden);
            double C[3][4];
            C[0][0] = mu * fma(p22 * p22, dv, 1.0);
            C[0][1] = -mu * p22 * g21 * dv;
            C[0][2] = C[0][1];
            C[0][3] = mu * fma(g21 * g21, dv, 1.0);
            C[2][0] = mu * fma(g12 * g12, dv, 1.0);
            C[2][1] = -mu * (p11 * g12 * dv);
            C[2][2] = C[2][1];
            C[2][3] = mu * fma(p11 * p11, dv, 1.0);
            C[1][0] = -mu * (p22 * g12 * dv);
            C[1][1] = mu * (v - 1.0 + p11 * p22 * dv);
            C[1][2] = mu * (1.0 - v + g12 * g21 * dv);
            C[1][3] = -mu * (p11 * g21 * dv);
            const double wzz = wz * zeta;
#pragma unroll
            for (int p = 0; p < 3; p++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    X[0][p][c] = fma(wj, C[p][c], X[0][p][c]);
                    X[1][p][c] = fma(wz, C[p][c], X[1][p][c]);
                    X[2][p][c] = fma(wzz, C[p][c], X[2][p][c]);
                }
            smu = fma(wj, mu, smu);
        }
    }
}

// Geometric blocks of one direction beta (Shell_1.cpp:1277-1302) and the
// column-4 blocks of Psi' (:1255-1270), written directly in GLOBAL axes:
// with R proper orthogonal, R^T skew(v) R = skew(R^T v) and V, d_V, Xi are
// frame-covariant, so G' = R^T G R and Psi' = Psi R are obtained by feeding the
// global-frame vectors (ag = interpolated nodal rotations, zg = R^T z,b,
// PA^T n = R^T Q n) instead of rotating every 3x3 block.  Results are PARKED in
// the record (slots of the upper-triangle area that are rewritten last).
GFA_DI void direction_blocks(double* rec, int slotYn, int slotYm, int slotGn, int slotGm, double* G44, double* f4,
                             const double* ag, const double* abg, const double* zg, const double* nb, const double* mb,
                             const double* PA, const double* Xig, double gg, bool first) {
    double Xib[9], tmp[9], Y[9], t3[3];
    d_xi(Xib, ag, abg, gg, Xig);
    skew_mul(tmp, zg, Xig); mm(Y, PA, tmp);                    // Qt Z,b Xi R
#pragma unroll
    for (int i = 0; i < 9; i++) rec[9 * slotYn + i] = Y[i];
    mtv(t3, Y, nb);
    if (first) { f4[0] = t3[0]; f4[1] = t3[1]; f4[2] = t3[2]; } else { f4[0] += t3[0]; f4[1] += t3[1]; f4[2] += t3[2]; }
    mm(Y, PA, Xib);                                            // Qt Xi,b R
#pragma unroll
    for (int i = 0; i < 9; i++) rec[9 * slotYm + i] = Y[i];
    mtv(t3, Y, mb);
    f4[0] += t3[0]; f4[1] += t3[1]; f4[2] += t3[2];

    double sn[3], sm[3], SnXi[9], V[9], zn[3];
    mtv(sn, PA, nb); mtv(sm, PA, mb);                          // R^T Q n, R^T Q m
    skew_mul(SnXi, sn, Xig);                                   // skew(n) Xi
#pragma unroll
    for (int i = 0; i < 9; i++) rec[9 * slotGn + i] = -SnXi[i];     // G(u,b ; alpha) = -skew(n) Xi
    v_op(V, ag, sm, gg);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rec[9 * slotGm + 3 * i + j] = V[3 * j + i];   // G(alpha,b ; alpha) = V(alpha, m)^T
    // G(alpha;alpha) += Xi^T (Z,b skew(n)) Xi - V(alpha, Z,b n) + dV(alpha, alpha,b, m) - Xi,b^T (skew(m) Xi)
    skew_mul(tmp, zg, SnXi);
    if (first) mtm(G44, Xig, tmp); else mtm_acc(G44, Xig, tmp);
    cross3(zn, zg, sn);
    v_op(V, ag, zn, gg);
#pragma unroll
    for (int i = 0; i < 9; i++) G44[i] -= V[i];
    dv_op(V, ag, abg, sm, gg);
#pragma unroll
    for (int i = 0; i < 9; i++) G44[i] += V[i];
    skew_mul(SnXi, sm, Xig);                                   // skew(m) Xi
    mtm(tmp, Xib, SnXi);
#pragma unroll
    for (int i = 0; i < 9; i++) G44[i] -= tmp[i];
}

// Phase A for one Gauss point: fills its shared-memory record.
__device__ void physics(const EvalArgs& A, int e, int g, double* rec) {
    // parking slots inside the upper-triangle area (rewritten by the last step)
    constexpr int P_Y0 = 0, P_Y1 = 1, P_Y2 = 2, P_Y3 = 3, P_G0 = 4, P_G1 = 5, P_G2 = 6, P_G3 = 7;
    int nd[6];
#pragma unroll
    for (int n = 0; n < 6; n++) nd[n] = __ldg(A.conn + 6 * (size_t)e + n);
    Frame fr;
    Kin kn;
    {
        // element frame and shape functions come from the PreCalc kernel
        // (Shell_1::PreCalc runs once per model, Database.cpp:713-714)
        Shape sh;
        load_precalc(A, e, g, fr, sh);
        interpolate(A, nd, fr, sh, kn);
#pragma unroll
        for (int i = 0; i < 6; i++) { rec[S_OFF + i] = sh.N1[i]; rec[S_OFF + 6 + i] = sh.N2[i]; }
#pragma unroll
        for (int i = 0; i < 3; i++) { rec[S_OFF + 12 + i] = sh.A1[i]; rec[S_OFF + 15 + i] = sh.A2[i]; rec[S_OFF + 18 + i] = sh.A0[i]; }
    }
    const double* pr = A.props + SHELL_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    const double lam = __ldg(pr), mu = __ldg(pr + 1), thick = __ldg(pr + 2), drill = __ldg(pr + 3);
    const size_t n_gp = (size_t)A.n_el * NGP, gp = (size_t)e * NGP + g;
    const double w = fr.area / 3.0;                           // alpha1 (:2364)

    double gg, Xi[9], Qt[9], z1[3], z2[3];
    Strains st;
    {
        double Qd[9], Qi[9], Q[9], t3[3];
#pragma unroll
        for (int i = 0; i < 9; i++) Qi[i] = __ldg(A.state + i * n_gp + gp);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            z1[i] = s_add(kn.u1[i], __ldg(A.state + (9 + i) * n_gp + gp));      // z,1 = u_delta,1 + z,1^i  (:1010)
            z2[i] = s_add(kn.u2[i], __ldg(A.state + (12 + i) * n_gp + gp));
        }
        s_rodrigues(kn.a, gg, Qd, Xi);
        s_mm(Q, Qd, Qi);
        m_transpose(Qt, Q);
        // back-rotated strains (:1017-1020), strict
        s_mv(st.eta1, Qt, z1); st.eta1[0] = s_sub(st.eta1[0], 1.0);
        s_mv(st.eta2, Qt, z2); st.eta2[1] = s_sub(st.eta2[1], 1.0);
        s_mtv(t3, Xi, kn.a1); s_mtv(st.kap1, Qi, t3);
        s_mtv(t3, Xi, kn.a2); s_mtv(st.kap2, Qi, t3);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            st.kap1[i] = s_add(st.kap1[i], __ldg(A.state + (15 + i) * n_gp + gp));
            st.kap2[i] = s_add(st.kap2[i], __ldg(A.state + (18 + i) * n_gp + gp));
        }
    }
    double X[3][3][4], smu = 0.0;
    double n1[3], n2[3], m1[3], m2[3];
    thickness<true>(st, lam, mu, thick, X, smu, n1, n2, m1, m2);
    m1[2] = s_mul(drill, st.kap1[2]);                             // drilling penalty (:1214-1217)
    m2[2] = s_mul(drill, st.kap2[2]);

    // Psi' = Psi (I5 (x) R) (:1239-1270): PA = Qt R -> blocks (0,0),(2,2); PB = Qt Xi R = PA Xi_g -> (1,1),(3,3)
    double PA[9], PB[9];
    {
        double Xig[9], G44[9], f4[3], t[3], zg1[3], zg2[3];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
            Xig[i] = gg * id;
        }
        // Xi in global axes: g (I + skew(ag)/2)
        Xig[1] = -0.5 * gg * kn.ga[2]; Xig[2] = 0.5 * gg * kn.ga[1];
        Xig[3] = 0.5 * gg * kn.ga[2];  Xig[5] = -0.5 * gg * kn.ga[0];
        Xig[6] = -0.5 * gg * kn.ga[1]; Xig[7] = 0.5 * gg * kn.ga[0];
        mm(PA, Qt, fr.R);
        mm(PB, PA, Xig);
        mtv(zg1, fr.R, z1); mtv(zg2, fr.R, z2);
        mtv(t, PA, n1);
#pragma unroll
        for (int i = 0; i < 3; i++) rec[F_OFF + i] = w * t[i];
        mtv(t, PA, n2);
#pragma unroll
        for (int i = 0; i < 3; i++) rec[F_OFF + 6 + i] = w * t[i];
        mtv(t, PB, m1);
#pragma unroll
        for (int i = 0; i < 3; i++) rec[F_OFF + 3 + i] = w * t[i];
        mtv(t, PB, m2);
#pragma unroll
        for (int i = 0; i < 3; i++) rec[F_OFF + 9 + i] = w * t[i];
        // the two directions run through ONE copy of the code (rolled loop, inputs selected per
        // pass): straight-line code is what this kernel is short of (instruction cache)
#pragma unroll
        for (int i = 0; i < 9; i++) G44[i] = 0.0;
        f4[0] = 0.0; f4[1] = 0.0; f4[2] = 0.0;
#pragma unroll 1
        for (int beta = 0; beta < 2; beta++) {
            double abg[3], zg[3], nb[3], mb[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                abg[i] = beta ? kn.ga2[i] : kn.ga1[i]; zg[i] = beta ? zg2[i] : zg1[i];
                nb[i] = beta ? n2[i] : n1[i]; mb[i] = beta ? m2[i] : m1[i];
            }
            direction_blocks(rec, beta ? P_Y2 : P_Y0, beta ? P_Y3 : P_Y1, beta ? P_G2 : P_G0, beta ? P_G3 : P_G1,
                             G44, f4, kn.ga, abg, zg, nb, mb, PA, Xig, gg, false);
        }
#pragma unroll
        for (int i = 0; i < 9; i++) rec[boff(4, 4) + i] = G44[i];       // parked in its final slot until column 4 is done
#pragma unroll
        for (int i = 0; i < 3; i++) rec[F_OFF + 12 + i] = w * f4[i];  // f = Psi'^T sigma (:1324)
    }

    // tangent moments (second pass over the thickness rule; fast arithmetic)
    thickness<false>(st, lam, mu, thick, X, smu, nullptr, nullptr, nullptr, nullptr);

    // C' = Psi'^T D Psi' + G': column 4 first (reads the parked Y / G' blocks) ...
#define GFA_PL(R_) ((R_) % 2 == 0 ? PA : PB)
    double tmp[9], tmp2[9];
    {
        double C44[9];
#pragma unroll
        for (int i = 0; i < 9; i++) C44[i] = rec[boff(4, 4) + i];
        // H_R = sum_s D(R,s) Y_s for the four row groups at once: every parked Y block is read once
        double H[4][9];
#define GFA_HCOL(S_, PY_, OP_)                                                        \
        { const double* Ys = rec + park(PY_);                                         \
          { const DBlk d = getD<0, S_>(X, smu, drill); OP_(H[0], d, Ys); }            \
          { const DBlk d = getD<1, S_>(X, smu, drill); OP_(H[1], d, Ys); }            \
          { const DBlk d = getD<2, S_>(X, smu, drill); OP_(H[2], d, Ys); }            \
          { const DBlk d = getD<3, S_>(X, smu, drill); OP_(H[3], d, Ys); } }
        GFA_HCOL(0, P_Y0, dmul) GFA_HCOL(1, P_Y1, dmul_acc) GFA_HCOL(2, P_Y2, dmul_acc) GFA_HCOL(3, P_Y3, dmul_acc)
#undef GFA_HCOL
#define GFA_COL4(R_, PG_, PY_)                                                        \
        { mtm(tmp2, GFA_PL(R_), H[R_]);                                               \
          _Pragma("unroll") for (int i = 0; i < 9; i++) tmp2[i] += rec[park(PG_) + i]; \
          mtm_acc(C44, rec + park(PY_), H[R_]);                                       \
          put_block<R_, 4>(rec, w, tmp2); }
        GFA_COL4(0, P_G0, P_Y0) GFA_COL4(1, P_G1, P_Y1) GFA_COL4(2, P_G2, P_Y2) GFA_COL4(3, P_G3, P_Y3)
#undef GFA_COL4
        put_block<4, 4>(rec, w, C44);
    }
    // ... then the 10 upper blocks among the first four groups (overwrite the parking area)
#define GFA_UPPER(R_, S_)                                                             \
    { const DBlk d = getD<R_, S_>(X, smu, drill); dmul(tmp, d, GFA_PL(S_));           \
      mtm(tmp2, GFA_PL(R_), tmp); put_block<R_, S_>(rec, w, tmp2); }
    GFA_UPPER(0, 0) GFA_UPPER(0, 1) GFA_UPPER(0, 2) GFA_UPPER(0, 3)
    GFA_UPPER(1, 1) GFA_UPPER(1, 2) GFA_UPPER(1, 3)
    GFA_UPPER(2, 2) GFA_UPPER(2, 3)
    GFA_UPPER(3, 3)
#undef GFA_UPPER
#undef GFA_PL
}

// C'[(p,ii),(q,jj)] from the record.  `c` carries the column: rec + jj, rec + 3 jj and the packed
// index of (ii, jj) in a symmetric diagonal block for ii = 0, 1, 2.
struct Col { const double* recJ; const double* rec3J; const double* sym[3]; };
GFA_DI Col col_of(const double* rec, int jj) {
    Col c;
    c.recJ = rec + jj; c.rec3J = rec + 3 * jj;
    c.sym[0] = rec + jj;                                  // (0,jj): 0 1 2
    c.sym[1] = rec + (jj == 0 ? 1 : 2 + jj);              // (1,jj): 1 3 4
    c.sym[2] = rec + (jj == 0 ? 2 : 3 + jj);              // (2,jj): 2 4 5
    return c;
}
template <int P, int II, int Q>
GFA_DI double c_at(const Col& c) {
    if (P == Q && P < 4) return c.sym[II][boff(P, P)];
    if (P <= Q) return c.recJ[boff(P, Q) + 3 * II];
    return c.rec3J[boff(Q, P) + II];
}

// 6-point Cowper rule factors sum_g w4[g] N_a(c_g) / area, a = 0..5 (:2185-2314)
GFA_DI double cowper_factor(int a, double area) {
    const double c[6][4] = {
        { 0.816847572980459, 0.091576213509771, 0.091576213509771, 0.109951743655322 },
        { 0.091576213509771, 0.816847572980459, 0.091576213509771, 0.109951743655322 },
        { 0.091576213509771, 0.091576213509771, 0.816847572980459, 0.109951743655322 },
        { 0.108103018168070, 0.445948490915965, 0.445948490915965, 0.223381589678011 },
        { 0.445948490915965, 0.108103018168070, 0.445948490915965, 0.223381589678011 },
        { 0.445948490915965, 0.445948490915965, 0.108103018168070, 0.223381589678011 } };
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < 6; g++) {
        const double L1 = c[g][0], L2 = c[g][1], L3 = c[g][2], w = area * c[g][3];
        double N;
        switch (a) {
        case 0: N = (2 * L1 - 1) * L1; break;
        case 1: N = (2 * L2 - 1) * L2; break;
        case 2: N = (2 * L3 - 1) * L3; break;
        case 3: N = 4 * L1 * L2; break;
        case 4: N = 4 * L2 * L3; break;
        default: N = 4 * L3 * L1; break;
        }
        s += w * N;
    }
    return s;
}

// ---- Phase B: K = sum_g dN^T C' dN, upper blocks only ---------------------
// self-weight entry of P for translational column (b, jj): applied twice, as the
// reference does (Shell_1.cpp:1340-1375)
GFA_DI double self_weight(const EvalArgs& A, int e, int b, int jj) {
    if (A.gx == 0.0 && A.gy == 0.0 && A.gz == 0.0) return 0.0;
    const double gk = jj == 0 ? A.gx : jj == 1 ? A.gy : A.gz;
    const double* pr = A.props + SHELL_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    const double rho_t = __ldg(pr + 4) * __ldg(pr + 2);
    const double one = cowper_factor(b, __ldg(A.geo + 9 * (size_t)A.n_el + e)) * (rho_t * gk);
    return one + one;
}

// Phase B is written as short runtime loops on purpose: straight-line code beyond the ~32 KB the
// SM's instruction cache holds is fetched at one instruction per ~3 cycles per warp
// (tools/icache_probe.cu), which bounded this kernel (profiles/r01_notes.md).  Its other budget is the
// L1 data pipe: shared-memory loads were 41 % of that pipe's cycles (63 % with the arena stores, ncu
// round 2), so every record value is read ONCE per lane -- the C' entries of a column depend on the
// component jj only, not on the node, and are kept in registers across the nodes of the column.
//
// Translational columns, component jj, of ALL six nodes: pass kk takes B1 = kk and B2 = 5 - kk, rows
// u_0..u_B1 of the first and u_0..u_B2 of the second (upper triangle of the symmetric u-u part) = 7
// blocks for every kk; the rows both columns have share their shape-function loads.
// BATCH: `ke` is the element's corner of its batch region and stored block n lies at ke + 72 n (gfa_device.h:
// shell_batch_offset); otherwise `ke` is the element's own region (shell_stored_offset).
template <bool BATCH>
__device__ void uu_items(const EvalArgs& A, int e, double* ke, const double* rec0, int jj, double* pe) {
    constexpr int GRP = BATCH ? 7 * 72 : 64, BLK = BATCH ? 72 : 9;
    const double* S0 = rec0 + S_OFF;
    double c00[NGP][3], c02[NGP][3], c20[NGP][3], c22[NGP][3], f0[NGP], f2[NGP];
#pragma unroll
    for (int g = 0; g < NGP; g++) {
        const double* rec = rec0 + g * REC;
        const Col cc = col_of(rec, jj);
        c00[g][0] = c_at<0, 0, 0>(cc); c00[g][1] = c_at<0, 1, 0>(cc); c00[g][2] = c_at<0, 2, 0>(cc);
        c02[g][0] = c_at<0, 0, 2>(cc); c02[g][1] = c_at<0, 1, 2>(cc); c02[g][2] = c_at<0, 2, 2>(cc);
        c20[g][0] = c_at<2, 0, 0>(cc); c20[g][1] = c_at<2, 1, 0>(cc); c20[g][2] = c_at<2, 2, 0>(cc);
        c22[g][0] = c_at<2, 0, 2>(cc); c22[g][1] = c_at<2, 1, 2>(cc); c22[g][2] = c_at<2, 2, 2>(cc);
        f0[g] = rec[jj + F_OFF + 0]; f2[g] = rec[jj + F_OFF + 6];
    }
#pragma unroll 1
    for (int kk = 0; kk < 3; kk++) {
        const int B1 = kk, B2 = 5 - kk;
        double* Ke_el = ke + GRP * kk + jj;
        // m[g] = sum_q S_g[q, column] C'_g[(p, .), (q, jj)] for the two columns, p in {u,1 ; u,2}
        double m10[NGP][3], m12[NGP][3], m20[NGP][3], m22[NGP][3];
        double F1 = 0.0, F2 = 0.0;
#pragma unroll
        for (int g = 0; g < NGP; g++) {
            const double p1 = S0[g * REC + B1], q1 = S0[g * REC + 6 + B1], p2 = S0[g * REC + B2], q2 = S0[g * REC + 6 + B2];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                m10[g][i] = fma(q1, c02[g][i], p1 * c00[g][i]); m12[g][i] = fma(q1, c22[g][i], p1 * c20[g][i]);
                m20[g][i] = fma(q2, c02[g][i], p2 * c00[g][i]); m22[g][i] = fma(q2, c22[g][i], p2 * c20[g][i]);
            }
            F1 = fma(q1, f2[g], fma(p1, f0[g], F1));
            F2 = fma(q2, f2[g], fma(p2, f0[g], F2));
        }
#pragma unroll 1
        for (int a = 0; a <= B2; a++) {            // block (u_a, u_B1) at 64 kk + 9 a, block (u_a, u_B2) at 64 kk + 9 (kk + 1 + a)
            double n1[NGP], n2[NGP];
#pragma unroll
            for (int g = 0; g < NGP; g++) { n1[g] = S0[g * REC + a]; n2[g] = S0[g * REC + 6 + a]; }
            if (a <= B1) {
                double k[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
                for (int g = 0; g < NGP; g++)
#pragma unroll
                    for (int i = 0; i < 3; i++) k[i] = fma(n2[g], m12[g][i], fma(n1[g], m10[g][i], k[i]));
                double* o = Ke_el + BLK * a;
                o[0] = k[0]; o[3] = k[1]; o[6] = k[2];
            }
            double k[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
            for (int g = 0; g < NGP; g++)
#pragma unroll
                for (int i = 0; i < 3; i++) k[i] = fma(n2[g], m22[g][i], fma(n1[g], m20[g][i], k[i]));
            double* o = Ke_el + BLK * (kk + 1 + a);
            o[0] = k[0]; o[3] = k[1]; o[6] = k[2];
        }
        pe[3 * B1 + jj] = F1 - self_weight(A, e, B1, jj);
        pe[3 * B2 + jj] = F2 - self_weight(A, e, B2, jj);
    }
}

// Rotational columns, component jj, of ALL three mid-side nodes b: all 27 rows -- the six u rows are the
// upper u-alpha blocks (their transposes are the alpha-u blocks), the three alpha rows are the
// non-symmetric alpha-alpha blocks, each stored on its own.  Column (b, jj) starts at 192 + 84 b + jj.
template <bool BATCH>
__device__ void rot_items(const EvalArgs& A, int e, double* ke, const double* rec0, int jj, double* pe) {
    (void)A; (void)e;
    constexpr int COL = BATCH ? 9 * 72 : 84, BLK = BATCH ? 72 : 9;
    double* Ke_el = ke + (BATCH ? 21 * 72 : 192) + jj;
    const double* S0 = rec0 + S_OFF;
    {   // rows u_a: gradient groups {u,1 ; u,2} x {alpha,1 ; alpha,2 ; alpha}.  The C' entries are read once; the three
        // nodes b run through ONE copy of the code (instruction cache), re-reading only the shape functions
        double c01[NGP][3], c03[NGP][3], c04[NGP][3], c21[NGP][3], c23[NGP][3], c24[NGP][3], f1[NGP], f3[NGP], f4[NGP];
#pragma unroll
        for (int g = 0; g < NGP; g++) {
            const double* rec = rec0 + g * REC;
            const double* recJ = rec + jj;
            const Col cc = col_of(rec, jj);
            c01[g][0] = c_at<0, 0, 1>(cc); c01[g][1] = c_at<0, 1, 1>(cc); c01[g][2] = c_at<0, 2, 1>(cc);
            c03[g][0] = c_at<0, 0, 3>(cc); c03[g][1] = c_at<0, 1, 3>(cc); c03[g][2] = c_at<0, 2, 3>(cc);
            c04[g][0] = c_at<0, 0, 4>(cc); c04[g][1] = c_at<0, 1, 4>(cc); c04[g][2] = c_at<0, 2, 4>(cc);
            c21[g][0] = c_at<2, 0, 1>(cc); c21[g][1] = c_at<2, 1, 1>(cc); c21[g][2] = c_at<2, 2, 1>(cc);
            c23[g][0] = c_at<2, 0, 3>(cc); c23[g][1] = c_at<2, 1, 3>(cc); c23[g][2] = c_at<2, 2, 3>(cc);
            c24[g][0] = c_at<2, 0, 4>(cc); c24[g][1] = c_at<2, 1, 4>(cc); c24[g][2] = c_at<2, 2, 4>(cc);
            f1[g] = recJ[F_OFF + 3]; f3[g] = recJ[F_OFF + 9]; f4[g] = recJ[F_OFF + 12];
        }
#pragma unroll 1
        for (int b = 0; b < 3; b++) {
            double m0[NGP][3], m2[NGP][3], F = 0.0;
#pragma unroll
            for (int g = 0; g < NGP; g++) {
                const double s1 = S0[g * REC + 12 + b], s3 = S0[g * REC + 15 + b], s4 = S0[g * REC + 18 + b];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    m0[g][i] = fma(s4, c04[g][i], fma(s3, c03[g][i], s1 * c01[g][i]));
                    m2[g][i] = fma(s4, c24[g][i], fma(s3, c23[g][i], s1 * c21[g][i]));
                }
                F = fma(s4, f4[g], fma(s3, f3[g], fma(s1, f1[g], F)));
            }
#pragma unroll 1
            for (int a = 0; a < 6; a++) {
                double k[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
                for (int g = 0; g < NGP; g++) {
                    const double n1 = S0[g * REC + a], n2 = S0[g * REC + 6 + a];
#pragma unroll
                    for (int i = 0; i < 3; i++) k[i] = fma(n2, m2[g][i], fma(n1, m0[g][i], k[i]));
                }
                double* o = Ke_el + COL * b + BLK * a;
                o[0] = k[0]; o[3] = k[1]; o[6] = k[2];
            }
            pe[18 + 3 * b + jj] = F;
        }
    }
    {   // rows alpha_a: gradient groups {alpha,1 ; alpha,2 ; alpha} on both sides; one Gauss point at a time, the 27
        // sums (row node a, column node b, component i) stay in registers -- same order of additions as a loop
        // over the points inside each sum
        double k[3][3][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int i = 0; i < 3; i++) k[a][b][i] = 0.0;
#pragma unroll 1
        for (int g = 0; g < NGP; g++) {
            const double* rec = rec0 + g * REC;
            const Col cc = col_of(rec, jj);
            const double* Sg = S0 + g * REC;
            double m1[3][3], m3[3][3], m4[3][3];
#define GFA_MROW(M_, P_)                                                                                              \
            { const double x1[3] = { c_at<P_, 0, 1>(cc), c_at<P_, 1, 1>(cc), c_at<P_, 2, 1>(cc) };                    \
              const double x3[3] = { c_at<P_, 0, 3>(cc), c_at<P_, 1, 3>(cc), c_at<P_, 2, 3>(cc) };                    \
              const double x4[3] = { c_at<P_, 0, 4>(cc), c_at<P_, 1, 4>(cc), c_at<P_, 2, 4>(cc) };                    \
              _Pragma("unroll") for (int b = 0; b < 3; b++) {                                                         \
                  const double s1 = Sg[12 + b], s3 = Sg[15 + b], s4 = Sg[18 + b];                                     \
                  _Pragma("unroll") for (int i = 0; i < 3; i++) M_[b][i] = fma(s4, x4[i], fma(s3, x3[i], s1 * x1[i])); } }
            GFA_MROW(m1, 1) GFA_MROW(m3, 3) GFA_MROW(m4, 4)
#undef GFA_MROW
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const double a1 = Sg[12 + a], a2 = Sg[15 + a], a0 = Sg[18 + a];
#pragma unroll
                for (int b = 0; b < 3; b++)
#pragma unroll
                    for (int i = 0; i < 3; i++) k[a][b][i] = fma(a0, m4[b][i], fma(a2, m3[b][i], fma(a1, m1[b][i], k[a][b][i])));
            }
        }
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) {
                double* o = Ke_el + COL * b + BLK * (6 + a);
                o[0] = k[a][b][0]; o[3] = k[a][b][1]; o[6] = k[a][b][2];
            }
    }
}

// One batch of `ne` <= EPW elements at list positions k0 .. k0 + ne - 1, evaluated by one warp; `smem` is the
// warp's own smem_bytes(EPW) region.  Shared by the classic evaluation kernel and the fused ring kernel.
template <int EPW, bool BATCH = false>
__device__ __forceinline__ void eval_batch(const EvalArgs& A, int k0, int ne, double* smem, int lane) {
    static_assert(EPW * 3 <= 32, "one lane per (element, component) item");
    static_assert(!BATCH || EPW == SHELL_BATCH, "the batch layout of the arena is the 8-element batch of this kernel");
    if (lane < ne * NGP) physics(A, eval_element(A, k0 + lane / NGP), lane % NGP, smem + lane * REC);
    __syncwarp();
    double* pe = smem + EPW * NGP * REC;      // the batch's P, written out in whole sectors below
    if (lane < ne * 3) {
        const int el = lane / 3, jj = lane % 3;
        const double* rec0 = smem + el * NGP * REC;
        const int e = eval_element(A, k0 + el);
        // BATCH: k0 is a multiple of 8 (classic arena, no element list): the batch's region, this element's corner
        double* ke = BATCH ? A.Ke + (size_t)(k0 / SHELL_BATCH) * (SHELL_BATCH * SHELL_ARENA) + 9 * el : eval_ke(A, k0 + el, SHELL_ARENA);
        uu_items<BATCH>(A, e, ke, rec0, jj, pe + 27 * el);
        rot_items<BATCH>(A, e, ke, rec0, jj, pe + 27 * el);
    }
    __syncwarp();
    if (!A.elist) { for (int i = lane; i < ne * 27; i += 32) A.Pe[(size_t)k0 * 27 + i] = pe[i]; }
    else { for (int i = lane; i < ne * 27; i += 32) A.Pe[(size_t)A.elist[k0 + i / 27] * 27 + i % 27] = pe[i]; }
    __syncwarp();
}

template <int EPW, bool BATCH = false>
__global__ void __launch_bounds__(32) eval_kernel(EvalArgs A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    for (long long batch = blockIdx.x; A.e_begin + batch * EPW < A.e_end; batch += gridDim.x) {
        const int k0 = A.e_begin + (int)(batch * EPW);
        eval_batch<EPW, BATCH>(A, k0, min(EPW, A.e_end - k0), smem, lane);
    }
}

// ShellLoad follower pressure on the current configuration (Shell_1::MountShellSpecialLoads, :1392-1467): 6-point
// Cowper rule (the w4 / N4 tables of PreCalc, :2185-2314), t_b = e_br + N,b u with u = copy - ref + displacements,
// n = t1 x t2 / |t1 x t2|, q = -p n; AreaUpdate 0: P -= w4 N^T q, K += w4 p N^T (I - n n^T)/|t1 x t2| (skew(t1) N,2 -
// skew(t2) N,1); AreaUpdate 1: the same with the Jacobian |t1 x t2| kept in the force and no projector.
// One thread per (entry, node pair a, b): the 3 x 3 block (a, b) of the load stiffness, thread b = 0 also the force.
__global__ void load_kernel(EvalArgs A, ShellLoadArgs Ld) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 36LL * Ld.n_entries) return;
    const int entry = (int)(t / 36), pair = (int)(t % 36), a = pair / 6, b = pair % 6;
    const int e = Ld.elem[entry], ld = Ld.load[entry];
    const double pressure = Ld.pressure[ld];
    const bool area_update = Ld.area_update[ld] != 0;
    int nd[6];
    double x[6][3], u[6][3];
    load_nodes(A, e, nd, x);
#pragma unroll
    for (int n = 0; n < 6; n++)
#pragma unroll
        for (int c = 0; c < 3; c++) u[n][c] = A.copy[6 * (size_t)nd[n] + c] - x[n][c] + A.disp[6 * (size_t)nd[n] + c];
    Frame fr; frame_of(x, fr);
    const double cw[6][4] = {
        { 0.816847572980459, 0.091576213509771, 0.091576213509771, 0.109951743655322 },
        { 0.091576213509771, 0.816847572980459, 0.091576213509771, 0.109951743655322 },
        { 0.091576213509771, 0.091576213509771, 0.816847572980459, 0.109951743655322 },
        { 0.108103018168070, 0.445948490915965, 0.445948490915965, 0.223381589678011 },
        { 0.445948490915965, 0.108103018168070, 0.445948490915965, 0.223381589678011 },
        { 0.445948490915965, 0.445948490915965, 0.108103018168070, 0.223381589678011 } };
    double K[9], P[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
    for (int i = 0; i < 9; i++) K[i] = 0.0;
    for (int g = 0; g < 6; g++) {
        const double L[3] = { cw[g][0], cw[g][1], cw[g][2] };
        const double w4 = fr.area * cw[g][3];
        double N[6], N1[6], N2[6];
        N[0] = (2 * L[0] - 1) * L[0]; N[1] = (2 * L[1] - 1) * L[1]; N[2] = (2 * L[2] - 1) * L[2];
        N[3] = 4 * L[0] * L[1]; N[4] = 4 * L[1] * L[2]; N[5] = 4 * L[2] * L[0];
#pragma unroll
        for (int k = 0; k < 3; k++) { N1[k] = 4 * fr.Lx[k] * L[k] - fr.Lx[k]; N2[k] = 4 * fr.Ly[k] * L[k] - fr.Ly[k]; }
        N1[3] = 4 * fr.Lx[0] * L[1] + 4 * L[0] * fr.Lx[1]; N1[4] = 4 * fr.Lx[1] * L[2] + 4 * L[1] * fr.Lx[2]; N1[5] = 4 * fr.Lx[2] * L[0] + 4 * L[2] * fr.Lx[0];
        N2[3] = 4 * fr.Ly[0] * L[1] + 4 * L[0] * fr.Ly[1]; N2[4] = 4 * fr.Ly[1] * L[2] + 4 * L[1] * fr.Ly[2]; N2[5] = 4 * fr.Ly[2] * L[0] + 4 * L[2] * fr.Ly[0];
        double t1[3], t2[3], c[3], n[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double s1 = fr.R[k], s2 = fr.R[3 + k];
#pragma unroll
            for (int m = 0; m < 6; m++) { s1 += N1[m] * u[m][k]; s2 += N2[m] * u[m][k]; }
            t1[k] = s1; t2[k] = s2;
        }
        cross3(c, t1, t2);
        const double jac = norm3(c);
#pragma unroll
        for (int k = 0; k < 3; k++) n[k] = c[k] / jac;
        // M = skew(t1) N2[b] - skew(t2) N1[b]
        double S1[9], S2[9], Mx[9];
        skew3(S1, t1); skew3(S2, t2);
#pragma unroll
        for (int i = 0; i < 9; i++) Mx[i] = S1[i] * N2[b] - S2[i] * N1[b];
        const double scale = area_update ? w4 * jac : w4;
        if (!area_update) {
            double Pm[9];                      // (I - n n^T) / jac * M
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    double v = 0.0;
#pragma unroll
                    for (int k = 0; k < 3; k++) v += ((i == k ? 1.0 : 0.0) - n[i] * n[k]) * Mx[3 * k + j];
                    Pm[3 * i + j] = v / jac;
                }
#pragma unroll
            for (int i = 0; i < 9; i++) K[i] += w4 * pressure * N[a] * Pm[i];
        } else {
#pragma unroll
            for (int i = 0; i < 9; i++) K[i] += w4 * pressure * N[a] * Mx[i];
        }
        if (b == 0)
#pragma unroll
            for (int k = 0; k < 3; k++) P[k] -= scale * N[a] * (-pressure * n[k]);
    }
    double* o = Ld.out + (size_t)entry * SHELL_LOAD_REC;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[(3 * a + i) * 18 + 3 * b + j] = K[3 * i + j];
    if (b == 0) { o[324 + 3 * a] = P[0]; o[324 + 3 * a + 1] = P[1]; o[324 + 3 * a + 2] = P[2]; }
}

// Shell_1::SaveLagrange (:1650-1664): one thread per Gauss point.
__global__ void commit_kernel(EvalArgs A) {
    const size_t n_gp = (size_t)A.n_el * NGP;
    const size_t gp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= n_gp) return;
    const int e = (int)(gp / NGP), g = (int)(gp % NGP);
    int nd[6];
#pragma unroll
    for (int n = 0; n < 6; n++) nd[n] = __ldg(A.conn + 6 * (size_t)e + n);
    Frame fr; Shape sh;
    load_precalc(A, e, g, fr, sh);
    Kin kn; interpolate(A, nd, fr, sh, kn);
    double gg, Qd[9], Xi[9], Qi[9], Qn[9], t3[3], dk[3];
    s_rodrigues(kn.a, gg, Qd, Xi);
#pragma unroll
    for (int i = 0; i < 9; i++) Qi[i] = A.state[i * n_gp + gp];
    s_mtv(t3, Xi, kn.a1); s_mtv(dk, Qi, t3);
#pragma unroll
    for (int i = 0; i < 3; i++) A.state[(15 + i) * n_gp + gp] = s_add(dk[i], A.state[(15 + i) * n_gp + gp]);
    s_mtv(t3, Xi, kn.a2); s_mtv(dk, Qi, t3);
#pragma unroll
    for (int i = 0; i < 3; i++) A.state[(18 + i) * n_gp + gp] = s_add(dk[i], A.state[(18 + i) * n_gp + gp]);
    s_mm(Qn, Qd, Qi);
#pragma unroll
    for (int i = 0; i < 9; i++) A.state[i * n_gp + gp] = Qn[i];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        A.state[(9 + i) * n_gp + gp] = s_add(kn.u1[i], A.state[(9 + i) * n_gp + gp]);
        A.state[(12 + i) * n_gp + gp] = s_add(kn.u2[i], A.state[(12 + i) * n_gp + gp]);
    }
}


// Gauss-point results Mount keeps for WriteResults / WriteVTK / WriteMonitor (Shell_1.cpp:624-707,
// Monitor.cpp:494): eta_r1, eta_r2, kappa_r1, kappa_r2, n_r1, n_r2, m_r1 (with the drilling term,
// :1214-1217), m_r2 per point and the element's strain energy (:1150-1161, :1326).  Computed on demand
// from the current increments and the committed state -- not on the per-iteration path.
// out[e * 73]: strain_energy, then per point g: eta1 eta2 kappa1 kappa2 n1 n2 m1 m2 (3 each).
__global__ void results_kernel(EvalArgs A, double* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n_el) return;
    int nd[6];
#pragma unroll
    for (int n = 0; n < 6; n++) nd[n] = __ldg(A.conn + 6 * (size_t)e + n);
    const double* pr = A.props + SHELL_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    const double lam = __ldg(pr), mu = __ldg(pr + 1), thick = __ldg(pr + 2), drill = __ldg(pr + 3);
    const size_t n_gp = (size_t)A.n_el * NGP;
    double energy = 0.0;
    double* o = out + (size_t)e * 73;
#pragma unroll 1
    for (int g = 0; g < NGP; g++) {
        const size_t gp = (size_t)e * NGP + g;
        Frame fr; Shape sh; Kin kn;
        load_precalc(A, e, g, fr, sh);
        interpolate(A, nd, fr, sh, kn);
        Strains st;
        {
            double gg, Xi[9], Qt[9], z1[3], z2[3], Qd[9], Qi[9], Q[9], t3[3];
#pragma unroll
            for (int i = 0; i < 9; i++) Qi[i] = A.state[i * n_gp + gp];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                z1[i] = s_add(kn.u1[i], A.state[(9 + i) * n_gp + gp]);
                z2[i] = s_add(kn.u2[i], A.state[(12 + i) * n_gp + gp]);
            }
            s_rodrigues(kn.a, gg, Qd, Xi);
            s_mm(Q, Qd, Qi);
            m_transpose(Qt, Q);
            s_mv(st.eta1, Qt, z1); st.eta1[0] = s_sub(st.eta1[0], 1.0);
            s_mv(st.eta2, Qt, z2); st.eta2[1] = s_sub(st.eta2[1], 1.0);
            s_mtv(t3, Xi, kn.a1); s_mtv(st.kap1, Qi, t3);
            s_mtv(t3, Xi, kn.a2); s_mtv(st.kap2, Qi, t3);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                st.kap1[i] = s_add(st.kap1[i], A.state[(15 + i) * n_gp + gp]);
                st.kap2[i] = s_add(st.kap2[i], A.state[(18 + i) * n_gp + gp]);
            }
        }
        double X[3][3][4], smu = 0.0, n1[3], n2[3], m1[3], m2[3];
        thickness<true>(st, lam, mu, thick, X, smu, n1, n2, m1, m2);
        m1[2] = s_mul(drill, st.kap1[2]);
        m2[2] = s_mul(drill, st.kap2[2]);
        // specific strain energy through the thickness (:1150-1161)
        double psi_t = 0.0;
        const double jac = thick / 2.0;
#pragma unroll 1
        for (int q = 0; q < 3; q++) {
            const double csi = (q == 0) ? -0.77459666924148337703585307995648 : (q == 1) ? 0.0 : 0.77459666924148337703585307995648;
            const double al2 = (q == 1) ? 0.88888888888888888888888888888889 : 0.55555555555555555555555555555556;
            const double zeta = thick * csi / 2.0;
            const double g11 = st.eta1[0] + zeta * st.kap1[1], g12 = st.eta1[1] - zeta * st.kap1[0], g13 = st.eta1[2];
            const double g21 = st.eta2[0] + zeta * st.kap2[1], g22 = st.eta2[1] - zeta * st.kap2[0], g23 = st.eta2[2];
            const double jb = (1.0 + g11) * (1.0 + g22) - g12 * g21;
            const double g33 = sqrt((lam + 2.0 * mu) / (lam * jb * jb + 2.0 * mu)) - 1.0;
            const double jF = jb * (1.0 + g33);
            const double I1 = (1.0 + g11) * (1.0 + g11) + g12 * g12 + g13 * g13 + g21 * g21 + (1.0 + g22) * (1.0 + g22) + g23 * g23 + (1.0 + g33) * (1.0 + g33);
            const double psi = 0.5 * lam * (0.5 * (jF * jF - 1.0) - log(jF)) + 0.5 * mu * (I1 - 3.0 - 2.0 * log(jF));
            psi_t += al2 * jac * psi;
        }
        energy += (fr.area / 3.0) * psi_t;                        // alpha1 (:2364), (:1326)
        double* og = o + 1 + 24 * g;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            og[i] = st.eta1[i]; og[3 + i] = st.eta2[i]; og[6 + i] = st.kap1[i]; og[9 + i] = st.kap2[i];
            og[12 + i] = n1[i]; og[15 + i] = n2[i]; og[18 + i] = m1[i]; og[21 + i] = m2[i];
        }
    }
    o[0] = energy;
}

} // namespace shell

// =========================================================================
// Beam_1
// =========================================================================
namespace beam {

constexpr int EPW = 16;
constexpr int NGP = 2;
constexpr int C_OFF = 0;       // full 9x9 C' (row-major), times weight
constexpr int F_OFF = 81;      // f (9)
constexpr int S_OFF = 90;      // dN[3], N[3]
constexpr int W_OFF = 96;      // jacobian * rho * A (gravity multiplier without l_factor*G)
constexpr int REC = 97;
constexpr int SMEM_BYTES = EPW * NGP * REC * 8;

struct Geo { double R[9], e3r[3], jac; double N[3], dN[3]; };

GFA_DI void geometry(const EvalArgs& A, int e, int g, const int* nd, const double* pr, Geo& go) {
#pragma unroll
    for (int i = 0; i < 9; i++) go.R[i] = __ldg(pr + 36 + i);
    double d[3];
#pragma unroll
    for (int c = 0; c < 3; c++) d[c] = s_sub(__ldg(A.xyz + 3 * (size_t)nd[2] + c), __ldg(A.xyz + 3 * (size_t)nd[0] + c));
    const double len = s_norm3(d);
    double e3[3];
    const double inv = 1.0 / len;
#pragma unroll
    for (int c = 0; c < 3; c++) e3[c] = s_mul(d[c], inv);
    s_mv(go.e3r, go.R, e3);                                       // Beam_1.cpp:609-614
    const double T0 = A.pret ? __ldg(A.pret + e) : 0.0;
    const double du0 = T0 / __ldg(pr + 14);                       // D(2,2) = EA  (:616-621)
    const double length = len / s_add(1.0, du0);
    go.jac = length / 2.0;
    const double xi = g == 0 ? -0.577350269189626 : 0.577350269189626;
    const double ij = 1.0 / go.jac;
    go.N[0] = s_mul(s_mul(0.5, xi), s_sub(xi, 1.0)); go.N[1] = s_sub(1.0, s_mul(xi, xi)); go.N[2] = s_mul(s_mul(0.5, xi), s_add(1.0, xi));
    go.dN[0] = s_mul(ij, s_sub(xi, 0.5)); go.dN[1] = s_mul(ij, s_mul(-2.0, xi)); go.dN[2] = s_mul(ij, s_add(0.5, xi));
}
struct Kin { double a[3], da[3], du[3]; };
GFA_DI void interpolate(const EvalArgs& A, const int* nd, const Geo& go, Kin& k) {
    double ga[3], gda[3], gdu[3];
#pragma unroll
    for (int n = 0; n < 3; n++) {
        // one 48-byte record per node, 16-byte aligned: three vector loads
        const double2* d = reinterpret_cast<const double2*>(A.disp + 6 * (size_t)nd[n]);
        const double2 u01 = __ldg(d), u2r0 = __ldg(d + 1), r12 = __ldg(d + 2);
        const double uu[3] = { u01.x, u01.y, u2r0.x }, rr[3] = { u2r0.y, r12.x, r12.y };
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double r = rr[c], u = uu[c];
            ga[c] = n == 0 ? s_mul(r, go.N[0]) : s_add(ga[c], s_mul(r, go.N[n]));
            gda[c] = n == 0 ? s_mul(r, go.dN[0]) : s_add(gda[c], s_mul(r, go.dN[n]));
            gdu[c] = n == 0 ? s_mul(u, go.dN[0]) : s_add(gdu[c], s_mul(u, go.dN[n]));
        }
    }
    s_mv(k.a, go.R, ga); s_mv(k.da, go.R, gda); s_mv(k.du, go.R, gdu);   // :743-746
}

// Section matrix of the in-scope sections: D = blockdiag(diag(d0), [[a, c, 0], [c, b, 0], [0, 0, t]])
// (Beam_1.cpp:560-580: GA, GA, EA | EI1, EI2, EI12, GIt; Pipe_1.cpp:1108-1114: diagonal); the property rows
// are built by gfa_create with exactly these non-zeros.
struct DSec { double d0[3], a, b, c, t; };
GFA_DI DSec load_dsec(const double* pr) {
    DSec D;
    D.d0[0] = __ldg(pr + 0); D.d0[1] = __ldg(pr + 7); D.d0[2] = __ldg(pr + 14);
    D.a = __ldg(pr + 21); D.c = __ldg(pr + 22); D.b = __ldg(pr + 28); D.t = __ldg(pr + 35);
    return D;
}
// o = D(0,0) P  and  o = D(1,1) P  (3x3 blocks of the 6x6 section matrix; D(0,1) = D(1,0) = 0)
GFA_DI void d00_mul(double* o, const DSec& D, const double* P) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = D.d0[i] * P[3 * i + j];
}
GFA_DI void d11_mul(double* o, const DSec& D, const double* P) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        o[j] = D.a * P[j] + D.c * P[3 + j];
        o[3 + j] = D.c * P[j] + D.b * P[3 + j];
        o[6 + j] = D.t * P[6 + j];
    }
}
GFA_DI void store9(double* rec, int p, int q, double w, const double* M) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) rec[C_OFF + 9 * (3 * p + i) + 3 * q + j] = w * M[3 * i + j];
}

// Beam_1::Mount at one Gauss point (:699-831), rotated to global axes
__device__ void physics(const EvalArgs& A, int e, int g, double* rec) {
    int nd[3];
#pragma unroll
    for (int n = 0; n < 3; n++) nd[n] = __ldg(A.conn + 3 * (size_t)e + n);
    const double* pr = A.props + BEAM_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    Geo go; geometry(A, e, g, nd, pr, go);
    Kin kn; interpolate(A, nd, go, kn);
    const DSec D = load_dsec(pr);
    const size_t n_gp = (size_t)A.n_el * NGP, gp = (size_t)e * NGP + g;
    double Qi[9], dz[3], ki[3];
#pragma unroll
    for (int i = 0; i < 9; i++) Qi[i] = A.state[i * n_gp + gp];
#pragma unroll
    for (int i = 0; i < 3; i++) { dz[i] = s_add(kn.du[i], A.state[(9 + i) * n_gp + gp]); ki[i] = A.state[(12 + i) * n_gp + gp]; }

    double gg, Qd[9], Xi[9], dXi[9], Q[9], Qt[9];
    s_rodrigues(kn.a, gg, Qd, Xi);
    d_xi(dXi, kn.a, kn.da, gg, Xi);
    s_mm(Q, Qd, Qi);
    m_transpose(Qt, Q);
    double eps[6], sig[6], t3[3];
    s_mv(eps, Qt, dz);
#pragma unroll
    for (int i = 0; i < 3; i++) eps[i] = s_sub(eps[i], go.e3r[i]);   // eta_r (:778), strict
    s_mtv(t3, Xi, kn.da); s_mtv(eps + 3, Qi, t3);
#pragma unroll
    for (int i = 0; i < 3; i++) eps[3 + i] = s_add(eps[3 + i], ki[i]);   // kappa_r (:779)
    // sigma_r = D eps (:788), strict; the zero entries of D add exact zeros in the reference's full sum
#pragma unroll
    for (int i = 0; i < 3; i++) sig[i] = s_mul(D.d0[i], eps[i]);
    sig[3] = s_add(s_mul(D.a, eps[3]), s_mul(D.c, eps[4]));
    sig[4] = s_add(s_mul(D.c, eps[3]), s_mul(D.b, eps[4]));
    sig[5] = s_mul(D.t, eps[5]);
    double n[3], m[3];
    mv(n, Q, sig); mv(m, Q, sig + 3);                            // spatial resultants (:796-797)

    // B' = B (I3 (x) R): B(0,0)=Qt, B(0,2)=Qt dZ Xi, B(1,1)=Qt Xi, B(1,2)=Qt dXi (:758-774)
    double XiR[9], B00[9], B02[9], B11[9], B12[9], tmp[9], tmp2[9];
    mm(B00, Qt, go.R);
    mm(XiR, Xi, go.R);
    mm(B11, Qt, XiR);
    skew_mul(tmp, dz, XiR); mm(B02, Qt, tmp);
    mm(tmp, dXi, go.R); mm(B12, Qt, tmp);

    const double w = 1.0 * go.jac;                                // alpha1 * jacobian (:826)
    {
        double f[9], t[3];
        mtv(f + 0, B00, sig); mtv(f + 3, B11, sig + 3);
        mtv(f + 6, B02, sig); mtv(t, B12, sig + 3);
        f[6] += t[0]; f[7] += t[1]; f[8] += t[2];
#pragma unroll
        for (int i = 0; i < 9; i++) rec[F_OFF + i] = w * f[i];    // (:828)
    }
    // geometric blocks (:799-822), rotated
    double G02[9], G20[9], G22[9], G12[9], G21[9];
    {
        const double h = gg;
        double SnXi[9], SmXi[9], V[9], zn[3], A1[9], XtSmXi[9];
        skew_mul(SnXi, n, Xi);
#pragma unroll
        for (int i = 0; i < 9; i++) G02[i] = -SnXi[i];            // -skew(n) Xi
        {   // Xi^T skew(n)
            double Sn[9]; skew3(Sn, n); mtm(G20, Xi, Sn);
        }
        skew_mul(A1, dz, SnXi); mtm(G22, Xi, A1);                 // Xi^T (dZ skew(n)) Xi
        cross3(zn, dz, n); v_op(V, kn.a, zn, h);
#pragma unroll
        for (int i = 0; i < 9; i++) G22[i] -= V[i];
        dv_op(V, kn.a, kn.da, m, h);
#pragma unroll
        for (int i = 0; i < 9; i++) G22[i] += V[i];
        skew_mul(SmXi, m, Xi);
        mtm(A1, dXi, SmXi);
#pragma unroll
        for (int i = 0; i < 9; i++) G22[i] -= A1[i];
        v_op(G21, kn.a, m, h);                                    // V(alpha, m)
        mtm(XtSmXi, Xi, SmXi);
#pragma unroll
        for (int i = 0; i < 9; i++) G12[i] = G21[i] - XtSmXi[i];
        mm(tmp, G02, go.R); mtm(G02, go.R, tmp);
        mm(tmp, G20, go.R); mtm(G20, go.R, tmp);
        mm(tmp, G22, go.R); mtm(G22, go.R, tmp);
        mm(tmp, G12, go.R); mtm(G12, go.R, tmp);
        mm(tmp, G21, go.R); mtm(G21, go.R, tmp);
    }
    // C' = B'^T D B' + G'  (:776, :824-826) with D = blockdiag(D00, D11): the blocks (1,0) and (0,1) of C' vanish
    // identically (they are neither stored nor read: congruence_cols skips them)
    double DB0[9], DB1[9];
    // column 0: B(.,0) = [B00; 0]
    d00_mul(DB0, D, B00);
    mtm(tmp, B00, DB0); store9(rec, 0, 0, w, tmp);
    mtm(tmp, B02, DB0);
#pragma unroll
    for (int i = 0; i < 9; i++) tmp[i] += G20[i];
    store9(rec, 2, 0, w, tmp);
    // column 1: B(.,1) = [0; B11]
    d11_mul(DB1, D, B11);
    mtm(tmp, B11, DB1); store9(rec, 1, 1, w, tmp);
    mtm(tmp, B12, DB1);
#pragma unroll
    for (int i = 0; i < 9; i++) tmp[i] += G21[i];
    store9(rec, 2, 1, w, tmp);
    // column 2: B(.,2) = [B02; B12]
    d00_mul(DB0, D, B02); d11_mul(DB1, D, B12);
    mtm(tmp, B00, DB0);
#pragma unroll
    for (int i = 0; i < 9; i++) tmp[i] += G02[i];
    store9(rec, 0, 2, w, tmp);
    mtm(tmp, B11, DB1);
#pragma unroll
    for (int i = 0; i < 9; i++) tmp[i] += G12[i];
    store9(rec, 1, 2, w, tmp);
    mtm(tmp2, B02, DB0); mtm_acc(tmp2, B12, DB1);
#pragma unroll
    for (int i = 0; i < 9; i++) tmp2[i] += G22[i];
    store9(rec, 2, 2, w, tmp2);

#pragma unroll
    for (int i = 0; i < 3; i++) { rec[S_OFF + i] = go.dN[i]; rec[S_OFF + 3 + i] = go.N[i]; }
    rec[W_OFF] = go.jac * __ldg(pr + 45);
}

// local DOF order: node-major [u_a(3), alpha_a(3)] (Beam_1.cpp:1439-1444)
// One item = component jj of the translational (ROT = false) or rotational columns of ALL three nodes: the C' entries
// of a column depend on jj and on the kind of column only, so they are read once and serve the three nodes (the
// congruence phases of all three element kernels are bound by the L1 data pipe, profiles/r02_notes.md).  Sums and
// their order are those of one column at a time.
template <bool ROT>
__device__ void congruence_cols(const EvalArgs& A, int e, double* ke, const double* rec0, int jj) {
    double K[3][18];
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
        for (int i = 0; i < 18; i++) K[b][i] = 0.0;
    double F[3] = { 0.0, 0.0, 0.0 }, fe[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
    for (int g = 0; g < NGP; g++) {
        const double* rec = rec0 + g * REC;
        const double* S = rec + S_OFF;
        // gradient groups 0: u', 1: alpha', 2: alpha; the blocks (1,0) and (0,1) of C' are identically zero
        double c0[3], c1a[3], c1b[3], c2a[3], c2b[3];
#pragma unroll
        for (int ii = 0; ii < 3; ii++) {
            const double* r0 = rec + C_OFF + 9 * ii + jj;
            const double* r1 = rec + C_OFF + 9 * (3 + ii) + jj;
            const double* r2 = rec + C_OFF + 9 * (6 + ii) + jj;
            if (ROT) { c0[ii] = r0[6]; c1a[ii] = r1[3]; c1b[ii] = r1[6]; c2a[ii] = r2[3]; c2b[ii] = r2[6]; }
            else { c0[ii] = r0[0]; c2a[ii] = r2[0]; c1a[ii] = 0.0; c1b[ii] = 0.0; c2b[ii] = 0.0; }
        }
        const double f0 = ROT ? rec[F_OFF + 3 + jj] : rec[F_OFF + jj], f1 = ROT ? rec[F_OFF + 6 + jj] : 0.0;
        double Sv[6];
#pragma unroll
        for (int i = 0; i < 6; i++) Sv[i] = S[i];
        const double wg = ROT ? 0.0 : rec[W_OFF];
#pragma unroll
        for (int b = 0; b < 3; b++) {
            double m0[3], m1[3], m2[3];
#pragma unroll
            for (int ii = 0; ii < 3; ii++) {
                if (ROT) {
                    m0[ii] = Sv[3 + b] * c0[ii];
                    m1[ii] = Sv[b] * c1a[ii] + Sv[3 + b] * c1b[ii];
                    m2[ii] = Sv[b] * c2a[ii] + Sv[3 + b] * c2b[ii];
                } else {
                    m0[ii] = Sv[b] * c0[ii];
                    m1[ii] = 0.0;
                    m2[ii] = Sv[b] * c2a[ii];
                }
            }
            F[b] += ROT ? Sv[b] * f0 + Sv[3 + b] * f1 : Sv[b] * f0;
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int ii = 0; ii < 3; ii++) {
                    K[b][6 * a + ii] += Sv[a] * m0[ii];
                    K[b][6 * a + 3 + ii] += ROT ? Sv[a] * m1[ii] + Sv[3 + a] * m2[ii] : Sv[3 + a] * m2[ii];
                }
            if (!ROT) {
                const double gk = jj == 0 ? A.gx : jj == 1 ? A.gy : A.gz;
                fe[b] += wg * Sv[3 + b] * gk;                     // mult * N_b * G (:868-878)
            }
        }
    }
#pragma unroll
    for (int b = 0; b < 3; b++) {
        const int col = 6 * b + (ROT ? 3 : 0) + jj;
        // stored blocks of this column block: rows 0..cb and, for a rotational column, the rotational rows below
        const int cb = col / 3;
        double* Ke = ke + (col % 3);
#pragma unroll
        for (int rb = 0; rb < 6; rb++)
            if (beam_is_stored(rb, cb)) {
                double* o = Ke + beam_stored_offset(rb, cb);
                o[0] = K[b][3 * rb]; o[3] = K[b][3 * rb + 1]; o[6] = K[b][3 * rb + 2];
            }
        A.Pe[(size_t)e * 18 + col] = F[b] - fe[b];
    }
}

// one batch of `ne` <= EPW beams at list positions k0 .. (see shell::eval_batch)
__device__ __forceinline__ void eval_batch(const EvalArgs& A, int k0, int ne, double* smem, int lane) {
    if (lane < ne * NGP) physics(A, eval_element(A, k0 + lane / NGP), lane % NGP, smem + lane * REC);
    __syncwarp();
    for (int it = lane; it < ne * 3; it += 32) {
        const int el = it / 3;
        congruence_cols<false>(A, eval_element(A, k0 + el), eval_ke(A, k0 + el, BEAM_ARENA), smem + el * NGP * REC, it % 3);
    }
    for (int it = lane; it < ne * 3; it += 32) {
        const int el = it / 3;
        congruence_cols<true>(A, eval_element(A, k0 + el), eval_ke(A, k0 + el, BEAM_ARENA), smem + el * NGP * REC, it % 3);
    }
    __syncwarp();
}
__global__ void __launch_bounds__(32) eval_kernel(EvalArgs A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    for (long long batch = blockIdx.x; A.e_begin + batch * EPW < A.e_end; batch += gridDim.x) {
        const int k0 = A.e_begin + (int)(batch * EPW);
        eval_batch(A, k0, min(EPW, A.e_end - k0), smem, lane);
    }
}

// PipeLoad internal pressure: Pipe_1::MountPipeSpecialLoads (Pipe_1.cpp:1443-1494), one thread per (load, element)
// entry.  The kinematics are those of Mount at the displacements of the last assembly (strict chain as in physics());
// the scalar g of the LAST Gauss point serves both points, as the member Mount leaves behind does in the reference
// (Pipe_1.cpp:887, 1470).  Record (SHELL_LOAD_REC doubles, global axes): what is ADDED to the stiffness, row-major
// 18 x 18 in the element's local DOF order, then what is added to P_loading.
__global__ void pipe_load_kernel(EvalArgs A, ShellLoadArgs Ld) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Ld.n_entries) return;
    const int e = Ld.elem[t];
    const double p0i = Ld.pressure[Ld.load[t]];
    double* out = Ld.out + (size_t)t * SHELL_LOAD_REC;
    for (int i = 0; i < SHELL_LOAD_REC; i++) out[i] = 0.0;
    int nd[3];
#pragma unroll
    for (int n = 0; n < 3; n++) nd[n] = A.conn[3 * (size_t)e + n];
    const double* pr = A.props + BEAM_PROP_STRIDE * (size_t)A.prop[e];
    const double Aint = pr[51];
    const size_t n_gp = (size_t)A.n_el * NGP;
    double g_last = 0.0;
    {   // the second point's g
        Geo go; geometry(A, e, 1, nd, pr, go);
        Kin kn; interpolate(A, nd, go, kn);
        double Qd[9], Xi[9];
        s_rodrigues(kn.a, g_last, Qd, Xi);
    }
#pragma unroll 1
    for (int g = 0; g < NGP; g++) {
        const size_t gp = (size_t)e * NGP + g;
        Geo go; geometry(A, e, g, nd, pr, go);
        Kin kn; interpolate(A, nd, go, kn);
        double Qi[9], dz[3], ki[3];
        for (int i = 0; i < 9; i++) Qi[i] = A.state[i * n_gp + gp];
        for (int i = 0; i < 3; i++) { dz[i] = s_add(kn.du[i], A.state[(9 + i) * n_gp + gp]); ki[i] = A.state[(12 + i) * n_gp + gp]; }
        double gg, Qd[9], Xi[9], dXi[9], Q[9], t3[3], kr[3];
        s_rodrigues(kn.a, gg, Qd, Xi);
        d_xi(dXi, kn.a, kn.da, gg, Xi);
        s_mm(Q, Qd, Qi);
        s_mtv(t3, Xi, kn.da); s_mtv(kr, Qi, t3);
        for (int i = 0; i < 3; i++) kr[i] = s_add(kr[i], ki[i]);                  // kappa_r (:920)
        double kip[3], e3ip[3], tf[3], tm[3], c[3], Xtc[3];
        mv(kip, Q, kr); mv(e3ip, Q, go.e3r);
        cross3(tf, kip, e3ip);
        cross3(c, dz, e3ip); mtv(tm, Xi, c);
        for (int i = 0; i < 3; i++) { tf[i] *= -p0i * Aint; tm[i] *= -p0i * Aint; }
        cross3(c, e3ip, dz); mtv(Xtc, Xi, c);
        double O1[9], Sc[9], K1ua[9], K1aa[9], K2ua[9], K2au[9], E3[9], tmp[9], tmp2[9];
        skew3(Sc, c); skew3(E3, e3ip);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) O1[3 * i + j] = -0.5 * g_last * (Xtc[i] * kn.a[j] - Sc[3 * i + j]);
        // K1ua = Kip E3 Xi + E3 dXi - E3 Kip Xi
        mm(K2ua, E3, Xi);                                   // E3 Xi
        skew_mul(K1ua, kip, K2ua);                          // Kip E3 Xi
        mm(tmp, E3, dXi);
        skew_mul(tmp2, kip, Xi); mm(Sc, E3, tmp2);          // E3 Kip Xi
        for (int i = 0; i < 9; i++) K1ua[i] = K1ua[i] + tmp[i] - 1.0 * Sc[i];
        skew_mul(tmp, dz, K2ua); mtm(K1aa, Xi, tmp);        // Xi^T dZ E3 Xi
        for (int i = 0; i < 9; i++) K1aa[i] += O1[i];
        mtm(K2au, Xi, E3);
        // rotate the four blocks and the load to global axes: T^T (.) T with T = blockdiag(R)
        const double* R = go.R;
        mm(tmp, K1ua, R); mtm(K1ua, R, tmp);
        mm(tmp, K1aa, R); mtm(K1aa, R, tmp);
        mm(tmp, K2ua, R); mtm(K2ua, R, tmp);
        mm(tmp, K2au, R); mtm(K2au, R, tmp);
        double tfg[3], tmg[3];
        mtv(tfg, R, tf); mtv(tmg, R, tm);
        const double mult = 1.0 * go.jac, ks = mult * p0i * Aint;
        for (int a = 0; a < 3; a++) {
            const double Na = go.N[a];
            for (int i = 0; i < 3; i++) {
                out[324 + 6 * a + i] -= mult * Na * tfg[i];
                out[324 + 6 * a + 3 + i] -= mult * Na * tmg[i];
            }
            for (int b = 0; b < 3; b++) {
                const double Nb = go.N[b], dNb = go.dN[b];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) {
                        // N^T Kext UpsilonN: (u, alpha) = K1ua N_b + K2ua N'_b, (alpha, u) = K2au N'_b, (alpha, alpha) = K1aa N_b
                        out[(6 * a + i) * 18 + 6 * b + 3 + j] -= ks * Na * (K1ua[3 * i + j] * Nb + K2ua[3 * i + j] * dNb);
                        out[(6 * a + 3 + i) * 18 + 6 * b + j] -= ks * Na * (K2au[3 * i + j] * dNb);
                        out[(6 * a + 3 + i) * 18 + 6 * b + 3 + j] -= ks * Na * (K1aa[3 * i + j] * Nb);
                    }
            }
        }
    }
}

// Beam_1::SaveLagrange (:1494-1506)
__global__ void commit_kernel(EvalArgs A) {
    const size_t n_gp = (size_t)A.n_el * NGP;
    const size_t gp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= n_gp) return;
    const int e = (int)(gp / NGP), g = (int)(gp % NGP);
    int nd[3];
#pragma unroll
    for (int n = 0; n < 3; n++) nd[n] = __ldg(A.conn + 3 * (size_t)e + n);
    const double* pr = A.props + BEAM_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    Geo go; geometry(A, e, g, nd, pr, go);
    Kin kn; interpolate(A, nd, go, kn);
    double gg, Qd[9], Xi[9], Qi[9], Qn[9], t3[3], kr[3];
    s_rodrigues(kn.a, gg, Qd, Xi);
#pragma unroll
    for (int i = 0; i < 9; i++) Qi[i] = A.state[i * n_gp + gp];
    s_mtv(t3, Xi, kn.da); s_mtv(kr, Qi, t3);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        A.state[(12 + i) * n_gp + gp] = s_add(kr[i], A.state[(12 + i) * n_gp + gp]);   // kappa_i_ref = kappa_r
        A.state[(9 + i) * n_gp + gp] = s_add(kn.du[i], A.state[(9 + i) * n_gp + gp]);  // dz_i = d_z
    }
    s_mm(Qn, Qd, Qi);
#pragma unroll
    for (int i = 0; i < 9; i++) A.state[i * n_gp + gp] = Qn[i];
}


// Gauss-point results of Beam_1::Mount for WriteResults / WriteMonitor (Beam_1.cpp:444-497, :781-794,
// :830): epsilon_r = [eta_r, kappa_r], sigma_r = D epsilon_r per point and the element's strain energy.
// out[e * 25]: strain_energy, then per point g: epsilon_r(6) sigma_r(6).
__global__ void results_kernel(EvalArgs A, double* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n_el) return;
    int nd[3];
#pragma unroll
    for (int n = 0; n < 3; n++) nd[n] = __ldg(A.conn + 3 * (size_t)e + n);
    const double* pr = A.props + BEAM_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    const size_t n_gp = (size_t)A.n_el * NGP;
    double energy = 0.0;
    double* o = out + (size_t)e * 25;
#pragma unroll 1
    for (int g = 0; g < NGP; g++) {
        const size_t gp = (size_t)e * NGP + g;
        Geo go; geometry(A, e, g, nd, pr, go);
        Kin kn; interpolate(A, nd, go, kn);
        double Qi[9], dz[3], ki[3];
#pragma unroll
        for (int i = 0; i < 9; i++) Qi[i] = A.state[i * n_gp + gp];
#pragma unroll
        for (int i = 0; i < 3; i++) { dz[i] = s_add(kn.du[i], A.state[(9 + i) * n_gp + gp]); ki[i] = A.state[(12 + i) * n_gp + gp]; }
        double gg, Qd[9], Xi[9], Q[9], Qt[9];
        s_rodrigues(kn.a, gg, Qd, Xi);
        s_mm(Q, Qd, Qi);
        m_transpose(Qt, Q);
        double eps[6], sig[6], t3[3];
        s_mv(eps, Qt, dz);
#pragma unroll
        for (int i = 0; i < 3; i++) eps[i] = s_sub(eps[i], go.e3r[i]);
        s_mtv(t3, Xi, kn.da); s_mtv(eps + 3, Qi, t3);
#pragma unroll
        for (int i = 0; i < 3; i++) eps[3 + i] = s_add(eps[3 + i], ki[i]);
        double se = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double sacc = s_mul(__ldg(pr + 6 * i), eps[0]);
#pragma unroll
            for (int j = 1; j < 6; j++) sacc = s_add(sacc, s_mul(__ldg(pr + 6 * i + j), eps[j]));
            sig[i] = sacc;
            se += sacc * eps[i];
        }
        energy += __ldg(pr + 46) * (0.5 * (1.0 * go.jac) * se);   // (:830); Pipe_1::Mount has no such line
#pragma unroll
        for (int i = 0; i < 6; i++) { o[1 + 12 * g + i] = eps[i]; o[1 + 12 * g + 6 + i] = sig[i]; }
    }
    o[0] = energy;
}

} // namespace beam

// =========================================================================
// Solid_1 -- builder-defined (reference bodies are empty, Solid_1.cpp:148-176):
// 8-node trilinear hexahedron, total Lagrangian, St.Venant-Kirchhoff, 2x2x2 Gauss.
// In gradient space (i,J): A_JL[i][k] = lambda F_iJ F_kL + mu F_iL F_kJ
//                                      + delta_JL mu (F F^T)_ik + delta_ik S_JL
// =========================================================================
namespace solid {

constexpr int EPW = 4;
constexpr int NGP = 8;
constexpr int C_OFF = 0;       // 6 upper 3x3 blocks (J<=L) of A, times weight
constexpr int F_OFF = 54;      // first Piola-Kirchhoff P_iJ stored as f[(J,i)], times weight
constexpr int S_OFF = 63;      // dN_a/dX_J  as S[J*8 + a]
constexpr int N_OFF = 87;      // N_a (8)
constexpr int W_OFF = 95;      // |J| * rho
constexpr int REC = 97;
constexpr int SMEM_BYTES = EPW * NGP * REC * 8;

__host__ __device__ constexpr int blk(int p, int q) { return p * 3 - (p * (p - 1)) / 2 + (q - p); }

__device__ void physics(const EvalArgs& A, int e, int g, double* rec) {
    const double sgn[8][3] = { {-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1} };
    const double gpc = 0.57735026918962576451;
    const double xi = gpc * sgn[g][0], et = gpc * sgn[g][1], ze = gpc * sgn[g][2];
    const double* pr = A.props + SOLID_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    const double lam = __ldg(pr), mu = __ldg(pr + 1), rho = __ldg(pr + 2);
    // strict arithmetic up to the second Piola-Kirchhoff stress: E = (F^T F - I)/2
    // is a difference of O(1) quantities (see gfa_math.cuh)
    double J[9], F[9], dNl[8][3], Nn[8], un[8][3];
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const int nd = __ldg(A.conn + 8 * (size_t)e + a);
        const double sx = sgn[a][0], sy = sgn[a][1], sz = sgn[a][2];
        const double fx = s_add(1.0, s_mul(sx, xi)), fy = s_add(1.0, s_mul(sy, et)), fz = s_add(1.0, s_mul(sz, ze));
        Nn[a] = s_mul(s_mul(s_mul(0.125, fx), fy), fz);
        dNl[a][0] = s_mul(s_mul(s_mul(0.125, sx), fy), fz);
        dNl[a][1] = s_mul(s_mul(s_mul(0.125, sy), fx), fz);
        dNl[a][2] = s_mul(s_mul(s_mul(0.125, sz), fx), fy);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double X = __ldg(A.xyz + 3 * (size_t)nd + i);
            un[a][i] = s_add(s_sub(__ldg(A.copy + 6 * (size_t)nd + i), X), __ldg(A.disp + 6 * (size_t)nd + i));
#pragma unroll
            for (int j = 0; j < 3; j++) J[3 * i + j] = a == 0 ? s_mul(X, dNl[0][j]) : s_add(J[3 * i + j], s_mul(X, dNl[a][j]));
        }
    }
#define GFA_M2(a_, b_, c_, d_) s_sub(s_mul(J[a_], J[b_]), s_mul(J[c_], J[d_]))
    const double det = s_add(s_sub(s_mul(J[0], GFA_M2(4, 8, 5, 7)), s_mul(J[1], GFA_M2(3, 8, 5, 6))), s_mul(J[2], GFA_M2(3, 7, 4, 6)));
    double Ji[9];
    Ji[0] = GFA_M2(4, 8, 5, 7) / det; Ji[1] = GFA_M2(2, 7, 1, 8) / det; Ji[2] = GFA_M2(1, 5, 2, 4) / det;
    Ji[3] = GFA_M2(5, 6, 3, 8) / det; Ji[4] = GFA_M2(0, 8, 2, 6) / det; Ji[5] = GFA_M2(2, 3, 0, 5) / det;
    Ji[6] = GFA_M2(3, 7, 4, 6) / det; Ji[7] = GFA_M2(1, 6, 0, 7) / det; Ji[8] = GFA_M2(0, 4, 1, 3) / det;
#undef GFA_M2
    F[0] = 1; F[1] = 0; F[2] = 0; F[3] = 0; F[4] = 1; F[5] = 0; F[6] = 0; F[7] = 0; F[8] = 1;
#pragma unroll
    for (int a = 0; a < 8; a++) {
        double dN[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double v = s_add(s_add(s_mul(dNl[a][0], Ji[j]), s_mul(dNl[a][1], Ji[3 + j])), s_mul(dNl[a][2], Ji[6 + j]));
            dN[j] = v;
            rec[S_OFF + 8 * j + a] = v;
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) F[3 * i + j] = s_add(F[3 * i + j], s_mul(un[a][i], dN[j]));
        rec[N_OFF + a] = Nn[a];
    }
    double Cg[9], Bm[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            Cg[3 * i + j] = s_add(s_add(s_mul(F[i], F[j]), s_mul(F[3 + i], F[3 + j])), s_mul(F[6 + i], F[6 + j]));
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) Bm[3 * i + k] = F[3 * i] * F[3 * k] + F[3 * i + 1] * F[3 * k + 1] + F[3 * i + 2] * F[3 * k + 2];
    const double E0 = s_mul(0.5, s_sub(Cg[0], 1.0)), E1 = s_mul(0.5, s_sub(Cg[4], 1.0)), E2 = s_mul(0.5, s_sub(Cg[8], 1.0));
    const double trE = s_add(s_add(E0, E1), E2);
    const double mu2 = s_mul(2.0, mu), ltr = s_mul(lam, trE);
    double S[9];
    S[0] = s_add(ltr, s_mul(mu2, E0)); S[4] = s_add(ltr, s_mul(mu2, E1)); S[8] = s_add(ltr, s_mul(mu2, E2));
    S[1] = S[3] = s_mul(mu, Cg[1]); S[5] = S[7] = s_mul(mu, Cg[5]); S[2] = S[6] = s_mul(mu, Cg[2]);
    double P[9];
    mm(P, F, S);
    const double w = det;      // Gauss weights are 1
#pragma unroll
    for (int Jd = 0; Jd < 3; Jd++)
#pragma unroll
        for (int i = 0; i < 3; i++) rec[F_OFF + 3 * Jd + i] = w * P[3 * i + Jd];
#pragma unroll
    for (int Jd = 0; Jd < 3; Jd++)
#pragma unroll
        for (int L = Jd; L < 3; L++)
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    double v = lam * F[3 * i + Jd] * F[3 * k + L] + mu * F[3 * i + L] * F[3 * k + Jd];
                    if (Jd == L) v += mu * Bm[3 * i + k];
                    if (i == k) v += S[3 * Jd + L];
                    rec[C_OFF + 9 * blk(Jd, L) + 3 * i + k] = w * v;
                }
    rec[W_OFF] = det * rho;
}

template <int P, int II, int Q>
GFA_DI double c_at(const double* recJ, const double* rec3J) {
    if (P <= Q) return recJ[C_OFF + 9 * blk(P, Q) + 3 * II];
    else return rec3J[C_OFF + 9 * blk(Q, P) + II];
}

// Columns b1 = p and b2 = 7 - p, component jj, in one item: the C' entries of a column depend on jj only and the
// gradient table on nothing but the point, so both are read once for the two columns -- the kernel's busiest unit
// was the L1 data pipe (77 %, 354 shared-memory loads per element; ncu, profiles/r02_notes.md), not the FP64 pipe.
__device__ void congruence_pair(const EvalArgs& A, int e, double* ke, const double* rec0, int p, int jj) {
    const int b1 = p, b2 = 7 - p;
    double K1[24], K2[24];
#pragma unroll
    for (int i = 0; i < 24; i++) { K1[i] = 0.0; K2[i] = 0.0; }
    double F1 = 0.0, F2 = 0.0, fe1 = 0.0, fe2 = 0.0;
    const double gk = jj == 0 ? A.gx : jj == 1 ? A.gy : A.gz;
#pragma unroll 1
    for (int g = 0; g < NGP; g++) {
        const double* rec = rec0 + g * REC;
        const double* recJ = rec + jj;
        const double* rec3J = rec + 3 * jj;
        const double* S = rec + S_OFF;
        const double s0 = S[b1], s1 = S[8 + b1], s2 = S[16 + b1];
        const double t0 = S[b2], t1 = S[8 + b2], t2 = S[16 + b2];
        double m[3][3], n[3][3];
#define GFA_ROW(P_, I_) { const double c0 = c_at<P_, I_, 0>(recJ, rec3J), c1 = c_at<P_, I_, 1>(recJ, rec3J), c2 = c_at<P_, I_, 2>(recJ, rec3J); \
                          m[P_][I_] = s0 * c0 + s1 * c1 + s2 * c2; n[P_][I_] = t0 * c0 + t1 * c1 + t2 * c2; }
        GFA_ROW(0, 0) GFA_ROW(0, 1) GFA_ROW(0, 2) GFA_ROW(1, 0) GFA_ROW(1, 1) GFA_ROW(1, 2) GFA_ROW(2, 0) GFA_ROW(2, 1) GFA_ROW(2, 2)
#undef GFA_ROW
        {
            const double f0 = recJ[F_OFF], f1 = recJ[F_OFF + 3], f2 = recJ[F_OFF + 6];
            F1 += s0 * f0 + s1 * f1 + s2 * f2;
            F2 += t0 * f0 + t1 * f1 + t2 * f2;
        }
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const double a0 = S[a], a1 = S[8 + a], a2 = S[16 + a];
#pragma unroll
            for (int ii = 0; ii < 3; ii++) {
                K1[3 * a + ii] += a0 * m[0][ii] + a1 * m[1][ii] + a2 * m[2][ii];
                K2[3 * a + ii] += a0 * n[0][ii] + a1 * n[1][ii] + a2 * n[2][ii];
            }
        }
        const double wg = rec[W_OFF];
        fe1 += wg * rec[N_OFF + b1] * gk;
        fe2 += wg * rec[N_OFF + b2] * gk;
    }
    // stored blocks of a column block: rows 0..b (the tangent is symmetric)
    double* Ke1 = ke + solid_stored_offset(0, b1) + jj;
    double* Ke2 = ke + solid_stored_offset(0, b2) + jj;
#pragma unroll
    for (int rb = 0; rb < 8; rb++) {
        if (rb <= b1) { Ke1[9 * rb] = K1[3 * rb]; Ke1[9 * rb + 3] = K1[3 * rb + 1]; Ke1[9 * rb + 6] = K1[3 * rb + 2]; }
        if (rb <= b2) { Ke2[9 * rb] = K2[3 * rb]; Ke2[9 * rb + 3] = K2[3 * rb + 1]; Ke2[9 * rb + 6] = K2[3 * rb + 2]; }
    }
    A.Pe[(size_t)e * 24 + 3 * b1 + jj] = F1 - fe1;
    A.Pe[(size_t)e * 24 + 3 * b2 + jj] = F2 - fe2;
}

// one batch of `ne` <= EPW solids at list positions k0 .. (see shell::eval_batch): 4 elements = 32 Gauss points, then 48
// column-pair items (1.5 rounds of the warp).  Measured on 4M solids: 13.4 ms with one column per item, 12.9 ms with
// the pair items; 8 elements a batch (three full rounds, but 4 warps per SM) 14.5 ms; halving the items over two
// lanes by Gauss points (three full rounds, one exchange) 14.8 ms -- the shuffles land on the same L1 data pipe
__device__ __forceinline__ void eval_batch(const EvalArgs& A, int k0, int ne, double* smem, int lane) {
    for (int gp = lane; gp < ne * NGP; gp += 32) physics(A, eval_element(A, k0 + gp / NGP), gp % NGP, smem + gp * REC);
    __syncwarp();
    for (int it = lane; it < ne * 12; it += 32) {
        const int el = it / 12, c = it % 12;
        congruence_pair(A, eval_element(A, k0 + el), eval_ke(A, k0 + el, SOLID_ARENA), smem + el * NGP * REC, c / 3, c % 3);
    }
    __syncwarp();
}
__global__ void __launch_bounds__(32) eval_kernel(EvalArgs A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    for (long long batch = blockIdx.x; A.e_begin + batch * EPW < A.e_end; batch += gridDim.x) {
        const int k0 = A.e_begin + (int)(batch * EPW);
        eval_batch(A, k0, min(EPW, A.e_end - k0), smem, lane);
    }
}

} // namespace solid

// =========================================================================
// Node::SaveConfiguration (Node.cpp:325-349): copy += disp, Rodrigues
// composition of rotations; then the increments are zeroed (Static.cpp:191).
// =========================================================================
__global__ void node_commit_kernel(int n_nodes, double* copy, double* disp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    double* c = copy + 6 * (size_t)i;
    double* d = disp + 6 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 3; k++) c[k] += d[k];
    const double a1[3] = { c[3], c[4], c[5] }, a2[3] = { d[3], d[4], d[5] };
    double cr[3];
    cross3(cr, a2, a1);
    const double s = 4.0 / (4.0 - dot3(a2, a1));
#pragma unroll
    for (int k = 0; k < 3; k++) c[3 + k] = s * (a2[k] + a1[k] + 0.5 * cr[k]);
#pragma unroll
    for (int k = 0; k < 6; k++) d[k] = 0.0;
}

// =========================================================================
// Scatter: MountGlobal + MountSparse for the free x free matrix and the
// residual vectors.
// =========================================================================
constexpr int SCATTER_THREADS = 128;     // measured on the 1M-shell plate: 64 / 96 / 128 / 192 / 256 / 512 threads = 1.80 / 1.79 / 1.77 / 1.83 / 1.83 / 2.04 ms (scatter + vectors)

// column c of the 3x3 destination patch: D(i,c) = S(i,c), or S(c,i) when the stored block is the transposed twin
GFA_DI void load_col(const double* Ke, unsigned off, bool tr, int c, double (&x)[3]) {
    const double* p = Ke + (size_t)off + (tr ? 3 * c : c);
    const int st = tr ? 1 : 3;
    x[0] = p[0]; x[1] = p[st]; x[2] = p[2 * st];
}

// One THREAD per CSR patch column: thread t owns column c = t % 3 of patch t / 3 (= one (group-node,
// neighbour) pair).  It gathers that column of the contributing 3x3 blocks straight from the arena --
// all loads of the first two issued before the first add, element-ascending adds in registers (the
// order the reference pushes and Eigen sums, Solution.cpp:327-328) -- and writes one value into each of
// the patch's <= 3 CSR rows.  The three threads of a patch read one contiguous 72-byte block together
// (24 contiguous bytes per load when the block is stored as is) and write 24 contiguous bytes of one
// CSR row per store; consecutive patches of a group-node are consecutive in its rows, so every CSR row
// is written once, in whole sectors.  No shared memory, no atomics, no synchronisation.
__global__ void __launch_bounds__(SCATTER_THREADS) scatter_kernel(ScatterArgs A) {
    const long long t = (long long)blockIdx.x * SCATTER_THREADS + threadIdx.x;
    const long long j = t / 3;
    if (j >= A.n_runs) return;
    const int c = (int)(t - 3 * j);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(A.runs) + j);
    const unsigned info = q.y;
    const int L = info & 0xffff, rm = (info >> 16) & 7, fm = (info >> 19) & 7, cnt = info >> 24;
    if (!((fm >> c) & 1)) return;
    double a[3];
    {
        unsigned s0 = q.z, s1 = q.w;
        bool t0 = (info >> 22) & 1, t1 = (info >> 23) & 1;
        if (cnt > 2) {
            const unsigned long long e0 = __ldg(A.ovf + q.z), e1 = __ldg(A.ovf + q.z + 1);
            s0 = (unsigned)e0; t0 = (e0 & SRC_T) != 0; s1 = (unsigned)e1; t1 = (e1 & SRC_T) != 0;
        }
        double x[3] = { 0.0, 0.0, 0.0 }, y[3] = { 0.0, 0.0, 0.0 };
        if (cnt > 0) load_col(A.Ke, s0, t0, c, x);           // count 0: a patch fed by other ranks only
        if (cnt > 1) load_col(A.Ke, s1, t1, c, y);
#pragma unroll
        for (int i = 0; i < 3; i++) a[i] = x[i] + y[i];
    }
    for (int k = 2; k < cnt; k++) {
        double x[3];
        const unsigned long long e = __ldg(A.ovf + q.z + k);
        load_col(A.Ke, (unsigned)e, (e & SRC_T) != 0, c, x);
#pragma unroll
        for (int i = 0; i < 3; i++) a[i] += x[i];
    }
    // rows = the group's free DOFs (consecutive CSR rows of equal length L), columns = the neighbour's free DOFs
    double* o = A.valAA + (size_t)(int)q.x + __popc(fm & ((1 << c) - 1));
    const int r1 = rm & 1, r2 = r1 + ((rm >> 1) & 1);
    if (rm & 1) o[0] = a[0];
    if (rm & 2) o[(size_t)r1 * L] = a[1];
    if (rm & 4) o[(size_t)r2 * L] = a[2];
}

// residual vectors: global_P_A / global_I_A (free) or global_P_B (fixed), element-ascending sums;
// one thread per (group-node, component)
__global__ void vector_kernel(ScatterArgs A) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * A.n_gn) return;
    const GnRec* gp = A.gn + t / 3;
    const int k = (int)(t % 3);
    const int gl = gp->gl[k];
    if (gl == 0) return;
    double s = 0.0;
    for (int i = gp->ib; i < gp->ie; i++) {
        const PInc in = A.inc[i];
        s += A.Pe[in.pe_off + 3 * in.la + k];
    }
    if (gl > 0) { A.PA[gl - 1] = s; A.IA[gl - 1] = s; }
    else A.PB[-gl - 1] = s;
}

// AB / BA / BB entries: explicit element-ascending gather lists, one thread per slot.
__global__ void gather_kernel(GatherArgs A) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_dest) return;
    double s = 0.0;
    for (long long k = A.seg[i]; k < A.seg[i + 1]; k++) s += A.Ke[A.src[k]];
    A.vals[A.dest[i]] = s;
}

__global__ void gather_add_kernel(double* vals, const long long* seg, const long long* src, const long long* dest, const double* from, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (long long k = seg[i]; k < seg[i + 1]; k++) s += from[src[k]];
    vals[dest[i]] += s;
}
__global__ void add_slots_kernel(double* vals, const long long* slots, const double* add, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vals[slots[i]] += add[i];
}
__global__ void pack_kernel(const double* vals, const long long* idx, double* buf, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = vals[idx[i]];
}
__global__ void unpack_nodes_kernel(double* disp, const int* nodes, const double* packed, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 6 * n) disp[6 * (size_t)nodes[i / 6] + i % 6] = packed[i];
}
__global__ void unpack_add_kernel(double* vals, const long long* idx, const double* buf, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vals[idx[i]] += buf[i];
}

// =========================================================================
// Fused ring pipeline: evaluation + scatter in one persistent kernel (FusedArgs, gfa_device.h)
// =========================================================================
namespace fused {

GFA_DI unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
GFA_DI unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
GFA_DI void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
GFA_DI unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// column c of a contributing block, read from L2 (the ring is rewritten while the kernel runs: L1 must not serve it)
GFA_DI void load_col_cg(const double* Ke, unsigned off, bool tr, int c, double (&x)[3]) {
    const double* p = Ke + (size_t)off + (tr ? 3 * c : c);
    const int st = tr ? 1 : 3;
    x[0] = __ldcg(p); x[1] = __ldcg(p + st); x[2] = __ldcg(p + 2 * st);
}

struct ShellT {
    static constexpr int EPW = 8, SMEM = shell::smem_bytes(8);
    static GFA_DI void batch(const EvalArgs& A, int k0, int ne, double* sm, int lane) { shell::eval_batch<8>(A, k0, ne, sm, lane); }
};
struct BeamT {
    static constexpr int EPW = beam::EPW, SMEM = beam::SMEM_BYTES;
    static GFA_DI void batch(const EvalArgs& A, int k0, int ne, double* sm, int lane) { beam::eval_batch(A, k0, ne, sm, lane); }
};
struct SolidT {
    static constexpr int EPW = solid::EPW, SMEM = solid::SMEM_BYTES;
    static GFA_DI void batch(const EvalArgs& A, int k0, int ne, double* sm, int lane) { solid::eval_batch(A, k0, ne, sm, lane); }
};

// a wait that gives up: returns true when the kernel has to leave (abort raised by anyone, or this warp waited too long)
GFA_DI bool backoff(unsigned* ctl, unsigned long long& t_wait, unsigned long long timeout_ns, unsigned who) {
    if (ld_relaxed(ctl + CTL_ABORT)) return true;
    const unsigned long long now = global_ns();
    if (!t_wait) t_wait = now;
    else if (now > t_wait && now - t_wait > timeout_ns) {
        if (atomicExch(ctl + CTL_ABORT, who) == 0u) ctl[5] = (unsigned)((now - t_wait) >> 10);     // who gave up first, after how many microseconds
        return true;
    }
    __nanosleep(128);
    return false;
}

// ---- evaluation kernel: one CTA per SM, every warp owns a record buffer and evaluates batches claimed in order ----
// Register budget: an SM's register file is four banks of 16 K, one per scheduler, and a CTA's warps are dealt
// round-robin to them.  Seven evaluation warps put two on three of the schedulers; the scatter CTA's eight warps
// need two slots of 32 registers on EVERY scheduler, so an evaluation thread may hold at most 256 - 32 = 224
// (tools/coresidency_probe.cu: at 254 registers the thin kernel never becomes resident beside five or more warps).
template <class T>
__global__ void __maxnreg__(FUSED_EVAL_REGS) eval_kernel(FusedArgs F) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* buf = smem + (size_t)warp * (T::SMEM / 8);
    unsigned* ctl = F.ctl;
    unsigned* edone = ctl + CTL_HDR;                         // batches evaluated, per chunk
    const unsigned* sdone = edone + F.total_chunks;          // tiles scattered, per ready chunk
    const EvalArgs& A = F.ev;
    const int n_list = A.e_end;
    const int n_batches = (n_list + T::EPW - 1) / T::EPW;
    // lane 0: chunks known to be scattered completely.  The scatter kernel runs on its own: it may still be draining the
    // last chunks of the previous element type's launch, but everything a full ring before this launch's first chunk
    // had to be scattered for that launch to get as far as it did
    int s_upto = max(0, A.chunk0 - A.ring_chunks);
    unsigned long long t_wait = 0;
    for (;;) {
        int b = 0;
        if (lane == 0) b = (int)atomicAdd(ctl + CTL_BATCH + F.type_slot, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= n_batches) break;
        const int k0 = b * T::EPW;
        const int c = A.chunk0 + k0 / A.chunk_el;
        // the ring slot of chunk c held chunk c - ring_chunks: everything that read it must have been scattered
        const int need = c - A.ring_chunks + F.span + 1;
        int stop = 0;
        if (lane == 0) {
            t_wait = 0;
            while (s_upto < need) {
                const unsigned have = ld_acquire(sdone + s_upto), want = (unsigned)(F.chunk_tile_ptr[s_upto + 1] - F.chunk_tile_ptr[s_upto]);
                if (have == want) { s_upto++; continue; }
                if (backoff(ctl, t_wait, F.timeout_ns, 1u + (unsigned)c * 16u)) {
                    unsigned* dbg = ctl + CTL_HDR + 3 * F.total_chunks;
                    if (atomicExch(dbg, 1u) == 0u) { dbg[1] = have; dbg[2] = want; dbg[3] = (unsigned)s_upto; dbg[4] = (unsigned)need; dbg[5] = (unsigned)b; dbg[6] = (unsigned)A.chunk0; dbg[7] = (unsigned)A.ring_chunks; }
                    stop = 1; break;
                }
            }
        }
        if (__shfl_sync(0xffffffffu, stop, 0)) break;
        T::batch(A, k0, min(T::EPW, n_list - k0), buf, lane);
        __syncwarp();               // the warp's arena stores are ordered before lane 0's release (cumulativity)
        if (lane == 0) { fence_acq_rel(); atomicAdd(edone + c, 1u); }
    }
}

// ---- scatter kernel: thin persistent warps beside the evaluation CTA of every SM ------------------------------
// Registers are what the evaluation kernel leaves least of (32 a thread here), and a gather that has to cover
// ~1 us of L2 latency with loads held in registers needs thousands of them per SM.  So the blocks do not pass
// through registers: a warp claims tiles of FUSED_TILE_PATCHES patches of one ready chunk (atomicAdd on that
// chunk's cursor, once the chunk and all earlier ones are evaluated), and for every tile
//   1. issues, for both possible sources of each patch, the five aligned 16-byte pieces that cover the 72-byte
//      block as cp.async.cg copies into its own staging buffer (L2 -> shared memory, no register, no L1: the ring
//      is rewritten while the kernel runs);
//   2. waits for the group, then sums each patch column from shared memory in element-ascending order and writes
//      the CSR rows with streaming stores -- the arithmetic of scatter_kernel, bit for bit.
// Patches fed by more than two blocks take their sources straight from L2 (overflow list).
constexpr int STAGE_BLOCK = 80;                                          // bytes staged per source block
constexpr int STAGE_BYTES = FUSED_TILE_PATCHES * 2 * STAGE_BLOCK;       // per warp
GFA_DI void cp_async16(unsigned dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst_smem), "l"(src) : "memory");
}
__global__ void __maxnreg__(32) scatter_kernel(FusedArgs F) {
    extern __shared__ __align__(16) unsigned char stage_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = stage_all + warp * STAGE_BYTES;
    const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
    unsigned* ctl = F.ctl;
    const unsigned* edone = ctl + CTL_HDR;
    unsigned* sdone = ctl + CTL_HDR + F.total_chunks;
    unsigned* tnext = sdone + F.total_chunks;
    const ScatterArgs& A = F.sc;
    const int chunk_end = F.total_chunks;
    int e_upto = 0, r_cur = 0;      // lane 0: evaluated prefix, first chunk that may hold unclaimed tiles
    unsigned long long t_wait = 0;
    for (;;) {
        int t0 = -1, nt = 0;
        if (lane == 0) {
            for (;;) {
                if (r_cur >= chunk_end) break;
                const int tiles = F.chunk_tile_ptr[r_cur + 1] - F.chunk_tile_ptr[r_cur];
                if (tiles == 0 || (int)ld_relaxed(tnext + r_cur) >= tiles) { r_cur++; continue; }
                while (e_upto <= r_cur && ld_acquire(edone + e_upto) == (unsigned)F.chunk_batches[e_upto]) e_upto++;
                if (e_upto <= r_cur) {                                    // not evaluated yet
                    if (backoff(ctl, t_wait, F.timeout_ns, 2u + (unsigned)r_cur * 16u)) break;
                    continue;
                }
                const int t = (int)atomicAdd(tnext + r_cur, (unsigned)F.tile_group);
                if (t >= tiles) { r_cur++; continue; }
                t0 = t; nt = min(F.tile_group, tiles - t); t_wait = 0;
                break;
            }
        }
        t0 = __shfl_sync(0xffffffffu, t0, 0);
        if (t0 < 0) break;                                                // every tile is claimed, or the watchdog fired
        nt = __shfl_sync(0xffffffffu, nt, 0);
        const int r = __shfl_sync(0xffffffffu, r_cur, 0);
        const long long run0 = F.chunk_run_ptr[r] + (long long)t0 * FUSED_TILE_PATCHES;
        const long long run1 = min(F.chunk_run_ptr[r + 1], run0 + (long long)nt * FUSED_TILE_PATCHES);
        // the slot-map entries of the claim come from DRAM: start them all now (128-byte lines)
        for (long long j = run0 + 8 * lane; j < run1; j += 256) asm volatile("prefetch.global.L2 [%0];" :: "l"(A.runs + j));
        uint4 qn = run0 + lane < run1 ? __ldcs(reinterpret_cast<const uint4*>(A.runs) + run0 + lane) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
        for (long long p0 = run0; p0 < run1; p0 += FUSED_TILE_PATCHES) {
            const uint4 q = qn;             // this lane's patch of the tile (info == 0 past the end)
            qn = p0 + FUSED_TILE_PATCHES + lane < run1 ? __ldcs(reinterpret_cast<const uint4*>(A.runs) + p0 + FUSED_TILE_PATCHES + lane) : make_uint4(0u, 0u, 0u, 0u);
            const unsigned info = q.y;
            const int cnt = info >> 24;
            // ---- 1. stage this lane's (at most two) source blocks: the five aligned 16-byte pieces around each
            if (cnt >= 1 && cnt <= 2) {
                const unsigned long long g = reinterpret_cast<unsigned long long>(A.Ke + q.z) & ~15ULL;
                const unsigned d = stage_s + lane * STAGE_BLOCK;
#pragma unroll
                for (int k = 0; k < 5; k++) cp_async16(d + 16 * k, reinterpret_cast<const void*>(g + 16 * k));
            }
            if (cnt == 2) {
                const unsigned long long g = reinterpret_cast<unsigned long long>(A.Ke + q.w) & ~15ULL;
                const unsigned d = stage_s + (32 + lane) * STAGE_BLOCK;
#pragma unroll
                for (int k = 0; k < 5; k++) cp_async16(d + 16 * k, reinterpret_cast<const void*>(g + 16 * k));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            // ---- 2. sum and store the patch's columns (a lane reads only what it staged itself)
            const int L = info & 0xffff, rm = (info >> 16) & 7, fm = (info >> 19) & 7;
            const int r1 = rm & 1, r2 = r1 + ((rm >> 1) & 1);
            const double* b0 = reinterpret_cast<const double*>(stage + lane * STAGE_BLOCK) + (q.z & 1u);
            const double* b1 = reinterpret_cast<const double*>(stage + (32 + lane) * STAGE_BLOCK) + (q.w & 1u);
            const bool tr0 = (info >> 22) & 1, tr1 = (info >> 23) & 1;
            double* o = A.valAA + (size_t)(int)q.x;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (!((fm >> c) & 1)) continue;
                double a[3];
                if (cnt <= 2) {
                    const int o0 = tr0 ? 3 * c : c, st0 = tr0 ? 1 : 3, o1 = tr1 ? 3 * c : c, st1 = tr1 ? 1 : 3;
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        const double x = cnt > 0 ? b0[o0 + i * st0] : 0.0, y = cnt > 1 ? b1[o1 + i * st1] : 0.0;
                        a[i] = x + y;
                    }
                } else {
                    const unsigned long long e0 = __ldg(A.ovf + q.z), e1 = __ldg(A.ovf + q.z + 1);
                    double v0[3], v1[3];
                    load_col_cg(A.Ke, (unsigned)e0, (e0 & SRC_T) != 0, c, v0);
                    load_col_cg(A.Ke, (unsigned)e1, (e1 & SRC_T) != 0, c, v1);
#pragma unroll
                    for (int i = 0; i < 3; i++) a[i] = v0[i] + v1[i];
#pragma unroll 1
                    for (int k = 2; k < cnt; k++) {
                        const unsigned long long e = __ldg(A.ovf + q.z + k);
                        load_col_cg(A.Ke, (unsigned)e, (e & SRC_T) != 0, c, v0);
#pragma unroll
                        for (int i = 0; i < 3; i++) a[i] += v0[i];
                    }
                }
                if (rm & 1) __stcs(o, a[0]);
                if (rm & 2) __stcs(o + (size_t)r1 * L, a[1]);
                if (rm & 4) __stcs(o + (size_t)r2 * L, a[2]);
                o++;
            }
        }
        __syncwarp();
        if (lane == 0) { fence_acq_rel(); atomicAdd(sdone + r, (unsigned)nt); }
    }
}

} // namespace fused

// =========================================================================
// Newton-loop vector steps either side of the assembly
// =========================================================================
// db.global_P_A = -1.0 * db.global_P_A (Static.cpp:210)
__global__ void negate_kernel(double* v, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = __dmul_rn(-1.0, v[i]);
}
// db.global_P_A = db.global_P_A - 1.0 * (db.global_stiffness_AB * db.global_X_B) (Static.cpp:216-217); the
// product is the reference's own row loop, y_i += a * x in column order (SparseMatrix.cpp:186-190).
// One thread per non-empty row of AB.
__global__ void sub_ab_xb_kernel(double* PA, const int* rows, const int* ptr, const int* inner, const double* vals, const double* XB, int n_rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_rows) return;
    double y = 0.0;
    for (int p = ptr[k]; p < ptr[k + 1]; p++) y = __dadd_rn(y, __dmul_rn(vals[p], XB[inner[p]]));
    const int r = rows[k];
    PA[r] = __dsub_rn(PA[r], __dmul_rn(1.0, y));
}
// Solution::UpdateDisps (Solution.cpp:390-402): displacements[j] += x(GL-1) for free active DOFs
__global__ void update_disps_kernel(const int* gls, double* disp, const double* x, int n_nodes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6LL * n_nodes) return;
    const int g = gls[i];
    if (g > 0) disp[i] = __dadd_rn(disp[i], x[g - 1]);
}
// The max-norms of EstablishResidualCriteria / CheckResidualConvergence / CheckGLConvergence
// (ConvergenceCriteria.cpp:200-217, 474-505, 305-340): pass 1 the maxima, pass 2 the first node that
// reaches them (the reference keeps the first, `value > max`).
__global__ void norms_max_kernel(const int* gls, const double* v, const double* disp, int n_nodes, NormAcc* acc) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mt = 0, mr = 0, dt = 0, dr = 0;
    int nan = 0;
    if (node < n_nodes) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int g = gls[6 * (size_t)node + k];
            if (g <= 0) continue;
            const double a = fabs(v[g - 1]);
            if (a != a) { nan = 1; continue; }
            const unsigned long long b = (unsigned long long)__double_as_longlong(a);
            if (k < 3) mt = max(mt, b); else mr = max(mr, b);
            if (disp) {
                const double d = fabs(disp[6 * (size_t)node + k]);
                if (d == d) { const unsigned long long db = (unsigned long long)__double_as_longlong(d); if (k < 3) dt = max(dt, db); else dr = max(dr, db); }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mt = max(mt, __shfl_xor_sync(0xffffffffu, mt, o)); mr = max(mr, __shfl_xor_sync(0xffffffffu, mr, o));
        dt = max(dt, __shfl_xor_sync(0xffffffffu, dt, o)); dr = max(dr, __shfl_xor_sync(0xffffffffu, dr, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mt) atomicMax(&acc->max_t, mt);
        if (mr) atomicMax(&acc->max_r, mr);
        if (dt) atomicMax(&acc->max_dt, dt);
        if (dr) atomicMax(&acc->max_dr, dr);
        if (nan) atomicOr(&acc->nan, 1);
    }
}
__global__ void norms_node_kernel(const int* gls, const double* v, int n_nodes, NormAcc* acc) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_nodes) return;
    const unsigned long long mt = acc->max_t, mr = acc->max_r;
    bool ht = false, hr = false;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int g = gls[6 * (size_t)node + k];
        if (g <= 0) continue;
        const double a = fabs(v[g - 1]);
        if (a != a) continue;
        const unsigned long long b = (unsigned long long)__double_as_longlong(a);
        if (k < 3) ht |= (b == mt && mt != 0); else hr |= (b == mr && mr != 0);
    }
    if (ht) atomicMin(&acc->node_t, node);
    if (hr) atomicMin(&acc->node_r, node);
}

// =========================================================================
// launchers
// =========================================================================
static int grid_for(long long items, int per_block, int cap) {
    long long g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}
static const int kSMs = 148;

int configure_kernels() {
    cudaError_t e;
    e = cudaFuncSetAttribute(shell::eval_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, shell::smem_bytes(10));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(shell::eval_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, shell::smem_bytes(8));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(shell::eval_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shell::smem_bytes(8));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(shell::eval_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, shell::smem_bytes(6));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(shell::eval_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, shell::smem_bytes(7));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(shell::eval_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, shell::smem_bytes(9));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(beam::eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, beam::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(solid::eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, solid::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fused::eval_kernel<fused::ShellT>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_buffers(0) * fused::ShellT::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fused::eval_kernel<fused::BeamT>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_buffers(1) * fused::BeamT::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fused::eval_kernel<fused::SolidT>, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_buffers(2) * fused::SolidT::SMEM);
    if (e != cudaSuccess) return (int)e;
    // the scatter kernel has to be resident BESIDE the evaluation CTA of every SM: an SM runs with one shared-memory
    // carve-out at a time, so the kernel without shared memory asks for the same (maximal) split
    e = cudaFuncSetAttribute(fused::scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SCATTER_WARPS * fused::STAGE_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fused::scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fused::eval_kernel<fused::ShellT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fused::eval_kernel<fused::BeamT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(fused::eval_kernel<fused::SolidT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    return (int)e;
}

// Elements per warp batch: 10 fills 30 of 32 lanes in the Gauss-point phase but
// leaves 5 resident warps per SM (shared memory); smaller batches trade lane
// use for occupancy.  GFA_SHELL_EPW overrides the default for experiments.
static int shell_epw() {
    static int v = 0;
    if (!v) {
        const char* s = getenv("GFA_SHELL_EPW");
        v = s ? atoi(s) : 8;             // measured best on B200 (profiles/r01_notes.md)
        if (v < 6 || v > 10) v = 8;
    }
    return v;
}
// Persistent grids are sized by what is resident on the CURRENT device (occupancy query x SM count), cached per
// device: handles on different GPUs of one process each get their own figure.
constexpr int kMaxDevices = 64;
static int current_device() { int d = 0; cudaGetDevice(&d); return d < 0 || d >= kMaxDevices ? 0 : d; }
static int sm_count() {
    static int sms[kMaxDevices] = { 0 };
    const int d = current_device();
    if (!sms[d] && (cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d) != cudaSuccess || sms[d] < 1)) sms[d] = kSMs;
    return sms[d];
}
// resident one-warp CTAs on the whole device for a kernel (persistent grid)
template <class K>
static int resident_ctas(K kernel, int smem) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 32, smem) != cudaSuccess || n < 1) n = 8;
    return sm_count() * n;
}
// the batch layout of the shell arena belongs to the 8-element kernel (GFA_SHELL_EPW experiments keep the compact one)
bool shell_batch_layout_available() { return shell_epw() == SHELL_BATCH; }
void launch_shell_eval(const EvalArgs& a, void* s) {
    if (a.e_end <= a.e_begin) return;
    const int epw = shell_epw();
    // persistent warps: exactly as many one-warp CTAs as are resident (7 per SM with 8-element
    // batches), each striding over the batches -- 5 % faster than 64 CTAs per SM taking turns
    // (profiles/r01_notes.md); GFA_SHELL_GRID overrides the CTAs per SM
    static int caps[kMaxDevices] = { 0 };
    int& cap = caps[current_device()];
    if (!cap) {
        const char* e = getenv("GFA_SHELL_GRID");
        const int forced = e ? atoi(e) : 0;
        if (forced >= 1) cap = sm_count() * forced;
        else switch (epw) {
            case 7: cap = resident_ctas(shell::eval_kernel<7>, shell::smem_bytes(7)); break;
            case 9: cap = resident_ctas(shell::eval_kernel<9>, shell::smem_bytes(9)); break;
            case 6: cap = resident_ctas(shell::eval_kernel<6>, shell::smem_bytes(6)); break;
            case 8: cap = resident_ctas(shell::eval_kernel<8>, shell::smem_bytes(8)); break;
            default: cap = resident_ctas(shell::eval_kernel<10>, shell::smem_bytes(10)); break;
        }
    }
    const int grid = grid_for(a.e_end - a.e_begin, epw, cap);
    cudaStream_t st = (cudaStream_t)s;
    switch (epw) {
    case 7: shell::eval_kernel<7><<<grid, 32, shell::smem_bytes(7), st>>>(a); break;
    case 9: shell::eval_kernel<9><<<grid, 32, shell::smem_bytes(9), st>>>(a); break;
    case 6: shell::eval_kernel<6><<<grid, 32, shell::smem_bytes(6), st>>>(a); break;
    case 8:
        if (a.batch_layout && !a.elist && a.ring_chunks == 0 && a.e_begin % SHELL_BATCH == 0) shell::eval_kernel<8, true><<<grid, 32, shell::smem_bytes(8), st>>>(a);
        else shell::eval_kernel<8><<<grid, 32, shell::smem_bytes(8), st>>>(a);
        break;
    default: shell::eval_kernel<10><<<grid, 32, shell::smem_bytes(10), st>>>(a); break;
    }
}
void launch_beam_eval(const EvalArgs& a, void* s) {
    if (a.e_end <= a.e_begin) return;
    static int caps[kMaxDevices] = { 0 };
    int& cap = caps[current_device()];
    if (!cap) cap = resident_ctas(beam::eval_kernel, beam::SMEM_BYTES);
    const int grid = grid_for(a.e_end - a.e_begin, beam::EPW, cap);
    beam::eval_kernel<<<grid, 32, beam::SMEM_BYTES, (cudaStream_t)s>>>(a);
}
void launch_solid_eval(const EvalArgs& a, void* s) {
    if (a.e_end <= a.e_begin) return;
    static int caps[kMaxDevices] = { 0 };
    int& cap = caps[current_device()];
    if (!cap) cap = resident_ctas(solid::eval_kernel, solid::SMEM_BYTES);
    const int grid = grid_for(a.e_end - a.e_begin, solid::EPW, cap);
    solid::eval_kernel<<<grid, 32, solid::SMEM_BYTES, (cudaStream_t)s>>>(a);
}
void launch_shell_precalc(const EvalArgs& a, double* geo, double* shp, void* s) {
    if (a.n_el <= 0) return;
    const long long n = (long long)a.n_el * shell::NGP;
    shell::precalc_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)s>>>(a, geo, shp);
}
void launch_shell_commit(const EvalArgs& a, void* s) {
    if (a.n_el <= 0) return;
    const long long n = (long long)a.n_el * shell::NGP;
    shell::commit_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)s>>>(a);
}
void launch_beam_commit(const EvalArgs& a, void* s) {
    if (a.n_el <= 0) return;
    const long long n = (long long)a.n_el * beam::NGP;
    beam::commit_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)s>>>(a);
}
void launch_negate(double* v, long long n, void* s) {
    if (n > 0) negate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(v, n);
}
void launch_sub_ab_xb(double* PA, const int* rows, const int* ptr, const int* inner, const double* vals, const double* XB, int n_rows, void* s) {
    if (n_rows > 0) sub_ab_xb_kernel<<<(n_rows + 127) / 128, 128, 0, (cudaStream_t)s>>>(PA, rows, ptr, inner, vals, XB, n_rows);
}
void launch_update_disps(const int* gls, double* disp, const double* x, int n_nodes, void* s) {
    if (n_nodes > 0) update_disps_kernel<<<(unsigned)((6LL * n_nodes + 255) / 256), 256, 0, (cudaStream_t)s>>>(gls, disp, x, n_nodes);
}
void launch_norms(const int* gls, const double* v, const double* disp, int n_nodes, NormAcc* acc, void* s) {
    if (n_nodes <= 0) return;
    norms_max_kernel<<<(n_nodes + 255) / 256, 256, 0, (cudaStream_t)s>>>(gls, v, disp, n_nodes, acc);
    norms_node_kernel<<<(n_nodes + 255) / 256, 256, 0, (cudaStream_t)s>>>(gls, v, n_nodes, acc);
}
void launch_shell_results(const EvalArgs& a, double* out, void* s) {
    if (a.n_el <= 0) return;
    shell::results_kernel<<<(a.n_el + 127) / 128, 128, 0, (cudaStream_t)s>>>(a, out);
}
void launch_beam_results(const EvalArgs& a, double* out, void* s) {
    if (a.n_el <= 0) return;
    beam::results_kernel<<<(a.n_el + 127) / 128, 128, 0, (cudaStream_t)s>>>(a, out);
}
void launch_node_commit(int n_nodes, double* copy, double* disp, void* s) {
    if (n_nodes <= 0) return;
    node_commit_kernel<<<(n_nodes + 255) / 256, 256, 0, (cudaStream_t)s>>>(n_nodes, copy, disp);
}
int launch_scatter(const ScatterArgs& a, void* s) {
    int launches = 0;
    if (a.n_runs > 0) {
        scatter_kernel<<<(unsigned)((3 * a.n_runs + SCATTER_THREADS - 1) / SCATTER_THREADS), SCATTER_THREADS, 0, (cudaStream_t)s>>>(a);
        launches++;
    }
    if (a.n_gn > 0) {
        vector_kernel<<<(unsigned)((3 * a.n_gn + 255) / 256), 256, 0, (cudaStream_t)s>>>(a);
        launches++;
    }
    return launches;
}
void launch_vectors(const ScatterArgs& a, void* s) {
    if (a.n_gn > 0) vector_kernel<<<(unsigned)((3 * a.n_gn + 255) / 256), 256, 0, (cudaStream_t)s>>>(a);
}
int fused_buffers(int slot) {
    const int per = slot == 0 ? fused::ShellT::SMEM : slot == 1 ? fused::BeamT::SMEM : fused::SolidT::SMEM;
    // what the scatter CTA's staging buffers and the two reserved kilobytes leave of an SM's 228 KB
    const int n = (226 * 1024 - FUSED_SCATTER_WARPS * fused::STAGE_BYTES) / per;
    return n < FUSED_WARPS ? n : FUSED_WARPS;
}
int launch_fused_eval(const FusedArgs& f, void* s) {
    const int nb = f.n_buf, sms = sm_count();
    cudaStream_t st = (cudaStream_t)s;
    if (f.type_slot == 0) fused::eval_kernel<fused::ShellT><<<sms, 32 * nb, nb * fused::ShellT::SMEM, st>>>(f);
    else if (f.type_slot == 1) fused::eval_kernel<fused::BeamT><<<sms, 32 * nb, nb * fused::BeamT::SMEM, st>>>(f);
    else fused::eval_kernel<fused::SolidT><<<sms, 32 * nb, nb * fused::SolidT::SMEM, st>>>(f);
    return (int)cudaGetLastError();
}
int launch_fused_scatter(const FusedArgs& f, void* s) {
    fused::scatter_kernel<<<sm_count() * f.scatter_ctas, 32 * FUSED_SCATTER_WARPS, FUSED_SCATTER_WARPS * fused::STAGE_BYTES, (cudaStream_t)s>>>(f);
    return (int)cudaGetLastError();
}
void launch_gather(const GatherArgs& a, void* s) {
    if (a.n_dest <= 0) return;
    gather_kernel<<<(unsigned)((a.n_dest + 255) / 256), 256, 0, (cudaStream_t)s>>>(a);
}
void launch_pipe_loads(const EvalArgs& a, const ShellLoadArgs& l, void* s) {
    if (l.n_entries <= 0) return;
    beam::pipe_load_kernel<<<(unsigned)((l.n_entries + 63) / 64), 64, 0, (cudaStream_t)s>>>(a, l);
}
void launch_shell_loads(const EvalArgs& a, const ShellLoadArgs& l, void* s) {
    if (l.n_entries <= 0) return;
    shell::load_kernel<<<(unsigned)((36LL * l.n_entries + 127) / 128), 128, 0, (cudaStream_t)s>>>(a, l);
}
void launch_gather_add(double* vals, const long long* seg, const long long* src, const long long* dest, const double* from, long long n, void* s) {
    if (n <= 0) return;
    gather_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(vals, seg, src, dest, from, n);
}
void launch_add_slots(double* vals, const long long* slots, const double* add, long long n, void* s) {
    if (n <= 0) return;
    add_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(vals, slots, add, n);
}
void launch_pack(const double* vals, const long long* idx, double* buf, long long n, void* s) {
    if (n <= 0) return;
    pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(vals, idx, buf, n);
}
void launch_unpack_nodes(double* disp, const int* nodes, const double* packed, long long n, void* s) {
    if (n <= 0) return;
    unpack_nodes_kernel<<<(unsigned)((6 * n + 255) / 256), 256, 0, (cudaStream_t)s>>>(disp, nodes, packed, n);
}
void launch_unpack_add(double* vals, const long long* idx, const double* buf, long long n, void* s) {
    if (n <= 0) return;
    unpack_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(vals, idx, buf, n);
}

} // namespace gfa
