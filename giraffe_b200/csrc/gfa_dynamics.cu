// Newmark dynamics on the device: Element::MountMass + MountDamping + MountDyn for Beam_1 and Shell_1
// (reference Beam_1.cpp:1537-1673, Shell_1.cpp:2367-2541, driven by Dynamic.cpp:323-337 through
// Solution.cpp:711-759) and Dynamic::UpdateDyn (Dynamic.cpp:480-556).
//
// The reference evaluates the Gauss-point inertial pseudo-forces dT and their tangent DdT with
// AceGen-generated code (Beam_1.cpp:1781-2361, Shell_1.cpp:2568-2890).  The formulation behind it
// (Newmark in the tangent space of the incremental rotation, Dynamic.cpp:509-545):
//     Q = Q(alpha_d) Q(alpha_i),  Xi = Xi(alpha_d)
//     omega  = Q(alpha_d) (a4 alpha_d + a5 omega_i + a6 domega_i)
//     domega = Q(alpha_d) (a1 alpha_d - a2 omega_i - a3 domega_i)
//     ddu    = a1 u_d - a2 du_i - a3 ddu_i
//     Beam_1 : f = rho A ddu,   mu = J domega + omega x (J omega),   J = Q Jr Q^T
//     Shell_1: f = coef1 ddu,   mu = coef2 e3 x (domega x e3 + omega x (omega x e3)),   e3 = Q e3r
//     dT = [f ; Xi^T mu],  DdT = d dT / d(u_d, alpha_d)
// (in-scope sections have br = 0 and Mr = rho A I, Beam_1.cpp:582-596, so the first-moment terms of
// the general code vanish and Q Mr Q^T = rho A I).  The rotational tangent is the forward-mode
// derivative of Xi^T mu with respect to alpha_d (three directions), i.e. the exact derivative AceGen
// generates symbolically; the translational tangent is a1 m I.
//
// Pipe_1 rides on the beam kernels with its own Mr = Rho I and Jr (Pipe_1.cpp:1131-1144, structural mass only).
// Two launches per element type:
//   gp kernel    one thread per element: Gauss-point quantities -> a small record per element
//                (node-pair scalars of the u-u blocks, 3x3 alpha-alpha blocks in global axes,
//                inertial_loading; with update_rayleigh also the modal mass of MountMassModal)
//   apply kernel one warp per element, lanes over the element's arena region (coalesced):
//                rayleigh_damping = alpha*mass_modal + beta*stiffness   (when updating; stored in a
//                second arena with the layout of the Ke arena)
//                stiffness += mass + a4*rayleigh_damping,  P_loading += inertial_loading +
//                rayleigh_damping * v   (one lane per row, columns ascending as the reference's GEMV)
#include <cuda_runtime.h>
#include <climits>

#include "gfa_device.h"
#include "gfa_math.cuh"

namespace gfa {

namespace {

// ---- forward-mode scalar: value + derivatives with respect to alpha_d(0..2) ----
struct D3 { double v, x, y, z; };
GFA_DI D3 mk(double v) { return D3{ v, 0.0, 0.0, 0.0 }; }
GFA_DI D3 operator+(const D3& a, const D3& b) { return D3{ a.v + b.v, a.x + b.x, a.y + b.y, a.z + b.z }; }
GFA_DI D3 operator-(const D3& a, const D3& b) { return D3{ a.v - b.v, a.x - b.x, a.y - b.y, a.z - b.z }; }
GFA_DI D3 operator-(const D3& a) { return D3{ -a.v, -a.x, -a.y, -a.z }; }
GFA_DI D3 operator*(const D3& a, const D3& b) { return D3{ a.v * b.v, a.x * b.v + a.v * b.x, a.y * b.v + a.v * b.y, a.z * b.v + a.v * b.z }; }
GFA_DI D3 operator*(double s, const D3& a) { return D3{ s * a.v, s * a.x, s * a.y, s * a.z }; }
GFA_DI D3 operator+(const D3& a, double s) { return D3{ a.v + s, a.x, a.y, a.z }; }
GFA_DI D3 recip4(const D3& d) {       // 4 / d
    const double q = 4.0 / d.v, m = -q / d.v;
    return D3{ q, m * d.x, m * d.y, m * d.z };
}
GFA_DI void crossD(D3* o, const D3* a, const D3* b) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
GFA_DI void mvD(D3* y, const D3* A, const D3* x) {
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
// Q = I + g (A + A A / 2), Xi = g (I + A / 2), g = 4 / (4 + |a|^2)
GFA_DI void rodriguesD(const D3* a, D3* Q, D3* Xi) {
    const D3 g = recip4(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + 4.0);
    const D3 z = mk(0.0);
    const D3 A[9] = { z, -a[2], a[1], a[2], z, -a[0], -a[1], a[0], z };
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const D3 s2 = A[3 * i] * A[j] + A[3 * i + 1] * A[3 + j] + A[3 * i + 2] * A[6 + j];
            const double id = i == j ? 1.0 : 0.0;
            Q[3 * i + j] = g * (A[3 * i + j] + 0.5 * s2) + id;
            Xi[3 * i + j] = g * (0.5 * A[3 * i + j] + id);
        }
}
GFA_DI void rotation_of(const double* a, double* Q) {
    double g, Xi[9];
    rodrigues(a, g, Q, Xi);
}

struct Newmark { double a1, a2, a3, a4, a5, a6; };

// omega, domega (forward-mode), Q = Q(alpha_d) Q(alpha_i), Xi(alpha_d); everything in the element frame
GFA_DI void newmark_rotation(const Newmark& nm, const double* ad, const double* ai, const double* om_i, const double* dom_i,
                             D3* Q, D3* Xi, D3* w, D3* dw) {
    D3 a[3] = { D3{ ad[0], 1.0, 0.0, 0.0 }, D3{ ad[1], 0.0, 1.0, 0.0 }, D3{ ad[2], 0.0, 0.0, 1.0 } };
    D3 Qd[9];
    rodriguesD(a, Qd, Xi);
    double Qi[9];
    rotation_of(ai, Qi);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Q[3 * i + j] = Qi[j] * Qd[3 * i] + Qi[3 + j] * Qd[3 * i + 1] + Qi[6 + j] * Qd[3 * i + 2];
    D3 wl[3], dwl[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        wl[k] = nm.a4 * a[k] + (om_i[k] * nm.a5 + dom_i[k] * nm.a6);
        dwl[k] = nm.a1 * a[k] + (-om_i[k] * nm.a2 - dom_i[k] * nm.a3);
    }
    mvD(w, Qd, wl); mvD(dw, Qd, dwl);
}
// dT(3..5) = Xi^T mu -> value P (3) and tangent K (3x3, row-major), element frame
GFA_DI void finish_moment(const D3* Xi, const D3* mu, double* P, double* K) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const D3 t = Xi[k] * mu[0] + Xi[3 + k] * mu[1] + Xi[6 + k] * mu[2];
        P[k] = t.v; K[3 * k] = t.x; K[3 * k + 1] = t.y; K[3 * k + 2] = t.z;
    }
}
// to global axes: Pg = R^T P, Kg = R^T K R
GFA_DI void to_global(const double* R, const double* P, const double* K, double* Pg, double* Kg) {
    double t[9];
    mtv(Pg, R, P);
    mtm(t, R, K);
    mm(Kg, t, R);
}
GFA_DI void load3(double* o, const double* arr, int node, int off) {
    const double* p = arr + 6 * (size_t)node + off;
    o[0] = __ldg(p); o[1] = __ldg(p + 1); o[2] = __ldg(p + 2);
}

// =========================================================================
// Beam_1
// =========================================================================
namespace beam {
constexpr int NGP = 2;
constexpr int UU = 0, AA = 9, PP = 90, MUU = 108, MAA = 117;

struct Geo { double R[9], jac, N[2][3]; };
GFA_DI void geometry(const EvalArgs& A, int e, const int* nd, const double* pr, Geo& go) {
#pragma unroll
    for (int i = 0; i < 9; i++) go.R[i] = __ldg(pr + 36 + i);
    double d[3];
#pragma unroll
    for (int c = 0; c < 3; c++) d[c] = __ldg(A.xyz + 3 * (size_t)nd[2] + c) - __ldg(A.xyz + 3 * (size_t)nd[0] + c);
    const double len = norm3(d);
    const double T0 = A.pret ? __ldg(A.pret + e) : 0.0;
    const double du0 = T0 / __ldg(pr + 14);                       // Beam_1.cpp:616-621
    go.jac = (len / (1.0 + du0)) / 2.0;
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const double xi = g == 0 ? -0.577350269189626 : 0.577350269189626;
        go.N[g][0] = 0.5 * xi * (xi - 1.0); go.N[g][1] = 1.0 - xi * xi; go.N[g][2] = 0.5 * xi * (1.0 + xi);
    }
}

// MountMass (Beam_1.cpp:1564-1636) and, when updating, MountMassModal (:1537-1552)
__global__ void gp_kernel(EvalArgs A, DynArgs D) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n_el) return;
    int nd[3];
#pragma unroll
    for (int n = 0; n < 3; n++) nd[n] = __ldg(A.conn + 3 * (size_t)e + n);
    const double* pr = A.props + BEAM_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    Geo go; geometry(A, e, nd, pr, go);
    const double rhoA = __ldg(pr + 45);
    const double J11 = __ldg(pr + 47), J22 = __ldg(pr + 48), J33 = __ldg(pr + 49), J12 = __ldg(pr + 50);
    const Newmark nm = { D.a1, D.a2, D.a3, D.a4, D.a5, D.a6 };
    const size_t n_gp = (size_t)A.n_el * NGP;
    double* rec = D.rec + (size_t)e * BEAM_DYN_REC;
    const double w = 1.0 * go.jac;                                   // alpha1 * jacobian

    double Kg[NGP][9], Pg[NGP][3], fu[NGP][3], Jm[NGP][9];
#pragma unroll
    for (int g = 0; g < NGP; g++) {
        double ga[3] = { 0, 0, 0 }, gu[3] = { 0, 0, 0 }, gom[3] = { 0, 0, 0 }, gdom[3] = { 0, 0, 0 }, gdu[3] = { 0, 0, 0 }, gddu[3] = { 0, 0, 0 };
#pragma unroll
        for (int n = 0; n < 3; n++) {
            double t[3];
            const double N = go.N[g][n];
            load3(t, A.disp, nd[n], 3);
#pragma unroll
            for (int c = 0; c < 3; c++) ga[c] += t[c] * N;
            load3(t, A.disp, nd[n], 0);
#pragma unroll
            for (int c = 0; c < 3; c++) gu[c] += t[c] * N;
            load3(t, D.copy_vel, nd[n], 3);
#pragma unroll
            for (int c = 0; c < 3; c++) gom[c] += t[c] * N;
            load3(t, D.copy_accel, nd[n], 3);
#pragma unroll
            for (int c = 0; c < 3; c++) gdom[c] += t[c] * N;
            load3(t, D.copy_vel, nd[n], 0);
#pragma unroll
            for (int c = 0; c < 3; c++) gdu[c] += t[c] * N;
            load3(t, D.copy_accel, nd[n], 0);
#pragma unroll
            for (int c = 0; c < 3; c++) gddu[c] += t[c] * N;
        }
        double ad[3], om[3], dom[3], ai[3];
        mv(ad, go.R, ga); mv(om, go.R, gom); mv(dom, go.R, gdom);          // to the element frame (:1618-1622, :743)
        const size_t gp = (size_t)e * NGP + g;
#pragma unroll
        for (int k = 0; k < 3; k++) ai[k] = D.alpha_i[k * n_gp + gp];
        // f = rho A ddu, in global axes directly (R^T R = I)
#pragma unroll
        for (int c = 0; c < 3; c++) fu[g][c] = rhoA * (nm.a1 * gu[c] - nm.a2 * gdu[c] - nm.a3 * gddu[c]);
        D3 Q[9], Xi[9], wv[3], dwv[3];
        newmark_rotation(nm, ad, ai, om, dom, Q, Xi, wv, dwv);
        // J x = Q Jr Q^T x
        D3 qw[3], qdw[3], jw[3], jdw[3], Jw[3], Jdw[3], wJw[3], mu[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            qw[k] = Q[k] * wv[0] + Q[3 + k] * wv[1] + Q[6 + k] * wv[2];
            qdw[k] = Q[k] * dwv[0] + Q[3 + k] * dwv[1] + Q[6 + k] * dwv[2];
        }
        jw[0] = J11 * qw[0] + J12 * qw[1]; jw[1] = J12 * qw[0] + J22 * qw[1]; jw[2] = J33 * qw[2];
        jdw[0] = J11 * qdw[0] + J12 * qdw[1]; jdw[1] = J12 * qdw[0] + J22 * qdw[1]; jdw[2] = J33 * qdw[2];
        mvD(Jw, Q, jw); mvD(Jdw, Q, jdw);
        crossD(wJw, wv, Jw);
#pragma unroll
        for (int k = 0; k < 3; k++) mu[k] = Jdw[k] + wJw[k];
        double P[3], K[9];
        finish_moment(Xi, mu, P, K);
        to_global(go.R, P, K, Pg[g], Kg[g]);
        if (D.update) {                                              // EvaluateMassModal (:1681-1777): J(alpha_i) in global axes
            double Qi[9], t[9], JrQt[9], Jl[9];
            rotation_of(ai, Qi);
            const double Jr[9] = { J11, J12, 0.0, J12, J22, 0.0, 0.0, 0.0, J33 };
            m_transpose(t, Qi);
            mm(JrQt, Jr, t);
            mm(Jl, Qi, JrQt);
            mtm(t, go.R, Jl);
            mm(Jm[g], t, go.R);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            rec[PP + 6 * a + c] = w * go.N[0][a] * fu[0][c] + w * go.N[1][a] * fu[1][c];
            rec[PP + 6 * a + 3 + c] = w * go.N[0][a] * Pg[0][c] + w * go.N[1][a] * Pg[1][c];
        }
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const double s0 = w * go.N[0][a] * go.N[0][b], s1 = w * go.N[1][a] * go.N[1][b];
            rec[UU + 3 * a + b] = (s0 + s1) * (nm.a1 * rhoA);
#pragma unroll
            for (int i = 0; i < 9; i++) rec[AA + 9 * (3 * a + b) + i] = s0 * Kg[0][i] + s1 * Kg[1][i];
            if (D.update) {
                rec[MUU + 3 * a + b] = (s0 + s1) * rhoA;
#pragma unroll
                for (int i = 0; i < 9; i++) rec[MAA + 9 * (3 * a + b) + i] = s0 * Jm[0][i] + s1 * Jm[1][i];
            }
        }
    }
}

// committed Rodrigues vector (Beam_1.cpp:1502-1503), one thread per Gauss point
__global__ void alpha_commit_kernel(EvalArgs A, double* alpha_i) {
    const size_t n_gp = (size_t)A.n_el * NGP;
    const size_t gp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= n_gp) return;
    const int e = (int)(gp / NGP), g = (int)(gp % NGP);
    int nd[3];
#pragma unroll
    for (int n = 0; n < 3; n++) nd[n] = __ldg(A.conn + 3 * (size_t)e + n);
    const double* pr = A.props + BEAM_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    Geo go; geometry(A, e, nd, pr, go);
    double ga[3] = { 0, 0, 0 }, ad[3], ai[3], cr[3];
#pragma unroll
    for (int n = 0; n < 3; n++) {
        double t[3];
        load3(t, A.disp, nd[n], 3);
#pragma unroll
        for (int c = 0; c < 3; c++) ga[c] += t[c] * (g == 0 ? go.N[0][n] : go.N[1][n]);
    }
    mv(ad, go.R, ga);
#pragma unroll
    for (int k = 0; k < 3; k++) ai[k] = alpha_i[k * n_gp + gp];
    cross3(cr, ad, ai);
    const double s = 4.0 / (4.0 - dot3(ad, ai));
#pragma unroll
    for (int k = 0; k < 3; k++) alpha_i[k * n_gp + gp] = s * (ad[k] + ai[k] + 0.5 * cr[k]);
}
} // namespace beam

// =========================================================================
// Shell_1
// =========================================================================
namespace shell {
constexpr int NGP = 3;
constexpr int UU = 0, AA = 36, PP = 117, MUU = 144, MAA = 180;

// area coordinates of the in-plane point g (located at mid-side node 4+g, Shell_1.cpp:2029-2039)
GFA_DI void area_coordinates(const double (&x)[6][3], double area, int g, double* L) {
    const double* xp = x[3 + g];
    double a[3], b[3], c[3], t[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { a[k] = x[0][k] - xp[k]; b[k] = x[1][k] - xp[k]; c[k] = x[2][k] - xp[k]; }
    cross3(t, b, c); L[0] = 0.5 * norm3(t) / area;
    cross3(t, c, a); L[1] = 0.5 * norm3(t) / area;
    cross3(t, a, b); L[2] = 0.5 * norm3(t) / area;
}
GFA_DI void shape_values(const double* L, double* Nu, double* Na) {
    Nu[0] = (2 * L[0] - 1) * L[0]; Nu[1] = (2 * L[1] - 1) * L[1]; Nu[2] = (2 * L[2] - 1) * L[2];
    Nu[3] = 4 * L[0] * L[1]; Nu[4] = 4 * L[1] * L[2]; Nu[5] = 4 * L[2] * L[0];
    Na[0] = 1 - 2 * L[2]; Na[1] = 1 - 2 * L[0]; Na[2] = 1 - 2 * L[1];
}

// MountMass (Shell_1.cpp:2406-2498) and, when updating, MountMassModal (:2367-2396)
__global__ void gp_kernel(EvalArgs A, DynArgs D) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.n_el) return;
    const size_t ne = (size_t)A.n_el, n_gp = ne * NGP;
    int nd[6];
    double x[6][3];
#pragma unroll
    for (int n = 0; n < 6; n++) {
        nd[n] = __ldg(A.conn + 6 * (size_t)e + n);
#pragma unroll
        for (int c = 0; c < 3; c++) x[n][c] = __ldg(A.xyz + 3 * (size_t)nd[n] + c);
    }
    double R[9];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = __ldg(A.geo + k * ne + e);
    const double area = __ldg(A.geo + 9 * ne + e);
    const double* pr = A.props + SHELL_PROP_STRIDE * (size_t)__ldg(A.prop + e);
    const double thick = __ldg(pr + 2), rho = __ldg(pr + 4);
    const double coef1 = thick * rho, coef2 = (1.0 / 12.0) * thick * thick * thick * rho;     // :2022-2024
    const double coef3 = rho * thick * area / (3 * 3.1415926535897932384626433832795);
    const Newmark nm = { D.a1, D.a2, D.a3, D.a4, D.a5, D.a6 };
    double* rec = D.rec + (size_t)e * SHELL_DYN_REC;
    const double w = area / 3.0;                                     // alpha1 (:2364)

    double Kg[NGP][9], Pg[NGP][3], fu[NGP][3], Nu[NGP][6], Na[NGP][3];
#pragma unroll
    for (int g = 0; g < NGP; g++) {
        double L[3];
        area_coordinates(x, area, g, L);
        shape_values(L, Nu[g], Na[g]);
        double gu[3] = { 0, 0, 0 }, gdu[3] = { 0, 0, 0 }, gddu[3] = { 0, 0, 0 }, ga[3] = { 0, 0, 0 }, gom[3] = { 0, 0, 0 }, gdom[3] = { 0, 0, 0 };
#pragma unroll
        for (int n = 0; n < 6; n++) {
            double t[3];
            load3(t, A.disp, nd[n], 0);
#pragma unroll
            for (int c = 0; c < 3; c++) gu[c] += t[c] * Nu[g][n];
            load3(t, D.copy_vel, nd[n], 0);
#pragma unroll
            for (int c = 0; c < 3; c++) gdu[c] += t[c] * Nu[g][n];
            load3(t, D.copy_accel, nd[n], 0);
#pragma unroll
            for (int c = 0; c < 3; c++) gddu[c] += t[c] * Nu[g][n];
            if (n >= 3) {
                load3(t, A.disp, nd[n], 3);
#pragma unroll
                for (int c = 0; c < 3; c++) ga[c] += t[c] * Na[g][n - 3];
                load3(t, D.copy_vel, nd[n], 3);
#pragma unroll
                for (int c = 0; c < 3; c++) gom[c] += t[c] * Na[g][n - 3];
                load3(t, D.copy_accel, nd[n], 3);
#pragma unroll
                for (int c = 0; c < 3; c++) gdom[c] += t[c] * Na[g][n - 3];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) fu[g][c] = coef1 * (nm.a1 * gu[c] - nm.a2 * gdu[c] - nm.a3 * gddu[c]);
        double ad[3], om[3], dom[3], ai[3];
        mv(ad, R, ga); mv(om, R, gom); mv(dom, R, gdom);               // :2471-2474, :996
        const size_t gp = (size_t)e * NGP + g;
#pragma unroll
        for (int k = 0; k < 3; k++) ai[k] = D.alpha_i[k * n_gp + gp];
        D3 Q[9], Xi[9], wv[3], dwv[3];
        newmark_rotation(nm, ad, ai, om, dom, Q, Xi, wv, dwv);
        const D3 e3[3] = { Q[2], Q[5], Q[8] };                        // Q e3r, e3r = (0,0,1)
        D3 we[3], wwe[3], dwe[3], acc[3], mu[3];
        crossD(we, wv, e3); crossD(wwe, wv, we); crossD(dwe, dwv, e3);
#pragma unroll
        for (int k = 0; k < 3; k++) acc[k] = dwe[k] + wwe[k];
        crossD(mu, e3, acc);
#pragma unroll
        for (int k = 0; k < 3; k++) mu[k] = coef2 * mu[k];
        double P[3], K[9];
        finish_moment(Xi, mu, P, K);
        to_global(R, P, K, Pg[g], Kg[g]);
    }
#pragma unroll
    for (int a = 0; a < 6; a++) {
#pragma unroll
        for (int c = 0; c < 3; c++) rec[PP + 3 * a + c] = w * Nu[0][a] * fu[0][c] + w * Nu[1][a] * fu[1][c] + w * Nu[2][a] * fu[2][c];
#pragma unroll
        for (int b = 0; b < 6; b++)
            rec[UU + 6 * a + b] = (w * Nu[0][a] * Nu[0][b] + w * Nu[1][a] * Nu[1][b] + w * Nu[2][a] * Nu[2][b]) * (nm.a1 * coef1);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int c = 0; c < 3; c++) rec[PP + 18 + 3 * a + c] = w * Na[0][a] * Pg[0][c] + w * Na[1][a] * Pg[1][c] + w * Na[2][a] * Pg[2][c];
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const double s0 = w * Na[0][a] * Na[0][b], s1 = w * Na[1][a] * Na[1][b], s2 = w * Na[2][a] * Na[2][b];
#pragma unroll
            for (int i = 0; i < 9; i++) rec[AA + 9 * (3 * a + b) + i] = s0 * Kg[0][i] + s1 * Kg[1][i] + s2 * Kg[2][i];
        }
    }
    if (!D.update) return;
    // MountMassModal: 6-point rule (Shell_1.cpp:2185-2246), alpha_i4 from the mid nodes' committed rotations
    double muu[36], maa[81];
#pragma unroll
    for (int i = 0; i < 36; i++) muu[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 81; i++) maa[i] = 0.0;
    double crot[3][3];
#pragma unroll
    for (int n = 0; n < 3; n++) load3(crot[n], A.copy, nd[3 + n], 3);
#pragma unroll 1
    for (int g = 0; g < 6; g++) {
        const double c1 = g < 3 ? 0.816847572980459 : 0.108103018168070, c2 = g < 3 ? 0.091576213509771 : 0.445948490915965;
        const int k = g % 3;
        const double L[3] = { k == 0 ? c1 : c2, k == 1 ? c1 : c2, k == 2 ? c1 : c2 };
        const double w4 = area * (g < 3 ? 0.109951743655322 : 0.223381589678011);
        double N4u[6], N4a[3];
        shape_values(L, N4u, N4a);
        double gai[3] = { 0, 0, 0 }, ai[3], Qa[9];
        for (int n = 0; n < 3; n++)
            for (int c = 0; c < 3; c++) gai[c] += crot[n][c] * N4a[n];
        mv(ai, R, gai);
        rotation_of(ai, Qa);
        const double e3[3] = { Qa[2], Qa[5], Qa[8] };
        // EvaluateMassModal (:3171-3221): J = coef2 (|e3|^2 I - e3 e3^T) + coef3 e3 e3^T, as written
        const double q0 = e3[0] * e3[0], q1 = e3[1] * e3[1], q2 = e3[2] * e3[2], dc = -coef2 + coef3;
        double Jl[9];
        Jl[0] = coef3 * q0 + coef2 * (q1 + q2); Jl[4] = coef3 * q1 + coef2 * (q0 + q2); Jl[8] = coef2 * (q1 + q0) + coef3 * q2;
        Jl[1] = e3[0] * (e3[1] * dc); Jl[2] = e3[0] * e3[2] * dc; Jl[5] = e3[2] * (e3[1] * dc);
        Jl[3] = Jl[1]; Jl[6] = Jl[2]; Jl[7] = Jl[5];
        double t[9], Jg[9];
        mtm(t, R, Jl);
        mm(Jg, t, R);
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) muu[6 * a + b] += w4 * N4u[a] * N4u[b] * coef1;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                const double s = w4 * N4a[a] * N4a[b];
                for (int i = 0; i < 9; i++) maa[9 * (3 * a + b) + i] += s * Jg[i];
            }
    }
    for (int i = 0; i < 36; i++) rec[MUU + i] = muu[i];
    for (int i = 0; i < 81; i++) rec[MAA + i] = maa[i];
}

// committed Rodrigues vector (Shell_1.cpp:1659-1660), one thread per Gauss point
__global__ void alpha_commit_kernel(EvalArgs A, double* alpha_i) {
    const size_t ne = (size_t)A.n_el, n_gp = ne * NGP;
    const size_t gp = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= n_gp) return;
    const int e = (int)(gp / NGP);
    double R[9], Na[3], ga[3] = { 0, 0, 0 }, ad[3], ai[3], cr[3];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = __ldg(A.geo + k * ne + e);
#pragma unroll
    for (int k = 0; k < 3; k++) Na[k] = __ldg(A.shp + (18 + k) * n_gp + gp);
#pragma unroll
    for (int n = 0; n < 3; n++) {
        double t[3];
        load3(t, A.disp, __ldg(A.conn + 6 * (size_t)e + 3 + n), 3);
#pragma unroll
        for (int c = 0; c < 3; c++) ga[c] += t[c] * Na[n];
    }
    mv(ad, R, ga);
#pragma unroll
    for (int k = 0; k < 3; k++) ai[k] = alpha_i[k * n_gp + gp];
    cross3(cr, ad, ai);
    const double s = 4.0 / (4.0 - dot3(ad, ai));
#pragma unroll
    for (int k = 0; k < 3; k++) alpha_i[k * n_gp + gp] = s * (ad[k] + ai[k] + 0.5 * cr[k]);
}
} // namespace shell

// =========================================================================
// apply: MountDamping + MountDyn on the element arena, one warp per element
// =========================================================================
// arena offset -> (row group a, column group b, i, j) packed a | b<<4 | i<<8 | j<<10; 0xFFFF = padding.
// Lanes look up different offsets, so the table is staged in shared memory (constant memory would
// serialise the divergent reads).
__device__ unsigned short g_shell_decode[SHELL_ARENA];
__device__ unsigned short g_beam_decode[BEAM_ARENA];

template <bool SHELL> struct Lay;
template <> struct Lay<true> {
    static constexpr int ARENA = SHELL_ARENA, NDOF = 27, REC = SHELL_DYN_REC, NN = 6, NG = 9;
    static constexpr int UU = shell::UU, AA = shell::AA, PP = shell::PP, MUU = shell::MUU, MAA = shell::MAA, NUU = 6;
    GFA_DI static const unsigned short* decode_table() { return g_shell_decode; }
    GFA_DI static int block_offset(int a, int b, bool& tr) { return shell_block_offset(a, b, tr); }
    // group -> (translation node | rotation node, is rotation)
    GFA_DI static void group(int grp, int& node, bool& rot) { rot = grp >= 6; node = rot ? grp - 6 : grp; }
    GFA_DI static int vel_index(const int* nd, int grp, int comp) { return grp < 6 ? 6 * nd[grp] + comp : 6 * nd[3 + grp - 6] + 3 + comp; }
};
template <> struct Lay<false> {
    static constexpr int ARENA = BEAM_ARENA, NDOF = 18, REC = BEAM_DYN_REC, NN = 3, NG = 6;
    static constexpr int UU = beam::UU, AA = beam::AA, PP = beam::PP, MUU = beam::MUU, MAA = beam::MAA, NUU = 3;
    GFA_DI static const unsigned short* decode_table() { return g_beam_decode; }
    GFA_DI static int block_offset(int a, int b, bool& tr) { return beam_block_offset(a, b, tr); }
    GFA_DI static void group(int grp, int& node, bool& rot) { rot = grp & 1; node = grp >> 1; }
    GFA_DI static int vel_index(const int* nd, int grp, int comp) { return 6 * nd[grp >> 1] + ((grp & 1) ? 3 : 0) + comp; }
};
// packed table entry -> (row group a, column group b, i, j); false for padding
GFA_DI bool decode_entry(const unsigned short* tab, int off, int& a, int& b, int& i, int& j) {
    const unsigned v = tab[off];
    if (v == 0xFFFFu) return false;
    a = v & 15; b = (v >> 4) & 15; i = (v >> 8) & 3; j = (v >> 10) & 3;
    return true;
}

template <bool SHELL>
__global__ void __launch_bounds__(128) apply_kernel(EvalArgs A, DynArgs D) {
    using L = Lay<SHELL>;
    __shared__ unsigned short tab[L::ARENA];
    __shared__ unsigned short ent[L::NG * L::NG];       // block (ra, cb) -> offset | transposed << 15
    for (int i = threadIdx.x; i < L::ARENA; i += blockDim.x) tab[i] = L::decode_table()[i];
    for (int i = threadIdx.x; i < L::NG * L::NG; i += blockDim.x) {
        bool tr;
        const int o = L::block_offset(i / L::NG, i % L::NG, tr);
        ent[i] = (unsigned short)(o | (tr ? 0x8000 : 0));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    // what this lane adds at each of its arena offsets does not depend on the element: record index of the
    // mass term (-1: none, -2: padding / beyond the region) and of the modal-mass term
    constexpr int NIT = (L::ARENA + 31) / 32;
    int aidx[NIT], midx[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int off = lane + 32 * it;
        int a, b, i, j;
        aidx[it] = -2; midx[it] = -1;
        if (off < L::ARENA && decode_entry(tab, off, a, b, i, j)) {
            int na, nb; bool ra, rb;
            L::group(a, na, ra); L::group(b, nb, rb);
            aidx[it] = -1;
            if (!ra && !rb) { if (i == j) { aidx[it] = L::UU + L::NUU * na + nb; midx[it] = L::MUU + L::NUU * na + nb; } }
            else if (ra && rb) { aidx[it] = L::AA + 9 * (3 * na + nb) + 3 * i + j; midx[it] = L::MAA + 9 * (3 * na + nb) + 3 * i + j; }
        }
    }
    for (long long e = warp; e < A.n_el; e += n_warps) {
        double* Ke = A.Ke + (size_t)e * L::ARENA;
        double* CR = D.CR ? D.CR + (size_t)e * L::ARENA : nullptr;
        const double* rec = D.rec + (size_t)e * L::REC;
        double k[NIT], cr[NIT], add[NIT];
#pragma unroll
        for (int it = 0; it < NIT; it++) {          // all loads of the element first
            const int off = lane + 32 * it;
            k[it] = 0.0; cr[it] = 0.0; add[it] = 0.0;
            if (aidx[it] != -2) {
                k[it] = Ke[off];
                if (aidx[it] >= 0) add[it] = rec[aidx[it]];
                if (CR) cr[it] = D.update ? (midx[it] >= 0 ? rec[midx[it]] : 0.0) : CR[off];
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; it++) {
            const int off = lane + 32 * it;
            if (aidx[it] == -2) continue;
            double c = cr[it];
            if (CR && D.update) { c = D.ray_alpha * c + D.ray_beta * k[it]; CR[off] = c; }   // rayleigh_damping (:1649, :2510)
            Ke[off] = k[it] + add[it] + D.a4 * c;                                            // MountDyn (:1672, :2540)
        }
        __syncwarp();
        if (lane < L::NDOF) {
            int nd[L::NN];
#pragma unroll
            for (int n = 0; n < L::NN; n++) nd[n] = __ldg(A.conn + L::NN * (size_t)e + n);
            double dl = 0.0, dl0 = 0.0, dl1 = 0.0, dl2 = 0.0;
            if (CR) {
                const int ra = lane / 3, i = lane % 3;
                for (int cb = 0; cb < L::NDOF / 3; cb++) {                                   // rayleigh_damping * v_ipp (:1663, :2531)
                    const unsigned v = ent[L::NG * ra + cb];
                    const int o0 = (v & 0x7fff) + ((v & 0x8000) ? i : 3 * i), st = (v & 0x8000) ? 3 : 1;
                    dl0 = fma(CR[o0], __ldg(D.vel + L::vel_index(nd, cb, 0)), dl0);
                    dl1 = fma(CR[o0 + st], __ldg(D.vel + L::vel_index(nd, cb, 1)), dl1);
                    dl2 = fma(CR[o0 + 2 * st], __ldg(D.vel + L::vel_index(nd, cb, 2)), dl2);
                }
                dl = dl0 + dl1 + dl2;
            }
            double* P = A.Pe + (size_t)e * L::NDOF + lane;
            *P = *P + rec[L::PP + lane] + dl;                                                // MountDyn (:1670, :2538)
        }
        __syncwarp();
    }
}

// =========================================================================
// Dynamic::UpdateDyn (Dynamic.cpp:480-556)
// =========================================================================
struct NodeKin { double va[3], aa[3]; };
// one node of the reference's loop acting on the running vel_aux / ace_aux
GFA_DI void update_rotation_step(const DynArgs& D, const int* gl, const double* d, const double* cv, const double* ca, NodeKin& k) {
    double gg, Qd[9], Xi[9];
    rodrigues(d + 3, gg, Qd, Xi);
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (gl[3 + j] > 0) {
            k.va[j] = d[3 + j] * D.a4 + cv[3 + j] * D.a5 + ca[3 + j] * D.a6;
            k.aa[j] = d[3 + j] * D.a1 - cv[3 + j] * D.a2 - ca[3 + j] * D.a3;
        }
    double t[3];
    mv(t, Qd, k.va); k.va[0] = t[0]; k.va[1] = t[1]; k.va[2] = t[2];
    mv(t, Qd, k.aa); k.aa[0] = t[0]; k.aa[1] = t[1]; k.aa[2] = t[2];
}
__global__ void update_dyn_kernel(DynArgs D, const int* gls, const double* disp, double* vel, double* accel, int n_nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    int gl[6]; double d[6], cv[6], ca[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        gl[k] = gls[6 * (size_t)i + k]; d[k] = disp[6 * (size_t)i + k];
        cv[k] = D.copy_vel[6 * (size_t)i + k]; ca[k] = D.copy_accel[6 * (size_t)i + k];
    }
#pragma unroll
    for (int j = 0; j < 3; j++)
        if (gl[j] > 0) {
            vel[6 * (size_t)i + j] = d[j] * D.a4 + cv[j] * D.a5 + ca[j] * D.a6;
            accel[6 * (size_t)i + j] = d[j] * D.a1 - cv[j] * D.a2 - ca[j] * D.a3;
        }
    // all three rotational DOFs free: the node does not see earlier nodes; none free: nothing is written;
    // partly free: replay kernel
    if (gl[3] > 0 && gl[4] > 0 && gl[5] > 0) {
        NodeKin k;
        update_rotation_step(D, gl, d, cv, ca, k);
#pragma unroll
        for (int j = 0; j < 3; j++) { vel[6 * (size_t)i + 3 + j] = k.va[j]; accel[6 * (size_t)i + 3 + j] = k.aa[j]; }
    }
}
// vel_aux / ace_aux are declared outside the node loop in the reference: a rotational DOF that is not free
// keeps what the previous node left.  One thread per partly-free node replays the loop from the last node
// that overwrote all three components (or from node 0 with zeros).
__global__ void update_dyn_replay_kernel(DynArgs D, const int* gls, const double* disp, double* vel, double* accel,
                                         const int* mixed, const int* start, int n_mixed) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mixed) return;
    const int node = mixed[m];
    NodeKin k = { { 0, 0, 0 }, { 0, 0, 0 } };
    for (int i = start[m]; i <= node; i++) {
        int gl[6]; double d[6], cv[6], ca[6];
#pragma unroll
        for (int q = 0; q < 6; q++) {
            gl[q] = gls[6 * (size_t)i + q]; d[q] = disp[6 * (size_t)i + q];
            cv[q] = D.copy_vel[6 * (size_t)i + q]; ca[q] = D.copy_accel[6 * (size_t)i + q];
        }
        update_rotation_step(D, gl, d, cv, ca, k);
        if (i == node)
#pragma unroll
            for (int j = 0; j < 3; j++)
                if (gl[3 + j] > 0) { vel[6 * (size_t)i + 3 + j] = k.va[j]; accel[6 * (size_t)i + 3 + j] = k.aa[j]; }
    }
}

inline int blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }

} // namespace

void launch_beam_dynamics(const EvalArgs& a, const DynArgs& d, void* stream) {
    if (a.n_el == 0) return;
    cudaStream_t s = (cudaStream_t)stream;
    beam::gp_kernel<<<blocks_for(a.n_el, 64), 64, 0, s>>>(a, d);
    apply_kernel<false><<<blocks_for((long long)a.n_el * 32, 128), 128, 0, s>>>(a, d);
}
void launch_shell_dynamics(const EvalArgs& a, const DynArgs& d, void* stream) {
    if (a.n_el == 0) return;
    cudaStream_t s = (cudaStream_t)stream;
    shell::gp_kernel<<<blocks_for(a.n_el, 64), 64, 0, s>>>(a, d);
    const long long want = blocks_for((long long)a.n_el * 32, 128);
    apply_kernel<true><<<(int)(want < 148 * 16 ? want : 148 * 16), 128, 0, s>>>(a, d);
}
void launch_beam_alpha_commit(const EvalArgs& a, double* alpha_i, void* stream) {
    if (a.n_el == 0) return;
    beam::alpha_commit_kernel<<<blocks_for((long long)a.n_el * beam::NGP, 128), 128, 0, (cudaStream_t)stream>>>(a, alpha_i);
}
void launch_shell_alpha_commit(const EvalArgs& a, double* alpha_i, void* stream) {
    if (a.n_el == 0) return;
    shell::alpha_commit_kernel<<<blocks_for((long long)a.n_el * shell::NGP, 128), 128, 0, (cudaStream_t)stream>>>(a, alpha_i);
}
void launch_update_dyn(const DynArgs& d, const int* gls, const double* disp, double* vel, double* accel, int n_nodes,
                       const int* mixed, const int* start, int n_mixed, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    update_dyn_kernel<<<blocks_for(n_nodes, 128), 128, 0, s>>>(d, gls, disp, vel, accel, n_nodes);
    if (n_mixed > 0) update_dyn_replay_kernel<<<blocks_for(n_mixed, 64), 64, 0, s>>>(d, gls, disp, vel, accel, mixed, start, n_mixed);
}

int configure_dynamics() {
    unsigned short tab[SHELL_ARENA];
    for (int i = 0; i < SHELL_ARENA; i++) tab[i] = 0xFFFF;
    for (int a = 0; a < 9; a++)
        for (int b = 0; b < 9; b++) {
            if (a > b && b < 6) continue;                // held as the transpose of (b, a)
            const int o = shell_stored_offset(a, b);
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) tab[o + 3 * i + j] = (unsigned short)(a | (b << 4) | (i << 8) | (j << 10));
        }
    cudaError_t e = cudaMemcpyToSymbol(g_shell_decode, tab, sizeof(tab));
    if (e != cudaSuccess) return (int)e;
    unsigned short tb[BEAM_ARENA];
    for (int i = 0; i < BEAM_ARENA; i++) tb[i] = 0xFFFF;
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            if (!beam_is_stored(a, b)) continue;         // held as the transpose of (b, a)
            const int o = beam_stored_offset(a, b);
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) tb[o + 3 * i + j] = (unsigned short)(a | (b << 4) | (i << 8) | (j << 10));
        }
    return (int)cudaMemcpyToSymbol(g_beam_decode, tb, sizeof(tb));
}

} // namespace gfa
