// 3x3 / 3-vector FP64 helpers kept in registers (row-major double[9]).
// These are the device-side counterparts of the reference's dense Matrix
// operations used inside Element::Mount (reference src/Matrix.cpp:168-365,
// 1793-2032), specialised to the 3x3 sub-structure of the element algebra.
#pragma once

#define GFA_DI __device__ __forceinline__

namespace gfa {

GFA_DI void m_zero(double* A) {
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = 0.0;
}
GFA_DI void m_copy(double* A, const double* B) {
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = B[i];
}
// C = A B
GFA_DI void mm(double* C, const double* A, const double* B) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A^T B
GFA_DI void mtm(double* C, const double* A, const double* B) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// C += A^T B
GFA_DI void mtm_acc(double* C, const double* A, const double* B) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] += A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// y = A x ; y = A^T x
GFA_DI void mv(double* y, const double* A, const double* x) {
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
GFA_DI void mtv(double* y, const double* A, const double* x) {
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
GFA_DI void m_transpose(double* T, const double* A) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * j + i];
}
GFA_DI double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
GFA_DI void cross3(double* c, const double* a, const double* b) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
GFA_DI double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
// skew(v): reference Matrix.cpp:1812-1832
GFA_DI void skew3(double* S, const double* v) {
    S[0] = 0.0;   S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2];  S[4] = 0.0;   S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0];  S[8] = 0.0;
}
// C = skew(v) B  (row i of C = (v x column)...); written out to avoid the zeros
GFA_DI void skew_mul(double* C, const double* v, const double* B) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        C[j]     = -v[2] * B[3 + j] + v[1] * B[6 + j];
        C[3 + j] =  v[2] * B[j]     - v[0] * B[6 + j];
        C[6 + j] = -v[1] * B[j]     + v[0] * B[3 + j];
    }
}

// Rodrigues-parameter rotation pieces at an integration point
// (reference Shell_1.cpp:1001-1009, Beam_1.cpp:748-754):
//   g = 4/(4+|a|^2), Xi = g (I + A/2), Qd = I + g (A + A^2/2), A = skew(a)
GFA_DI void rodrigues(const double* a, double& g, double* Qd, double* Xi) {
    const double al2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    // |a| enters only squared; the reference squares norm(a) again
    const double al = sqrt(al2);
    g = 4.0 / (4.0 + al * al);
    double A[9], AA[9];
    skew3(A, a);
    mm(AA, A, A);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        Qd[i] = id + g * (A[i] + 0.5 * AA[i]);
        Xi[i] = g * (id + 0.5 * A[i]);
    }
}
// dXi = -g/2 ((a.da) Xi - skew(da))   (Shell_1.cpp:1008, Beam_1.cpp:754)
GFA_DI void d_xi(double* dXi, const double* a, const double* da, double g, const double* Xi) {
    const double ad = dot3(a, da);
    double S[9];
    skew3(S, da);
#pragma unroll
    for (int i = 0; i < 9; i++) dXi[i] = (-0.5 * g) * (ad * Xi[i] - S[i]);
}

// V(x,t) = (h8 t - h4 x^t) (x) x + h2 skew(t)          (Matrix.cpp:1999-2012)
GFA_DI void v_op(double* V, const double* x, const double* t, double h) {
    const double h2 = 0.5 * h, h4 = -0.25 * h * h, h8 = -0.5 * h * h;
    double xt[3], w[3], S[9];
    cross3(xt, x, t);
#pragma unroll
    for (int i = 0; i < 3; i++) w[i] = h8 * t[i] - h4 * xt[i];
    skew3(S, t);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) V[3 * i + j] = w[i] * x[j] + h2 * S[3 * i + j];
}
// d_V(x,dx,t)                                           (Matrix.cpp:2014-2032)
GFA_DI void dv_op(double* V, const double* x, const double* dx, const double* t, double h) {
    const double h4 = -0.25 * h * h, h6 = 0.25 * h * h * h, h8 = -0.5 * h * h, h9 = 0.5 * h * h * h;
    const double xd = dot3(x, dx);
    double xt[3], dxt[3], w1[3], w2[3], w3[3], S[9];
    cross3(xt, x, t);
    cross3(dxt, dx, t);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        w1[i] = h9 * t[i] - h6 * xt[i];
        w2[i] = h8 * t[i] - h4 * xt[i];
        w3[i] = (-h4) * dxt[i];
    }
    skew3(S, t);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            V[3 * i + j] = xd * (w1[i] * x[j]) + w2[i] * dx[j] + w3[i] * x[j] + (h4 * xd) * S[3 * i + j];
}

// -------------------------------------------------------------------------
// "Strict" arithmetic: separately rounded multiply / add in a fixed order.
// The back-rotated strains eta = Q^T z' - e (and J - 1 in the shell's
// constitutive law) are differences of O(1) quantities; for strains of 1e-5
// one ulp of Q^T z' is 1e-11 of the strain, hence of the internal force.  To
// agree with the reference to 1e-12 the kinematic chain from the nodal
// displacements up to the stress resultants must round exactly like the
// reference's mul-then-add Matrix arithmetic (no FMA contraction, k-inner
// accumulation as in Matrix.cpp:218-250).  Everything downstream of the
// resultants (the bulk of the flops) uses FMA freely.
// -------------------------------------------------------------------------
GFA_DI double s_mul(double a, double b) { return __dmul_rn(a, b); }
GFA_DI double s_add(double a, double b) { return __dadd_rn(a, b); }
GFA_DI double s_sub(double a, double b) { return __dsub_rn(a, b); }
GFA_DI double s_dot3(const double* a, const double* b) {
    return s_add(s_add(s_mul(a[0], b[0]), s_mul(a[1], b[1])), s_mul(a[2], b[2]));
}
GFA_DI double s_norm3(const double* a) { return sqrt(s_dot3(a, a)); }
GFA_DI void s_cross3(double* c, const double* a, const double* b) {
    c[0] = s_sub(s_mul(a[1], b[2]), s_mul(a[2], b[1]));
    c[1] = s_sub(s_mul(a[2], b[0]), s_mul(a[0], b[2]));
    c[2] = s_sub(s_mul(a[0], b[1]), s_mul(a[1], b[0]));
}
GFA_DI void s_mv(double* y, const double* A, const double* x) {
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = s_add(s_add(s_mul(A[3 * i], x[0]), s_mul(A[3 * i + 1], x[1])), s_mul(A[3 * i + 2], x[2]));
}
GFA_DI void s_mtv(double* y, const double* A, const double* x) {
#pragma unroll
    for (int i = 0; i < 3; i++) y[i] = s_add(s_add(s_mul(A[i], x[0]), s_mul(A[3 + i], x[1])), s_mul(A[6 + i], x[2]));
}
GFA_DI void s_mm(double* C, const double* A, const double* B) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = s_add(s_add(s_mul(A[3 * i], B[j]), s_mul(A[3 * i + 1], B[3 + j])), s_mul(A[3 * i + 2], B[6 + j]));
}
// Rodrigues pieces with the reference's rounding (Shell_1.cpp:1001-1005)
GFA_DI void s_rodrigues(const double* a, double& g, double* Qd, double* Xi) {
    const double al = s_norm3(a);
    g = 4.0 / s_add(4.0, s_mul(al, al));
    double A[9], AA[9];
    skew3(A, a);
    s_mm(AA, A, A);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        Qd[i] = s_add(id, s_mul(g, s_add(A[i], s_mul(0.5, AA[i]))));
        Xi[i] = s_mul(g, s_add(id, s_mul(0.5, A[i])));
    }
}

} // namespace gfa
