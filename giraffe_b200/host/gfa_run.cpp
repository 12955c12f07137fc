// Small driver over GfaHost: reads a GIRAFFE .inp (in-scope subset), runs the
// set-up steps in the reference's order and one Newton-iteration assembly at
// the first time increment, and prints a JSON summary.
//   gfa_run --parse-only model.inp      (no GPU needed: reader + DOF numbering)
//   gfa_run model.inp                   (needs a CUDA device)
//   gfa_run --solve model.inp           Static::Solve's loop (Static.cpp:161-236) to end_time: per increment a fixed
//                                       number of Newton iterations -- assembly on the device, MountLoads, the sign
//                                       flip and residual norms of gfa_residual, a dense LU solve on the host (the
//                                       sparse solve is not part of the path; small models only), UpdateDisps,
//                                       SaveConfiguration -- and prints the final copy_coordinates
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "GfaHost.h"

// dense LU with partial pivoting, in place; returns false for a singular matrix
static bool lu_solve(std::vector<double>& A, std::vector<double>& b, int n) {
    for (int k = 0; k < n; k++) {
        int piv = k;
        for (int i = k + 1; i < n; i++) if (fabs(A[(size_t)i * n + k]) > fabs(A[(size_t)piv * n + k])) piv = i;
        if (A[(size_t)piv * n + k] == 0.0) return false;
        if (piv != k) { for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)piv * n + j]); std::swap(b[k], b[piv]); }
        for (int i = k + 1; i < n; i++) {
            const double f = A[(size_t)i * n + k] / A[(size_t)k * n + k];
            if (f == 0.0) continue;
            for (int j = k; j < n; j++) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
            b[i] -= f * b[k];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int j = i + 1; j < n; j++) s -= A[(size_t)i * n + j] * b[j];
        b[i] = s / A[(size_t)i * n + i];
    }
    return true;
}

static int solve_static(GfaHost& host, int iterations) {
    const int n = host.n_GL_free;
    if (n > 4000) { fprintf(stderr, "--solve uses a dense LU: %d free DOFs are too many\n", n); return 1; }
    double t = 0.0;
    int increments = 0;
    double last_dx = 0.0, last_res = 0.0;
    while (t < host.end_time - 1e-12 * host.end_time) {
        host.last_converged_time = t;
        host.current_time_step = std::min(host.time_step, host.end_time - t);
        for (int it = 0; it < iterations; it++) {
            host.Clear();
            if (!host.MountLocal()) { fprintf(stderr, "MountLocal: %s\n", host.last_error().c_str()); return 1; }
            host.MountElementLoads();
            if (!host.MountLoads()) { fprintf(stderr, "MountLoads: %s\n", host.last_error().c_str()); return 1; }
            host.MountGlobal();
            gfa_norms_t norms;
            if (gfa_residual(host.handle(), nullptr, &norms) != GFA_OK) { fprintf(stderr, "gfa_residual: %s\n", gfa_last_error()); return 1; }   // P_A = -P_A (Static.cpp:210)
            host.MountSparse();
            std::vector<int> outer, inner;
            std::vector<double> val, rhs;
            if (!host.GetCSR(GFA_AA, outer, inner, val) || !host.GetVector(GFA_P_A, rhs)) { fprintf(stderr, "%s\n", host.last_error().c_str()); return 1; }
            std::vector<double> A((size_t)n * n, 0.0);
            for (int r = 0; r < n; r++) for (int k = outer[r]; k < outer[r + 1]; k++) A[(size_t)r * n + inner[k]] = val[k];
            if (!lu_solve(A, rhs, n)) { fprintf(stderr, "singular tangent\n"); return 1; }
            host.UpdateDisps(rhs.data());
            last_dx = 0.0; for (double v : rhs) last_dx = fmax(last_dx, fabs(v));
            last_res = fmax(norms.max_force, norms.max_moment);
        }
        if (!host.SaveConfiguration()) { fprintf(stderr, "SaveConfiguration: %s\n", host.last_error().c_str()); return 1; }
        t += host.current_time_step;
        increments++;
    }
    std::vector<double> copy(6 * (size_t)host.number_nodes());
    if (gfa_copy_coordinates(host.handle(), copy.data()) != GFA_OK) { fprintf(stderr, "%s\n", gfa_last_error()); return 1; }
    printf("{\"increments\": %d, \"iterations_per_increment\": %d, \"last_max_dx\": %.6e, \"last_max_residual\": %.6e, \"copy_coordinates\": [", increments, iterations, last_dx, last_res);
    for (size_t i = 0; i < copy.size(); i++) printf("%s%.17g", i ? ", " : "", copy[i]);
    printf("]}\n");
    return 0;
}

int main(int argc, char** argv) {
    bool parse_only = false, solve = false;
    const char* path = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--parse-only")) parse_only = true;
        else if (!strcmp(argv[i], "--solve")) solve = true;
        else path = argv[i];
    }
    if (!path) { fprintf(stderr, "usage: gfa_run [--parse-only] model.inp\n"); return 2; }
    GfaHost host;
    if (!host.ReadFile(path)) { fprintf(stderr, "ReadFile: %s\n", host.last_error().c_str()); return 1; }
    host.DOFsActive();
    host.SetGlobalDOFs();
    if (parse_only) {
        printf("{\"nodes\": %d, \"elements\": %d, \"n_GL_free\": %d, \"n_GL_fixed\": %d, \"loads\": %zu, \"node_sets\": %zu, \"time_step\": %.17g, \"end_time\": %.17g, "
               "\"dynamic\": %d, \"alpha\": %.17g, \"beta\": %.17g, \"update\": %d, \"beta_new\": %.17g, \"gamma_new\": %.17g, \"shell_loads\": %zu, \"pipe_loads\": %zu, \"element_sets\": %zu}\n",
               host.number_nodes(), host.number_elements(), host.n_GL_free, host.n_GL_fixed, host.loads.size(), host.node_sets.size(), host.time_step, host.end_time,
               host.dynamic ? 1 : 0, host.alpha, host.beta, host.update, host.beta_new, host.gamma_new, host.shell_loads.size(), host.pipe_loads.size(), host.element_sets.size());
        return 0;
    }
    if (!host.PreCalc(0)) { fprintf(stderr, "PreCalc: %s\n", host.last_error().c_str()); return 1; }
    if (!host.SetGlobalSize()) { fprintf(stderr, "SetGlobalSize: %s\n", host.last_error().c_str()); return 1; }
    if (solve) {
        if (host.dynamic) { fprintf(stderr, "--solve runs Static solution steps\n"); return 2; }
        return solve_static(host, 8);
    }
    host.last_converged_time = 0.0;
    host.current_time_step = host.time_step;
    if (host.dynamic) {
        // first Newton iteration of Dynamic::Solve (Dynamic.cpp:303-340) from rest
        host.CalculateNewmarkCoeff(host.time_step);
        host.Clear();
        if (!host.UpdateDyn()) { fprintf(stderr, "UpdateDyn: %s\n", host.last_error().c_str()); return 1; }
        if (!host.MountLocalDynamic(true)) { fprintf(stderr, "MountLocalDynamic: %s\n", host.last_error().c_str()); return 1; }
        host.MountElementLoads();
        if (!host.MountLoads()) { fprintf(stderr, "MountLoads: %s\n", host.last_error().c_str()); return 1; }
        host.MountMass(); host.MountDamping(true); host.MountDyn();
        host.MountGlobal();
        host.MountSparse();
    } else {
        // one Newton iteration of Static::Solve (Static.cpp:203-212)
        host.Clear();
        if (!host.MountLocal()) { fprintf(stderr, "MountLocal: %s\n", host.last_error().c_str()); return 1; }
        host.MountElementLoads();
        if (!host.MountLoads()) { fprintf(stderr, "MountLoads: %s\n", host.last_error().c_str()); return 1; }
        host.MountGlobal();
        host.MountSparse();
    }
    std::vector<int> outer, inner;
    std::vector<double> val, pa;
    if (!host.GetCSR(GFA_AA, outer, inner, val) || !host.GetVector(GFA_P_A, pa)) { fprintf(stderr, "%s\n", host.last_error().c_str()); return 1; }
    double sum = 0.0, amax = 0.0, pmax = 0.0;
    for (double v : val) { sum += v; amax = fmax(amax, fabs(v)); }
    for (double v : pa) pmax = fmax(pmax, fabs(v));
    printf("{\"nodes\": %d, \"elements\": %d, \"n_GL_free\": %d, \"n_GL_fixed\": %d, \"nnz_AA\": %zu, \"sum_AA\": %.17g, \"max_AA\": %.17g, \"max_P_A\": %.17g, \"shell_loads\": %zu, \"pipe_loads\": %zu}\n",
           host.number_nodes(), host.number_elements(), host.n_GL_free, host.n_GL_fixed, val.size(), sum, amax, pmax, host.shell_loads.size(), host.pipe_loads.size());
    return 0;
}
