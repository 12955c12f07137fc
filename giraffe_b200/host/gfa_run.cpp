// Small driver over GfaHost: reads a GIRAFFE .inp (in-scope subset), runs the
// set-up steps in the reference's order and one Newton-iteration assembly at
// the first time increment, and prints a JSON summary.
//   gfa_run --parse-only model.inp      (no GPU needed: reader + DOF numbering)
//   gfa_run model.inp                   (needs a CUDA device)
#include <cmath>
#include <cstdio>
#include <cstring>

#include "GfaHost.h"

int main(int argc, char** argv) {
    bool parse_only = false;
    const char* path = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--parse-only")) parse_only = true;
        else path = argv[i];
    }
    if (!path) { fprintf(stderr, "usage: gfa_run [--parse-only] model.inp\n"); return 2; }
    GfaHost host;
    if (!host.ReadFile(path)) { fprintf(stderr, "ReadFile: %s\n", host.last_error().c_str()); return 1; }
    host.DOFsActive();
    host.SetGlobalDOFs();
    if (parse_only) {
        printf("{\"nodes\": %d, \"elements\": %d, \"n_GL_free\": %d, \"n_GL_fixed\": %d, \"loads\": %zu, \"node_sets\": %zu, \"time_step\": %.17g, \"end_time\": %.17g, "
               "\"dynamic\": %d, \"alpha\": %.17g, \"beta\": %.17g, \"update\": %d, \"beta_new\": %.17g, \"gamma_new\": %.17g, \"shell_loads\": %zu, \"element_sets\": %zu}\n",
               host.number_nodes(), host.number_elements(), host.n_GL_free, host.n_GL_fixed, host.loads.size(), host.node_sets.size(), host.time_step, host.end_time,
               host.dynamic ? 1 : 0, host.alpha, host.beta, host.update, host.beta_new, host.gamma_new, host.shell_loads.size(), host.element_sets.size());
        return 0;
    }
    if (!host.PreCalc(0)) { fprintf(stderr, "PreCalc: %s\n", host.last_error().c_str()); return 1; }
    if (!host.SetGlobalSize()) { fprintf(stderr, "SetGlobalSize: %s\n", host.last_error().c_str()); return 1; }
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
    printf("{\"nodes\": %d, \"elements\": %d, \"n_GL_free\": %d, \"n_GL_fixed\": %d, \"nnz_AA\": %zu, \"sum_AA\": %.17g, \"max_AA\": %.17g, \"max_P_A\": %.17g}\n",
           host.number_nodes(), host.number_elements(), host.n_GL_free, host.n_GL_fixed, val.size(), sum, amax, pmax);
    return 0;
}
