// Host-side mirror (C++, the reference's own language) of the slice of GIRAFFE's
// Database / IO / Solution API that drives the assembly path, sitting on top of
// the C-ABI in include/gfa.h.  Method names, argument meaning and call order are
// the reference's (src/Solution.h:22-43, src/Static.cpp:161-212) so that a
// maintainer can read this next to the reference and the parity tests read like
// reference code:
//
//   host.ReadFile(path)          IO::ReadFile            (IO.cpp:193-770, in-scope blocks of SURVEY.md App. B)
//   host.PreCalc()               Database::PreCalc       (Database.cpp:704-759)  -> gfa_create
//   host.DOFsActive()            Solution::DOFsActive    (Solution.cpp:121-224)
//   host.SetGlobalDOFs()         Solution::SetGlobalDOFs (Solution.cpp:40-118)
//   host.SetGlobalSize()         Solution::SetGlobalSize (Solution.cpp:577-654)  -> gfa_set_dofs
//   host.Clear()                 Solution::Clear         (Solution.cpp:833-848)
//   host.MountLocal()            Solution::MountLocal + MountElementLoads + MountGlobal + MountSparse
//                                for the device element types                    -> gfa_assemble
//   host.MountLoads()            Solution::MountLoads    (NodalLoad::Mount, NodalLoad.cpp:322-401; ShellLoad::Mount ->
//                                Shell_1::MountShellSpecialLoads, Shell_1.cpp:1392-1467; host side)
//   host.UpdateDisps(x)          Solution::UpdateDisps   (Solution.cpp:390-402)
//   host.SaveConfiguration()     Solution::SaveConfiguration (Solution.cpp:426-454) -> gfa_commit_state
//   host.GetGaussPointResults()  what WriteResults / WriteMonitor read from the elements   -> gfa_gauss_point_results
// Dynamic::Solve (src/Dynamic.cpp:255-340), when the first solution step of the input is `Dynamic`:
//   host.CalculateNewmarkCoeff(dt)  Dynamic::CalculateNewmarkCoeff (Dynamic.cpp:582-590)
//   host.UpdateDyn()                Dynamic::UpdateDyn            (Dynamic.cpp:480-556)   -> gfa_update_dyn
//   host.MountLocalDynamic(update)  MountLocal + MountElementLoads + MountMass + MountDamping(update) + MountDyn
//                                   + MountGlobal + MountSparse                            -> gfa_assemble_dynamic
//
// Errors follow the reference: Read* return false on a malformed block, the
// rest report through last_error() and leave the state untouched.
#pragma once
#include <string>
#include <vector>

#include "../../include/gfa.h"

struct GfaNodalLoad {              // NodalLoad with a numeric table (NodalLoad.h, Table.h)
    int node_set = 0, cs = 0;
    bool follower = false;         // NodalFollowerLoad (NodalFollowerLoad.h): same table, loads follow the node's rotation
    std::vector<double> table;     // rows: time FX FY FZ MX MY MZ
    double GetValueAt(double t, int column) const;   // Table::GetValueAt, linear interpolation
};

struct GfaShellLoad {              // ShellLoad with a numeric table (ShellLoad.h): follower pressure on an ElementSet
    int element_set = 0;
    bool area_update = false;
    std::vector<double> table;     // rows: time pressure
    double GetValueAt(double t) const;
};

struct GfaPipeLoad {               // PipeLoad with a numeric table (PipeLoad.h): internal pressure on an ElementSet of Pipe_1
    int element_set = 0;
    std::vector<double> table;     // rows: time P0I P0E RhoI RhoE
    double GetValueAt(double t, int column) const;
};

class GfaHost {
public:
    ~GfaHost();

    // ---- Database --------------------------------------------------------
    std::vector<double> ref_coordinates;          // [n][3]   Node::ref_coordinates
    std::vector<double> displacements;            // [n][6]   Node::displacements
    std::vector<int> constraints;                 // [n] 6-bit mask, Node::constraints
    std::vector<int> GLs;                         // [n][6]   Node::GLs
    std::vector<double> hooke;                    // [m][3]   E nu rho
    std::vector<double> section_defs;             // [s][3]   kind(0 Rectangle,1 Tube) a b
    std::vector<double> sections;                 // [s][6]   A I11 I22 I12 I33 It (after PreCalc)
    std::vector<double> shell_thickness;
    std::vector<double> pipe_sections;            // [p][11]  EA EI GJ GA Rho CDt CDn CAt CAn De Di (PipeSection.h:13-23)
    std::vector<double> cs_defs;                  // [c][6]   E1 E3 as read
    std::vector<double> cs;                       // [c][9]   E1 E2 E3 normalised
    std::vector<int> elem_type, elem_material, elem_section, elem_cs, elem_node_ptr, elem_nodes;
    std::vector<double> pretension;
    std::vector<std::vector<int> > node_sets;
    struct NodalConstraint { int node_set; int mask; };
    std::vector<NodalConstraint> nodal_constraints;
    std::vector<GfaNodalLoad> loads;
    std::vector<std::vector<int> > element_sets;  // ElementSet::el_list, 1-based (ElementSet.h)
    std::vector<GfaShellLoad> shell_loads;        // ShellLoad (host contributor: Shell_1::MountShellSpecialLoads)
    std::vector<double> copy_cache;               // Node::copy_coordinates [n][6], fetched for NodalFollowerLoad, dropped at SaveConfiguration
    std::vector<GfaPipeLoad> pipe_loads;          // PipeLoad (Pipe_1::MountPipeSpecialLoads, evaluated on the device)
    bool g_exist = false;
    double G[3] = { 0, 0, 0 };
    double end_time = 1.0, time_step = 1.0;       // first solution step (Static.cpp:43-130, Dynamic.cpp:65-222)
    bool dynamic = false;                         // the first solution step is `Dynamic`
    double alpha = 0.0, beta = 0.0;               // Dynamic::alpha, beta (RayleighDamping)
    int update = 0;                               // Dynamic::update
    double beta_new = 0.3, gamma_new = 0.5;       // NewmarkCoefficients (defaults of Dynamic::Dynamic, Dynamic.cpp:40-41)
    double a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0;
    int n_GL_free = 0, n_GL_fixed = 0;
    double last_converged_time = 0.0, current_time_step = 0.0;

    int number_nodes() const { return (int)(ref_coordinates.size() / 3); }
    int number_elements() const { return (int)elem_type.size(); }

    // ---- IO ----------------------------------------------------------------
    bool ReadFile(const char* path);

    // ---- Database / Solution steps ------------------------------------------
    bool PreCalc(int device = 0);
    void DOFsActive();
    void SetGlobalDOFs();
    bool SetGlobalSize();
    void Clear() {}                               // every device slot is rewritten by MountLocal()
    bool MountLocal();                            // device: Mount + MountElementLoads + MountGlobal + MountSparse
    void MountElementLoads() {}                   // folded into MountLocal()
    void MountGlobal() {}                         // folded into MountLocal()
    void MountSparse() {}                         // folded into MountLocal()
    bool MountLoads();                            // host NodalLoad / ShellLoad -> gfa_add_host_triplets / gfa_add_host_vector
    void UpdateDisps(const double* x_A);          // displacements[j] += x(GL-1) for free active DOFs
    bool SaveConfiguration();
    // Dynamic::Solve
    void CalculateNewmarkCoeff(double dt);
    bool SetKinematics(const double* vel, const double* accel, const double* copy_vel, const double* copy_accel);   // InitialCondition
    bool UpdateDyn();                             // device: vel / accel of every free DOF from `displacements`
    bool MountLocalDynamic(bool update_rayleigh); // device: MountLocal .. MountMass, MountDamping(update_rayleigh), MountDyn .. MountSparse
    void MountMass() {}                           // folded into MountLocalDynamic()
    void MountDamping(bool) {}
    void MountDyn() {}
    bool GetKinematics(std::vector<double>& vel, std::vector<double>& accel);

    // ---- global system -----------------------------------------------------------
    bool GetCSR(int which, std::vector<int>& outer, std::vector<int>& inner, std::vector<double>& values);
    bool GetVector(int which, std::vector<double>& v);
    bool GetGaussPointResults(int element_type, std::vector<double>& out);   // strain energy + per-point strains / resultants
    double LoadFactor() const;                    // BoolTable::GetLinearFactorAtCurrentTime for a first step
    const std::string& last_error() const { return err; }
    gfa_t* handle() { return h; }

private:
    gfa_t* h = nullptr;
    std::string err;
    std::vector<unsigned char> active_GL;         // [n][6] Node::active_GL
    bool fail(const std::string& m) { err = m; return false; }
    void CollectLoadPattern(std::vector<int>& m, std::vector<int>& r, std::vector<int>& c);
};
