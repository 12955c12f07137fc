/* TEST INFRASTRUCTURE (oracle) -- C interface of the builder's CPU restatement
 * of GIRAFFE's per-Newton-iteration element assembly.
 *
 * PARITY STATUS
 *   Beam_1, Shell_1, DOF numbering, triplet order, CSR build: pinned against
 *   the reference's own sources compiled here (oracle/_ref, see
 *   oracle/Makefile) by tests/test_oracle_vs_ref.py, and against the committed
 *   fixtures under tests/golden/ that the same build generated.
 *   Newmark dynamics (MountMass / MountDamping / MountDyn / UpdateDyn of Beam_1, Pipe_1, Shell_1): the
 *   AceGen-generated inertia code of the reference is restated from its formulation with a forward-mode
 *   tangent; pinned against the reference's own Dynamic object the same two ways
 *   (tests/test_oracle_vs_ref.py, tests/golden/dynamic_*.npz), agreement 2e-15.
 *   Solid_1: PARITY UNPINNED -- the reference's Solid_1::Mount/MountGlobal are
 *   empty bodies (reference Solid_1.cpp:148-176); the formulation restated
 *   here is builder-defined (8-node trilinear hexahedron, total-Lagrangian
 *   St.Venant-Kirchhoff with Hooke constants, 2x2x2 Gauss).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product never does.
 */
#ifndef GFA_ORACLE_H
#define GFA_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* model tables (same meaning as include/gfa.h gfa_model_t) */
int gfo_reset(void);
int gfo_set_nodes(int n, const double* xyz);
int gfo_set_materials(int n, const double* hooke3);
int gfo_set_sections(int n, const double* sec6);
int gfo_set_shell_sections(int n, const double* thickness);
int gfo_set_cs(int n, const double* e123);
int gfo_set_pipe_sections(int n, const double* v11);   /* EA EI GJ GA Rho CDt CDn CAt CAn De Di; element type 2 = Pipe_1 */
int gfo_set_elements(int n, const int* type, const int* mat, const int* sec, const int* cs,
                     const int* node_ptr, const int* nodes, const double* pretension);
int gfo_set_gravity(int on, double gx, double gy, double gz);
int gfo_set_constraint_mask(const int* mask_per_node);
int gfo_set_threads(int n);

/* PreCalc + DOFsActive + SetGlobalDOFs */
int gfo_precalc(void);
int gfo_n_free(void);
int gfo_n_fixed(void);
int gfo_get_gls(int* gls);

/* one assembly: Clear, MountLocal, MountElementLoads, MountGlobal, MountSparse.
 * gravity_factor = BoolTable::GetLinearFactorAtCurrentTime().
 * extra triplets (e.g. NodalLoad blocks) are pushed BEFORE the elements, as
 * MountLoads precedes MountGlobal (reference Static.cpp:207-208). */
int gfo_set_extra_triplets(int which, long n, const int* rows, const int* cols, const double* vals);
int gfo_assemble(const double* disp6, double gravity_factor, double* seconds5);

long gfo_triplet_count(int which);
int  gfo_csr_rows(int which);
int  gfo_csr_cols(int which);
long gfo_csr_nnz(int which);
int  gfo_csr_get(int which, int* outer, int* inner, double* val);
int  gfo_get_vectors(double* PA, double* IA, double* PB);
int  gfo_get_element(int e, double* K_rowmajor, double* P, double* energy);
int  gfo_get_state(int e, double* out);
/* Gauss-point results in the layout of gfa_gauss_point_results (include/gfa.h); returns doubles written */
int  gfo_get_results(int e, double* out);
int  gfo_commit(void);
int  gfo_get_copy_coordinates(double* c6);


/* Newmark dynamics (Dynamic.cpp:303-340): Beam_1, Shell_1 and Pipe_1 without ocean data (Solid_1 returns -7).
 * newmark6 = Dynamic::a1..a6; kinematics arrays are Node::vel / accel / copy_vel / copy_accel [n_nodes*6]
 * (NULL = leave / skip).  gfo_assemble_dynamic = Clear, MountLocal, MountElementLoads, MountMass,
 * MountDamping(update_rayleigh), MountDyn, MountGlobal, MountSparse.  gfo_commit also copies vel/accel
 * (Node.cpp:375-380) and updates alpha_i. */
int gfo_set_dynamic(const double* newmark6, double rayleigh_alpha, double rayleigh_beta);
int gfo_set_kinematics(const double* vel, const double* accel, const double* copy_vel, const double* copy_accel);
int gfo_get_kinematics(double* vel, double* accel, double* copy_vel, double* copy_accel);
int gfo_update_dyn(const double* disp6);      /* Dynamic::UpdateDyn */
int gfo_assemble_dynamic(const double* disp6, double gravity_factor, int update_rayleigh);
int gfo_get_alpha_i(int e, double* out);

#ifdef __cplusplus
}
#endif
#endif
