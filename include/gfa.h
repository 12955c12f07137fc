/* gfa.h -- C-ABI of the B200 element-assembly library (libgfa.so).
 *
 * Drop-in boundary for GIRAFFE's per-Newton-iteration element assembly:
 *   Solution::MountLocal -> MountElementLoads -> MountGlobal -> MountSparse
 *   (reference src/Solution.cpp:227-265, 322-349, 851-863), for Beam_1,
 *   Shell_1 and Solid_1, as driven by Static::Solve (src/Static.cpp:161-163,
 *   203-212) and Dynamic::Solve (src/Dynamic.cpp:323-340).
 *
 * The reference has no FFI; the seam is the virtual Element API
 * (src/Element.h:61-76) plus the global system in Database
 * (src/Database.h:394-405).  Each entry point below names the reference
 * interface it replaces.  Plain pointers and sizes only; every function
 * returns 0 on success or a negative GFA_E* code, never throws, and leaves a
 * message retrievable with gfa_last_error().  A handle is bound to one CUDA
 * device and is not thread-safe (the reference calls this path from its
 * single main thread).
 *
 * There is no CPU fallback: without a CUDA device gfa_create fails with
 * GFA_ENODEVICE.
 */
/* Limits of the slot map (gfa_set_dofs refuses what exceeds them with GFA_EUNSUPPORTED; none is reached by the
 * BASELINE configs):
 *   - DOF numbering: the reference's own (src/Solution.cpp:53-72) -- free ids ascending with (node, DOF), the free
 *     DOFs of one node's translations (rotations) consecutive; DOFs beyond the node table (Lagrange multipliers,
 *     super nodes) may take any ids that leave those groups intact and enter through gfa_set_dofs' extra positions;
 *   - AA non-zeros per rank < 2^31 (Eigen / PARDISO 32-bit indices; the reference's own `int size_AA` overflows
 *     at ~2.9 M shells, src/Solution.cpp:579);
 *   - a CSR row of an element-only group-node holds at most 65 535 entries (16-bit row stride of a patch);
 *   - at most 255 elements of one rank share a pair of 3-DOF groups (8-bit source count of a patch);
 *   - element arena of one rank < 2^32 doubles (32-bit block offsets: 34 GB, ~9.6 M shells or 13 M solids);
 *   - at most 64 ranks (interface bookkeeping keeps rank sets in 64 bits). */
#ifndef GFA_H
#define GFA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFA_VERSION 100

/* element type ids = the reference's (src/Element.h:8-15) */
#define GFA_BEAM_1  1
#define GFA_PIPE_1  2   /* evaluated by the Beam_1 kernel: Pipe_1::Mount is Beam_1::Mount (src/Pipe_1.cpp:836-974) */
#define GFA_SHELL_1 3
#define GFA_SOLID_1 7

/* global matrices (src/Database.h:394-397) */
#define GFA_AA 0   /* free  x free  : db.global_stiffness_AA */
#define GFA_AB 1   /* free  x fixed : db.global_stiffness_AB */
#define GFA_BA 2   /* fixed x free  : db.global_stiffness_BA */
#define GFA_BB 3   /* fixed x fixed : db.global_stiffness_BB */

/* global vectors (src/Database.h:400-403) */
#define GFA_P_A 0  /* db.global_P_A, n_free  (sign as accumulated by MountGlobal: +P) */
#define GFA_I_A 1  /* db.global_I_A, n_free  */
#define GFA_P_B 2  /* db.global_P_B, n_fixed */

#define GFA_OK            0
#define GFA_EINVAL       -1   /* bad argument / inconsistent model        */
#define GFA_ENODEVICE    -2   /* no usable CUDA device                    */
#define GFA_ECUDA        -3   /* CUDA runtime error (see gfa_last_error)  */
#define GFA_ESTATE       -4   /* call out of order (e.g. assemble before set_dofs) */
#define GFA_EPATTERN     -5   /* host triplet outside the registered pattern */
#define GFA_ENOMEM       -6
#define GFA_EUNSUPPORTED -7

typedef struct gfa_handle gfa_t;

/* Model tables = what Database owns for the in-scope entities after
 * IO::ReadFile and Database::PreCalc (src/Database.cpp:704-759).  All arrays
 * are host pointers, copied by gfa_create; ids are 1-based as in the .inp. */
typedef struct gfa_model {
    int32_t n_nodes;
    const double* ref_coordinates;   /* [n_nodes*3]  Node::ref_coordinates[0..2] (src/Node.h:17) */
    const double* copy_coordinates;  /* [n_nodes*6] or NULL (= ref, zero rotation) Node::copy_coordinates */

    int32_t n_materials;
    const double* hooke;             /* [n_materials*3] E, nu, rho (src/Hooke.h:9, Material.h:11) */

    int32_t n_sections;
    const double* sections;          /* [n_sections*6] A I11 I22 I12 I33 It after Section::PreCalc (src/Section.h:14) */

    int32_t n_shell_sections;
    const double* shell_thickness;   /* [n_shell_sections] ShellSection::thickness (src/ShellSection.h:14) */

    int32_t n_cs;
    const double* cs;                /* [n_cs*9] E1,E2,E3 normalised (src/CoordinateSystem.h:13-17) */

    int32_t n_elements;
    const int32_t* elem_type;        /* [n_elements] GFA_BEAM_1 / GFA_PIPE_1 / GFA_SHELL_1 / GFA_SOLID_1 */
    const int32_t* elem_material;    /* Element::material */
    const int32_t* elem_section;     /* Element::section (beam: Sections id, pipe: PipeSections id, shell: ShellSections id, solid: unused) */
    const int32_t* elem_cs;          /* Element::cs (beam; shells with homogeneous sections ignore it) */
    const int32_t* elem_node_ptr;    /* [n_elements+1] offsets into elem_nodes */
    const int32_t* elem_nodes;       /* Element::nodes, 1-based */
    const double*  beam_pretension;  /* [n_elements] Beam_1::T0 or NULL */

    int32_t gravity_on;              /* Environment::g_exist */
    double  gravity[3];              /* Environment::G */

    /* mesh partition for multi-GPU runs: this handle evaluates only the
     * elements e with part_begin <= rank-local index < part_end of EACH type
     * is derived from (rank, world); single-GPU callers pass 0 and 1. */
    int32_t part_rank;
    int32_t part_world;

    int32_t n_pipe_sections;
    const double* pipe_sections;     /* [n_pipe_sections*11] EA EI GJ GA Rho CDt CDn CAt CAn De Di (src/PipeSection.h:13-23);
                                      * Pipe_1 has no material: elem_material is ignored for it */
} gfa_model_t;

/* Per-iteration inputs (src/Static.cpp:200-212; src/Solution.cpp:390-402). */
typedef struct gfa_step {
    const double* displacements;     /* [n_nodes*6] Node::displacements of every node, node-major; NULL = keep the
                                      * device copy (after gfa_update_displacements) */
    int32_t displacements_on_device; /* 0: host pointer (copied H2D inside the call); 1: device pointer */
    double  gravity_factor;          /* BoolTable::GetLinearFactorAtCurrentTime() of Environment::bool_g */
} gfa_step_t;

const char* gfa_last_error(void);
int gfa_device_count(void);

/* PreCalc of every element (src/Database.cpp:713-714): builds device tables,
 * element constants and the initial Gauss-point state (Shell_1.cpp:2357-2362,
 * LagrangeSave.cpp:41-50, Beam_1.cpp:616-619). */
int gfa_create(const gfa_model_t* model, int device, gfa_t** out);
int gfa_destroy(gfa_t* h);

/* DOFsActive + SetGlobalDOFs done by the library from the per-node
 * constraint masks (bit k = Node::constraints[k]); writes Node::GLs
 * (src/Solution.cpp:40-118,121-224).  Optional helper: hosts that number DOFs
 * themselves skip it. */
int gfa_number_dofs(gfa_t* h, const int32_t* constraint_mask, int32_t* GLs_out,
                    int32_t* n_free, int32_t* n_fixed);

/* SetGlobalSize (src/Solution.cpp:577-654): fixes the DOF map for a solution
 * step and builds the CSR patterns of AA/AB/BA/BB (union of all element
 * blocks, explicit zeros kept, columns sorted -- what setFromTriplets yields,
 * src/SparseMatrix.cpp:67-71) plus the element->slot maps.  `extra_*` lists
 * positions other host contributors will push (e.g. NodalLoad 3x3 blocks,
 * src/NodalLoad.cpp:384-397) so that they are part of the pattern. */
int gfa_set_dofs(gfa_t* h, const int32_t* GLs /* [n_nodes*6] */, int32_t n_free, int32_t n_fixed,
                 int64_t n_extra, const int32_t* extra_matrix, const int32_t* extra_rows, const int32_t* extra_cols);

/* CSR pattern, Eigen row-major layout: outer[rows+1], inner[nnz] (0-based). */
int gfa_csr_dims(gfa_t* h, int which, int32_t* rows, int32_t* cols, int64_t* nnz);
int gfa_csr_pattern(gfa_t* h, int which, int32_t* outer, int32_t* inner);

/* One Newton-iteration assembly: Clear + MountLocal + MountElementLoads +
 * MountGlobal + MountSparse for the supported element types.  Returns after
 * the results are complete on the device (stream-synchronised). */
int gfa_assemble(gfa_t* h, const gfa_step_t* step);
/* The same work, only ENQUEUED on gfa_stream(): returns at once so that the caller can queue what follows
 * (interface pack / NCCL / unpack in a multi-GPU run, its own consumers) behind it without leaving the GPU idle
 * while the host catches up.  step->displacements must be a device pointer or NULL.  Every read entry point
 * (gfa_csr_values, gfa_vector, gfa_element_block, gfa_last_timing ...) waits for the stream. */
int gfa_assemble_enqueue(gfa_t* h, const gfa_step_t* step);

/* Contributions of host-side contributors (loads, joints, contacts, element
 * types without a kernel), summed into existing slots AFTER the elements.
 * Replaces their SparseMatrix::setValue calls (src/SparseMatrix.cpp:58-66). */
int gfa_add_host_triplets(gfa_t* h, int which, int64_t n, const int32_t* rows, const int32_t* cols, const double* vals);
int gfa_add_host_vector(gfa_t* h, int which_vector, int64_t n, const int32_t* index, const double* vals);

/* ShellLoad follower pressure on the device (src/ShellLoad.cpp:133-148 -> Shell_1::MountShellSpecialLoads,
 * src/Shell_1.cpp:1392-1467): the one Load of the reference that is element arithmetic -- a 6-point integration
 * over the current configuration of every shell of an element set, with a non-symmetric load stiffness on the u-u
 * blocks.  gfa_set_shell_loads registers the loads after gfa_set_dofs (a new gfa_set_dofs drops them): load l
 * applies to the Shell_1 elements load_elements[load_ptr[l] .. load_ptr[l+1]) (0-based element ids; elements of
 * other ranks' partitions are skipped), area_update[l] = ShellLoad::area_update.  gfa_apply_shell_loads, called
 * after gfa_assemble (MountLoads runs after the elements are mounted), evaluates them for the displacements of that
 * assembly and the committed configuration with pressures[l] = ShellLoad::GetValueAt(time) and adds stiffness and
 * load vector into the CSR values and P_A / I_A / P_B, contributions summed in registration order. */
int gfa_set_shell_loads(gfa_t* h, int32_t n_loads, const int32_t* load_ptr, const int32_t* load_elements, const int32_t* area_update);
int gfa_apply_shell_loads(gfa_t* h, const double* pressures /* [n_loads] */);

/* ---- PipeLoad internal pressure on the device ("Morison/pressure loads ... become a second
 * kernel"; src/PipeLoad.cpp:117-133 -> Pipe_1::MountPipeSpecialLoads, src/Pipe_1.cpp:1443-1494).  Register the
 * loads once per DOF map: load l acts on load_elements[load_ptr[l] .. load_ptr[l+1]) (0-based global element
 * indices, all GFA_PIPE_1; the ElementSet of the reference, src/PipeLoad.cpp:91-106).  After every gfa_assemble,
 * gfa_apply_pipe_loads(h, p0i) takes P0I = Load::GetValueAt(last_converged_time + current_time_step, 0) of every
 * load and folds the pressure's end-cap / curvature force and its non-symmetric u-alpha, alpha-u, alpha-alpha load
 * stiffness into the CSR values and P_A / I_A / P_B, at the displacements of that assembly.  P0E, RhoI and RhoE of
 * the table are read by the reference and not used (src/Pipe_1.cpp:1446-1449); neither are they here.  Buoyancy and
 * the Morison sea-current loads (src/Pipe_1.cpp:1278-1441: they need Environment::OceanData) stay host
 * contributors. */
int gfa_set_pipe_loads(gfa_t* h, int32_t n_loads, const int32_t* load_ptr, const int32_t* load_elements);
int gfa_apply_pipe_loads(gfa_t* h, const double* p0i /* [n_loads] host */);

/* Results.  Device pointers stay valid until the next gfa_set_dofs/destroy. */
int gfa_csr_values(gfa_t* h, int which, double* host_out /* [nnz] */);
int gfa_csr_values_device(gfa_t* h, int which, double** dev_ptr);
int gfa_vector(gfa_t* h, int which_vector, double* host_out);
int gfa_vector_device(gfa_t* h, int which_vector, double** dev_ptr);

/* Element block after Mount + MountElementLoads, for inspection:
 * K row-major nDOF x nDOF in the element's local DOF order
 * (Shell_1.cpp:1523-1557, Beam_1.cpp:1439-1444), P = P_loading. */
int gfa_element_block(gfa_t* h, int32_t element /* 0-based */, double* K, double* P);

/* Post-convergence: Node::SaveConfiguration + Element::SaveLagrange for the
 * displacements of the LAST gfa_assemble call (src/Solution.cpp:426-454),
 * then the increments are zero for the next time step (src/Static.cpp:191).
 * Assembly itself never changes committed state, so a diverged increment
 * needs no rollback call. */
int gfa_commit_state(gfa_t* h);

/* Committed Gauss-point state of one element:
 *   Shell_1: 3 x [Q_i(9,row-major) z_x1_i(3) z_x2_i(3) kappa_r1_i(3) kappa_r2_i(3)]  (src/Shell_1.h:117-127)
 *   Beam_1 : 2 x [Q_i(9) dz_i(3) kappa_i_ref(3)]                                     (src/LagrangeSave.h:11-17)
 * returns the number of doubles written. */
int gfa_element_state(gfa_t* h, int32_t element, double* out);

/* Result read-back (on demand, sampled steps only): the Gauss-point quantities
 * Element::Mount leaves in its members for WriteResults / WriteVTK_XMLBase /
 * WriteMonitor (src/Shell_1.cpp:624-707, src/Beam_1.cpp:444-497) and
 * Element::strain_energy (src/Element.h:47, summed by src/Monitor.cpp:494),
 * evaluated for the displacements of the LAST gfa_assemble call and the
 * committed state.  One record per element of `element_type` in this rank's
 * partition, ascending element order:
 *   GFA_SHELL_1 (73): strain_energy, then per Gauss point g = 0..2
 *                     eta_r1 eta_r2 kappa_r1 kappa_r2 n_r1 n_r2 m_r1 m_r2 (3 each; src/Shell_1.cpp:1017-1020,
 *                     1146-1149, m_r*(2) = stiff_drill * kappa_r*(2) as at :1214-1217)
 *   GFA_BEAM_1  (25): strain_energy, then per point g = 0..1 epsilon_r(6) sigma_r(6) (src/Beam_1.cpp:781-794, 830)
 *   GFA_PIPE_1  (25): the same record; strain_energy is 0 as in the reference (Pipe_1::Mount never adds to it)
 * gfa_results_stride returns the record length (0 for a type without results:
 * Solid_1 keeps none in the reference).  `capacity` is the length of host_out
 * in doubles; returns the number of records written or a negative error. */
int gfa_results_stride(int element_type);
int64_t gfa_gauss_point_results(gfa_t* h, int element_type, double* host_out, int64_t capacity);
int gfa_copy_coordinates(gfa_t* h, double* host_out /* [n_nodes*6] */);

/* ---- the Newton-loop vector steps either side of the assembly, on the device copies ----
 * (they remove the per-iteration D2H of the residual). */
typedef struct gfa_norms {
    double  max_force, max_moment;   /* max |v(GL-1)| over free translational / rotational node DOFs:
                                      * v = P_A in gfa_residual (ConvergenceCriteria.cpp:200-217, 474-505),
                                      * v = the increment x_A in gfa_update_displacements (:305-340) */
    int32_t node_force, node_moment; /* 1-based node of the first maximum in node order, 0 if none (node_force / node_moment,
                                      * node_disp / node_rot of the reference) */
    double  max_disp_value, max_rot_value; /* gfa_update_displacements only: max |Node::displacements| after the update */
    int32_t nan_detected;            /* NaNDetector hit (the reference sets `diverged`) */
} gfa_norms_t;

/* Static.cpp:210-217: db.global_P_A = -1.0*db.global_P_A and, when X_B is given (first iteration of an
 * increment), db.global_P_A -= 1.0*(db.global_stiffness_AB*db.global_X_B) with the reference's own row loop
 * (SparseMatrix.cpp:186-190); then the max-norms EstablishResidualCriteria / CheckResidualConvergence read.
 * Call after gfa_assemble and the host contributions; gfa_vector(GFA_P_A) afterwards returns the right-hand
 * side handed to the solver.  In a partitioned run (after the interface exchange) the vector steps act on the
 * entries this rank stores and the norms cover the rows it OWNS: the caller takes the maximum over ranks (and
 * the smallest node among equal maxima), one tiny all-reduce, where the reference has one loop. */
int gfa_residual(gfa_t* h, const double* X_B /* [n_fixed] host, or NULL */, gfa_norms_t* out);

/* Solution::UpdateDisps (src/Solution.cpp:390-402) for node DOFs on the device copy of Node::displacements
 * (the array of the last gfa_assemble): displacements[j] += x_A(GL-1) where GL > 0, and the norms
 * CheckGLConvergence reads.  A following gfa_assemble with step.displacements == NULL evaluates the updated
 * device copy; gfa_displacements copies it back. */
int gfa_update_displacements(gfa_t* h, const double* x_A /* [n_free] host */, gfa_norms_t* out);
int gfa_displacements(gfa_t* h, double* host_out /* [n_nodes*6] */);

/* ---- Newmark dynamics: the element contributions Dynamic::Solve adds per Newton iteration -------------
 * (SURVEY.md 8f rank 1).  Beam_1, Shell_1 and Pipe_1 with its structural mass (src/Pipe_1.cpp:1131-1144; the
 * model tables carry no ocean data, so the added-mass branch of src/Pipe_1.cpp:1758-1795 stays on the host);
 * a model that holds Solid_1 elements gets GFA_EUNSUPPORTED from gfa_assemble_dynamic. */
typedef struct gfa_dynamic {
    double a1, a2, a3, a4, a5, a6;   /* Dynamic::a1..a6 of the current time step (src/Dynamic.cpp:582-590) */
    double rayleigh_alpha;           /* Dynamic::alpha (mass-proportional)       (src/Dynamic.h:26-27) */
    double rayleigh_beta;            /* Dynamic::beta  (stiffness-proportional) */
    int32_t update_rayleigh;         /* the argument of Solution::MountDamping: recompute Element::rayleigh_damping =
                                      * alpha*mass_modal + beta*stiffness from this iteration's stiffness
                                      * (first iteration of the solution step, or every iteration when Dynamic::update == 1;
                                      * src/Dynamic.cpp:329-335); otherwise the stored matrix is used */
} gfa_dynamic_t;

/* Node::vel / accel / copy_vel / copy_accel, [n_nodes*6] host arrays each (src/Node.h); NULL = leave the
 * device copy as it is (all four start at zero).  InitialCondition / prescribed motions write them on the
 * host and push them here; gfa_update_dyn and gfa_commit_state maintain them afterwards. */
int gfa_set_kinematics(gfa_t* h, const double* vel, const double* accel, const double* copy_vel, const double* copy_accel);
int gfa_kinematics(gfa_t* h, double* vel, double* accel, double* copy_vel, double* copy_accel);

/* Dynamic::UpdateDyn for node DOFs (src/Dynamic.cpp:480-556): vel / accel of every free DOF from the
 * displacement increments (`displacements` host [n_nodes*6], or NULL = the device copy left by
 * gfa_assemble / gfa_update_displacements), rotations rotated by Q(alpha_delta) as the reference does.
 * A rotational DOF that is not free keeps, in the reference's loop, the value the previous node left in
 * vel_aux / ace_aux; nodes with partly-free rotations reproduce that by replaying the loop. */
int gfa_update_dyn(gfa_t* h, const double* displacements, const gfa_dynamic_t* dyn);

/* One Newton iteration of Dynamic::Solve up to MountSparse (src/Dynamic.cpp:323-340): gfa_assemble plus
 * Element::MountMass (src/Beam_1.cpp:1564-1636, src/Shell_1.cpp:2406-2498), MountDamping
 * (src/Beam_1.cpp:1639-1664, src/Shell_1.cpp:2501-2533) and MountDyn (src/Beam_1.cpp:1667-1672,
 * src/Shell_1.cpp:2536-2541) folded into the element blocks before the scatter: the CSR values hold
 * stiffness + mass + a4*rayleigh_damping, the vectors P_loading + inertial_loading + damping_loading.
 * gfa_commit_state afterwards also saves alpha_i and copies vel/accel (src/Node.cpp:375-380). */
int gfa_assemble_dynamic(gfa_t* h, const gfa_step_t* step, const gfa_dynamic_t* dyn);

/* Committed Rodrigues rotation vector alpha_i of every Gauss point of one element, element frame
 * (src/Shell_1.h:121, src/LagrangeSave.h:12): 3 per point; returns the number of doubles written. */
int gfa_element_alpha_i(gfa_t* h, int32_t element, double* out);

/* Timing of the last gfa_assemble, milliseconds from CUDA events on the
 * library's stream: [0] H2D of displacements, [1] element evaluation
 * (MountLocal+MountElementLoads), [2] scatter (MountGlobal+MountSparse),
 * [3] whole call on the device. */
int gfa_last_timing(gfa_t* h, double* ms4);
/* Number of kernels the last gfa_assemble launched. */
int gfa_last_launch_count(gfa_t* h);
/* Which pipeline gfa_set_dofs chose for this DOF map: returns 0 for the classic evaluation + scatter kernels (the
 * default), 1 for the ring pipeline, negative on error; `buf` receives a one-line description.  In the ring
 * pipeline the element blocks go through an L2-resident ring of arena slots: persistent evaluation kernels (one
 * per element type) fill it chunk by chunk and a persistent scatter kernel, co-resident on every SM, drains it
 * behind them (release / acquire counters per chunk, a watchdog instead of a hang).  It halves the DRAM traffic
 * of a step and gives bitwise the classic results, but is slower on B200 today (profiles/r02_notes.md), hence
 * opt-in: GFA_RING=1 (always) or 2 (when the arena is larger than the ring); GFA_RING_CHUNK_KB and
 * GFA_RING_CHUNKS size the ring (default 9 x 7 MB).  gfa_last_timing then reports [1] = pre-pass of the pinned
 * elements, [2] = ring kernels + vectors.
 * The description also names the layout of the classic Shell_1 arena: by default the blocks of eight consecutive
 * elements are interleaved (one store instruction of the evaluation kernel covers 576 contiguous bytes);
 * GFA_ARENA_LAYOUT=0 keeps one region per element, and a handle switches to that layout by itself at its first
 * gfa_assemble_dynamic (the Newmark kernels walk an element's own region; the DOF map and registered loads stay). */
int gfa_pipeline_info(gfa_t* h, char* buf, int32_t capacity);

/* ---- multi-GPU (mesh partition by element range) ----------------------
 * Every rank holds the full DOF map but evaluates only its element partition
 * and stores only the AA rows its elements touch (gfa_local_rows; with one
 * rank these are all rows, in order, i.e. exactly the reference's CSR).  A
 * stored row always carries its complete global column set, so rows of nodes
 * on partition interfaces have the same layout on every rank that holds them
 * and receive partial sums there.  The exchange step moves only those rows:
 *   gfa_interface_counts : per peer rank, number of doubles this rank sends / receives
 *   gfa_interface_pack   : gathers this rank's partial interface values into
 *                          send_buf (device), segments ordered by peer rank
 *   gfa_interface_unpack : adds received partials (device recv_buf, segments
 *                          ordered by peer rank, peers ascending => fixed
 *                          summation order) into the owned rows
 * The transport between pack and unpack is the caller's (NCCL send/recv).  gfa_assemble scatters the rows of
 * partition interfaces FIRST and records an event; pack and unpack only ENQUEUE their kernels on
 * gfa_interface_stream() -- pack behind that event, so with gfa_assemble_enqueue the exchange runs while the interior
 * rows are still being scattered on gfa_stream().  The caller issues its transport on gfa_interface_stream() between
 * the two calls.  gfa_interface_unpack makes gfa_stream() wait for the exchange, so gfa_csr_values / gfa_vector / the
 * next gfa_assemble see the summed rows. */
int gfa_interface_counts(gfa_t* h, int64_t* send_counts /* [world] */, int64_t* recv_counts /* [world] */);
int gfa_interface_pack(gfa_t* h, double* send_buf_device);
int gfa_interface_unpack(gfa_t* h, const double* recv_buf_device);
/* global ids of the AA rows stored on this rank (ascending; row i of
 * gfa_csr_pattern(GFA_AA) is global row rows_out[i]) */
int gfa_local_rows(gfa_t* h, int64_t* n_rows, int32_t* rows_out /* may be NULL */);
/* global ids of the AA rows (= entries of P_A/I_A) that are complete on this
 * rank after the exchange; every row is owned by exactly one rank */
int gfa_owned_rows(gfa_t* h, int64_t* n_rows, int32_t* rows_out /* may be NULL */);

/* ---- partition-local transfers: what one rank of a partitioned run has to move per iteration ----------
 * gfa_touched_nodes: the nodes this rank's elements reference (0-based indices into the displacement array,
 * ascending; every node with one rank).  gfa_set_displacements_packed uploads Node::displacements of exactly those
 * nodes ([n_touched*6], host or device pointer) into the device copy -- a following gfa_assemble with
 * step.displacements == NULL evaluates them -- so that the host-to-device bytes of an iteration do not grow with the
 * number of ranks.  gfa_vector_owned returns the entries of P_A / I_A for the rows this rank owns (order of
 * gfa_owned_rows; with GFA_P_B: the whole vector, it is small): the residual is row-distributed like the matrix. */
int gfa_touched_nodes(gfa_t* h, int64_t* n_nodes, int32_t* nodes_out /* may be NULL */);
int gfa_set_displacements_packed(gfa_t* h, const double* packed, int32_t on_device);
int gfa_vector_owned(gfa_t* h, int which_vector, double* host_out);

/* Raw stream the library launches on (cudaStream_t), for callers that time
 * or order their own work against it. */
int gfa_stream(gfa_t* h, void** stream_out);
/* Stream of the interface exchange (pack, the caller's transport, unpack). */
int gfa_interface_stream(gfa_t* h, void** stream_out);

#ifdef __cplusplus
}
#endif
#endif /* GFA_H */
