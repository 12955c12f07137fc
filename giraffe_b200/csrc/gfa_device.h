// Device-side argument blocks shared by gfa_api.cu (host) and gfa_kernels.cu.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GFA_HD_INLINE __host__ __device__ __forceinline__
#else
#define GFA_HD_INLINE inline
#endif

namespace gfa {

// ---- element evaluation -------------------------------------------------
// One block per element type.  `n_el` counts the elements of that type in
// THIS rank's partition; `conn` holds 0-based node ids.  Gauss-point state is
// structure-of-arrays: state[k * n_gp + gp], gp = element * NGP + point.
struct EvalArgs {
    int n_el;                // elements of this type on this rank (array extents)
    int e_begin, e_end;      // range evaluated by this launch
    const int* conn;
    const int* prop;         // per element index into props
    const double* props;     // per-type property rows (see *_PROP_STRIDE)
    const double* pret;      // Beam_1::T0 per element or nullptr
    const double* xyz;       // [n_nodes*3] Node::ref_coordinates
    const double* copy;      // [n_nodes*6] Node::copy_coordinates
    const double* disp;      // [n_nodes*6] Node::displacements
    const double* geo;       // Shell_1 PreCalc: R(9), area per element (SoA)
    const double* shp;       // Shell_1 PreCalc: 21 shape values per Gauss point (SoA)
    double* state;           // committed Gauss-point state (read by eval, written by commit)
    double* Ke;              // element arena of contiguous, row-major 3x3 blocks (reference local DOF order).
                             // Beam_1: BEAM_ARENA, Solid_1: SOLID_ARENA doubles per element
                             // (see beam_block_offset(), solid_block_offset()).
                             // Shell_1: SHELL_ARENA doubles per element (see shell_block_offset()).
    double* Pe;              // [n_el * ndof]  P_loading = Fint - Fext
    double gx, gy, gz;       // Environment::G * l_factor (zero when no gravity)
    // Evaluation order and arena placement.  The launch evaluates list positions k in [e_begin, e_end);
    // position k is element elist[k] (identity when elist == nullptr).  Its blocks go to arena slot k
    // (classic: the arena holds every element) or, in ring mode (ring_chunks > 0), into the slot of its
    // chunk inside an L2-resident ring of ring_chunks x chunk_doubles doubles that the scatter role of the
    // fused kernel drains behind the evaluation (see FusedArgs).
    const int* elist;
    int ring_chunks;         // 0: classic arena
    int chunk_el;            // elements per chunk (multiple of the type's batch size)
    int chunk0;              // global index of this type's first chunk (ring slot = (chunk0 + k / chunk_el) % ring_chunks)
    long long chunk_doubles; // doubles per ring slot
    // Shell_1, classic arena only: 1 = the blocks of a batch of SHELL_BATCH consecutive elements are interleaved
    // (shell_batch_offset()) so that one store instruction of the evaluation kernel covers 9 * 64 contiguous bytes
    int batch_layout;
};
GFA_HD_INLINE int eval_element(const EvalArgs& A, int k) { return A.elist ? A.elist[k] : k; }
// first arena double of list position k; `arena` = doubles per element of the type
GFA_HD_INLINE double* eval_ke(const EvalArgs& A, int k, int arena) {
    if (A.ring_chunks == 0) return A.Ke + (size_t)k * arena;
    const int c = k / A.chunk_el;
    return A.Ke + (size_t)((A.chunk0 + c) % A.ring_chunks) * (size_t)A.chunk_doubles + (size_t)(k - c * A.chunk_el) * arena;
}

constexpr int SHELL_PROP_STRIDE = 5;   // lambda, mu, thickness, stiff_drill, rho
constexpr int BEAM_PROP_STRIDE = 52;   // D(36 row-major), R=CS triad(9 row-major E1,E2,E3), rho*A, strain-energy switch (1 Beam_1, 0 Pipe_1),
                                       // Jr = rho*(I11, I22, I33, I12) (Beam_1.cpp:587-591), Aint = pi Di^2 / 4 (Pipe_1.cpp:1151; 0 for Beam_1)
constexpr int SOLID_PROP_STRIDE = 3;   // lambda, mu, rho

constexpr int SHELL_RESULTS = 73;      // strain_energy + 3 x (eta1 eta2 kappa1 kappa2 n1 n2 m1 m2)
constexpr int BEAM_RESULTS = 25;       // strain_energy + 2 x (epsilon_r(6) sigma_r(6))
constexpr int SHELL_STATE = 21;        // Q_i(9) z_x1_i(3) z_x2_i(3) kappa_r1_i(3) kappa_r2_i(3)
constexpr int BEAM_STATE = 15;         // Q_i(9) dz_i(3) kappa_i_ref(3)

// Shell_1 stored blocks over its 9 group-nodes (0-5: u of nodes 1-6, 6-8: alpha of nodes 4-6):
// the 45 blocks (a <= b) of the upper triangle plus the 3 strictly-lower rotation-rotation blocks
// (the only part of the tangent that is not symmetric, Shell_1.cpp:1277-1302).  Block (a > b)
// outside the rotation corner is the transpose of stored block (b, a).
// Arena layout of one element (SHELL_ARENA doubles): the blocks are grouped by the evaluation
// work item that produces them, each group padded to whole 32-byte sectors, so that every sector
// is completed by one item within a few instructions (a sector left partially written is
// evicted early and costs a DRAM read-modify-write, profiles/r01_notes.md):
//   [64 K, 64 K + 63)       K = 0..2: translational columns K (rows u_0..u_K) then 5-K (rows u_0..u_{5-K})
//   [192 + 84 c, .. + 81)   c = 0..2: rotational column c, rows u_0..u_5, alpha_0..alpha_2
constexpr int SHELL_STORED = 48;
constexpr int SHELL_ARENA = 444;
// offset of stored block (a, b) inside the element's arena region
__host__ __device__ constexpr int shell_stored_offset(int a, int b) {
    return b < 6 ? (b < 3 ? 64 * b + 9 * a : 64 * (5 - b) + 9 * (6 - b + a)) : 192 + 84 * (b - 6) + 9 * a;
}
// offset of the stored block that holds block (a, b); `transposed` tells whether it holds (b, a)
__host__ __device__ inline int shell_block_offset(int a, int b, bool& transposed) {
    transposed = a > b && b < 6;
    return transposed ? shell_stored_offset(b, a) : shell_stored_offset(a, b);
}

// Batch layout of the classic shell arena (the default; the compact per-element layout above remains for the ring
// pipeline, for element lists and for the Newmark kernels, which walk an element's region).  Elements are taken in
// batches of SHELL_BATCH = the evaluation kernel's 8 elements per warp; a batch owns SHELL_BATCH * SHELL_ARENA
// doubles, and stored block number n (shell_stored_index) of its element number r sits at 72 n + 9 r: the same
// block of the eight elements is 576 contiguous bytes, which one store instruction of the kernel (24 lanes = 8
// elements x 3 components, 8 bytes each, per block row) covers in 4.5 lines instead of eight scattered ones
// (measured: evaluation 2.40 -> 2.28 ms; every block is still 72 contiguous bytes, the scatter does not change).
constexpr int SHELL_BATCH = 8;
__host__ __device__ constexpr int shell_stored_index(int a, int b) {
    return b < 6 ? (b < 3 ? 7 * b + a : 7 * (5 - b) + (6 - b + a)) : 21 + 9 * (b - 6) + a;
}
// offset of the stored block holding block (a, b) of the element at position `local` of its type, from the type's base
__host__ __device__ inline long long shell_batch_offset(long long local, int a, int b, bool& transposed) {
    transposed = a > b && b < 6;
    const int n = transposed ? shell_stored_index(b, a) : shell_stored_index(a, b);
    return (local / SHELL_BATCH) * (long long)(SHELL_BATCH * SHELL_ARENA) + 72 * n + 9 * (local % SHELL_BATCH);
}

// Beam_1 / Pipe_1 stored blocks over the 6 group-nodes (2 * node + rot): the 21 blocks (a <= b) of the upper
// triangle plus the 3 strictly-lower rotation-rotation blocks (3,1), (5,1), (5,3) -- as for Shell_1 the
// rotation-rotation part is the only non-symmetric one (Beam_1.cpp:799-822); block (a > b) elsewhere is the
// transpose of stored block (b, a).  Column block b holds its rows a = 0..b and then, for a rotational
// column, the rotational rows below it: 1, 4, 3, 5, 5, 6 blocks = 24 blocks, BEAM_ARENA doubles per element.
constexpr int BEAM_STORED = 24;
constexpr int BEAM_ARENA = 216;
__host__ __device__ constexpr int beam_col_base(int b) { return b == 0 ? 0 : b == 1 ? 1 : b == 2 ? 5 : b == 3 ? 8 : b == 4 ? 13 : 18; }
__host__ __device__ constexpr bool beam_is_stored(int a, int b) { return a <= b || ((a & 1) && (b & 1)); }
// offset of stored block (a, b) inside the element's arena region (beam_is_stored(a, b) must hold)
__host__ __device__ constexpr int beam_stored_offset(int a, int b) {
    return 9 * (beam_col_base(b) + (a <= b ? a : b + 1 + (a - b - 2) / 2));
}
// offset of the stored block that holds block (a, b); `transposed` tells whether it holds (b, a)
__host__ __device__ inline int beam_block_offset(int a, int b, bool& transposed) {
    transposed = !beam_is_stored(a, b);
    return transposed ? beam_stored_offset(b, a) : beam_stored_offset(a, b);
}

// Solid_1 (builder-defined hexahedron): the tangent is symmetric, so only the 36 blocks (a <= b) over the 8 nodes
// are stored, column block b holding its rows a = 0..b; block (a > b) is the transpose of stored block (b, a).
constexpr int SOLID_ARENA = 324;
__host__ __device__ constexpr int solid_stored_offset(int a, int b) { return 9 * (b * (b + 1) / 2 + a); }
__host__ __device__ inline int solid_block_offset(int a, int b, bool& transposed) {
    transposed = a > b;
    return transposed ? solid_stored_offset(b, a) : solid_stored_offset(a, b);
}

// ---- scatter ------------------------------------------------------------
// "Group-node" = one 3-DOF group of a node (translations or rotations); every
// in-scope element block is 3x3-structured over group-nodes in the
// reference's own local DOF order (Shell_1.cpp:1523-1557, Beam_1.cpp:1439-1444).
// All rows of a group-node share one column layout, made of "runs": the free
// DOFs of one neighbouring group-node are consecutive columns.  The slot map is
// therefore one entry per (group-node, neighbour) pair listing the element
// blocks that contribute to that 3x3 patch of the CSR, in ascending element
// order (the order the reference pushes and Eigen sums, Solution.cpp:327-328).
struct RunEnt {         // 16 bytes, one per CSR patch that is summed by the scatter kernel
    int dst;            // valAA offset of the patch's first entry (first free row, first free column)
    unsigned info;      // row stride (bits 0-15) | free mask of the row group (16-18) | free mask of the
                        // column group (19-21) | src0 / src1 read transposed (22, 23) | number of
                        // contributing blocks (24-31)
    unsigned src0, src1;// count <= 2: the sources themselves; count > 2: src0 = start in the overflow list
};
// source encoding: offset (in doubles, 32 bits) of the contiguous 3x3 block in the Ke arena; overflow
// list entries are 64-bit, bit 63 = read transposed
constexpr unsigned long long SRC_T = 1ULL << 63;
struct PInc {           // (element, local block) incidences of a group-node, for the residual vectors
    int pe_off;         // offset of the element's P in the Pe arena
    int la;             // local block index
};
// one record per group-node this rank's elements touch (residual vectors)
struct GnRec {
    int gl[3];          // global DOF ids (Node::GLs) of the group's 3 DOFs
    int ib, ie;         // incidences [ib, ie), element-ascending
};

struct ScatterArgs {
    long long n_runs;
    const RunEnt* runs;
    const unsigned long long* ovf;   // overflow source lists (patches fed by more than two blocks)
    long long n_gn;
    const GnRec* gn;
    const PInc* inc;
    const double* Ke;            // arena
    const double* Pe;            // arena
    double* valAA;
    double* PA; double* IA; double* PB;
};

// ---- ring pipeline: two co-resident persistent kernels -----------------------------------------------
// Evaluation: one CTA per SM, FUSED_WARPS (or as many as record buffers fit) independent warps at the full
// register budget; each claims element batches in order from a counter and evaluates them into an L2-resident
// RING of arena slots.  Scatter: one thin CTA per SM (FUSED_SCATTER_WARPS warps, 32 registers a thread -- the two
// kernels split an SM's register file exactly), running beside it on a second stream: the slot map is sorted by the
// chunk after which a group-node's patches are complete ("ready chunk") and cut into tiles of FUSED_TILE_COLS
// patch columns; a warp claims tiles of a chunk as soon as the evaluation of it (and of all earlier chunks) has
// finished.  A batch of chunk c may overwrite its ring slot once every tile that reads chunk c - ring_chunks has
// been scattered.  Completion is published through the control block with release / acquire counters; no atomics
// touch the results and the sums keep the reference's element-ascending order, so the output is bitwise that of
// the two-kernel classic path.  A watchdog (globaltimer) turns a wait that never ends into an error code.
constexpr int FUSED_WARPS = 7;                        // evaluation warps per CTA at most (7 x 255 registers leave exactly the scatter CTA's share)
constexpr int FUSED_SCATTER_WARPS = 6;
constexpr int FUSED_EVAL_REGS = 224;                  // registers per evaluation thread (see fused::eval_kernel)
constexpr int FUSED_TILE_PATCHES = 32;                 // patches per tile: one per lane, staged through shared memory together
constexpr int CTL_ABORT = 0, CTL_BATCH = 2 /* + type slot */, CTL_HDR = 8;
struct FusedArgs {
    EvalArgs ev;                     // ring-mode evaluation arguments (e_begin = 0, e_end = list length)
    int total_chunks;                // chunks of the whole step (all types)
    int span;                        // a patch whose ready chunk is r reads chunks [r - span, r]
    int n_buf;                       // evaluation warps (record buffers) per CTA
    int tile_group;                  // tiles claimed per atomic
    int scatter_ctas;                // scatter CTAs per SM
    int type_slot;
    ScatterArgs sc;                  // ring slot map
    const long long* chunk_run_ptr;  // [total_chunks + 1] first run (patch) whose ready chunk is c
    const int* chunk_tile_ptr;       // [total_chunks + 1] first tile of ready chunk c
    const int* chunk_batches;        // [total_chunks] evaluation batches of chunk c
    unsigned* ctl;                   // [CTL_HDR + 3 * total_chunks]: header, edone[], sdone[], tnext[]; zeroed per step
    unsigned long long timeout_ns;   // watchdog: a warp that waits longer sets CTL_ABORT and everyone leaves
};

// ---- ShellLoad follower pressure (ShellLoad.cpp:133-148 -> Shell_1::MountShellSpecialLoads, Shell_1.cpp:1392-1467) ----
// one entry per (load, Shell_1 element of its set): the 18 x 18 load stiffness on the element's u DOFs (not symmetric)
// and the 18 load-vector entries, SHELL_LOAD_REC doubles: K row-major, then P
constexpr int SHELL_LOAD_REC = 324 + 18;
struct ShellLoadArgs {
    int n_entries;
    const int* elem;             // [n_entries] local index of the Shell_1 element
    const int* load;             // [n_entries] index of the load
    const double* pressure;      // [n_loads] ShellLoad::GetValueAt(time) of every load
    const int* area_update;      // [n_loads] ShellLoad::area_update
    double* out;                 // [n_entries * SHELL_LOAD_REC]
};
bool shell_batch_layout_available();
void launch_shell_loads(const EvalArgs& a, const ShellLoadArgs& l, void* stream);
void launch_pipe_loads(const EvalArgs& a, const ShellLoadArgs& l, void* stream);     // PipeLoad: `pressure` = P0I per load, same record layout
// vals[dest[i]] += sum of src[seg[i] .. seg[i+1]) in list order (one thread per destination: fixed summation order)
void launch_gather_add(double* vals, const long long* seg, const long long* src, const long long* dest, const double* from, long long n_dest, void* stream);

// entries that involve a fixed DOF (AB, BA, BB): explicit gather lists
struct GatherArgs {
    long long n_dest;
    const long long* seg;        // [n_dest+1]
    const long long* src;        // Ke-arena offsets (entry granularity), element-ascending inside a segment
    const long long* dest;       // index into `vals`
    const double* Ke;
    double* vals;                // AB | BA | BB value arrays, one arena
};

// ---- Newton-loop vector steps either side of the assembly (Static.cpp:210-217, Solution.cpp:390-402,
//      ConvergenceCriteria.cpp:187-380, 460-598) -------------------------------------------------
struct NormAcc {                    // filled by the norm kernels; doubles are >= 0 and compared as bit patterns
    unsigned long long max_t, max_r;        // max |v(GL-1)| over free translational / rotational node DOFs
    unsigned long long max_dt, max_dr;      // max |Node::displacements| over the same DOFs (update only)
    int node_t, node_r;                     // 0-based node of the first maximum in node order (INT_MAX: none)
    int nan;                                // NaN seen in v
    int pad;
};
// ---- Newmark dynamics: MountMass + MountDamping + MountDyn (Beam_1.cpp:1564-1673, Shell_1.cpp:2406-2541),
//      Dynamic::UpdateDyn (Dynamic.cpp:480-556) ------------------------------------------------------
// Phase 1 (one thread per element) writes a small record per element; phase 2 (one warp per element) folds
// it and the stored Rayleigh matrix into the element arena in place, coalesced.
//   record = [uu NU*NU | aa 3*3*9 | P ndof | modal uu NU*NU | modal aa 3*3*9]
//   uu(a,b): scalar s of the translation-translation block s*I3 of nodes (a,b); aa(a,b): 3x3 rotation-rotation
//   block of the rotational nodes (Shell_1: nodes 4-6), global axes; P: inertial_loading in local DOF order.
constexpr int BEAM_DYN_REC = 9 + 81 + 18 + 9 + 81;
constexpr int SHELL_DYN_REC = 36 + 81 + 27 + 36 + 81;
struct DynArgs {
    double a1, a2, a3, a4, a5, a6;      // Dynamic::a1..a6
    double ray_alpha, ray_beta;         // Dynamic::alpha, beta
    int update;                         // MountDamping(update_rayleigh)
    const double* vel;                  // [n_nodes*6] Node::vel
    const double* copy_vel;             // Node::copy_vel
    const double* copy_accel;           // Node::copy_accel
    double* alpha_i;                    // committed Rodrigues vector per Gauss point, SoA [3][n_gp], element frame
    double* rec;                        // [n_el * *_DYN_REC]
    double* CR;                         // Element::rayleigh_damping in the layout of the Ke arena, or nullptr (= zero)
};
void launch_shell_dynamics(const EvalArgs& a, const DynArgs& d, void* stream);   // 2 launches
void launch_beam_dynamics(const EvalArgs& a, const DynArgs& d, void* stream);    // 2 launches
void launch_shell_alpha_commit(const EvalArgs& a, double* alpha_i, void* stream);
void launch_beam_alpha_commit(const EvalArgs& a, double* alpha_i, void* stream);
// vel/accel of every node from the displacement increments; `mixed` lists the nodes whose rotational DOFs are
// partly free (they see the values the reference's loop leaves from earlier nodes), `start` the node each
// replay begins at
void launch_update_dyn(const DynArgs& d, const int* gls, const double* disp, double* vel, double* accel, int n_nodes,
                       const int* mixed, const int* start, int n_mixed, void* stream);
int configure_dynamics();   // constant tables; returns cudaError_t as int

void launch_negate(double* v, long long n, void* stream);
void launch_sub_ab_xb(double* PA, const int* rows, const int* ptr, const int* inner, const double* vals, const double* XB, int n_rows, void* stream);
void launch_update_disps(const int* gls, double* disp, const double* x, int n_nodes, void* stream);
void launch_norms(const int* gls, const double* v, const double* disp, int n_nodes, NormAcc* acc, void* stream);

void launch_shell_eval(const EvalArgs& a, void* stream);
void launch_beam_eval(const EvalArgs& a, void* stream);
void launch_solid_eval(const EvalArgs& a, void* stream);
void launch_shell_precalc(const EvalArgs& a, double* geo, double* shp, void* stream);
void launch_shell_commit(const EvalArgs& a, void* stream);
void launch_beam_commit(const EvalArgs& a, void* stream);
void launch_shell_results(const EvalArgs& a, double* out, void* stream);     // out[n_el * GFA_SHELL_RESULTS]
void launch_beam_results(const EvalArgs& a, double* out, void* stream);      // out[n_el * GFA_BEAM_RESULTS]
void launch_node_commit(int n_nodes, double* copy, double* disp, void* stream);
int launch_scatter(const ScatterArgs& a, void* stream);      // returns the number of kernels launched
void launch_vectors(const ScatterArgs& a, void* stream);
int launch_fused_eval(const FusedArgs& f, void* stream);     // returns cudaError_t as int
int launch_fused_scatter(const FusedArgs& f, void* stream);
int fused_buffers(int type_slot);                            // record buffers (evaluating warps) per CTA of the fused kernel
void launch_gather(const GatherArgs& a, void* stream);
void launch_add_slots(double* vals, const long long* slots, const double* add, long long n, void* stream);
void launch_pack(const double* vals, const long long* idx, double* buf, long long n, void* stream);
void launch_unpack_nodes(double* disp, const int* nodes, const double* packed, long long n_nodes, void* stream);   // disp[nodes[i]*6 + k] = packed[i*6 + k]
void launch_unpack_add(double* vals, const long long* idx, const double* buf, long long n, void* stream);
int configure_kernels();   // opt-in shared memory sizes; returns cudaError_t as int

} // namespace gfa
