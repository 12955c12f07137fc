// C-ABI implementation (include/gfa.h): host-side model tables, DOF map,
// CSR pattern + element->slot maps, and the per-iteration launch sequence.
// Host code mirrors the reference's set-up steps:
//   gfa_create      <-> Element::PreCalc loop            (Database.cpp:713-714)
//   gfa_number_dofs <-> DOFsActive + SetGlobalDOFs        (Solution.cpp:40-224)
//   gfa_set_dofs    <-> SetGlobalSize                     (Solution.cpp:577-654)
//   gfa_assemble    <-> Clear, MountLocal, MountElementLoads, MountGlobal,
//                       MountSparse                       (Static.cpp:203-212)
//   gfa_commit_state<-> SaveConfiguration                 (Solution.cpp:426-454)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gfa.h"
#include "gfa_device.h"

using namespace gfa;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) return fail(GFA_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T>& h) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
};

struct TypeInfo { int type, nn, nb, ndof, ngp, nstate; };
const TypeInfo kTypes[3] = {
    { GFA_SHELL_1, 6, 9, 27, 3, SHELL_STATE },
    { GFA_BEAM_1, 3, 6, 18, 2, BEAM_STATE },
    { GFA_SOLID_1, 8, 8, 24, 8, 0 },
};
// doubles per element in the Ke arena: Shell_1 stores the upper triangle plus the non-symmetric
// rotation corner in sector-padded groups (gfa_device.h: shell_stored_offset), the others every block
inline int arena_doubles(int slot) { return slot == 0 ? SHELL_ARENA : slot == 1 ? BEAM_ARENA : SOLID_ARENA; }
// doubles of the classic arena a type's n elements take: whole batches for Shell_1 (batch layout, gfa_device.h)
inline long long type_region(int slot, size_t n) {
    return slot == 0 ? (long long)((n + SHELL_BATCH - 1) / SHELL_BATCH) * (SHELL_BATCH * SHELL_ARENA) : (long long)n * arena_doubles(slot);
}
// Pipe_1 shares the Beam_1 block: same Mount / MountGlobal / SaveLagrange (Pipe_1.cpp:836-974, 1027-1104),
// other constants (PreCalc, Pipe_1.cpp:1106-1146)
inline int type_slot(int t) { return t == GFA_SHELL_1 ? 0 : (t == GFA_BEAM_1 || t == GFA_PIPE_1) ? 1 : t == GFA_SOLID_1 ? 2 : -1; }
// local 3-DOF block -> (local node, DOF group 0 = translations / 1 = rotations)
// in the reference's local DOF order (Shell_1.cpp:1523-1557, Beam_1.cpp:1439-1444)
inline void block_node(int slot, int b, int& a, int& grp) {
    if (slot == 0) { if (b < 6) { a = b; grp = 0; } else { a = 3 + (b - 6); grp = 1; } }
    else if (slot == 1) { a = b / 2; grp = b % 2; }
    else { a = b; grp = 0; }
}

struct TypeBlock {
    std::vector<int> elems;          // global element ids of this rank's partition, ascending
    std::vector<int> conn, prop;
    std::vector<double> props, pret;
    bool any_pret = false;
    long long ke_base = 0;           // classic arena: first double of this type's block
    int pe_base = 0;
    // arena placement of every local element (offset of its first double in d_Ke), set by gfa_set_dofs
    std::vector<long long> ke_off;
    // ring pipeline: elements evaluated into the ring (ascending) and the pinned ones (own regions behind it)
    std::vector<int> ring_list, pin_list;
    int chunk_el = 0, chunk0 = 0, n_chunks = 0;
    long long pin_base = 0;          // first double of this type's pinned regions in d_Ke
    DevBuf<int> d_ring_list, d_pin_list;
    DevBuf<int> d_conn, d_prop;
    DevBuf<double> d_props, d_pret, d_state, d_geo, d_shp;
    // Newmark dynamics: committed Rodrigues vector per Gauss point (always kept: SaveLagrange updates it in
    // static steps too), per-element record and Element::rayleigh_damping (allocated on first use)
    DevBuf<double> d_alpha_i, d_dynrec, d_CR;
};

// std::vector whose resize() leaves new elements uninitialised: the 2 GB column array of a 1M-shell AA is written
// once, in parallel, right after it is sized -- zero-filling it first costs half a second of one core
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    template <class U, class... Args> void construct(U* p, Args&&... args) {
        if constexpr (sizeof...(Args) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<Args>(args)...);
    }
};

struct HostCsr {
    std::vector<long long> rowptr;
    std::vector<int, NoInitAlloc<int> > inner;
    int rows = 0, cols = 0;
    // AA only: rows stored on this rank (all rows when world == 1) and the
    // inverse map global row -> local row (-1 = row lives on other ranks only)
    std::vector<int> row_ids, row_local;
};

} // namespace

struct gfa_handle {
    int device = 0;
    cudaStream_t stream = nullptr;        // the stream every kernel of the path runs on (callers may time / order against it)
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    // multi-GPU: the rows of partition interfaces are scattered first; pack / transport / unpack run on their own
    // stream behind ev_iface while the interior rows are still being scattered on `stream`
    cudaStream_t stream_if = nullptr;
    cudaEvent_t ev_iface = nullptr, ev_unpacked = nullptr;
    long long n_iface_runs = 0, n_iface_gn = 0;
    int rank = 0, world = 1;

    int n_nodes = 0, n_el = 0;
    std::vector<int> el_type, el_ptr, el_nodes;      // global tables (all ranks), nodes 0-based
    std::vector<int> el_owner_slot, el_local;        // element -> type slot / index inside this rank's block (-1 if not owned)
    TypeBlock tb[3];
    std::vector<long long> type_count = std::vector<long long>(3, 0);
    bool gravity_on = false;
    double grav[3] = { 0, 0, 0 };

    DevBuf<double> d_xyz, d_copy, d_disp, d_Ke, d_Pe;
    // Newmark dynamics: Node::vel / accel / copy_vel / copy_accel (allocated on first use), nodes whose
    // rotational DOFs are partly free (Dynamic::UpdateDyn replay) and where each replay starts
    DevBuf<double> d_vel, d_accel, d_cvel, d_caccel;
    DevBuf<int> d_mixed, d_mixed_start;
    int n_mixed = 0;

    // DOF map / pattern
    bool dofs_set = false;
    int n_free = 0, n_fixed = 0;
    std::vector<int> gls;
    HostCsr csr[4];
    long long arena_off[4] = { 0, 0, 0, 0 };         // value offsets of AA, AB, BA, BB in the arena
    long long vec_off[3] = { 0, 0, 0 };              // PA, IA, PB
    long long arena_size = 0;
    DevBuf<double> d_arena;
    DevBuf<GnRec> d_gn;
    DevBuf<RunEnt> d_runs;
    DevBuf<unsigned long long> d_ovf;
    // Newton-loop vector steps: DOF map and the non-empty rows of AB on the device
    DevBuf<int> d_gls, d_ab_rows, d_ab_ptr, d_ab_inner;
    DevBuf<NormAcc> d_norm;
    int n_ab_rows = 0;
    DevBuf<PInc> d_inc;
    long long n_runs = 0, n_gn_local = 0;
    DevBuf<long long> d_gseg, d_gsrc, d_gdest;
    long long n_gdest = 0;
    // interface exchange
    std::vector<long long> send_cnt, recv_cnt;
    DevBuf<long long> d_send_idx, d_recv_idx;
    std::vector<int> owned_rows;
    // partition-local transfers: nodes this rank's elements reference, staging for the packed upload, owned-row gather
    std::vector<int> touched_nodes;
    DevBuf<int> d_touched_nodes;
    DevBuf<double> d_packed;
    DevBuf<long long> d_owned_idx;

    // ring pipeline (fused evaluation + scatter, FusedArgs in gfa_device.h); `ring` false = classic two-kernel path
    bool ring = false, force_classic = false;
    bool ring_serial = false;             // ring placement, but classic kernels launched group by group (GFA_RING=3)
    int ring_group = 4;                   // chunks per group of the serial variant
    std::vector<long long> chunk_run_ptr_h;
    int ring_chunks = 0, ring_span = 0, total_chunks = 0;
    long long chunk_doubles = 0, ring_doubles = 0;
    long long n_pre_runs = 0;             // runs fed by pinned elements only: scattered by the classic kernel before the fused launches
    DevBuf<long long> d_chunk_run_ptr;
    DevBuf<int> d_chunk_tile_ptr, d_chunk_batches;
    cudaStream_t stream_sc = nullptr;     // the scatter kernel of the ring pipeline runs here, beside the evaluation kernels on `stream`
    cudaEvent_t ev_ctl = nullptr, ev_scattered = nullptr;
    DevBuf<unsigned> d_ctl;
    DevBuf<double> d_scratch_ke;          // one element's blocks, for gfa_element_block in ring mode
    DevBuf<int> d_one;
    std::string ring_note;                // why the classic path was chosen, or the ring geometry
    // arguments of the last gfa_set_dofs (replayed when a handle has to leave ring mode)
    std::vector<int> ex_mat, ex_rows, ex_cols;
    long long n_vec_untouched = 0;        // vector entries no element of this rank writes (zeroed per assembly ...
    bool vec_dirty = false;               // ... when something has been added to the vectors since they were last zero)
    // persistent staging of gfa_add_host_*
    DevBuf<long long> d_stage_slots; DevBuf<double> d_stage_vals;
    std::vector<long long> stage_slots; std::vector<double> stage_vals;
    std::vector<std::pair<long long, double> > stage_items;
    DevBuf<double> d_xb, d_xa;            // X_B / x_A of the Newton-loop vector steps
    DevBuf<int> d_gls_owned;              // DOF map with the free ids of rows other ranks own removed (norms of a partitioned run)

    int fused_eval_warps[3] = { 0, 0, 0 };
    int fused_tile_group = 8, fused_scatter_ctas = 1;
    unsigned long long fused_timeout_ns = 4000000000ULL;
    bool abort_check_pending = false;     // the watchdog flag of the last fused launches has not been read yet
    double last_gfac = 0.0;

    // element loads evaluated on the device: ShellLoad follower pressure (gfa_set_shell_loads / gfa_apply_shell_loads)
    // and PipeLoad internal pressure (gfa_set_pipe_loads / gfa_apply_pipe_loads)
    struct LoadSet {
        int n_loads = 0, n_entries = 0;
        long long n_dest = 0;
        DevBuf<int> d_elem, d_of, d_flag;
        DevBuf<double> d_value, d_out;
        DevBuf<long long> d_seg, d_src, d_dest;
    } shell_loads, pipe_loads;

    bool replaying = false;               // gfa_set_dofs called by leave_ring_mode with the handle's own arguments
    bool batch_layout = false;            // classic shell arena interleaved by batches of 8 elements (gfa_device.h: shell_batch_offset)
    bool force_compact = false;           // the Newmark kernels walk an element's own region: compact layout from then on
    bool assembled = false;
    bool timing_pending = false;          // events of the last assembly not read yet (gfa_assemble_enqueue)
    float last_ms[4] = { 0, 0, 0, 0 };
    int last_launches = 0;
};

namespace {

enum EvalPass { PASS_ALL, PASS_PINNED, PASS_RING };

EvalArgs eval_args(gfa_t* h, int slot, double gfac, EvalPass pass = PASS_ALL) {
    TypeBlock& t = h->tb[slot];
    EvalArgs a;
    a.n_el = (int)t.elems.size();
    a.e_begin = 0; a.e_end = a.n_el;
    a.elist = nullptr; a.ring_chunks = 0; a.chunk_el = 1; a.chunk0 = 0; a.chunk_doubles = 0;
    a.batch_layout = (slot == 0 && h->batch_layout) ? 1 : 0;
    a.conn = t.d_conn.p; a.prop = t.d_prop.p; a.props = t.d_props.p;
    a.pret = t.any_pret ? t.d_pret.p : nullptr;
    a.xyz = h->d_xyz.p; a.copy = h->d_copy.p; a.disp = h->d_disp.p;
    a.state = t.d_state.p;
    a.geo = t.d_geo.p; a.shp = t.d_shp.p;
    a.Ke = h->d_Ke.p + t.ke_base;
    a.Pe = h->d_Pe.p + t.pe_base;
    const double f = h->gravity_on ? gfac : 0.0;
    a.gx = h->grav[0] * f; a.gy = h->grav[1] * f; a.gz = h->grav[2] * f;
    if (pass == PASS_PINNED) {          // ring mode: the pinned elements, each into its own region behind the ring
        a.elist = t.d_pin_list.p; a.e_end = (int)t.pin_list.size();
        a.Ke = h->d_Ke.p + t.pin_base;
    } else if (pass == PASS_RING) {
        a.elist = t.pin_list.empty() ? nullptr : t.d_ring_list.p; a.e_end = (int)t.ring_list.size();
        a.Ke = h->d_Ke.p;
        a.ring_chunks = h->ring_chunks; a.chunk_el = t.chunk_el; a.chunk0 = t.chunk0; a.chunk_doubles = h->chunk_doubles;
    }
    return a;
}

// offset (in doubles) of block (la, b) of a local element in the Ke arena; `tr` = stored transposed
inline long long arena_block(const gfa_t* h, int slot, int local, int la, int b, bool& tr) {
    tr = false;
    if (slot == 0 && h->batch_layout) return h->tb[0].ke_base + shell_batch_offset(local, la, b, tr);
    const long long base = h->tb[slot].ke_off[local];
    if (slot == 0) return base + shell_block_offset(la, b, tr);
    if (slot == 1) return base + beam_block_offset(la, b, tr);
    return base + solid_block_offset(la, b, tr);
}

} // namespace

extern "C" {

const char* gfa_last_error(void) { return g_err.c_str(); }

int gfa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gfa_create(const gfa_model_t* m, int device, gfa_t** out) {
    if (!m || !out) return fail(GFA_EINVAL, "gfa_create: null argument");
    *out = nullptr;
    int ndev = gfa_device_count();
    if (ndev <= 0) return fail(GFA_ENODEVICE, "gfa_create: no CUDA device is visible (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(GFA_EINVAL, "gfa_create: device %d out of range (%d visible)", device, ndev);
    if (m->n_nodes <= 0 || m->n_elements < 0 || !m->ref_coordinates) return fail(GFA_EINVAL, "gfa_create: empty model");
    if (m->part_world < 1 || m->part_rank < 0 || m->part_rank >= m->part_world)
        return fail(GFA_EINVAL, "gfa_create: bad partition rank %d of %d", m->part_rank, m->part_world);
    if (m->part_world > 64) return fail(GFA_EUNSUPPORTED, "gfa_create: %d ranks; the interface bookkeeping holds rank sets in 64 bits", m->part_world);
    CUDA_TRY(cudaSetDevice(device));
    int cfg = configure_kernels();
    if (cfg == 0) cfg = configure_dynamics();
    if (cfg != 0) return fail(GFA_ECUDA, "kernel configuration: %s", cudaGetErrorString((cudaError_t)cfg));

    gfa_t* h = new gfa_handle();
    h->device = device;
    h->rank = m->part_rank; h->world = m->part_world;
    h->n_nodes = m->n_nodes; h->n_el = m->n_elements;
    h->gravity_on = m->gravity_on != 0;
    for (int k = 0; k < 3; k++) h->grav[k] = m->gravity[k];
#define FAIL_FREE(code, ...) do { int c_ = fail(code, __VA_ARGS__); delete h; return c_; } while (0)

    // ---- global connectivity tables, validation ------------------------
    h->el_type.assign(m->elem_type, m->elem_type + m->n_elements);
    h->el_ptr.assign(m->elem_node_ptr, m->elem_node_ptr + m->n_elements + 1);
    h->el_nodes.resize((size_t)h->el_ptr[m->n_elements]);
    for (size_t i = 0; i < h->el_nodes.size(); i++) {
        int nd = m->elem_nodes[i];
        if (nd < 1 || nd > m->n_nodes) FAIL_FREE(GFA_EINVAL, "element connectivity references node %d (model has %d)", nd, m->n_nodes);
        h->el_nodes[i] = nd - 1;
    }
    for (int e = 0; e < m->n_elements; e++) {
        int s = type_slot(h->el_type[e]);
        if (s < 0) FAIL_FREE(GFA_EUNSUPPORTED, "element %d: type %d has no kernel (Beam_1=1, Pipe_1=2, Shell_1=3, Solid_1=7)", e + 1, h->el_type[e]);
        if (h->el_ptr[e + 1] - h->el_ptr[e] != kTypes[s].nn) FAIL_FREE(GFA_EINVAL, "element %d: expected %d nodes", e + 1, kTypes[s].nn);
        const int* nd = &h->el_nodes[h->el_ptr[e]];
        for (int a = 0; a < kTypes[s].nn; a++)
            for (int b = a + 1; b < kTypes[s].nn; b++)
                if (nd[a] == nd[b]) FAIL_FREE(GFA_EINVAL, "element %d repeats node %d", e + 1, nd[a] + 1);
        int mat = m->elem_material[e];
        if (h->el_type[e] != GFA_PIPE_1 && (mat < 1 || mat > m->n_materials)) FAIL_FREE(GFA_EINVAL, "element %d: material %d out of range", e + 1, mat);
        h->type_count[s]++;
    }
    // ---- partition: contiguous range of every type's elements ----------
    h->el_owner_slot.assign(m->n_elements, -1);
    h->el_local.assign(m->n_elements, -1);
    {
        long long seen[3] = { 0, 0, 0 };
        for (int e = 0; e < m->n_elements; e++) {
            int s = type_slot(h->el_type[e]);
            long long k = seen[s]++;
            long long lo = h->type_count[s] * h->rank / h->world, hi = h->type_count[s] * (h->rank + 1) / h->world;
            if (k >= lo && k < hi) {
                h->el_owner_slot[e] = s;
                h->el_local[e] = (int)h->tb[s].elems.size();
                h->tb[s].elems.push_back(e);
            }
        }
    }
    {   // nodes referenced by this rank's elements (gfa_touched_nodes)
        std::vector<unsigned char> used((size_t)m->n_nodes, 0);
        for (int sl = 0; sl < 3; sl++)
            for (int e : h->tb[sl].elems)
                for (int k = h->el_ptr[e]; k < h->el_ptr[e + 1]; k++) used[(size_t)h->el_nodes[k]] = 1;
        for (int i = 0; i < m->n_nodes; i++) if (used[i] || h->world == 1) h->touched_nodes.push_back(i);
    }
    // ---- per-type tables and property rows -------------------------------
    std::vector<double> state_init[3];
    for (int s = 0; s < 3; s++) {
        TypeBlock& t = h->tb[s];
        const TypeInfo& ti = kTypes[s];
        const size_t ne = t.elems.size();
        t.conn.resize(ne * ti.nn);
        t.prop.resize(ne);
        t.pret.assign(ne, 0.0);
        std::map<std::vector<int>, int> combos;
        for (size_t k = 0; k < ne; k++) {
            const int e = t.elems[k];
            for (int a = 0; a < ti.nn; a++) t.conn[k * ti.nn + a] = h->el_nodes[h->el_ptr[e] + a];
            const int mat = m->elem_material[e], sec = m->elem_section[e], cs = m->elem_cs[e];
            std::vector<int> key;
            if (s == 0) {
                if (sec < 1 || sec > m->n_shell_sections) FAIL_FREE(GFA_EINVAL, "shell element %d: shell section %d out of range", e + 1, sec);
                key = { mat, sec };
            } else if (s == 1 && h->el_type[e] == GFA_PIPE_1) {
                if (sec < 1 || sec > m->n_pipe_sections || !m->pipe_sections) FAIL_FREE(GFA_EINVAL, "pipe element %d: pipe section %d out of range", e + 1, sec);
                if (cs < 1 || cs > m->n_cs) FAIL_FREE(GFA_EINVAL, "pipe element %d: CS %d out of range", e + 1, cs);
                key = { -1, sec, cs };
            } else if (s == 1) {
                if (sec < 1 || sec > m->n_sections) FAIL_FREE(GFA_EINVAL, "beam element %d: section %d out of range", e + 1, sec);
                if (cs < 1 || cs > m->n_cs) FAIL_FREE(GFA_EINVAL, "beam element %d: CS %d out of range", e + 1, cs);
                key = { mat, sec, cs };
                if (m->beam_pretension && m->beam_pretension[e] != 0.0) { t.pret[k] = m->beam_pretension[e]; t.any_pret = true; }
            } else key = { mat };
            auto it = combos.find(key);
            if (it == combos.end()) {
                const int id = (int)combos.size();
                combos[key] = id;
                const bool pipe = h->el_type[e] == GFA_PIPE_1;
                const double E = pipe ? 0.0 : m->hooke[3 * (mat - 1)], nu = pipe ? 0.0 : m->hooke[3 * (mat - 1) + 1], rho = pipe ? 0.0 : m->hooke[3 * (mat - 1) + 2];
                if (s == 0) {           // Shell_1::PreCalc, Shell_1.cpp:2016-2021
                    const double th = m->shell_thickness[sec - 1];
                    const double mu = E / (2.0 * (1 + nu));
                    const double lambda = 2.0 * mu * nu / (1 - 2.0 * nu);
                    const double row[SHELL_PROP_STRIDE] = { lambda, mu, th, E * th * th * th, rho };
                    t.props.insert(t.props.end(), row, row + SHELL_PROP_STRIDE);
                } else if (s == 1 && pipe) {   // Pipe_1::PreCalc, Pipe_1.cpp:1108-1114, 1130-1133
                    const double* ps = m->pipe_sections + 11 * (size_t)(sec - 1);      // EA EI GJ GA Rho ...
                    double row[BEAM_PROP_STRIDE];
                    for (int i = 0; i < BEAM_PROP_STRIDE; i++) row[i] = 0.0;
                    row[0] = ps[3]; row[7] = ps[3]; row[14] = ps[0]; row[21] = ps[1]; row[28] = ps[1]; row[35] = ps[2];
                    for (int i = 0; i < 9; i++) row[36 + i] = m->cs[9 * (size_t)(cs - 1) + i];
                    row[45] = ps[4];            // mass per unit length (gravity: Pipe_1.cpp:1311-1330 without ocean data)
                    row[46] = 0.0;              // Pipe_1::Mount leaves strain_energy at zero
                    // Jr as written in Pipe_1.cpp:1137-1139 (the squared radii are subtracted); Mr = Rho I without ocean data
                    const double rr = (ps[9] / 2.0) * (ps[9] / 2.0) - (ps[10] / 2.0) * (ps[10] / 2.0);
                    row[47] = (ps[4] * rr / 4.0); row[48] = (ps[4] * rr / 4.0); row[49] = (ps[4] * rr / 2.0); row[50] = 0.0;
                    row[51] = 3.1415926535897932384626433832795 * ps[10] * ps[10] / 4.0;   // Aint (Pipe_1.cpp:1151), for the PipeLoad pressure
                    t.props.insert(t.props.end(), row, row + BEAM_PROP_STRIDE);
                } else if (s == 1) {    // Beam_1::PreCalc, Beam_1.cpp:560-580
                    const double* sc = m->sections + 6 * (size_t)(sec - 1);
                    const double G = E / (2 * (1 + nu)), sf = 1.0;
                    double row[BEAM_PROP_STRIDE];
                    for (int i = 0; i < BEAM_PROP_STRIDE; i++) row[i] = 0.0;
                    row[0] = sf * G * sc[0]; row[7] = sf * G * sc[0]; row[14] = E * sc[0];
                    row[21] = E * sc[1]; row[28] = E * sc[2]; row[22] = E * sc[3]; row[27] = E * sc[3]; row[35] = G * sc[5];
                    for (int i = 0; i < 9; i++) row[36 + i] = m->cs[9 * (size_t)(cs - 1) + i];
                    row[45] = rho * sc[0];
                    row[46] = 1.0;
                    row[47] = rho * sc[1]; row[48] = rho * sc[2]; row[49] = rho * sc[4]; row[50] = rho * sc[3];   // Jr, Beam_1.cpp:587-591
                    t.props.insert(t.props.end(), row, row + BEAM_PROP_STRIDE);
                } else {                // builder-defined Solid_1: Lame constants of the Hooke material
                    const double mu = E / (2.0 * (1 + nu));
                    const double lambda = E * nu / ((1 + nu) * (1 - 2.0 * nu));
                    const double row[SOLID_PROP_STRIDE] = { lambda, mu, rho };
                    t.props.insert(t.props.end(), row, row + SOLID_PROP_STRIDE);
                }
                t.prop[k] = id;
            } else t.prop[k] = it->second;
        }
        // initial committed state (Shell_1.cpp:2357-2362; LagrangeSave.cpp:41-50, Beam_1.cpp:616-619)
        const size_t ngp = ne * ti.ngp;
        state_init[s].assign(ngp * ti.nstate, 0.0);
        for (size_t gp = 0; gp < ngp && ti.nstate; gp++) {
            double* st = state_init[s].data();
            st[0 * ngp + gp] = 1.0; st[4 * ngp + gp] = 1.0; st[8 * ngp + gp] = 1.0;     // Q_i = I
            if (s == 0) { st[9 * ngp + gp] = 1.0; st[13 * ngp + gp] = 1.0; }            // z,1 = e1, z,2 = e2
            else {
                const size_t k = gp / ti.ngp;
                const double EA = t.props[BEAM_PROP_STRIDE * (size_t)t.prop[k] + 14];
                st[11 * ngp + gp] = 1.0 + t.pret[k] / EA;                              // dz_i(2) = 1 + T0/EA
            }
        }
    }
    // ---- device uploads ----------------------------------------------------
    {
        std::vector<double> xyz(m->ref_coordinates, m->ref_coordinates + 3 * (size_t)m->n_nodes);
        std::vector<double> copy(6 * (size_t)m->n_nodes, 0.0);
        if (m->copy_coordinates) copy.assign(m->copy_coordinates, m->copy_coordinates + 6 * (size_t)m->n_nodes);
        else for (int i = 0; i < m->n_nodes; i++) for (int k = 0; k < 3; k++) copy[6 * (size_t)i + k] = xyz[3 * (size_t)i + k];
        cudaError_t e = h->d_xyz.upload(xyz);
        if (e == cudaSuccess) e = h->d_copy.upload(copy);
        if (e == cudaSuccess) e = h->d_disp.alloc(6 * (size_t)m->n_nodes);
        if (e == cudaSuccess) e = cudaMemset(h->d_disp.p, 0, 6 * (size_t)m->n_nodes * sizeof(double));
        long long ke = 0; long long pe = 0;
        for (int s = 0; s < 3 && e == cudaSuccess; s++) {
            TypeBlock& t = h->tb[s];
            t.ke_base = ke; t.pe_base = (int)pe;
            ke += type_region(s, t.elems.size());
            pe += (long long)t.elems.size() * kTypes[s].ndof;
            e = t.d_conn.upload(t.conn);
            if (e == cudaSuccess) e = t.d_prop.upload(t.prop);
            if (e == cudaSuccess) e = t.d_props.upload(t.props);
            if (e == cudaSuccess && t.any_pret) e = t.d_pret.upload(t.pret);
            if (e == cudaSuccess) e = t.d_state.upload(state_init[s]);
            if (e == cudaSuccess && kTypes[s].nstate) {                    // alpha_i = 0 (LagrangeSave.cpp:16-18, Shell_1.cpp:163)
                e = t.d_alpha_i.alloc(3 * t.elems.size() * kTypes[s].ngp);
                if (e == cudaSuccess && t.d_alpha_i.n) e = cudaMemset(t.d_alpha_i.p, 0, t.d_alpha_i.n * sizeof(double));
            }
        }
        if (pe > 0x7fffffffLL) FAIL_FREE(GFA_EUNSUPPORTED, "element force arena exceeds 2^31 entries");
        // the element arena is sized by gfa_set_dofs (classic: every element; ring: ring slots + pinned elements)
        if (e == cudaSuccess) e = h->d_Pe.alloc((size_t)pe);
        if (e == cudaSuccess) e = h->tb[0].d_geo.alloc(10 * h->tb[0].elems.size());
        if (e == cudaSuccess) e = h->tb[0].d_shp.alloc(21 * 3 * h->tb[0].elems.size());
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream_if, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream_sc, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_ctl, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_scattered, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_iface, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_unpacked, cudaEventDisableTiming);
        for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&h->ev[i]);
        if (e != cudaSuccess) FAIL_FREE(e == cudaErrorMemoryAllocation ? GFA_ENOMEM : GFA_ECUDA, "device set-up: %s", cudaGetErrorString(e));
    }
#undef FAIL_FREE
    // Shell_1::PreCalc on the device (frames, areas, shape functions)
    launch_shell_precalc(eval_args(h, 0, 0.0), h->tb[0].d_geo.p, h->tb[0].d_shp.p, h->stream);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        delete h;
        return fail(GFA_ECUDA, "Shell_1 PreCalc kernel failed");
    }
    *out = h;
    return GFA_OK;
}

int gfa_destroy(gfa_t* h) {
    if (!h) return GFA_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->stream_if) cudaStreamSynchronize(h->stream_if);
    if (h->stream_sc) { cudaStreamSynchronize(h->stream_sc); cudaStreamDestroy(h->stream_sc); }
    if (h->ev_ctl) cudaEventDestroy(h->ev_ctl);
    if (h->ev_scattered) cudaEventDestroy(h->ev_scattered);
    for (int i = 0; i < 4; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->ev_iface) cudaEventDestroy(h->ev_iface);
    if (h->ev_unpacked) cudaEventDestroy(h->ev_unpacked);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream_if) cudaStreamDestroy(h->stream_if);
    delete h;
    return GFA_OK;
}

// DOFsActive (Solution.cpp:121-224) + SetGlobalDOFs (:40-118) for nodes.
int gfa_number_dofs(gfa_t* h, const int32_t* cmask, int32_t* GLs, int32_t* n_free, int32_t* n_fixed) {
    if (!h || !GLs) return fail(GFA_EINVAL, "gfa_number_dofs: null argument");
    std::vector<unsigned char> active((size_t)h->n_nodes * 6, 0);
    for (int e = 0; e < h->n_el; e++) {
        const int s = type_slot(h->el_type[e]);
        for (int b = 0; b < kTypes[s].nb; b++) {
            int a, grp; block_node(s, b, a, grp);
            const size_t nd = (size_t)h->el_nodes[h->el_ptr[e] + a];
            for (int k = 0; k < 3; k++) active[6 * nd + 3 * grp + k] = 1;
        }
    }
    int nf = 0, nx = 0;
    for (int i = 0; i < h->n_nodes; i++)
        for (int k = 0; k < 6; k++) {
            int g = 0;
            if (active[6 * (size_t)i + k]) g = (cmask && ((cmask[i] >> k) & 1)) ? -(++nx) : ++nf;
            GLs[6 * (size_t)i + k] = g;
        }
    if (n_free) *n_free = nf;
    if (n_fixed) *n_fixed = nx;
    return GFA_OK;
}

int gfa_set_dofs(gfa_t* h, const int32_t* GLs, int32_t n_free, int32_t n_fixed,
                 int64_t n_extra, const int32_t* ex_mat, const int32_t* ex_rows, const int32_t* ex_cols) {
    if (!h || !GLs) return fail(GFA_EINVAL, "gfa_set_dofs: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    h->dofs_set = false; h->assembled = false; h->vec_dirty = false;
    h->d_owned_idx.release();
    if (!h->replaying) {          // registered against the old DOF map (a replay keeps the map: the load routes stay valid)
        h->shell_loads.n_loads = h->shell_loads.n_entries = 0; h->shell_loads.n_dest = 0;
        h->pipe_loads.n_loads = h->pipe_loads.n_entries = 0; h->pipe_loads.n_dest = 0;
    }
    h->n_free = n_free; h->n_fixed = n_fixed;
    h->gls.assign(GLs, GLs + 6 * (size_t)h->n_nodes);
    const std::vector<int>& gls = h->gls;
    for (size_t i = 0; i < gls.size(); i++)
        if (gls[i] > n_free || -gls[i] > n_fixed) return fail(GFA_EINVAL, "GLs[%zu] = %d outside [-%d, %d]", i, gls[i], n_fixed, n_free);
    {   // The pattern builder and the slot map rely on the reference's node-major numbering (Solution.cpp:53-72): free ids
        // ascend with (node, DOF) and the free DOFs of one 3-DOF group are consecutive ids, so that a group's rows are
        // consecutive CSR rows and columns ascend with the group-node id.  Anything else is refused, not mis-assembled.
        int last = 0;
        std::vector<unsigned char> seen_fixed((size_t)n_fixed, 0);
        for (size_t gn = 0; gn < gls.size() / 3; gn++) {
            int prev = 0;
            for (int k = 0; k < 3; k++) {
                const int g = gls[3 * gn + k];
                if (g > 0) {
                    if (g <= last) return fail(GFA_EUNSUPPORTED, "GLs are not in the reference's node-major ascending order at node %zu (free id %d after %d)", gn / 2 + 1, g, last);
                    if (prev && g != prev + 1) return fail(GFA_EUNSUPPORTED, "free ids %d and %d of node %zu are not consecutive (another DOF is numbered in between)", prev, g, gn / 2 + 1);
                    last = g; prev = g;
                } else if (g < 0) {
                    if (seen_fixed[(size_t)(-g - 1)]) return fail(GFA_EINVAL, "fixed id %d appears twice in GLs", g);
                    seen_fixed[(size_t)(-g - 1)] = 1;
                }
            }
        }
    }
    // keep the arguments: a handle that has to leave ring mode replays the call (from copies, see leave_ring_mode)
    h->ex_mat.assign(ex_mat, ex_mat + (n_extra > 0 ? n_extra : 0));
    h->ex_rows.assign(ex_rows, ex_rows + (n_extra > 0 ? n_extra : 0));
    h->ex_cols.assign(ex_cols, ex_cols + (n_extra > 0 ? n_extra : 0));

    const bool timing = getenv("GFA_SETUP_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gfa] set_dofs %-28s %7.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    // launchers such as torchrun pin OMP_NUM_THREADS to 1 per process: take this rank's share of the host cores instead
    const int host_threads = std::max(1, (int)std::thread::hardware_concurrency() / std::max(1, h->world));
    // ---- group-node adjacency over ALL elements (every rank builds the same pattern)
    const size_t n_gn_all = (size_t)h->n_nodes * 2;
    std::vector<int> gptr(n_gn_all + 1, 0);
    for (int e = 0; e < h->n_el; e++) {
        const int s = type_slot(h->el_type[e]);
        for (int b = 0; b < kTypes[s].nb; b++) {
            int a, grp; block_node(s, b, a, grp);
            gptr[(size_t)h->el_nodes[h->el_ptr[e] + a] * 2 + grp + 1]++;
        }
    }
    for (size_t i = 0; i < n_gn_all; i++) gptr[i + 1] += gptr[i];
    std::vector<int> ginc_e((size_t)gptr[n_gn_all]), ginc_b((size_t)gptr[n_gn_all]);
    {
        std::vector<int> fill(gptr.begin(), gptr.end() - 1);
        for (int e = 0; e < h->n_el; e++) {
            const int s = type_slot(h->el_type[e]);
            for (int b = 0; b < kTypes[s].nb; b++) {
                int a, grp; block_node(s, b, a, grp);
                const int p = fill[(size_t)h->el_nodes[h->el_ptr[e] + a] * 2 + grp]++;
                ginc_e[p] = e; ginc_b[p] = b;
            }
        }
    }
    auto free_mask = [&](size_t gn) { int mk = 0; for (int k = 0; k < 3; k++) if (gls[3 * gn + k] > 0) mk |= 1 << k; return mk; };
    auto fix_mask = [&](size_t gn) { int mk = 0; for (int k = 0; k < 3; k++) if (gls[3 * gn + k] < 0) mk |= 1 << k; return mk; };
    // note: gls is [node][6] = [group-node][3] with gn = node*2 + grp

    lap("adjacency");
    // ---- which rank evaluates which element (same rule as gfa_create) -----
    std::vector<int> el_rank;
    if (h->world > 1) {
        el_rank.resize(h->n_el);
        long long seen[3] = { 0, 0, 0 };
        for (int e = 0; e < h->n_el; e++) {
            const int s = type_slot(h->el_type[e]);
            const long long k = seen[s]++;
            int r = (int)((k * h->world) / std::max<long long>(h->type_count[s], 1));
            while (r > 0 && k < h->type_count[s] * r / h->world) r--;
            while (r < h->world - 1 && k >= h->type_count[s] * (r + 1) / h->world) r++;
            el_rank[e] = r;
        }
    }
    // group-nodes whose rows this rank stores: those its own elements touch
    std::vector<unsigned char> need(n_gn_all, 0);
    for (size_t gn = 0; gn < n_gn_all; gn++) {
        if (gptr[gn] == gptr[gn + 1]) continue;
        if (h->world == 1) { need[gn] = 1; continue; }
        for (int p = gptr[gn]; p < gptr[gn + 1]; p++) if (el_rank[ginc_e[p]] == h->rank) { need[gn] = 1; break; }
    }

    // ---- neighbour lists (sorted group-node ids), over ALL incident elements so
    //      that a stored row carries its complete global column set ------------
    std::vector<long long> nptr(n_gn_all + 1, 0);
    {
        std::vector<int> cnt(n_gn_all, 0);
#pragma omp parallel num_threads(host_threads)
        {
            std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096)
            for (long long gn = 0; gn < (long long)n_gn_all; gn++) {
                if (!need[gn]) continue;
                tmp.clear();
                for (int p = gptr[gn]; p < gptr[gn + 1]; p++) {
                    const int e = ginc_e[p], s = type_slot(h->el_type[e]);
                    for (int b = 0; b < kTypes[s].nb; b++) {
                        int a, grp; block_node(s, b, a, grp);
                        tmp.push_back(h->el_nodes[h->el_ptr[e] + a] * 2 + grp);
                    }
                }
                std::sort(tmp.begin(), tmp.end());
                cnt[gn] = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
            }
        }
        for (size_t i = 0; i < n_gn_all; i++) nptr[i + 1] = nptr[i] + cnt[i];
    }
    std::vector<int> nbr((size_t)nptr[n_gn_all]);
#pragma omp parallel num_threads(host_threads)
    {
        std::vector<int> tmp;
#pragma omp for schedule(dynamic, 4096)
        for (long long gn = 0; gn < (long long)n_gn_all; gn++) {
            if (!need[gn]) continue;
            tmp.clear();
            for (int p = gptr[gn]; p < gptr[gn + 1]; p++) {
                const int e = ginc_e[p], s = type_slot(h->el_type[e]);
                for (int b = 0; b < kTypes[s].nb; b++) {
                    int a, grp; block_node(s, b, a, grp);
                    tmp.push_back(h->el_nodes[h->el_ptr[e] + a] * 2 + grp);
                }
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            std::copy(tmp.begin(), tmp.end(), nbr.begin() + nptr[gn]);
        }
    }

    lap("neighbour lists");
    // ---- host positions outside the element pattern (Solution.cpp:268-281, 322-349: joints, contacts and other
    //      contributors push arbitrary triplets into the same lists).  An extra AA position whose column group is not a
    //      neighbour of its row group -- or that involves a DOF beyond the node table (Lagrange multipliers, super
    //      nodes) -- lengthens that row only.  Group-nodes with such rows are "irregular": their rows no longer share
    //      one layout, so they leave the patch slot map and are filled through explicit per-slot source lists (the
    //      mechanism of AB / BA / BB); every other group-node is untouched.
    std::vector<int> dof_gn((size_t)n_free, -1);             // free id -> group-node (-1: not a node DOF)
    for (size_t gn = 0; gn < n_gn_all; gn++)
        for (int k = 0; k < 3; k++) if (gls[3 * gn + k] > 0) dof_gn[(size_t)gls[3 * gn + k] - 1] = (int)gn;
    std::vector<std::pair<int, int> > xtra;                  // (row, column) pairs outside the element pattern, sorted, unique
    std::vector<unsigned char> irregular(n_gn_all, 0);
    for (int64_t i = 0; i < n_extra; i++) {
        if (ex_mat[i] != GFA_AA) continue;
        const int r = ex_rows[i], c = ex_cols[i];
        if (r < 0 || r >= n_free || c < 0 || c >= n_free) return fail(GFA_EINVAL, "extra pattern entry %lld out of range", (long long)i);
        const int gr = dof_gn[(size_t)r], gc = dof_gn[(size_t)c];
        bool inside = false;
        if (gr >= 0 && gc >= 0 && need[(size_t)gr]) inside = std::binary_search(nbr.begin() + nptr[gr], nbr.begin() + nptr[gr + 1], gc);
        if (!inside) xtra.emplace_back(r, c);
    }
    std::sort(xtra.begin(), xtra.end());
    xtra.erase(std::unique(xtra.begin(), xtra.end()), xtra.end());
    for (const std::pair<int, int>& rc : xtra) if (dof_gn[(size_t)rc.first] >= 0) irregular[(size_t)dof_gn[(size_t)rc.first]] = 1;
    // extra columns of a row that are not in its group's element columns
    auto row_extras = [&](int r) { return std::equal_range(xtra.begin(), xtra.end(), std::make_pair(r, 0), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first < y.first; }); };


    lap("extras");
    // ---- arena placement -------------------------------------------------------------------------
    // classic: every local element owns a region of one big arena (written by the evaluation kernel, read back
    // by the scatter kernel through DRAM).  ring: elements are evaluated chunk by chunk into an L2-resident ring
    // that the scatter role of the fused kernel drains behind the evaluation (gfa_device.h: FusedArgs).  Elements
    // whose blocks must outlive their chunk are PINNED to regions of their own behind the ring and evaluated
    // first: those that touch a fixed DOF (explicit AB/BA/BB gather lists), a partition interface (their rows
    // are scattered, packed and sent first) or a group-node whose incident elements lie more than the ring's
    // reach apart in evaluation order.
    std::vector<unsigned char> gn_iface(n_gn_all, 0);
    if (h->world > 1)
        for (size_t gn = 0; gn < n_gn_all; gn++) {
            unsigned long long rs = 0;
            for (int p = gptr[gn]; p < gptr[gn + 1]; p++) rs |= 1ULL << el_rank[ginc_e[p]];
            gn_iface[gn] = __builtin_popcountll(rs) > 1;
        }
    long long classic_doubles = 0;
    for (int s = 0; s < 3; s++) classic_doubles += type_region(s, h->tb[s].elems.size());
    {
        const char* env = getenv("GFA_RING");
        const char* ekb = getenv("GFA_RING_CHUNK_KB");
        const char* ek = getenv("GFA_RING_CHUNKS");
        const long long chunk_bytes = (ekb && atoll(ekb) > 0 ? atoll(ekb) : 7168) * 1024LL;
        int K = ek && atoi(ek) >= 3 ? atoi(ek) : 9;
        // 1: ring pipeline, 2: ring pipeline when the arena exceeds the ring; default 0 -- measured on B200 the thin
        // scatter kernel cannot keep up beside the register-hungry evaluation kernel (profiles/r02_notes.md), so the
        // classic two-kernel path stays the default until the evaluation kernel leaves more of the SM free
        const int mode = env ? atoi(env) : 0;
        h->ring_serial = mode == 3 || mode == 4;             // 3: serial ring always, 4: when the arena exceeds the ring
        if (const char* eg = getenv("GFA_RING_GROUP")) if (atoi(eg) >= 1) h->ring_group = atoi(eg);
        if (h->ring_serial && K < h->ring_group + 2) K = h->ring_group + 2;
        bool ring = !h->force_classic && mode != 0 && (mode == 1 || mode == 3 || classic_doubles * 8 > (long long)K * chunk_bytes);
        h->batch_layout = false;                             // set below when the classic arena is chosen
        h->ring_note = h->force_classic ? "classic: dynamics / explicit request" : mode == 0 ? "classic (GFA_RING=1 selects the ring pipeline)" : "classic: the element arena fits the ring";
        std::vector<std::vector<unsigned char> > pinned(3);
        int ce[3] = { 1, 1, 1 };
        for (int s = 0; s < 3; s++) {
            const int epw = s == 0 ? 8 : s == 1 ? 16 : 4;          // batch sizes of the evaluation kernels
            long long n = chunk_bytes / (8LL * arena_doubles(s)) / epw * epw;
            ce[s] = (int)std::max<long long>(n, epw);
            pinned[s].assign(h->tb[s].elems.size(), 0);
        }
        auto chunk_of = [&](const int* c0, int s, long long pos) { return c0[s] + (int)(pos / ce[s]); };
        if (ring) {
            int c0p[3]; long long acc = 0;                          // provisional chunks over all local elements
            for (int s = 0; s < 3; s++) { c0p[s] = (int)acc; acc += ((long long)h->tb[s].elems.size() + ce[s] - 1) / ce[s]; }
            const int reach = std::max(1, K / 3);
            for (size_t gn = 0; gn < n_gn_all; gn++) {
                int lo = 0x7fffffff, hi = -1; bool fixed = false;
                for (int k = 0; k < 3; k++) fixed |= gls[3 * gn + k] < 0;
                for (int p = gptr[gn]; p < gptr[gn + 1]; p++) {
                    const int e = ginc_e[p], s = h->el_owner_slot[e];
                    if (s < 0) continue;
                    const int c = chunk_of(c0p, s, h->el_local[e]);
                    lo = std::min(lo, c); hi = std::max(hi, c);
                }
                if (hi < 0) continue;
                if (fixed || gn_iface[gn] || irregular[gn] || hi - lo > reach)
                    for (int p = gptr[gn]; p < gptr[gn + 1]; p++) { const int e = ginc_e[p], s = h->el_owner_slot[e]; if (s >= 0) pinned[s][h->el_local[e]] = 1; }
            }
        }
        // lists, chunks, per-element offsets
        int c0[3] = { 0, 0, 0 }; long long tot_chunks = 0, chunk_doubles = 0, n_pinned = 0, n_ring = 0;
        for (int s = 0; s < 3; s++) {
            TypeBlock& t = h->tb[s];
            t.ring_list.clear(); t.pin_list.clear();
            for (size_t l = 0; l < t.elems.size(); l++) (ring && !pinned[s][l] ? t.ring_list : t.pin_list).push_back((int)l);
            if (!ring) t.pin_list.clear();
            t.chunk_el = ce[s]; t.chunk0 = c0[s] = (int)tot_chunks;
            t.n_chunks = (int)(((long long)t.ring_list.size() + ce[s] - 1) / ce[s]);
            tot_chunks += t.n_chunks;
            if (!t.ring_list.empty()) chunk_doubles = std::max(chunk_doubles, (long long)ce[s] * arena_doubles(s));
            n_pinned += (long long)t.pin_list.size(); n_ring += (long long)t.ring_list.size();
        }
        int span = 0;
        if (ring) {
            std::vector<int> fc[3];                                  // final chunk of a local element, -1 = pinned
            for (int s = 0; s < 3; s++) {
                fc[s].assign(h->tb[s].elems.size(), -1);
                for (size_t k = 0; k < h->tb[s].ring_list.size(); k++) fc[s][h->tb[s].ring_list[k]] = chunk_of(c0, s, (long long)k);
            }
            for (size_t gn = 0; gn < n_gn_all; gn++) {
                int lo = 0x7fffffff, hi = -1;
                for (int p = gptr[gn]; p < gptr[gn + 1]; p++) {
                    const int e = ginc_e[p], s = h->el_owner_slot[e];
                    if (s < 0 || fc[s][h->el_local[e]] < 0) continue;
                    lo = std::min(lo, fc[s][h->el_local[e]]); hi = std::max(hi, fc[s][h->el_local[e]]);
                }
                if (hi >= 0) span = std::max(span, hi - lo);
            }
            if (n_ring == 0) { ring = false; h->ring_note = "classic: every element is pinned (fixed DOFs / interfaces / numbering without locality)"; }
            else if (span > K - 2) { ring = false; h->ring_note = "classic: patches reach further than the ring"; }
        }
        h->ring = ring;
        long long total = 0;
        if (ring) {
            h->ring_chunks = K; h->ring_span = span; h->total_chunks = (int)tot_chunks; h->chunk_doubles = chunk_doubles;
            h->ring_doubles = (long long)K * chunk_doubles;
            total = h->ring_doubles;
            for (int s = 0; s < 3; s++) {
                TypeBlock& t = h->tb[s];
                t.pin_base = total; total += (long long)t.pin_list.size() * arena_doubles(s);
                t.ke_off.assign(t.elems.size(), 0);
                for (size_t k = 0; k < t.ring_list.size(); k++)
                    t.ke_off[t.ring_list[k]] = (long long)((c0[s] + (int)(k / ce[s])) % K) * chunk_doubles + (long long)(k % ce[s]) * arena_doubles(s);
                for (size_t k = 0; k < t.pin_list.size(); k++) t.ke_off[t.pin_list[k]] = t.pin_base + (long long)k * arena_doubles(s);
            }
            char buf[256];
            snprintf(buf, sizeof(buf), "%s: %d chunks of %.1f MB (%lld total), span %d, %lld elements through the ring, %lld pinned",
                     h->ring_serial ? "serial ring" : "ring", K, chunk_doubles * 8 / 1048576.0, tot_chunks, span, n_ring, n_pinned);
            h->ring_note = buf;
        } else {
            {   // GFA_ARENA_LAYOUT=0 keeps the compact per-element layout (experiments)
                const char* lay = getenv("GFA_ARENA_LAYOUT");
                h->batch_layout = !h->force_compact && shell_batch_layout_available() && !(lay && atoi(lay) == 0);
                if (!h->tb[0].elems.empty()) h->ring_note += h->batch_layout ? "; shell arena in batches of 8 elements" : "; compact shell arena";
            }
            h->ring_chunks = 0; h->ring_span = 0; h->total_chunks = 0; h->chunk_doubles = 0; h->ring_doubles = 0;
            for (int s = 0; s < 3; s++) {
                TypeBlock& t = h->tb[s];
                t.ring_list.clear(); t.pin_list.clear(); t.n_chunks = 0; t.pin_base = 0;
                t.ke_off.resize(t.elems.size());
                for (size_t l = 0; l < t.elems.size(); l++) t.ke_off[l] = t.ke_base + (long long)l * arena_doubles(s);
            }
            total = classic_doubles;
        }
        if (total >= (1LL << 32)) return fail(GFA_EUNSUPPORTED, "element arena of %lld doubles is too large for the 32-bit block offsets of the slot map%s", total, ring ? "" : " (classic path)");
        if (h->d_Ke.n != (size_t)total + 2) {
            cudaError_t e = h->d_Ke.alloc((size_t)total + 2);      // the scatter kernel stages whole 16-byte pieces: one past the last block
            if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? GFA_ENOMEM : GFA_ECUDA, "element arena: %s", cudaGetErrorString(e));
        }
    }
    // ready chunk of a group-node: the last chunk that holds one of its incident ring elements (-1: pinned sources only)
    auto ready_chunk = [&](size_t gn) {
        int rc = -1;
        if (!h->ring) return rc;
        for (int p = gptr[gn]; p < gptr[gn + 1]; p++) {
            const int e = ginc_e[p], s = h->el_owner_slot[e];
            if (s < 0) continue;
            const TypeBlock& t = h->tb[s];
            const long long off = t.ke_off[h->el_local[e]];
            if (off >= h->ring_doubles) continue;            // pinned
            const int* pos = std::lower_bound(t.ring_list.data(), t.ring_list.data() + t.ring_list.size(), h->el_local[e]);
            rc = std::max(rc, t.chunk0 + (int)((pos - t.ring_list.data()) / t.chunk_el));
        }
        return rc;
    };

    lap("placement");
    // ---- AA pattern: rows of a group-node share one column layout.  Stored
    //      rows are numbered in ascending global order (all rows when world == 1,
    //      so the single-GPU CSR is exactly the reference's). ------------------
    HostCsr& AA = h->csr[GFA_AA];
    AA.cols = n_free;
    AA.row_local.assign((size_t)n_free, -1);
    AA.row_ids.clear();
    if (h->world == 1) {
        AA.row_ids.resize((size_t)n_free);
        for (int r = 0; r < n_free; r++) { AA.row_ids[r] = r; AA.row_local[r] = r; }
    } else {
        for (size_t gn = 0; gn < n_gn_all; gn++)
            if (need[gn]) for (int k = 0; k < 3; k++) { const int g = gls[3 * gn + k]; if (g > 0) AA.row_ids.push_back(g - 1); }
        if (h->rank == 0)       // rows of DOFs beyond the node table exist through host positions only: rank 0 keeps them
            for (size_t i = 0; i < xtra.size(); i++)
                if (dof_gn[(size_t)xtra[i].first] < 0 && (i == 0 || xtra[i - 1].first != xtra[i].first)) AA.row_ids.push_back(xtra[i].first);
        std::sort(AA.row_ids.begin(), AA.row_ids.end());
        for (size_t i = 0; i < AA.row_ids.size(); i++) AA.row_local[AA.row_ids[i]] = (int)i;
    }
    AA.rows = (int)AA.row_ids.size();
    AA.rowptr.assign((size_t)AA.rows + 1, 0);
    for (size_t gn = 0; gn < n_gn_all; gn++) {
        if (!need[gn]) continue;
        long long L = 0;
        for (long long q = nptr[gn]; q < nptr[gn + 1]; q++) L += __builtin_popcount(free_mask((size_t)nbr[q]));
        for (int k = 0; k < 3; k++) { const int g = gls[3 * gn + k]; if (g > 0) AA.rowptr[AA.row_local[g - 1] + 1] = L; }
    }
    for (const std::pair<int, int>& rc : xtra) { const int lr = AA.row_local[(size_t)rc.first]; if (lr >= 0) AA.rowptr[(size_t)lr + 1]++; }
    for (int r = 0; r < AA.rows; r++) AA.rowptr[r + 1] += AA.rowptr[r];
    const long long nnzAA = AA.rowptr[AA.rows];
    if (nnzAA > 0x7fffffffLL) return fail(GFA_EUNSUPPORTED, "AA has %lld non-zeros on this rank; 32-bit CSR (PARDISO/Eigen int) cannot hold it", nnzAA);
    AA.inner.resize((size_t)nnzAA);
#pragma omp parallel for schedule(dynamic, 4096) num_threads(host_threads)
    for (long long gn = 0; gn < (long long)n_gn_all; gn++) {
        if (!need[gn]) continue;
        for (int k = 0; k < 3; k++) {
            const int g = gls[3 * gn + k];
            if (g <= 0) continue;
            int* out = AA.inner.data() + AA.rowptr[AA.row_local[g - 1]];
            if (!irregular[gn]) {
                for (long long q = nptr[gn]; q < nptr[gn + 1]; q++)
                    for (int c = 0; c < 3; c++) { const int gc = gls[3 * (size_t)nbr[q] + c]; if (gc > 0) *out++ = gc - 1; }
            } else {            // element columns merged with the row's host positions, ascending
                std::vector<int> cols;
                for (long long q = nptr[gn]; q < nptr[gn + 1]; q++)
                    for (int c = 0; c < 3; c++) { const int gc = gls[3 * (size_t)nbr[q] + c]; if (gc > 0) cols.push_back(gc - 1); }
                const auto ex = row_extras(g - 1);
                auto a = cols.begin(); auto b = ex.first;
                while (a != cols.end() || b != ex.second) {
                    if (b == ex.second || (a != cols.end() && *a < b->second)) *out++ = *a++;
                    else *out++ = (b++)->second;
                }
            }
        }
    }
    for (size_t i = 0; i < xtra.size(); i++) {      // rows of DOFs beyond the node table
        const int r = xtra[i].first, lr = AA.row_local[(size_t)r];
        if (dof_gn[(size_t)r] >= 0 || lr < 0 || (i > 0 && xtra[i - 1].first == r)) continue;
        int* out = AA.inner.data() + AA.rowptr[lr];
        for (size_t j = i; j < xtra.size() && xtra[j].first == r; j++) *out++ = xtra[j].second;
    }

    lap("AA pattern");
    // ---- AB / BA / BB: explicit entry lists of elements that touch a fixed DOF
    struct Ent { int mat, row, col; long long src; int rank; };
    std::vector<Ent> ents;      // pattern from all elements; src = -1 when the element is not this rank's
    std::vector<Ent> aa_ents;   // AA slots of rows that carry host positions: element sources, element-ascending
    for (int e = 0; e < h->n_el; e++) {
        const int s = type_slot(h->el_type[e]);
        const TypeInfo& ti = kTypes[s];
        int gl[27]; bool any_fixed = false;
        for (int b = 0; b < ti.nb; b++) {
            int a, grp; block_node(s, b, a, grp);
            const size_t gn = (size_t)h->el_nodes[h->el_ptr[e] + a] * 2 + grp;
            for (int k = 0; k < 3; k++) { gl[3 * b + k] = gls[3 * gn + k]; any_fixed |= gl[3 * b + k] < 0; }
        }
        bool any_irregular = false;
        for (int b = 0; b < ti.nb; b++) {
            int a, grp; block_node(s, b, a, grp);
            any_irregular |= irregular[(size_t)h->el_nodes[h->el_ptr[e] + a] * 2 + grp] != 0;
        }
        if (!any_fixed && !any_irregular) continue;
        const bool mine = h->el_owner_slot[e] >= 0;
        for (int i = 0; i < ti.ndof; i++)
            for (int j = 0; j < ti.ndof; j++) {
                const int g1 = gl[i], g2 = gl[j];
                if (g1 > 0 && g2 > 0) {
                    // free x free: only the rows of irregular group-nodes are filled through explicit lists
                    if (!mine || !irregular[(size_t)dof_gn[(size_t)g1 - 1]] || AA.row_local[(size_t)g1 - 1] < 0) continue;
                    bool tr;
                    const long long blk = arena_block(h, s, h->el_local[e], i / 3, j / 3, tr);
                    aa_ents.push_back({ GFA_AA, g1 - 1, g2 - 1, blk + (tr ? (j % 3) * 3 + (i % 3) : (i % 3) * 3 + (j % 3)), -1 });
                    continue;
                }
                if (g1 == 0 || g2 == 0) continue;
                Ent en;
                en.mat = (g1 > 0) ? GFA_AB : (g2 > 0 ? GFA_BA : GFA_BB);
                en.row = std::abs(g1) - 1; en.col = std::abs(g2) - 1;
                en.src = -1;
                if (mine) {
                    bool tr;
                    const long long blk = arena_block(h, s, h->el_local[e], i / 3, j / 3, tr);
                    en.src = blk + (tr ? (j % 3) * 3 + (i % 3) : (i % 3) * 3 + (j % 3));
                }
                en.rank = h->world > 1 ? el_rank[e] : 0;
                ents.push_back(en);
            }
    }
    // extra host positions of the small matrices are taken as they come (AA: see the pattern above)
    for (int64_t i = 0; i < n_extra; i++) {
        const int w = ex_mat[i], r = ex_rows[i], c = ex_cols[i];
        if (w < 0 || w > 3) return fail(GFA_EINVAL, "extra pattern entry %lld: matrix %d", (long long)i, w);
        const int nr = (w == GFA_AA || w == GFA_AB) ? n_free : n_fixed, nc = (w == GFA_AA || w == GFA_BA) ? n_free : n_fixed;
        if (r < 0 || r >= nr || c < 0 || c >= nc) return fail(GFA_EINVAL, "extra pattern entry %lld out of range", (long long)i);
        if (w == GFA_AA) continue;           // in the element pattern already, or added to its row above
        { Ent en; en.mat = w; en.row = r; en.col = c; en.src = -1; en.rank = -1; ents.push_back(en); }
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& x, const Ent& y) {
        if (x.mat != y.mat) return x.mat < y.mat;
        if (x.row != y.row) return x.row < y.row;
        return x.col < y.col;
    });
    for (int w = 1; w < 4; w++) {
        HostCsr& M = h->csr[w];
        M.rows = (w == GFA_AB) ? n_free : n_fixed; M.cols = (w == GFA_BA) ? n_free : n_fixed;
        M.rowptr.assign((size_t)M.rows + 1, 0); M.inner.clear();
    }
    std::vector<long long> gseg, gsrc, gdest;
    std::vector<unsigned long long> granks;   // per unique dest: set of ranks whose elements contribute
    {
        size_t i = 0;
        while (i < ents.size()) {
            size_t j = i;
            HostCsr& M = h->csr[ents[i].mat];
            const long long slot = (long long)M.inner.size();
            M.inner.push_back(ents[i].col);
            M.rowptr[ents[i].row + 1]++;
            gseg.push_back((long long)gsrc.size());
            gdest.push_back(((long long)ents[i].mat << 56) | slot);   // patched to arena offsets below
            unsigned long long rs = 0;
            while (j < ents.size() && ents[j].mat == ents[i].mat && ents[j].row == ents[i].row && ents[j].col == ents[i].col) {
                if (ents[j].src >= 0) gsrc.push_back(ents[j].src);
                if (ents[j].rank >= 0) rs |= 1ULL << ents[j].rank;
                j++;
            }
            granks.push_back(rs);
            i = j;
        }
        // AA rows that carry host positions: every slot of the row is rewritten per assembly (zero when no element feeds it)
        for (size_t gn = 0; gn < n_gn_all; gn++) {
            if (!irregular[gn]) continue;
            for (int k = 0; k < 3; k++) {
                const int g = gls[3 * gn + k];
                if (g <= 0 || AA.row_local[(size_t)g - 1] < 0) continue;
                const int lr = AA.row_local[(size_t)g - 1];
                for (long long q = AA.rowptr[lr]; q < AA.rowptr[lr + 1]; q++) aa_ents.push_back({ GFA_AA, g - 1, AA.inner[(size_t)q], -1, -1 });
            }
        }
        for (size_t i = 0; i < xtra.size(); i++) {
            const int r = xtra[i].first;
            if (dof_gn[(size_t)r] < 0 && AA.row_local[(size_t)r] >= 0) aa_ents.push_back({ GFA_AA, r, xtra[i].second, -1, -1 });
        }
        std::stable_sort(aa_ents.begin(), aa_ents.end(), [](const Ent& x, const Ent& y) { return x.row != y.row ? x.row < y.row : x.col < y.col; });
        for (size_t a = 0; a < aa_ents.size();) {
            size_t b = a;
            const int lr = AA.row_local[(size_t)aa_ents[a].row];
            const int* c0 = AA.inner.data() + AA.rowptr[lr]; const int* c1 = AA.inner.data() + AA.rowptr[lr + 1];
            const long long slot = AA.rowptr[lr] + (std::lower_bound(c0, c1, aa_ents[a].col) - c0);
            gseg.push_back((long long)gsrc.size());
            gdest.push_back(slot);                            // matrix 0: the AA values start the arena
            granks.push_back(0ULL);
            for (; b < aa_ents.size() && aa_ents[b].row == aa_ents[a].row && aa_ents[b].col == aa_ents[a].col; b++)
                if (aa_ents[b].src >= 0) gsrc.push_back(aa_ents[b].src);
            a = b;
        }
        gseg.push_back((long long)gsrc.size());
    }
    for (int w = 1; w < 4; w++) { HostCsr& M = h->csr[w]; for (int r = 0; r < M.rows; r++) M.rowptr[r + 1] += M.rowptr[r]; }

    lap("fixed-DOF entries");
    // ---- value arena: [AA | AB | BA | BB | P_A | I_A | P_B] ----------------
    long long off = 0;
    for (int w = 0; w < 4; w++) { h->arena_off[w] = off; off += (long long)h->csr[w].inner.size(); }
    h->vec_off[GFA_P_A] = off; off += n_free;
    h->vec_off[GFA_I_A] = off; off += n_free;
    h->vec_off[GFA_P_B] = off; off += n_fixed;
    h->arena_size = off;
    for (size_t i = 0; i < gdest.size(); i++) {
        const int w = (int)(gdest[i] >> 56);
        gdest[i] = h->arena_off[w] + (gdest[i] & 0x00ffffffffffffffLL);
    }
    h->n_gdest = (long long)gdest.size();
    // AB/BA/BB slots fed by several ranks: partial sums travel to the lowest contributing rank
    std::vector<std::vector<long long> > send_small(h->world), recv_small(h->world);
    if (h->world > 1)
        for (size_t i = 0; i < gdest.size(); i++) {
            const unsigned long long rs = granks[i];
            if (__builtin_popcountll(rs) < 2 || !((rs >> h->rank) & 1)) continue;
            const int owner = __builtin_ctzll(rs);
            if (owner == h->rank) { for (int r = 0; r < h->world; r++) if (r != h->rank && ((rs >> r) & 1)) recv_small[r].push_back(gdest[i]); }
            else send_small[owner].push_back(gdest[i]);
        }

    lap("arena offsets");
    // ---- interface ownership (owner = lowest rank with an incidence), ascending group-node
    //      order on every rank so that the send and receive lists pair up ----------------------
    std::vector<std::vector<long long> > send_idx(h->world), recv_idx(h->world);
    h->owned_rows.clear();
    std::vector<int> rowL(n_gn_all, 0);              // CSR row length of the group's rows
    std::vector<int> touched_gn;                     // group-nodes with a local incidence
    std::vector<long long> touched_key;              // position of the first local incident element (locality key)
    std::vector<char> touched_iface;                 // shared with another rank: scattered first (see gfa_assemble)
    std::vector<int> touched_rc;                     // ring mode: chunk after which the group-node's patches are complete
    for (size_t gn = 0; gn < n_gn_all; gn++) {
        if (gptr[gn] == gptr[gn + 1]) continue;
        int owner = 0; bool touched = h->world == 1; unsigned long long rank_set = 0;
        if (h->world > 1) {
            owner = h->world;
            for (int p = gptr[gn]; p < gptr[gn + 1]; p++) { const int r = el_rank[ginc_e[p]]; rank_set |= 1ULL << r; owner = std::min(owner, r); if (r == h->rank) touched = true; }
        }
        // AA row geometry shared by the group's rows
        long long L = 0;
        if (need[gn]) for (long long q = nptr[gn]; q < nptr[gn + 1]; q++) L += __builtin_popcount(free_mask((size_t)nbr[q]));
        if (L > 0xffff) return fail(GFA_EUNSUPPORTED, "a row of AA has %lld entries; the slot map holds 16-bit row strides", L);
        rowL[gn] = (int)L;
        if (owner == h->rank || h->world == 1)
            for (int k = 0; k < 3; k++) if (gls[3 * gn + k] > 0) h->owned_rows.push_back(gls[3 * gn + k] - 1);
        if (h->world > 1 && touched && __builtin_popcountll(rank_set) > 1) {
            // values exchanged for this group-node: its free AA rows, and its P_A / I_A / P_B entries
            auto list_for = [&](std::vector<long long>& dst) {
                for (int k = 0; k < 3; k++) {
                    const int g = gls[3 * gn + k];
                    if (g > 0) {
                        const int lr = AA.row_local[g - 1];
                        for (long long p = AA.rowptr[lr]; p < AA.rowptr[lr + 1]; p++) dst.push_back(h->arena_off[GFA_AA] + p);
                        dst.push_back(h->vec_off[GFA_P_A] + g - 1);
                        dst.push_back(h->vec_off[GFA_I_A] + g - 1);
                    } else if (g < 0) dst.push_back(h->vec_off[GFA_P_B] + (-g - 1));
                }
            };
            if (owner == h->rank) { for (int r = 0; r < h->world; r++) if (r != h->rank && ((rank_set >> r) & 1)) list_for(recv_idx[r]); }
            else list_for(send_idx[owner]);
        }
        if (!touched) continue;
        long long key = -1;
        for (int p = gptr[gn]; p < gptr[gn + 1] && key < 0; p++) {
            const int e = ginc_e[p];
            if (h->el_owner_slot[e] >= 0) key = (long long)h->tb[h->el_owner_slot[e]].pe_base + (long long)h->el_local[e] * kTypes[h->el_owner_slot[e]].ndof;
        }
        if (key < 0) continue;
        touched_gn.push_back((int)gn);
        touched_key.push_back(key);
        touched_iface.push_back(h->world > 1 && __builtin_popcountll(rank_set) > 1 ? 1 : 0);
        touched_rc.push_back(ready_chunk(gn));
    }
    if (h->rank == 0)
        for (size_t i = 0; i < xtra.size(); i++)
            if (dof_gn[(size_t)xtra[i].first] < 0 && (i == 0 || xtra[i - 1].first != xtra[i].first)) h->owned_rows.push_back(xtra[i].first);
    std::sort(h->owned_rows.begin(), h->owned_rows.end());
    {   // patches are processed in the order of their group-node's first incident element, so that the
        // blocks read next to each other were written next to each other (node-id order would visit an
        // element's corner and mid-side nodes millions of group-nodes apart)
        std::vector<size_t> order(touched_gn.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        // group-nodes on a partition interface come first: their rows are complete (and can be packed and sent)
        // while the interior rows are still being scattered
        // (ring mode: then by ready chunk -- the fused kernel scatters a chunk's group-nodes as soon as it is evaluated)
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) {
            if (touched_iface[x] != touched_iface[y]) return touched_iface[x] > touched_iface[y];
            if (touched_rc[x] != touched_rc[y]) return touched_rc[x] < touched_rc[y];
            return touched_key[x] < touched_key[y];
        });
        std::vector<int> sorted(touched_gn.size()), sorted_rc(touched_gn.size());
        for (size_t i = 0; i < order.size(); i++) { sorted[i] = touched_gn[order[i]]; sorted_rc[i] = touched_rc[order[i]]; }
        touched_gn.swap(sorted); touched_rc.swap(sorted_rc);
    }
    size_t n_iface_touched = 0;
    for (char c : touched_iface) n_iface_touched += c ? 1 : 0;

    lap("ownership + ordering");
    // ---- slot map of this rank's group-nodes ----------------------------------
    // For every (group-node, neighbour) patch: the local element blocks feeding it, element-ascending.  Built in
    // parallel: contiguous pieces of the (ordered) group-node list go to per-piece buffers that are concatenated in
    // order afterwards, indices shifted by the pieces' offsets.
    std::vector<PInc> incs;
    std::vector<RunEnt> runs;
    std::vector<unsigned long long> ovf;
    std::vector<GnRec> gn_recs(touched_gn.size());
    std::vector<int> run_count(touched_gn.size(), 0);        // patches emitted per group-node
    h->n_iface_runs = 0; h->n_iface_gn = 0;
    std::vector<long long> chunk_run_ptr((size_t)h->total_chunks + 1, 0);
    {
        const size_t T = touched_gn.size();
        const int n_pieces = (int)std::max<size_t>(1, std::min<size_t>((size_t)host_threads * 8, (T + 4095) / 4096));
        struct Piece { std::vector<PInc> incs; std::vector<RunEnt> runs; std::vector<unsigned long long> ovf; std::vector<size_t> ovf_runs; int err = 0; };
        std::vector<Piece> pieces((size_t)n_pieces);
#pragma omp parallel for schedule(dynamic, 1) num_threads(host_threads)
        for (int pc = 0; pc < n_pieces; pc++) {
            Piece& P = pieces[(size_t)pc];
            std::vector<std::vector<unsigned long long> > run_src;   // scratch: sources per patch of the current group-node
            const size_t t0 = T * (size_t)pc / (size_t)n_pieces, t1 = T * (size_t)(pc + 1) / (size_t)n_pieces;
            for (size_t ti_ = t0; ti_ < t1 && !P.err; ti_++) {
                const size_t gn = (size_t)touched_gn[ti_];
                const int first_inc = (int)P.incs.size();
                const int* nb0 = nbr.data() + nptr[gn]; const int* nb1 = nbr.data() + nptr[gn + 1];
                const int n_runs = (int)(nb1 - nb0);
                if ((int)run_src.size() < n_runs) run_src.resize(n_runs);
                for (int j = 0; j < n_runs; j++) run_src[j].clear();
                for (int p = gptr[gn]; p < gptr[gn + 1]; p++) {
                    const int e = ginc_e[p];
                    if (h->el_owner_slot[e] < 0) continue;
                    const int sl = h->el_owner_slot[e], local = h->el_local[e];
                    const TypeInfo& ti = kTypes[sl];
                    const int la = ginc_b[p];
                    PInc in;
                    in.pe_off = h->tb[sl].pe_base + local * ti.ndof;
                    in.la = la;
                    P.incs.push_back(in);
                    for (int bb = 0; bb < ti.nb; bb++) {
                        int an, grp; block_node(sl, bb, an, grp);
                        const int other = h->el_nodes[h->el_ptr[e] + an] * 2 + grp;
                        const int j = (int)(std::lower_bound(nb0, nb1, other) - nb0);
                        bool tr;
                        const long long blk = arena_block(h, sl, local, la, bb, tr);
                        if (blk >= (1LL << 32)) { P.err = 1; break; }
                        run_src[j].push_back((unsigned long long)blk | (tr ? SRC_T : 0ULL));
                    }
                }
                GnRec rec;
                int rm = 0, first_row = -1;
                for (int k = 0; k < 3; k++) {
                    const int g = gls[3 * gn + k];
                    rec.gl[k] = g;
                    if (g > 0) { rm |= 1 << k; if (first_row < 0) first_row = AA.row_local[g - 1]; }
                }
                rec.ib = first_inc; rec.ie = (int)P.incs.size();      // piece-local: shifted below
                gn_recs[ti_] = rec;
                if (!rm || irregular[gn]) continue;      // rows with host positions are filled through the explicit lists
                const long long row0 = AA.rowptr[first_row];
                int col = 0;
                const size_t runs_before = P.runs.size();
                for (int j = 0; j < n_runs; j++) {
                    const int fm = free_mask((size_t)nb0[j]);
                    const std::vector<unsigned long long>& src = run_src[j];
                    const long long dst = row0 + col;
                    col += __builtin_popcount(fm);
                    if (!fm) continue;
                    if (src.empty() && h->world == 1) continue;      // cannot happen: every stored patch has a local source
                    if (src.size() > 255) { P.err = 2; break; }
                    RunEnt r;
                    r.dst = (int)dst;
                    r.info = (unsigned)rowL[gn] | ((unsigned)rm << 16) | ((unsigned)fm << 19) | ((unsigned)src.size() << 24);
                    r.src0 = src.size() > 0 ? (unsigned)src[0] : 0u; r.src1 = src.size() > 1 ? (unsigned)src[1] : 0u;
                    if (src.size() > 0 && (src[0] & SRC_T)) r.info |= 1u << 22;
                    if (src.size() > 1 && (src[1] & SRC_T)) r.info |= 1u << 23;
                    if (src.size() > 2) {
                        r.src0 = (unsigned)P.ovf.size();                  // piece-local: shifted below
                        P.ovf_runs.push_back(P.runs.size());
                        P.ovf.insert(P.ovf.end(), src.begin(), src.end());
                    }
                    P.runs.push_back(r);
                }
                run_count[ti_] = (int)(P.runs.size() - runs_before);
            }
        }
        for (const Piece& P : pieces) {
            if (P.err == 1) return fail(GFA_EUNSUPPORTED, "element arena too large for the 32-bit block offsets of the slot map");
            if (P.err == 2) return fail(GFA_EUNSUPPORTED, "group-node with more than 255 incident elements");
        }
        std::vector<size_t> inc_off((size_t)n_pieces + 1, 0), run_off((size_t)n_pieces + 1, 0), ovf_off((size_t)n_pieces + 1, 0);
        for (int pc = 0; pc < n_pieces; pc++) {
            inc_off[(size_t)pc + 1] = inc_off[(size_t)pc] + pieces[(size_t)pc].incs.size();
            run_off[(size_t)pc + 1] = run_off[(size_t)pc] + pieces[(size_t)pc].runs.size();
            ovf_off[(size_t)pc + 1] = ovf_off[(size_t)pc] + pieces[(size_t)pc].ovf.size();
        }
        if (inc_off[(size_t)n_pieces] > 0x7fffffffULL) return fail(GFA_EUNSUPPORTED, "more than 2^31 (group-node, element) incidences on one rank");
        if (ovf_off[(size_t)n_pieces] >= (1ULL << 32)) return fail(GFA_EUNSUPPORTED, "overflow source list exceeds 2^32 entries");
        incs.resize(inc_off[(size_t)n_pieces]); runs.resize(run_off[(size_t)n_pieces]); ovf.resize(ovf_off[(size_t)n_pieces]);
#pragma omp parallel for schedule(dynamic, 1) num_threads(host_threads)
        for (int pc = 0; pc < n_pieces; pc++) {
            Piece& P = pieces[(size_t)pc];
            for (size_t r : P.ovf_runs) P.runs[r].src0 += (unsigned)ovf_off[(size_t)pc];
            std::copy(P.incs.begin(), P.incs.end(), incs.begin() + inc_off[(size_t)pc]);
            std::copy(P.runs.begin(), P.runs.end(), runs.begin() + run_off[(size_t)pc]);
            std::copy(P.ovf.begin(), P.ovf.end(), ovf.begin() + ovf_off[(size_t)pc]);
            const size_t t0 = T * (size_t)pc / (size_t)n_pieces, t1 = T * (size_t)(pc + 1) / (size_t)n_pieces;
            for (size_t ti_ = t0; ti_ < t1; ti_++) { gn_recs[ti_].ib += (int)inc_off[(size_t)pc]; gn_recs[ti_].ie += (int)inc_off[(size_t)pc]; }
        }
        // interface prefix and, in ring mode, the first run of every ready chunk
        long long run_prefix = 0;
        int crp_filled = 0;                              // chunk_run_ptr[0 .. crp_filled) are final
        for (size_t ti_ = 0; ti_ < T; ti_++) {
            if (ti_ == n_iface_touched) { h->n_iface_runs = run_prefix; h->n_iface_gn = (long long)ti_; }
            if (h->ring) for (; crp_filled <= touched_rc[ti_]; crp_filled++) chunk_run_ptr[crp_filled] = run_prefix;
            run_prefix += run_count[ti_];
        }
        if (n_iface_touched == T) { h->n_iface_runs = (long long)runs.size(); h->n_iface_gn = (long long)T; }
        for (; crp_filled <= h->total_chunks; crp_filled++) chunk_run_ptr[crp_filled] = (long long)runs.size();
    }
    (void)fix_mask;
    {   // Solution::Clear zeroes the whole vectors (Solution.cpp:833-848); entries no local element writes -- DOFs beyond
        // the node table, rows of other ranks -- are zeroed at the head of every assembly when there are any
        long long covered = 0;
        for (const GnRec& r : gn_recs) for (int k = 0; k < 3; k++) covered += r.gl[k] != 0;
        h->n_vec_untouched = (long long)n_free + n_fixed - covered;
    }
    h->n_pre_runs = h->ring ? chunk_run_ptr[0] : (long long)runs.size();
    if (h->ring) {
        std::vector<int> chunk_tile_ptr((size_t)h->total_chunks + 1, 0);
        for (int c = 0; c < h->total_chunks; c++) {
            const long long tiles = (chunk_run_ptr[c + 1] - chunk_run_ptr[c] + FUSED_TILE_PATCHES - 1) / FUSED_TILE_PATCHES;
            if (chunk_tile_ptr[c] + tiles > 0x7fffffffLL) return fail(GFA_EUNSUPPORTED, "too many scatter tiles");
            chunk_tile_ptr[c + 1] = chunk_tile_ptr[c] + (int)tiles;
        }
        CUDA_TRY(h->d_chunk_run_ptr.upload(chunk_run_ptr));
        h->chunk_run_ptr_h = chunk_run_ptr;
        CUDA_TRY(h->d_chunk_tile_ptr.upload(chunk_tile_ptr));
        std::vector<int> chunk_batches((size_t)h->total_chunks, 0);
        for (int sl = 0; sl < 3; sl++) {
            const TypeBlock& t = h->tb[sl];
            const int epw = sl == 0 ? 8 : sl == 1 ? 16 : 4;
            for (int c = 0; c < t.n_chunks; c++) {
                const long long el = std::min<long long>(t.chunk_el, (long long)t.ring_list.size() - (long long)c * t.chunk_el);
                chunk_batches[(size_t)t.chunk0 + c] = (int)((el + epw - 1) / epw);
            }
        }
        CUDA_TRY(h->d_chunk_batches.upload(chunk_batches));
        CUDA_TRY(h->d_ctl.alloc((size_t)CTL_HDR + 3 * (size_t)h->total_chunks + 8));
        for (int sl = 0; sl < 3; sl++) {
            CUDA_TRY(h->tb[sl].d_ring_list.upload(h->tb[sl].ring_list));
            CUDA_TRY(h->tb[sl].d_pin_list.upload(h->tb[sl].pin_list));
        }
        if (h->d_scratch_ke.n == 0) { CUDA_TRY(h->d_scratch_ke.alloc(SHELL_ARENA)); CUDA_TRY(h->d_one.alloc(1)); }
        for (int sl = 0; sl < 3; sl++) {
            const char* ew = getenv("GFA_FUSED_EVAL_WARPS");
            int nb = fused_buffers(sl);
            if (ew && atoi(ew) >= 1 && atoi(ew) <= fused_buffers(sl)) nb = atoi(ew);
            h->fused_eval_warps[sl] = nb;
        }
        if (const char* tg = getenv("GFA_FUSED_TILE_GROUP")) if (atoi(tg) >= 1 && atoi(tg) <= 64) h->fused_tile_group = atoi(tg);
        if (const char* sc = getenv("GFA_FUSED_SCATTER_CTAS")) if (atoi(sc) >= 1 && atoi(sc) <= 8) h->fused_scatter_ctas = atoi(sc);
        if (const char* tm = getenv("GFA_FUSED_TIMEOUT_MS")) if (atoll(tm) > 0) h->fused_timeout_ns = 1000000ULL * (unsigned long long)atoll(tm);
        // keep the ring in L2: persisting access-policy window over it on the library's stream
        const char* pe = getenv("GFA_RING_PERSIST");
        if (!pe || atoi(pe) != 0) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, h->device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
            const size_t ring_bytes = (size_t)h->ring_doubles * sizeof(double);
            if (max_persist > 0 && max_window > 0) {
                const size_t carve = std::min(ring_bytes, (size_t)max_persist);
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
                cudaStreamAttrValue av;
                std::memset(&av, 0, sizeof(av));
                av.accessPolicyWindow.base_ptr = h->d_Ke.p;
                av.accessPolicyWindow.num_bytes = std::min(ring_bytes, (size_t)max_window);
                av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)av.accessPolicyWindow.num_bytes);
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
                char buf[96];
                snprintf(buf, sizeof(buf), "; L2 window %.0f MB persisting (carve-out %.0f MB)", av.accessPolicyWindow.num_bytes / 1048576.0, carve / 1048576.0);
                h->ring_note += buf;
            }
        }
    } else {
        cudaStreamAttrValue av;
        std::memset(&av, 0, sizeof(av));
        if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
    }

    lap("slot map");
    // ---- uploads ----------------------------------------------------------
    CUDA_TRY(h->d_arena.alloc((size_t)h->arena_size));
    CUDA_TRY(cudaMemset(h->d_arena.p, 0, (size_t)h->arena_size * sizeof(double)));
    CUDA_TRY(h->d_gn.upload(gn_recs));
    CUDA_TRY(h->d_runs.upload(runs));
    if (ovf.empty()) ovf.push_back(0);
    CUDA_TRY(h->d_ovf.upload(ovf));
    CUDA_TRY(h->d_inc.upload(incs));
    h->n_runs = (long long)runs.size();
    h->n_gn_local = (long long)gn_recs.size();
    CUDA_TRY(h->d_gseg.upload(gseg));
    CUDA_TRY(h->d_gsrc.upload(gsrc));
    CUDA_TRY(h->d_gdest.upload(gdest));
    h->send_cnt.assign(h->world, 0); h->recv_cnt.assign(h->world, 0);
    {
        std::vector<long long> s_all, r_all;
        for (int r = 0; r < h->world; r++) {
            send_idx[r].insert(send_idx[r].end(), send_small[r].begin(), send_small[r].end());
            recv_idx[r].insert(recv_idx[r].end(), recv_small[r].begin(), recv_small[r].end());
            h->send_cnt[r] = (long long)send_idx[r].size(); h->recv_cnt[r] = (long long)recv_idx[r].size();
            s_all.insert(s_all.end(), send_idx[r].begin(), send_idx[r].end());
            r_all.insert(r_all.end(), recv_idx[r].begin(), recv_idx[r].end());
        }
        CUDA_TRY(h->d_send_idx.upload(s_all));
        CUDA_TRY(h->d_recv_idx.upload(r_all));
    }
    {   // Newton-loop vector steps
        CUDA_TRY(h->d_gls.upload(h->gls));
        if (h->world > 1) {
            std::vector<unsigned char> own((size_t)n_free, 0);
            for (int r : h->owned_rows) own[(size_t)r] = 1;
            std::vector<int> g2(h->gls);
            for (size_t i = 0; i < g2.size(); i++) if (g2[i] > 0 && !own[(size_t)g2[i] - 1]) g2[i] = 0;
            CUDA_TRY(h->d_gls_owned.upload(g2));
        }
        const HostCsr& AB = h->csr[GFA_AB];
        // rows of AB that hold entries; CSR storage is contiguous, so row rows[k] spans [ptr[k], ptr[k+1]) of
        // the value array when the empty rows in between are skipped
        std::vector<int> rows, ptr, inner(AB.inner.begin(), AB.inner.end());
        for (int r = 0; r < AB.rows; r++)
            if (AB.rowptr[r + 1] > AB.rowptr[r]) { rows.push_back(r); ptr.push_back((int)AB.rowptr[r]); }
        ptr.push_back((int)AB.inner.size());
        h->n_ab_rows = (int)rows.size();
        if (rows.empty()) { rows.push_back(0); inner.push_back(0); }
        CUDA_TRY(h->d_ab_rows.upload(rows));
        CUDA_TRY(h->d_ab_ptr.upload(ptr));
        CUDA_TRY(h->d_ab_inner.upload(inner));
        CUDA_TRY(h->d_norm.alloc(1));
    }
    {   // Dynamic::UpdateDyn (Dynamic.cpp:493-556): nodes whose rotational DOFs are partly free see the values
        // earlier nodes left in vel_aux / ace_aux; each replay starts at the last node that overwrote all three
        std::vector<int> mixed, start;
        int last_full = 0;
        for (int i = 0; i < h->n_nodes; i++) {
            const int* g = &gls[6 * (size_t)i];
            const int nf = (g[3] > 0) + (g[4] > 0) + (g[5] > 0);
            if (nf == 3) last_full = i;
            else if (nf > 0) { mixed.push_back(i); start.push_back(last_full); }
        }
        h->n_mixed = (int)mixed.size();
        if (h->n_mixed) { CUDA_TRY(h->d_mixed.upload(mixed)); CUDA_TRY(h->d_mixed_start.upload(start)); }
    }
    lap("uploads");
    h->dofs_set = true;
    return GFA_OK;
}

int gfa_csr_dims(gfa_t* h, int which, int32_t* rows, int32_t* cols, int64_t* nnz) {
    if (!h || which < 0 || which > 3) return fail(GFA_EINVAL, "gfa_csr_dims: bad argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_csr_dims before gfa_set_dofs");
    if (rows) *rows = h->csr[which].rows;
    if (cols) *cols = h->csr[which].cols;
    if (nnz) *nnz = (int64_t)h->csr[which].inner.size();
    return GFA_OK;
}

int gfa_csr_pattern(gfa_t* h, int which, int32_t* outer, int32_t* inner) {
    if (!h || which < 0 || which > 3) return fail(GFA_EINVAL, "gfa_csr_pattern: bad argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_csr_pattern before gfa_set_dofs");
    const HostCsr& M = h->csr[which];
    if (outer) for (int r = 0; r <= M.rows; r++) outer[r] = (int32_t)M.rowptr[r];
    if (inner && !M.inner.empty()) std::memcpy(inner, M.inner.data(), M.inner.size() * sizeof(int));
    return GFA_OK;
}

} // extern "C"

namespace {

void resolve_timing(gfa_t* h) {
    if (!h->timing_pending) return;
    cudaEventSynchronize(h->ev[3]);
    cudaEventElapsedTime(&h->last_ms[0], h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&h->last_ms[1], h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&h->last_ms[2], h->ev[2], h->ev[3]);
    cudaEventElapsedTime(&h->last_ms[3], h->ev[0], h->ev[3]);
    h->timing_pending = false;
}

int ensure_kinematics(gfa_t* h) {
    const size_t n = 6 * (size_t)h->n_nodes;
    if (h->d_vel.n == n) return GFA_OK;
    DevBuf<double>* bufs[4] = { &h->d_vel, &h->d_accel, &h->d_cvel, &h->d_caccel };
    for (DevBuf<double>* b : bufs) {
        CUDA_TRY(b->alloc(n));
        CUDA_TRY(cudaMemset(b->p, 0, n * sizeof(double)));
    }
    return GFA_OK;
}

DynArgs dyn_args(gfa_t* h, int slot, const gfa_dynamic_t* d) {
    DynArgs a;
    a.a1 = d->a1; a.a2 = d->a2; a.a3 = d->a3; a.a4 = d->a4; a.a5 = d->a5; a.a6 = d->a6;
    a.ray_alpha = d->rayleigh_alpha; a.ray_beta = d->rayleigh_beta; a.update = d->update_rayleigh != 0;
    a.vel = h->d_vel.p; a.copy_vel = h->d_cvel.p; a.copy_accel = h->d_caccel.p;
    a.alpha_i = slot >= 0 ? h->tb[slot].d_alpha_i.p : nullptr;
    a.rec = slot >= 0 ? h->tb[slot].d_dynrec.p : nullptr;
    a.CR = slot >= 0 ? h->tb[slot].d_CR.p : nullptr;
    return a;
}

// A handle in ring mode keeps no complete element arena; paths that need one (the Newmark kernels work on it in
// place) rebuild the slot map for the classic two-kernel path, once.
int leave_ring_mode(gfa_t* h) {
    if (!h->ring && !h->batch_layout) return GFA_OK;
    h->force_classic = true;
    h->force_compact = true;      // the Newmark kernels walk an element's own arena region
    const std::vector<int> gl = h->gls, em = h->ex_mat, er = h->ex_rows, ec = h->ex_cols;
    h->replaying = true;
    const int rc = gfa_set_dofs(h, gl.data(), h->n_free, h->n_fixed, (int64_t)em.size(), em.data(), er.data(), ec.data());
    h->replaying = false;
    return rc;
}

// the fused kernel's watchdog: a warp that waited longer than the time-out raised CTL_ABORT and every warp left
int check_abort(gfa_t* h) {
    if (!h->abort_check_pending) return GFA_OK;
    h->abort_check_pending = false;
    unsigned flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, h->d_ctl.p + CTL_ABORT, sizeof(flag), cudaMemcpyDeviceToHost));
    if (flag && getenv("GFA_FUSED_DEBUG")) {
        std::vector<unsigned> c(h->d_ctl.n);
        cudaMemcpy(c.data(), h->d_ctl.p, c.size() * sizeof(unsigned), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[gfa] watchdog: batches claimed %u %u %u, chunks %d; first to give up: %s waiting on chunk %u after %u us\n", c[CTL_BATCH], c[CTL_BATCH + 1], c[CTL_BATCH + 2], h->total_chunks, (c[0] & 15) == 1 ? "evaluation warp" : "scatter warp", c[0] >> 4, c[5]);
        { const unsigned* g = c.data() + CTL_HDR + 3 * h->total_chunks;
          fprintf(stderr, "[gfa]  giving-up evaluation warp: saw %u, wanted %u, s_upto %u, need %u, batch %u, chunk0 %u, ring_chunks %u\n", g[1], g[2], g[3], g[4], g[5], g[6], g[7]); }
        std::vector<int> tp(h->d_chunk_tile_ptr.n), cb(h->d_chunk_batches.n);
        cudaMemcpy(tp.data(), h->d_chunk_tile_ptr.p, tp.size() * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(cb.data(), h->d_chunk_batches.p, cb.size() * sizeof(int), cudaMemcpyDeviceToHost);
        for (int k = 0; k < h->total_chunks && k < 14; k++)
            fprintf(stderr, "[gfa]  chunk %d: evaluated %u of %d, scattered %u of %d, tiles claimed %u\n", k, c[CTL_HDR + k], cb[k], c[CTL_HDR + h->total_chunks + k], tp[k + 1] - tp[k], c[CTL_HDR + 2 * h->total_chunks + k]);
    }
    if (flag) { h->assembled = false; return fail(GFA_ECUDA, "fused assembly kernel gave up after waiting %.1f s for a chunk (watchdog); results are incomplete", h->fused_timeout_ns * 1e-9); }
    return GFA_OK;
}

int assemble_impl(gfa_t* h, const gfa_step_t* st, const gfa_dynamic_t* dyn, bool wait = true) {
    if (!h || !st) return fail(GFA_EINVAL, "gfa_assemble: null argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_assemble before gfa_set_dofs");
    CUDA_TRY(cudaSetDevice(h->device));
    if (dyn) {
        // MountMass / MountDamping exist on the device for Beam_1, Shell_1 and Pipe_1 (structural mass Rho only:
        // the model tables carry no ocean data, so the added-mass branch of Pipe_1.cpp:1758-1795 is never taken);
        // Solid_1 has no arithmetic in the reference
        if (!h->tb[2].elems.empty()) return fail(GFA_EUNSUPPORTED, "gfa_assemble_dynamic: Solid_1 has no dynamic contributions");
        int rc = ensure_kinematics(h);
        if (rc != GFA_OK) return rc;
        for (int slot = 0; slot < 2; slot++) {
            TypeBlock& t = h->tb[slot];
            if (t.elems.empty()) continue;
            const size_t rec = (slot == 0 ? SHELL_DYN_REC : BEAM_DYN_REC) * t.elems.size();
            if (t.d_dynrec.n != rec) CUDA_TRY(t.d_dynrec.alloc(rec));
            const bool rayleigh = dyn->update_rayleigh && (dyn->rayleigh_alpha != 0.0 || dyn->rayleigh_beta != 0.0);
            const size_t cr = (size_t)arena_doubles(slot) * t.elems.size();
            if (rayleigh && t.d_CR.n != cr) {           // Element::rayleigh_damping, zero until the first update
                CUDA_TRY(t.d_CR.alloc(cr));
                CUDA_TRY(cudaMemset(t.d_CR.p, 0, cr * sizeof(double)));
            }
        }
    }
    if (dyn && (h->ring || h->batch_layout)) {       // the mass / damping kernels fold their terms into a complete, compact arena in place
        int rc = leave_ring_mode(h);
        if (rc != GFA_OK) return rc;
    }
    cudaStream_t s = h->stream;
    const size_t nd = 6 * (size_t)h->n_nodes * sizeof(double);
    int launches = 0;
    h->last_gfac = st->gravity_factor;
    CUDA_TRY(cudaEventRecord(h->ev[0], s));
    if (st->displacements)       // NULL: keep the device copy (gfa_update_displacements)
        CUDA_TRY(cudaMemcpyAsync(h->d_disp.p, st->displacements, nd,
                                 st->displacements_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    if (h->n_vec_untouched > 0 && h->vec_dirty) {   // Solution::Clear for the vector entries no element writes: they start at
        // zero (gfa_set_dofs) and only gfa_add_host_vector can have changed them (shell loads, the interface unpack and
        // K_AB X_B add into rows that elements write; negating a zero leaves a zero)
        CUDA_TRY(cudaMemsetAsync(h->d_arena.p + h->vec_off[GFA_P_A], 0, (size_t)(h->arena_size - h->vec_off[GFA_P_A]) * sizeof(double), s));
        h->vec_dirty = false;
    }
    CUDA_TRY(cudaEventRecord(h->ev[1], s));
    ScatterArgs sa;
    sa.n_runs = h->n_runs; sa.runs = h->d_runs.p; sa.ovf = h->d_ovf.p;
    sa.n_gn = h->n_gn_local; sa.gn = h->d_gn.p; sa.inc = h->d_inc.p; sa.Ke = h->d_Ke.p; sa.Pe = h->d_Pe.p;
    sa.valAA = h->d_arena.p + h->arena_off[GFA_AA];
    sa.PA = h->d_arena.p + h->vec_off[GFA_P_A]; sa.IA = h->d_arena.p + h->vec_off[GFA_I_A]; sa.PB = h->d_arena.p + h->vec_off[GFA_P_B];
    GatherArgs g;
    g.n_dest = h->n_gdest; g.seg = h->d_gseg.p; g.src = h->d_gsrc.p; g.dest = h->d_gdest.p;
    g.Ke = h->d_Ke.p; g.vals = h->d_arena.p;
    if (!h->ring) {
        // MountLocal + MountElementLoads: one evaluation launch per element type
        for (int slot = 0; slot < 3; slot++) {
            if (h->tb[slot].elems.empty()) continue;
            const EvalArgs ea = eval_args(h, slot, st->gravity_factor);
            if (slot == 0) launch_shell_eval(ea, s); else if (slot == 1) launch_beam_eval(ea, s); else launch_solid_eval(ea, s);
            launches++;
            if (dyn) {      // MountMass + MountDamping + MountDyn folded into the element blocks before the scatter
                const DynArgs da = dyn_args(h, slot, dyn);
                if (slot == 0) launch_shell_dynamics(ea, da, s); else launch_beam_dynamics(ea, da, s);
                launches += 2;
            }
        }
        CUDA_TRY(cudaEventRecord(h->ev[2], s));
        // MountGlobal + MountSparse: rows of partition interfaces first (n_iface_* are zero with one rank), then the
        // fixed-DOF entries: after ev_iface everything the interface exchange packs is final, and the interior rows
        // follow behind it
        ScatterArgs si = sa;
        si.n_runs = h->n_iface_runs; si.n_gn = h->n_iface_gn;
        launches += launch_scatter(si, s);
        if (h->n_gdest > 0) { launch_gather(g, s); launches++; }
        CUDA_TRY(cudaEventRecord(h->ev_iface, s));
        sa.runs += h->n_iface_runs; sa.n_runs -= h->n_iface_runs;
        sa.gn += h->n_iface_gn; sa.n_gn -= h->n_iface_gn;
        launches += launch_scatter(sa, s);
    } else {
        // ---- ring pipeline.  Pre-pass: the pinned elements (fixed DOFs, partition interfaces, far-reaching
        // group-nodes) into their own regions, then everything only they feed -- interface rows and vectors
        // first, the AB / BA / BB entries, ev_iface, the remaining pinned-only patches.
        for (int slot = 0; slot < 3; slot++) {
            if (h->tb[slot].pin_list.empty()) continue;
            const EvalArgs ea = eval_args(h, slot, st->gravity_factor, PASS_PINNED);
            if (slot == 0) launch_shell_eval(ea, s); else if (slot == 1) launch_beam_eval(ea, s); else launch_solid_eval(ea, s);
            launches++;
        }
        ScatterArgs si = sa;
        si.n_runs = h->n_iface_runs; si.n_gn = h->n_iface_gn;
        launches += launch_scatter(si, s);
        if (h->n_gdest > 0) { launch_gather(g, s); launches++; }
        CUDA_TRY(cudaEventRecord(h->ev_iface, s));
        ScatterArgs sp = sa;
        sp.runs += h->n_iface_runs; sp.n_runs = h->n_pre_runs - h->n_iface_runs; sp.n_gn = 0;
        launches += launch_scatter(sp, s);
        CUDA_TRY(cudaEventRecord(h->ev[2], s));
        if (h->ring_serial) {
            // ---- serial ring: the classic kernels, launched group by group.  A group of chunks is evaluated into the
            // ring and the group-nodes it completes are scattered right behind it, while their blocks are still in L2;
            // stream order is the only synchronisation (a slot is rewritten ring_chunks chunks later, after every
            // scatter launch that read it)
            for (int slot = 0; slot < 3; slot++) {
                const TypeBlock& t = h->tb[slot];
                if (t.ring_list.empty()) continue;
                EvalArgs ea = eval_args(h, slot, st->gravity_factor, PASS_RING);
                for (int c0 = 0; c0 < t.n_chunks; c0 += h->ring_group) {
                    const int c1 = std::min(t.n_chunks, c0 + h->ring_group);
                    ea.e_begin = c0 * t.chunk_el; ea.e_end = (int)std::min<long long>((long long)c1 * t.chunk_el, (long long)t.ring_list.size());
                    if (slot == 0) launch_shell_eval(ea, s); else if (slot == 1) launch_beam_eval(ea, s); else launch_solid_eval(ea, s);
                    ScatterArgs sg = sa;
                    sg.runs += h->chunk_run_ptr_h[(size_t)t.chunk0 + c0];
                    sg.n_runs = h->chunk_run_ptr_h[(size_t)t.chunk0 + c1] - h->chunk_run_ptr_h[(size_t)t.chunk0 + c0];
                    sg.n_gn = 0;
                    launches += 1 + launch_scatter(sg, s);
                }
            }
        } else {
        // ---- ring pipeline: the scatter kernel on its own stream, the evaluation kernels (one per element type)
        // beside it; the two talk through the control block
        CUDA_TRY(cudaMemsetAsync(h->d_ctl.p, 0, h->d_ctl.n * sizeof(unsigned), s));
        CUDA_TRY(cudaEventRecord(h->ev_ctl, s));
        CUDA_TRY(cudaStreamWaitEvent(h->stream_sc, h->ev_ctl, 0));
        FusedArgs f;
        f.total_chunks = h->total_chunks; f.span = h->ring_span; f.tile_group = h->fused_tile_group; f.scatter_ctas = h->fused_scatter_ctas;
        f.sc = sa;
        f.chunk_run_ptr = h->d_chunk_run_ptr.p; f.chunk_tile_ptr = h->d_chunk_tile_ptr.p; f.chunk_batches = h->d_chunk_batches.p;
        f.ctl = h->d_ctl.p; f.timeout_ns = h->fused_timeout_ns;
        // the evaluation kernel goes first: it configures every SM for the large shared-memory carve-out, under
        // which the scatter CTA (no shared memory, 32 registers a thread) still fits beside it -- an SM that
        // started with the scatter kernel's small carve-out could not take the evaluation CTA until it drained
        bool scatter_launched = false;
        for (int slot = 0; slot < 3; slot++) {
            const TypeBlock& t = h->tb[slot];
            if (t.ring_list.empty()) continue;
            f.ev = eval_args(h, slot, st->gravity_factor, PASS_RING);
            f.n_buf = h->fused_eval_warps[slot]; f.type_slot = slot;
            int e = launch_fused_eval(f, s);
            if (e != 0) return fail(GFA_ECUDA, "evaluation kernel launch: %s", cudaGetErrorString((cudaError_t)e));
            launches++;
            if (!scatter_launched) {
                e = launch_fused_scatter(f, h->stream_sc);
                if (e != 0) return fail(GFA_ECUDA, "scatter kernel launch: %s", cudaGetErrorString((cudaError_t)e));
                launches++;
                CUDA_TRY(cudaEventRecord(h->ev_scattered, h->stream_sc));
                scatter_launched = true;
            }
        }
        CUDA_TRY(cudaStreamWaitEvent(s, h->ev_scattered, 0));
        h->abort_check_pending = true;
        }
        // residual vectors of the remaining group-nodes (the element force arena holds every element)
        ScatterArgs sv = sa;
        sv.gn += h->n_iface_gn; sv.n_gn -= h->n_iface_gn;
        if (sv.n_gn > 0) { launch_vectors(sv, s); launches++; }
    }
    CUDA_TRY(cudaEventRecord(h->ev[3], s));
    CUDA_TRY(cudaGetLastError());
    h->last_launches = launches;
    h->assembled = true;
    h->timing_pending = true;
    if (wait) {
        CUDA_TRY(cudaStreamSynchronize(s));
        resolve_timing(h);
        return check_abort(h);
    }
    return GFA_OK;
}

} // namespace

extern "C" {

int gfa_assemble(gfa_t* h, const gfa_step_t* st) { return assemble_impl(h, st, nullptr); }

int gfa_assemble_enqueue(gfa_t* h, const gfa_step_t* st) {
    if (st && st->displacements && !st->displacements_on_device)
        return fail(GFA_EINVAL, "gfa_assemble_enqueue: displacements must be a device pointer or NULL (the call returns before they are read)");
    return assemble_impl(h, st, nullptr, false);
}

int gfa_assemble_dynamic(gfa_t* h, const gfa_step_t* st, const gfa_dynamic_t* dyn) {
    if (!dyn) return fail(GFA_EINVAL, "gfa_assemble_dynamic: null argument");
    return assemble_impl(h, st, dyn);
}

int gfa_set_kinematics(gfa_t* h, const double* vel, const double* accel, const double* copy_vel, const double* copy_accel) {
    if (!h) return fail(GFA_EINVAL, "gfa_set_kinematics: null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = ensure_kinematics(h);
    if (rc != GFA_OK) return rc;
    const size_t nb = 6 * (size_t)h->n_nodes * sizeof(double);
    const double* src[4] = { vel, accel, copy_vel, copy_accel };
    double* dst[4] = { h->d_vel.p, h->d_accel.p, h->d_cvel.p, h->d_caccel.p };
    for (int k = 0; k < 4; k++)
        if (src[k]) CUDA_TRY(cudaMemcpyAsync(dst[k], src[k], nb, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return GFA_OK;
}

int gfa_kinematics(gfa_t* h, double* vel, double* accel, double* copy_vel, double* copy_accel) {
    if (!h) return fail(GFA_EINVAL, "gfa_kinematics: null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = ensure_kinematics(h);
    if (rc != GFA_OK) return rc;
    const size_t nb = 6 * (size_t)h->n_nodes * sizeof(double);
    double* dst[4] = { vel, accel, copy_vel, copy_accel };
    const double* src[4] = { h->d_vel.p, h->d_accel.p, h->d_cvel.p, h->d_caccel.p };
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 4; k++)
        if (dst[k]) CUDA_TRY(cudaMemcpy(dst[k], src[k], nb, cudaMemcpyDeviceToHost));
    return GFA_OK;
}

int gfa_update_dyn(gfa_t* h, const double* displacements, const gfa_dynamic_t* dyn) {
    if (!h || !dyn) return fail(GFA_EINVAL, "gfa_update_dyn: null argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_update_dyn before gfa_set_dofs");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = ensure_kinematics(h);
    if (rc != GFA_OK) return rc;
    if (displacements)
        CUDA_TRY(cudaMemcpyAsync(h->d_disp.p, displacements, 6 * (size_t)h->n_nodes * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_update_dyn(dyn_args(h, -1, dyn), h->d_gls.p, h->d_disp.p, h->d_vel.p, h->d_accel.p, h->n_nodes,
                      h->d_mixed.p, h->d_mixed_start.p, h->n_mixed, h->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return GFA_OK;
}

int gfa_element_alpha_i(gfa_t* h, int32_t e, double* out) {
    if (!h || e < 0 || e >= h->n_el || !out) return fail(GFA_EINVAL, "gfa_element_alpha_i: bad argument");
    const int s = h->el_owner_slot[e];
    if (s < 0) return fail(GFA_EINVAL, "element %d belongs to another rank's partition", e + 1);
    if (kTypes[s].nstate == 0) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const size_t ngp = h->tb[s].elems.size() * kTypes[s].ngp;
    int w = 0;
    for (int g = 0; g < kTypes[s].ngp; g++)
        for (int k = 0; k < 3; k++) {
            CUDA_TRY(cudaMemcpy(out + w, h->tb[s].d_alpha_i.p + (size_t)k * ngp + (size_t)h->el_local[e] * kTypes[s].ngp + g, sizeof(double), cudaMemcpyDeviceToHost));
            w++;
        }
    return w;
}

namespace {
// Sum duplicates in push order (what Eigen's setFromTriplets does with repeated positions) and add the result
// into the value arena: slots sorted with a stable sort, staging buffers kept in the handle -- nothing is
// allocated on the per-iteration path once they have grown to the largest call.
int add_staged(gfa_t* h, std::vector<std::pair<long long, double> >& items) {
    std::stable_sort(items.begin(), items.end(), [](const std::pair<long long, double>& x, const std::pair<long long, double>& y) { return x.first < y.first; });
    h->stage_slots.clear(); h->stage_vals.clear();
    for (size_t i = 0; i < items.size();) {
        size_t j = i; double acc = items[i].second;
        for (j = i + 1; j < items.size() && items[j].first == items[i].first; j++) acc += items[j].second;
        h->stage_slots.push_back(items[i].first); h->stage_vals.push_back(acc);
        i = j;
    }
    const size_t n = h->stage_slots.size();
    if (h->d_stage_slots.n < n) {
        const size_t cap = std::max(n, 2 * h->d_stage_slots.n + 1024);
        CUDA_TRY(h->d_stage_slots.alloc(cap)); CUDA_TRY(h->d_stage_vals.alloc(cap));
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_stage_slots.p, h->stage_slots.data(), n * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_stage_vals.p, h->stage_vals.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_add_slots(h->d_arena.p, h->d_stage_slots.p, h->d_stage_vals.p, (long long)n, h->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));      // the host staging vectors are reused by the next call
    return GFA_OK;
}
} // namespace

int gfa_add_host_triplets(gfa_t* h, int which, int64_t n, const int32_t* rows, const int32_t* cols, const double* vals) {
    if (!h || which < 0 || which > 3 || (n > 0 && (!rows || !cols || !vals))) return fail(GFA_EINVAL, "gfa_add_host_triplets: bad argument");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_add_host_triplets before gfa_assemble");
    if (n <= 0) return GFA_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const HostCsr& M = h->csr[which];
    std::vector<std::pair<long long, double> >& items = h->stage_items;
    items.clear();
    const int n_rows_global = (which == GFA_AA || which == GFA_AB) ? h->n_free : h->n_fixed;
    for (int64_t i = 0; i < n; i++) {
        const int r = rows[i], c = cols[i];
        if (r < 0 || r >= n_rows_global) return fail(GFA_EPATTERN, "host triplet %lld: row %d outside the matrix", (long long)i, r);
        int lr = r;
        if (which == GFA_AA) {
            lr = M.row_local[r];
            if (lr < 0) return fail(GFA_EPATTERN, "host triplet row %d is not stored on rank %d", r, h->rank);
        }
        const int* b = M.inner.data() + M.rowptr[lr]; const int* e = M.inner.data() + M.rowptr[lr + 1];
        const int* p = std::lower_bound(b, e, c);
        if (p == e || *p != c) return fail(GFA_EPATTERN, "host triplet (%d,%d) is not in the registered pattern of matrix %d; list it in gfa_set_dofs", r, c, which);
        items.emplace_back(h->arena_off[which] + M.rowptr[lr] + (p - b), vals[i]);
    }
    return add_staged(h, items);
}

int gfa_add_host_vector(gfa_t* h, int wv, int64_t n, const int32_t* index, const double* vals) {
    if (!h || wv < 0 || wv > 2 || (n > 0 && (!index || !vals))) return fail(GFA_EINVAL, "gfa_add_host_vector: bad argument");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_add_host_vector before gfa_assemble");
    if (n <= 0) return GFA_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const int len = wv == GFA_P_B ? h->n_fixed : h->n_free;
    h->vec_dirty = true;
    std::vector<std::pair<long long, double> >& items = h->stage_items;
    items.clear();
    for (int64_t i = 0; i < n; i++) {
        if (index[i] < 0 || index[i] >= len) return fail(GFA_EINVAL, "host vector entry %lld: index %d out of range", (long long)i, index[i]);
        items.emplace_back(h->vec_off[wv] + index[i], vals[i]);
    }
    return add_staged(h, items);
}

// Shared by the two element-load families: for every (load, element) entry a record of 18 x 18 + 18 doubles
// (LOAD_REC) is written by the load kernel; `slots` routes every record entry to its CSR value / vector position,
// sources of a destination in registration order.  `shell`: local DOFs are the translations of the six nodes;
// otherwise (Pipe_1) the six DOFs of the three nodes.
static int build_load_set(gfa_t* h, gfa_t::LoadSet& L, bool shell, int32_t n_loads, const int32_t* load_ptr, const int32_t* load_elements,
                          const int32_t* flags, const char* what) {
    std::vector<int> elem, load, flag;
    if (flags) flag.assign(flags, flags + n_loads); else flag.assign((size_t)std::max(n_loads, 1), 0);
    struct Slot { long long dest, src; };
    std::vector<Slot> slots;
    const int slot_of_type = shell ? 0 : 1, nn = shell ? 6 : 3, per = shell ? 3 : 6;
    for (int l = 0; l < n_loads; l++)
        for (int k = load_ptr[l]; k < load_ptr[l + 1]; k++) {
            const int e = load_elements[k];
            const bool ok = e >= 0 && e < h->n_el && (shell ? h->el_type[e] == GFA_SHELL_1 : h->el_type[e] == GFA_PIPE_1);
            if (!ok) return fail(GFA_EINVAL, "%s %d: element %d is not a %s element", what, l + 1, e + 1, shell ? "Shell_1" : "Pipe_1");
            if (h->el_owner_slot[e] != slot_of_type) continue;              // another rank's partition
            const long long entry = (long long)elem.size();
            elem.push_back(h->el_local[e]); load.push_back(l);
            int gl[18];
            for (int a = 0; a < nn; a++)
                for (int c = 0; c < per; c++) gl[per * a + c] = h->gls[6 * (size_t)h->el_nodes[h->el_ptr[e] + a] + c];
            for (int i = 0; i < 18; i++) {
                const int g1 = gl[i];
                if (g1 == 0) continue;
                slots.push_back({ (g1 > 0 ? h->vec_off[GFA_P_A] + g1 - 1 : h->vec_off[GFA_P_B] + (-g1 - 1)), entry * SHELL_LOAD_REC + 324 + i });
                if (g1 > 0) slots.push_back({ h->vec_off[GFA_I_A] + g1 - 1, entry * SHELL_LOAD_REC + 324 + i });
                for (int j = 0; j < 18; j++) {
                    const int g2 = gl[j];
                    if (g2 == 0) continue;
                    const int w = g1 > 0 ? (g2 > 0 ? GFA_AA : GFA_AB) : (g2 > 0 ? GFA_BA : GFA_BB);
                    const HostCsr& M = h->csr[w];
                    int lr = std::abs(g1) - 1;
                    if (w == GFA_AA) lr = M.row_local[(size_t)lr];
                    const int col = std::abs(g2) - 1;
                    const int* b = M.inner.data() + M.rowptr[lr]; const int* en = M.inner.data() + M.rowptr[lr + 1];
                    const int* p = std::lower_bound(b, en, col);
                    if (p == en || *p != col) return fail(GFA_EPATTERN, "%s position (%d,%d) of matrix %d is not in the pattern", what, lr, col, w);
                    slots.push_back({ h->arena_off[w] + M.rowptr[lr] + (p - b), entry * SHELL_LOAD_REC + i * 18 + j });
                }
            }
        }
    // one destination per slot, its sources in registration order (stable sort)
    std::stable_sort(slots.begin(), slots.end(), [](const Slot& x, const Slot& y) { return x.dest < y.dest; });
    std::vector<long long> seg, src, dest;
    for (size_t i = 0; i < slots.size();) {
        size_t j = i;
        seg.push_back((long long)src.size()); dest.push_back(slots[i].dest);
        for (; j < slots.size() && slots[j].dest == slots[i].dest; j++) src.push_back(slots[j].src);
        i = j;
    }
    seg.push_back((long long)src.size());
    L.n_loads = n_loads; L.n_entries = (int)elem.size(); L.n_dest = (long long)dest.size();
    CUDA_TRY(L.d_elem.upload(elem)); CUDA_TRY(L.d_of.upload(load)); CUDA_TRY(L.d_flag.upload(flag));
    CUDA_TRY(L.d_value.alloc((size_t)std::max(n_loads, 1)));
    CUDA_TRY(L.d_out.alloc((size_t)std::max<size_t>(elem.size(), 1) * SHELL_LOAD_REC));
    CUDA_TRY(L.d_seg.upload(seg)); CUDA_TRY(L.d_src.upload(src)); CUDA_TRY(L.d_dest.upload(dest));
    return GFA_OK;
}

static int apply_load_set(gfa_t* h, gfa_t::LoadSet& L, bool shell, const double* values) {
    if (L.n_entries == 0) return GFA_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(L.d_value.p, values, (size_t)L.n_loads * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ShellLoadArgs la;
    la.n_entries = L.n_entries; la.elem = L.d_elem.p; la.load = L.d_of.p;
    la.pressure = L.d_value.p; la.area_update = L.d_flag.p; la.out = L.d_out.p;
    if (shell) launch_shell_loads(eval_args(h, 0, 0.0), la, h->stream);
    else launch_pipe_loads(eval_args(h, 1, 0.0), la, h->stream);
    launch_gather_add(h->d_arena.p, L.d_seg.p, L.d_src.p, L.d_dest.p, L.d_out.p, L.n_dest, h->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));      // `values` is the caller's
    return GFA_OK;
}

int gfa_set_shell_loads(gfa_t* h, int32_t n_loads, const int32_t* load_ptr, const int32_t* load_elements, const int32_t* area_update) {
    if (!h || n_loads < 0 || (n_loads > 0 && (!load_ptr || !load_elements || !area_update))) return fail(GFA_EINVAL, "gfa_set_shell_loads: bad argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_set_shell_loads before gfa_set_dofs");
    CUDA_TRY(cudaSetDevice(h->device));
    return build_load_set(h, h->shell_loads, true, n_loads, load_ptr, load_elements, area_update, "shell load");
}

int gfa_apply_shell_loads(gfa_t* h, const double* pressures) {
    if (!h || (h->shell_loads.n_loads > 0 && !pressures)) return fail(GFA_EINVAL, "gfa_apply_shell_loads: bad argument");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_apply_shell_loads before gfa_assemble");
    return apply_load_set(h, h->shell_loads, true, pressures);
}

int gfa_set_pipe_loads(gfa_t* h, int32_t n_loads, const int32_t* load_ptr, const int32_t* load_elements) {
    if (!h || n_loads < 0 || (n_loads > 0 && (!load_ptr || !load_elements))) return fail(GFA_EINVAL, "gfa_set_pipe_loads: bad argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_set_pipe_loads before gfa_set_dofs");
    CUDA_TRY(cudaSetDevice(h->device));
    return build_load_set(h, h->pipe_loads, false, n_loads, load_ptr, load_elements, nullptr, "pipe load");
}

int gfa_apply_pipe_loads(gfa_t* h, const double* p0i) {
    if (!h || (h->pipe_loads.n_loads > 0 && !p0i)) return fail(GFA_EINVAL, "gfa_apply_pipe_loads: bad argument");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_apply_pipe_loads before gfa_assemble");
    return apply_load_set(h, h->pipe_loads, false, p0i);
}

int gfa_csr_values(gfa_t* h, int which, double* out) {
    if (!h || which < 0 || which > 3 || !out) return fail(GFA_EINVAL, "gfa_csr_values: bad argument");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_csr_values before gfa_assemble");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));      // interface unpack and host additions are stream-ordered
    if (int rc = check_abort(h)) return rc;
    const size_t n = h->csr[which].inner.size();
    if (n) CUDA_TRY(cudaMemcpy(out, h->d_arena.p + h->arena_off[which], n * sizeof(double), cudaMemcpyDeviceToHost));
    return GFA_OK;
}
int gfa_csr_values_device(gfa_t* h, int which, double** p) {
    if (!h || which < 0 || which > 3 || !p) return fail(GFA_EINVAL, "gfa_csr_values_device: bad argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_csr_values_device before gfa_set_dofs");
    *p = h->d_arena.p + h->arena_off[which];
    return GFA_OK;
}
int gfa_vector(gfa_t* h, int wv, double* out) {
    if (!h || wv < 0 || wv > 2 || !out) return fail(GFA_EINVAL, "gfa_vector: bad argument");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_vector before gfa_assemble");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));      // interface unpack and host additions are stream-ordered
    if (int rc = check_abort(h)) return rc;
    const size_t n = wv == GFA_P_B ? h->n_fixed : h->n_free;
    if (n) CUDA_TRY(cudaMemcpy(out, h->d_arena.p + h->vec_off[wv], n * sizeof(double), cudaMemcpyDeviceToHost));
    return GFA_OK;
}
int gfa_vector_device(gfa_t* h, int wv, double** p) {
    if (!h || wv < 0 || wv > 2 || !p) return fail(GFA_EINVAL, "gfa_vector_device: bad argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_vector_device before gfa_set_dofs");
    *p = h->d_arena.p + h->vec_off[wv];
    return GFA_OK;
}

int gfa_element_block(gfa_t* h, int32_t e, double* K, double* P) {
    if (!h || e < 0 || e >= h->n_el) return fail(GFA_EINVAL, "gfa_element_block: element out of range");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_element_block before gfa_assemble");
    const int s = h->el_owner_slot[e];
    if (s < 0) return fail(GFA_EINVAL, "element %d belongs to another rank's partition", e + 1);
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));      // interface unpack and host additions are stream-ordered
    const int n = kTypes[s].ndof;
    if (K) {
        const bool batch = s == 0 && h->batch_layout && !h->ring;
        const int nd = batch ? SHELL_BATCH * SHELL_ARENA : arena_doubles(s), local = h->el_local[e];
        std::vector<double> blk((size_t)nd);
        // batch layout: fetch the whole batch region and pick this element's blocks out of it
        const double* src = batch ? h->d_Ke.p + h->tb[0].ke_base + (long long)(local / SHELL_BATCH) * (SHELL_BATCH * SHELL_ARENA)
                                  : h->d_Ke.p + h->tb[s].ke_off[local];
        if (h->ring) {
            // the ring keeps only the last chunks: evaluate this one element again, into a scratch region
            CUDA_TRY(cudaMemcpyAsync(h->d_one.p, &local, sizeof(int), cudaMemcpyHostToDevice, h->stream));
            EvalArgs ea = eval_args(h, s, h->last_gfac);
            ea.elist = h->d_one.p; ea.e_begin = 0; ea.e_end = 1; ea.Ke = h->d_scratch_ke.p;
            if (s == 0) launch_shell_eval(ea, h->stream); else if (s == 1) launch_beam_eval(ea, h->stream); else launch_solid_eval(ea, h->stream);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            src = h->d_scratch_ke.p;
        }
        CUDA_TRY(cudaMemcpy(blk.data(), src, sizeof(double) * nd, cudaMemcpyDeviceToHost));
        // device layout is block-wise (and upper-triangular for Shell_1); hand back plain row-major
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                bool tr = false;
                const int off = batch ? (int)shell_batch_offset(local % SHELL_BATCH, i / 3, j / 3, tr)
                              : s == 0 ? shell_block_offset(i / 3, j / 3, tr) : s == 1 ? beam_block_offset(i / 3, j / 3, tr) : solid_block_offset(i / 3, j / 3, tr);
                K[i * n + j] = blk[(size_t)off + (tr ? (j % 3) * 3 + (i % 3) : (i % 3) * 3 + (j % 3))];
            }
    }
    if (P) CUDA_TRY(cudaMemcpy(P, h->d_Pe.p + h->tb[s].pe_base + (size_t)h->el_local[e] * n, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return n;
}

int gfa_commit_state(gfa_t* h) {
    if (!h) return fail(GFA_EINVAL, "gfa_commit_state: null handle");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_commit_state before gfa_assemble");
    CUDA_TRY(cudaSetDevice(h->device));
    launch_shell_commit(eval_args(h, 0, 0.0), h->stream);
    launch_beam_commit(eval_args(h, 1, 0.0), h->stream);
    launch_shell_alpha_commit(eval_args(h, 0, 0.0), h->tb[0].d_alpha_i.p, h->stream);      // alpha_i (Shell_1.cpp:1659, Beam_1.cpp:1502)
    launch_beam_alpha_commit(eval_args(h, 1, 0.0), h->tb[1].d_alpha_i.p, h->stream);
    launch_node_commit(h->n_nodes, h->d_copy.p, h->d_disp.p, h->stream);
    if (h->d_vel.n) {       // copy_vel = vel, copy_accel = accel (Node.cpp:375-380)
        const size_t nb = h->d_vel.n * sizeof(double);
        CUDA_TRY(cudaMemcpyAsync(h->d_cvel.p, h->d_vel.p, nb, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_caccel.p, h->d_accel.p, nb, cudaMemcpyDeviceToDevice, h->stream));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return GFA_OK;
}

int gfa_element_state(gfa_t* h, int32_t e, double* out) {
    if (!h || e < 0 || e >= h->n_el || !out) return fail(GFA_EINVAL, "gfa_element_state: bad argument");
    const int s = h->el_owner_slot[e];
    if (s < 0) return fail(GFA_EINVAL, "element %d belongs to another rank's partition", e + 1);
    CUDA_TRY(cudaSetDevice(h->device));
    const TypeInfo& ti = kTypes[s];
    const size_t ngp = h->tb[s].elems.size() * ti.ngp;
    int w = 0;
    for (int g = 0; g < ti.ngp && ti.nstate; g++)
        for (int k = 0; k < ti.nstate; k++) {
            CUDA_TRY(cudaMemcpy(out + w, h->tb[s].d_state.p + (size_t)k * ngp + (size_t)h->el_local[e] * ti.ngp + g, sizeof(double), cudaMemcpyDeviceToHost));
            w++;
        }
    return w;
}

int gfa_results_stride(int element_type) {
    return element_type == GFA_SHELL_1 ? SHELL_RESULTS : (element_type == GFA_BEAM_1 || element_type == GFA_PIPE_1) ? BEAM_RESULTS : 0;
}

int64_t gfa_gauss_point_results(gfa_t* h, int element_type, double* out, int64_t capacity) {
    if (!h || !out) return fail(GFA_EINVAL, "gfa_gauss_point_results: null argument");
    const int s = type_slot(element_type);
    const int stride = gfa_results_stride(element_type);
    if (s < 0 || stride == 0) return fail(GFA_EUNSUPPORTED, "element type %d keeps no Gauss-point results", element_type);
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_gauss_point_results before gfa_assemble");
    const size_t n = h->tb[s].elems.size();
    size_t n_out = 0;                      // Beam_1 and Pipe_1 share a block: hand back the asked type only
    for (size_t k = 0; k < n; k++) if (h->el_type[h->tb[s].elems[k]] == element_type) n_out++;
    if ((int64_t)(n_out * stride) > capacity) return fail(GFA_EINVAL, "gfa_gauss_point_results: %zu records of %d doubles need capacity %zu, got %lld", n_out, stride, n_out * stride, (long long)capacity);
    if (n_out == 0) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    DevBuf<double> d;
    cudaError_t e = d.alloc(n * stride);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? GFA_ENOMEM : GFA_ECUDA, "gfa_gauss_point_results: %s", cudaGetErrorString(e));
    const EvalArgs ea = eval_args(h, s, 0.0);
    if (s == 0) launch_shell_results(ea, d.p, h->stream); else launch_beam_results(ea, d.p, h->stream);
    CUDA_TRY(cudaGetLastError());
    if (n_out == n) {
        CUDA_TRY(cudaMemcpyAsync(out, d.p, n * stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    } else {
        std::vector<double> all(n * stride);
        CUDA_TRY(cudaMemcpyAsync(all.data(), d.p, n * stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        size_t w = 0;
        for (size_t k = 0; k < n; k++)
            if (h->el_type[h->tb[s].elems[k]] == element_type) { std::memcpy(out + w * stride, all.data() + k * stride, stride * sizeof(double)); w++; }
    }
    return (int64_t)n_out;
}

namespace {
int read_norms(gfa_t* h, bool with_values, gfa_norms_t* out) {
    NormAcc a;
    CUDA_TRY(cudaMemcpyAsync(&a, h->d_norm.p, sizeof(a), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    auto dbl = [](unsigned long long b) { double d; std::memcpy(&d, &b, sizeof(d)); return d; };
    std::memset(out, 0, sizeof(*out));
    out->max_force = dbl(a.max_t); out->max_moment = dbl(a.max_r);
    out->node_force = a.node_t == 0x7fffffff ? 0 : a.node_t + 1;
    out->node_moment = a.node_r == 0x7fffffff ? 0 : a.node_r + 1;
    if (with_values) { out->max_disp_value = dbl(a.max_dt); out->max_rot_value = dbl(a.max_dr); }
    out->nan_detected = a.nan;
    return GFA_OK;
}
int reset_norms(gfa_t* h) {
    NormAcc z; std::memset(&z, 0, sizeof(z)); z.node_t = 0x7fffffff; z.node_r = 0x7fffffff;
    CUDA_TRY(cudaMemcpyAsync(h->d_norm.p, &z, sizeof(z), cudaMemcpyHostToDevice, h->stream));
    return GFA_OK;
}
} // namespace

int gfa_residual(gfa_t* h, const double* X_B, gfa_norms_t* out) {
    if (!h) return fail(GFA_EINVAL, "gfa_residual: null handle");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_residual before gfa_assemble");
    CUDA_TRY(cudaSetDevice(h->device));
    double* PA = h->d_arena.p + h->vec_off[GFA_P_A];
    launch_negate(PA, h->n_free, h->stream);
    if (X_B && h->n_fixed > 0 && h->n_ab_rows > 0) {
        if (h->d_xb.n < (size_t)h->n_fixed) CUDA_TRY(h->d_xb.alloc((size_t)h->n_fixed));
        CUDA_TRY(cudaMemcpyAsync(h->d_xb.p, X_B, (size_t)h->n_fixed * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        launch_sub_ab_xb(PA, h->d_ab_rows.p, h->d_ab_ptr.p, h->d_ab_inner.p, h->d_arena.p + h->arena_off[GFA_AB], h->d_xb.p, h->n_ab_rows, h->stream);
    }
    if (out) {
        if (int rc = reset_norms(h)) return rc;
        // a partitioned run: this rank's norms cover the rows it owns (complete after the interface exchange)
        launch_norms(h->world > 1 ? h->d_gls_owned.p : h->d_gls.p, PA, nullptr, h->n_nodes, h->d_norm.p, h->stream);
        CUDA_TRY(cudaGetLastError());
        return read_norms(h, false, out);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return GFA_OK;
}

int gfa_update_displacements(gfa_t* h, const double* x_A, gfa_norms_t* out) {
    if (!h || !x_A) return fail(GFA_EINVAL, "gfa_update_displacements: null argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_update_displacements before gfa_set_dofs");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->d_xa.n < (size_t)std::max(h->n_free, 1)) CUDA_TRY(h->d_xa.alloc((size_t)std::max(h->n_free, 1)));
    CUDA_TRY(cudaMemcpyAsync(h->d_xa.p, x_A, (size_t)h->n_free * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_update_disps(h->d_gls.p, h->d_disp.p, h->d_xa.p, h->n_nodes, h->stream);
    if (out) {
        if (int rc = reset_norms(h)) return rc;
        launch_norms(h->world > 1 ? h->d_gls_owned.p : h->d_gls.p, h->d_xa.p, h->d_disp.p, h->n_nodes, h->d_norm.p, h->stream);
        CUDA_TRY(cudaGetLastError());
        return read_norms(h, true, out);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return GFA_OK;
}

int gfa_displacements(gfa_t* h, double* out) {
    if (!h || !out) return fail(GFA_EINVAL, "gfa_displacements: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));      // interface unpack and host additions are stream-ordered
    CUDA_TRY(cudaMemcpy(out, h->d_disp.p, 6 * (size_t)h->n_nodes * sizeof(double), cudaMemcpyDeviceToHost));
    return GFA_OK;
}

int gfa_copy_coordinates(gfa_t* h, double* out) {
    if (!h || !out) return fail(GFA_EINVAL, "gfa_copy_coordinates: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpy(out, h->d_copy.p, 6 * (size_t)h->n_nodes * sizeof(double), cudaMemcpyDeviceToHost));
    return GFA_OK;
}

int gfa_last_timing(gfa_t* h, double* ms4) {
    if (!h || !ms4) return fail(GFA_EINVAL, "gfa_last_timing: bad argument");
    resolve_timing(h);
    for (int i = 0; i < 4; i++) ms4[i] = h->last_ms[i];
    return check_abort(h);
}
int gfa_pipeline_info(gfa_t* h, char* buf, int32_t capacity) {
    if (!h || !buf || capacity < 1) return fail(GFA_EINVAL, "gfa_pipeline_info: bad argument");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_pipeline_info before gfa_set_dofs");
    snprintf(buf, (size_t)capacity, "%s", h->ring_note.c_str());
    return h->ring ? 1 : 0;
}
int gfa_last_launch_count(gfa_t* h) { return h ? h->last_launches : 0; }

int gfa_interface_counts(gfa_t* h, int64_t* sc, int64_t* rc) {
    if (!h) return fail(GFA_EINVAL, "gfa_interface_counts: null handle");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_interface_counts before gfa_set_dofs");
    for (int r = 0; r < h->world; r++) { if (sc) sc[r] = h->send_cnt[r]; if (rc) rc[r] = h->recv_cnt[r]; }
    return GFA_OK;
}
int gfa_interface_pack(gfa_t* h, double* buf) {
    if (!h) return fail(GFA_EINVAL, "gfa_interface_pack: null handle");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_interface_pack before gfa_assemble");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamWaitEvent(h->stream_if, h->ev_iface, 0));      // interface rows and fixed-DOF entries are final
    launch_pack(h->d_arena.p, h->d_send_idx.p, buf, (long long)h->d_send_idx.n, h->stream_if);
    CUDA_TRY(cudaGetLastError());
    return GFA_OK;       // asynchronous: ordered on gfa_interface_stream()
}
int gfa_interface_unpack(gfa_t* h, const double* buf) {
    if (!h) return fail(GFA_EINVAL, "gfa_interface_unpack: null handle");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_interface_unpack before gfa_assemble");
    CUDA_TRY(cudaSetDevice(h->device));
    // peers ascending, one launch per peer segment => fixed summation order
    long long off = 0;
    for (int r = 0; r < h->world; r++) {
        launch_unpack_add(h->d_arena.p, h->d_recv_idx.p + off, buf + off, h->recv_cnt[r], h->stream_if);
        off += h->recv_cnt[r];
    }
    CUDA_TRY(cudaGetLastError());
    // everything that follows on gfa_stream() -- reads, host additions, the next assembly -- waits for the exchange
    CUDA_TRY(cudaEventRecord(h->ev_unpacked, h->stream_if));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_unpacked, 0));
    return GFA_OK;       // asynchronous: ordered on gfa_interface_stream()
}
int gfa_local_rows(gfa_t* h, int64_t* n_rows, int32_t* rows_out) {
    if (!h) return fail(GFA_EINVAL, "gfa_local_rows: null handle");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_local_rows before gfa_set_dofs");
    const std::vector<int>& r = h->csr[GFA_AA].row_ids;
    if (n_rows) *n_rows = (int64_t)r.size();
    if (rows_out && !r.empty()) std::memcpy(rows_out, r.data(), r.size() * sizeof(int));
    return GFA_OK;
}
int gfa_owned_rows(gfa_t* h, int64_t* n_rows, int32_t* rows_out) {
    if (!h) return fail(GFA_EINVAL, "gfa_owned_rows: null handle");
    if (!h->dofs_set) return fail(GFA_ESTATE, "gfa_owned_rows before gfa_set_dofs");
    if (n_rows) *n_rows = (int64_t)h->owned_rows.size();
    if (rows_out && !h->owned_rows.empty()) std::memcpy(rows_out, h->owned_rows.data(), h->owned_rows.size() * sizeof(int));
    return GFA_OK;
}
int gfa_touched_nodes(gfa_t* h, int64_t* n_nodes, int32_t* nodes_out) {
    if (!h) return fail(GFA_EINVAL, "gfa_touched_nodes: null handle");
    if (n_nodes) *n_nodes = (int64_t)h->touched_nodes.size();
    if (nodes_out && !h->touched_nodes.empty()) std::memcpy(nodes_out, h->touched_nodes.data(), h->touched_nodes.size() * sizeof(int));
    return GFA_OK;
}
int gfa_set_displacements_packed(gfa_t* h, const double* packed, int32_t on_device) {
    if (!h || !packed) return fail(GFA_EINVAL, "gfa_set_displacements_packed: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n = h->touched_nodes.size();
    if (h->d_touched_nodes.n != n) CUDA_TRY(h->d_touched_nodes.upload(h->touched_nodes));
    const double* src = packed;
    if (!on_device) {
        if (h->d_packed.n != 6 * n) CUDA_TRY(h->d_packed.alloc(6 * n));
        CUDA_TRY(cudaMemcpyAsync(h->d_packed.p, packed, 6 * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        src = h->d_packed.p;
    }
    launch_unpack_nodes(h->d_disp.p, h->d_touched_nodes.p, src, (long long)n, h->stream);
    CUDA_TRY(cudaGetLastError());
    return GFA_OK;       // stream-ordered: the next gfa_assemble (displacements == NULL) runs behind it
}
int gfa_vector_owned(gfa_t* h, int wv, double* out) {
    if (!h || wv < 0 || wv > 2 || !out) return fail(GFA_EINVAL, "gfa_vector_owned: bad argument");
    if (!h->assembled) return fail(GFA_ESTATE, "gfa_vector_owned before gfa_assemble");
    if (wv == GFA_P_B || h->world == 1) return gfa_vector(h, wv, out);
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n = h->owned_rows.size();
    if (n == 0) return GFA_OK;
    if (h->d_owned_idx.n != n) {
        std::vector<long long> idx(n);
        for (size_t i = 0; i < n; i++) idx[i] = (long long)h->owned_rows[i];
        CUDA_TRY(h->d_owned_idx.upload(idx));
    }
    if (h->d_packed.n < n) CUDA_TRY(h->d_packed.alloc(std::max(n, 6 * h->touched_nodes.size())));
    launch_pack(h->d_arena.p + h->vec_off[wv], h->d_owned_idx.p, h->d_packed.p, (long long)n, h->stream);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, h->d_packed.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_abort(h);
}
int gfa_interface_stream(gfa_t* h, void** out) {
    if (!h || !out) return fail(GFA_EINVAL, "gfa_interface_stream: bad argument");
    *out = (void*)h->stream_if;
    return GFA_OK;
}

int gfa_stream(gfa_t* h, void** out) {
    if (!h || !out) return fail(GFA_EINVAL, "gfa_stream: bad argument");
    *out = (void*)h->stream;
    return GFA_OK;
}

} // extern "C"
