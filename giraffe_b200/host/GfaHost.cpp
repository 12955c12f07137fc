// See GfaHost.h.  Reader and set-up logic restate the in-scope parts of the
// reference's IO / Database / Solution classes on plain arrays; the assembly
// itself is the C-ABI (no element arithmetic lives here).
#include "GfaHost.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace {

const double kPi = 3.1415926535897932384626433832795;

// Whitespace tokens with // and /* */ comments removed (IO.cpp:683-752).
std::vector<std::string> tokenize(const std::string& text) {
    std::string clean;
    clean.reserve(text.size());
    for (size_t i = 0; i < text.size();) {
        if (text.compare(i, 2, "//") == 0) { while (i < text.size() && text[i] != '\n') i++; }
        else if (text.compare(i, 2, "/*") == 0) { size_t e = text.find("*/", i + 2); i = e == std::string::npos ? text.size() : e + 2; clean += ' '; }
        else clean += text[i++];
    }
    std::vector<std::string> tk;
    std::istringstream ss(clean);
    std::string t;
    while (ss >> t) tk.push_back(t);
    return tk;
}

bool is_top(const std::string& s) {
    static const char* k[] = { "Nodes", "Elements", "Materials", "Sections", "PipeSections", "ShellSections", "CoordinateSystems", "NodeSets",
        "Constraints", "Loads", "Environment", "SolutionSteps", "SolverOptions", "Monitor", "PostFiles",
        "ConvergenceCriteria", "ElementSets", "ExecutionData", "EOF" };
    for (const char* w : k) if (s == w) return true;
    return false;
}

} // namespace

double GfaNodalLoad::GetValueAt(double t, int column) const {
    const int n = (int)(table.size() / 7);
    if (n == 0) return 0.0;
    if (t <= table[0]) return table[column];
    for (int r = 0; r + 1 < n; r++) {
        const double t0 = table[7 * r], t1 = table[7 * (r + 1)];
        if (t <= t1) return table[7 * r + column] + (table[7 * (r + 1) + column] - table[7 * r + column]) * (t - t0) / (t1 - t0);
    }
    return table[7 * (n - 1) + column];
}

double GfaPipeLoad::GetValueAt(double t, int column) const {
    const int n = (int)(table.size() / 5);
    if (n == 0) return 0.0;
    if (t <= table[0]) return table[1 + column];
    for (int r = 0; r + 1 < n; r++) {
        const double t0 = table[5 * r], t1 = table[5 * (r + 1)];
        if (t <= t1) return table[5 * r + 1 + column] + (table[5 * (r + 1) + 1 + column] - table[5 * r + 1 + column]) * (t - t0) / (t1 - t0);
    }
    return table[5 * (n - 1) + 1 + column];
}

double GfaShellLoad::GetValueAt(double t) const {
    const int n = (int)(table.size() / 2);
    if (n == 0) return 0.0;
    if (t <= table[0]) return table[1];
    for (int r = 0; r + 1 < n; r++) {
        const double t0 = table[2 * r], t1 = table[2 * (r + 1)];
        if (t <= t1) return table[2 * r + 1] + (table[2 * (r + 1) + 1] - table[2 * r + 1]) * (t - t0) / (t1 - t0);
    }
    return table[2 * (n - 1) + 1];
}

GfaHost::~GfaHost() { gfa_destroy(h); }

// IO::ReadFile for the in-scope keyword blocks; other blocks are skipped up to
// the next known keyword (the reference would parse them on the host anyway).
bool GfaHost::ReadFile(const char* path) {
    std::ifstream f(path);
    if (!f) return fail(std::string("cannot open ") + path);
    std::stringstream buf;
    buf << f.rdbuf();
    const std::vector<std::string> tk = tokenize(buf.str());
    size_t i = 0;
    auto num = [&](size_t j) { return j < tk.size() ? atof(tk[j].c_str()) : 0.0; };
    auto integer = [&](size_t j) { return j < tk.size() ? atoi(tk[j].c_str()) : 0; };
    elem_node_ptr.assign(1, 0);
    while (i < tk.size()) {
        const std::string& kw = tk[i];
        if (kw == "EOF") break;
        if (kw == "Nodes") {
            const int n = integer(i + 1); i += 2;
            ref_coordinates.assign(3 * (size_t)n, 0.0);
            for (int r = 0; r < n; r++, i += 5) {
                if (tk[i] != "Node") return fail("Error reading Nodes block");
                const int id = integer(i + 1);
                if (id < 1 || id > n) return fail("Node ids must be consecutive");
                for (int k = 0; k < 3; k++) ref_coordinates[3 * (size_t)(id - 1) + k] = num(i + 2 + k);
            }
        } else if (kw == "Materials") {
            const int n = integer(i + 1); i += 2;
            for (int r = 0; r < n; r++, i += 8) {
                if (tk[i] != "Hooke") return fail("material " + tk[i] + " is outside the accelerated path");
                hooke.push_back(num(i + 3)); hooke.push_back(num(i + 5)); hooke.push_back(num(i + 7));
            }
        } else if (kw == "Sections") {
            const int n = integer(i + 1); i += 2;
            for (int r = 0; r < n; r++) {
                const int kind = tk[i] == "Rectangle" ? 0 : tk[i] == "Tube" ? 1 : -1;
                if (kind < 0) return fail("section " + tk[i] + " is outside the accelerated path");
                section_defs.push_back(kind); section_defs.push_back(num(i + 3)); section_defs.push_back(num(i + 5));
                i += 6;
                if (i < tk.size() && tk[i] == "AD") i += 7;
            }
        } else if (kw == "PipeSections") {     // IO::ReadPipeSections (IO.cpp:1282-1305), PipeSection::Read
            const int n = integer(i + 1); i += 2;
            static const char* names[11] = { "EA", "EI", "GJ", "GA", "Rho", "CDt", "CDn", "CAt", "CAn", "De", "Di" };
            for (int r = 0; r < n; r++, i += 24) {
                if (tk[i] != "PS") return fail("Error reading Pipe Section " + std::to_string(r + 1));
                for (int k = 0; k < 11; k++) {
                    if (tk[i + 2 + 2 * k] != names[k]) return fail("Error reading Pipe Section " + std::to_string(r + 1));
                    pipe_sections.push_back(num(i + 3 + 2 * k));
                }
            }
        } else if (kw == "ShellSections") {
            const int n = integer(i + 1); i += 2;
            for (int r = 0; r < n; r++, i += 4) {
                if (tk[i] != "Homogeneous") return fail("shell section " + tk[i] + " is outside the accelerated path");
                shell_thickness.push_back(num(i + 3));
            }
        } else if (kw == "CoordinateSystems") {
            const int n = integer(i + 1); i += 2;
            for (int r = 0; r < n; r++, i += 10) {
                if (tk[i] != "CS") return fail("Error reading CoordinateSystems block");
                for (int k = 0; k < 3; k++) cs_defs.push_back(num(i + 3 + k));
                for (int k = 0; k < 3; k++) cs_defs.push_back(num(i + 7 + k));
            }
        } else if (kw == "NodeSets") {
            const int n = integer(i + 1); i += 2;
            node_sets.assign(n, std::vector<int>());
            for (int r = 0; r < n; r++) {
                if (tk[i] != "NodeSet") return fail("Error reading NodeSets block");
                const int id = integer(i + 1), cnt = integer(i + 3);
                std::vector<int>& s = node_sets[id - 1];
                if (tk[i + 4] == "List") { for (int k = 0; k < cnt; k++) s.push_back(integer(i + 5 + k)); i += 5 + cnt; }
                else { const int a = integer(i + 6), inc = integer(i + 8); for (int k = 0; k < cnt; k++) s.push_back(a + k * inc); i += 9; }
            }
        } else if (kw == "ElementSets") {          // ElementSet id Elements n List ... | Sequence Initial a Increment k (ElementSet.cpp:27-95)
            const int n = integer(i + 1); i += 2;
            element_sets.assign(n, std::vector<int>());
            for (int r = 0; r < n; r++) {
                if (tk[i] != "ElementSet") return fail("Error reading ElementSets block");
                const int id = integer(i + 1), cnt = integer(i + 3);
                std::vector<int>& s = element_sets[id - 1];
                if (tk[i + 4] == "List") { for (int k = 0; k < cnt; k++) s.push_back(integer(i + 5 + k)); i += 5 + cnt; }
                else { const int a = integer(i + 6), inc = integer(i + 8); for (int k = 0; k < cnt; k++) s.push_back(a + k * inc); i += 9; }
            }
        } else if (kw == "Elements") {
            const int n = integer(i + 1); i += 2;
            for (int r = 0; r < n; r++) {
                const std::string& ty = tk[i];
                int type = 0, nn = 0, mat = ty == "Pipe_1" ? 0 : integer(i + 3), sec = 0, c = 0;
                double T0 = 0.0;
                size_t nodes_at = 0;
                if (ty == "Beam_1") { type = GFA_BEAM_1; nn = 3; sec = integer(i + 5); c = integer(i + 7); nodes_at = i + 9; i += 12;
                    if (i < tk.size() && tk[i] == "PreTension") { T0 = num(i + 1); i += 2; } }
                else if (ty == "Pipe_1") {      // Pipe_1 id PipeSec s CS c Nodes a b c (Pipe_1.cpp:535-572)
                    if (tk[i + 2] != "PipeSec" || tk[i + 4] != "CS" || tk[i + 6] != "Nodes") return fail("Error reading Pipe_1 element");
                    type = GFA_PIPE_1; nn = 3; sec = integer(i + 3); c = integer(i + 5); nodes_at = i + 7; i += 10; }
                else if (ty == "Shell_1") { type = GFA_SHELL_1; nn = 6; sec = integer(i + 5); i += 6;
                    if (tk[i] == "CS") { c = integer(i + 1); i += 2; }
                    nodes_at = i + 1; i += 7; }
                else if (ty == "Solid_1") { type = GFA_SOLID_1; nn = 8; c = integer(i + 5); nodes_at = i + 7; i += 15; }
                else return fail("element type " + ty + " is outside the accelerated path");
                elem_type.push_back(type); elem_material.push_back(mat); elem_section.push_back(sec); elem_cs.push_back(c);
                pretension.push_back(T0);
                for (int k = 0; k < nn; k++) elem_nodes.push_back(integer(nodes_at + k));
                elem_node_ptr.push_back((int)elem_nodes.size());
            }
        } else if (kw == "Constraints") {
            const int n = integer(i + 1); i += 2;
            static const char* names[6] = { "UX", "UY", "UZ", "ROTX", "ROTY", "ROTZ" };
            for (int r = 0; r < n; r++) {
                if (tk[i] != "NodalConstraint") return fail("constraint " + tk[i] + " stays on the host");
                NodalConstraint nc; nc.node_set = integer(i + 3); nc.mask = 0; i += 4;
                for (;;) {
                    int k = -1;
                    for (int q = 0; q < 6 && i < tk.size(); q++) if (tk[i] == names[q]) k = q;
                    if (k < 0) break;
                    i += 2;                                    // keyword + "BoolTable"
                    bool first = true;
                    while (i < tk.size() && isdigit((unsigned char)tk[i][0])) { if (first && integer(i) == 1) nc.mask |= 1 << k; first = false; i++; }
                }
                nodal_constraints.push_back(nc);
            }
        } else if (kw == "Loads") {
            const int n = integer(i + 1); i += 2;
            for (int r = 0; r < n; r++) {
                if (tk[i] == "ShellLoad") {        // ShellLoad id ElementSet s AreaUpdate b NTimes n (ShellLoad.cpp:28-87)
                    GfaShellLoad l; l.element_set = integer(i + 3); l.area_update = integer(i + 5) != 0;
                    const int nt = integer(i + 7); i += 8;
                    for (int k = 0; k < 2 * nt; k++) l.table.push_back(num(i + k));
                    i += 2 * (size_t)nt;
                    shell_loads.push_back(l);
                    continue;
                }
                if (tk[i] == "PipeLoad") {         // PipeLoad id ElementSet s NTimes n, rows time P0I P0E RhoI RhoE (PipeLoad.cpp:44-88)
                    GfaPipeLoad l; l.element_set = integer(i + 3);
                    const int nt = integer(i + 5); i += 6;
                    for (int k = 0; k < 5 * nt; k++) l.table.push_back(num(i + k));
                    i += 5 * (size_t)nt;
                    pipe_loads.push_back(l);
                    continue;
                }
                if (tk[i] != "NodalLoad" && tk[i] != "NodalFollowerLoad") return fail("load " + tk[i] + " is outside the subset this reader keeps");
                GfaNodalLoad l; l.follower = tk[i] == "NodalFollowerLoad"; l.node_set = integer(i + 3); l.cs = integer(i + 5);
                const int nt = integer(i + 7); i += 8;
                for (int k = 0; k < 7 * nt; k++) l.table.push_back(num(i + k));
                i += 7 * (size_t)nt;
                loads.push_back(l);
            }
        } else if (kw == "Environment") {
            i++;
            if (i < tk.size() && tk[i] == "GravityData") {
                g_exist = true;
                for (int k = 0; k < 3; k++) G[k] = num(i + 2 + k);
                i += 5;
                if (i < tk.size() && tk[i] == "BoolTable") { i++; while (i < tk.size() && isdigit((unsigned char)tk[i][0])) i++; }
            }
        } else if (kw == "SolutionSteps") {
            i += 2;
            if (i < tk.size() && tk[i] == "Static") {
                for (size_t j = i + 2; j + 1 < tk.size() && j < i + 20; j += 2) {
                    if (tk[j] == "EndTime") end_time = num(j + 1);
                    if (tk[j] == "TimeStep") time_step = num(j + 1);
                }
                i += 20;
            } else if (i < tk.size() && tk[i] == "Dynamic") {      // Dynamic::Read (Dynamic.cpp:65-222)
                dynamic = true;
                size_t j = i + 2;
                for (; j + 1 < tk.size() && j < i + 20; j += 2) {
                    if (tk[j] == "EndTime") end_time = num(j + 1);
                    if (tk[j] == "TimeStep") time_step = num(j + 1);
                }
                if (j + 6 < tk.size() && tk[j] == "RayleighDamping") { alpha = num(j + 2); beta = num(j + 4); update = integer(j + 6); j += 7; }
                else return fail("Error reading Dynamic solution step: RayleighDamping expected");
                if (j + 4 < tk.size() && tk[j] == "NewmarkCoefficients") { beta_new = num(j + 2); gamma_new = num(j + 4); j += 5; }
                else return fail("Error reading Dynamic solution step: NewmarkCoefficients expected");
                i = j;
            }
        } else {
            i++;
            while (i < tk.size() && !is_top(tk[i])) i++;
        }
    }
    const int n = number_nodes();
    displacements.assign(6 * (size_t)n, 0.0);
    constraints.assign(n, 0);
    for (size_t k = 0; k < elem_nodes.size(); k++)
        if (elem_nodes[k] < 1 || elem_nodes[k] > n) return fail("element references a node that does not exist");
    return n > 0 && !elem_type.empty();
}

// Section::PreCalc (SecRectangle.cpp:82-96, SecTube.cpp:86-94), CoordinateSystem
// normalisation (CoordinateSystem.cpp:64-77), then Element::PreCalc on the device.
bool GfaHost::PreCalc(int device) {
    sections.clear();
    for (size_t s = 0; s + 2 < section_defs.size() + 0 && s < section_defs.size(); s += 3) {
        const int kind = (int)section_defs[s];
        const double a = section_defs[s + 1], b = section_defs[s + 2];
        double A, I11, I22, I33, It;
        if (kind == 0) {
            A = a * b; I11 = a * b * b * b / 12.0; I22 = b * a * a * a / 12.0; I33 = I11 + I22;
            double temp = 0;
            for (int n = 1; n < 22; n = n + 2) temp += (1.0 / (pow((double)n, 5))) * tanh(n * kPi * b / (2 * a));
            It = (1.0 / 3.0) * a * a * a * b * (1.0 - 192.0 * a * temp / (pow(kPi, 5) * b));
        } else {
            A = (kPi / 4.0) * (a * a - b * b);
            I11 = (kPi / 64.0) * (a * a * a * a - b * b * b * b); I22 = I11;
            I33 = (kPi / 32.0) * (a * a * a * a - b * b * b * b); It = I33;
        }
        const double row[6] = { A, I11, I22, 0.0, I33, It };
        sections.insert(sections.end(), row, row + 6);
    }
    cs.clear();
    for (size_t c = 0; c < cs_defs.size(); c += 6) {
        double e1[3] = { cs_defs[c], cs_defs[c + 1], cs_defs[c + 2] }, e3[3] = { cs_defs[c + 3], cs_defs[c + 4], cs_defs[c + 5] };
        double e2[3] = { e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0] };
        double* v[3] = { e1, e2, e3 };
        for (int k = 0; k < 3; k++) {
            const double nrm = sqrt(v[k][0] * v[k][0] + v[k][1] * v[k][1] + v[k][2] * v[k][2]);
            if (nrm != 1.0) for (int q = 0; q < 3; q++) v[k][q] = (1.0 / nrm) * v[k][q];
            cs.insert(cs.end(), v[k], v[k] + 3);
        }
    }
    gfa_model_t m;
    memset(&m, 0, sizeof(m));
    m.n_nodes = number_nodes(); m.ref_coordinates = ref_coordinates.data();
    m.n_materials = (int)(hooke.size() / 3); m.hooke = hooke.data();
    m.n_sections = (int)(sections.size() / 6); m.sections = sections.data();
    m.n_shell_sections = (int)shell_thickness.size(); m.shell_thickness = shell_thickness.data();
    m.n_cs = (int)(cs.size() / 9); m.cs = cs.data();
    m.n_elements = number_elements();
    m.elem_type = elem_type.data(); m.elem_material = elem_material.data(); m.elem_section = elem_section.data();
    m.elem_cs = elem_cs.data(); m.elem_node_ptr = elem_node_ptr.data(); m.elem_nodes = elem_nodes.data();
    m.beam_pretension = pretension.data();
    m.gravity_on = g_exist ? 1 : 0;
    for (int k = 0; k < 3; k++) m.gravity[k] = G[k];
    m.part_rank = 0; m.part_world = 1;
    m.n_pipe_sections = (int)(pipe_sections.size() / 11); m.pipe_sections = pipe_sections.empty() ? nullptr : pipe_sections.data();
    gfa_destroy(h); h = nullptr;
    if (gfa_create(&m, device, &h) != GFA_OK) return fail(gfa_last_error());
    return true;
}

// Solution::DOFsActive: element DOF masks OR-ed per node, then the nodal constraints
// of the current solution step (NodalConstraint::Mount, NodalConstraint.cpp:152-173).
void GfaHost::DOFsActive() {
    const int n = number_nodes();
    active_GL.assign(6 * (size_t)n, 0);
    constraints.assign(n, 0);
    for (int e = 0; e < number_elements(); e++) {
        const int nn = elem_node_ptr[e + 1] - elem_node_ptr[e];
        for (int a = 0; a < nn; a++) {
            const size_t nd = (size_t)(elem_nodes[elem_node_ptr[e] + a] - 1);
            for (int k = 0; k < 3; k++) active_GL[6 * nd + k] = 1;
            const bool rot = elem_type[e] == GFA_BEAM_1 || elem_type[e] == GFA_PIPE_1 || (elem_type[e] == GFA_SHELL_1 && a > 2);   // Shell_1.cpp:51-68
            if (rot) for (int k = 3; k < 6; k++) active_GL[6 * nd + k] = 1;
        }
    }
    for (const NodalConstraint& c : nodal_constraints)
        for (int nd : node_sets[c.node_set - 1]) constraints[nd - 1] |= c.mask;
}

// Solution::SetGlobalDOFs: node-major, DOF-minor; free ids 1.., fixed ids -1, -2, ...
void GfaHost::SetGlobalDOFs() {
    const int n = number_nodes();
    GLs.assign(6 * (size_t)n, 0);
    int GL_free = 0, GL_fixed = 0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 6; j++) {
            if (!active_GL[6 * (size_t)i + j]) continue;
            if ((constraints[i] >> j) & 1) GLs[6 * (size_t)i + j] = --GL_fixed;
            else GLs[6 * (size_t)i + j] = ++GL_free;
        }
    n_GL_free = GL_free; n_GL_fixed = -GL_fixed;
}

// positions NodalLoad::Mount pushes: 3x3 rotational block of every loaded node (NodalLoad.cpp:384-397)
void GfaHost::CollectLoadPattern(std::vector<int>& m, std::vector<int>& r, std::vector<int>& c) {
    for (const GfaNodalLoad& l : loads)
        for (int nd : node_sets[l.node_set - 1])
            for (int lin = l.follower ? -3 : 0; lin < 3; lin++)          // NodalFollowerLoad pushes the node's whole 6 x 6 block
                for (int col = l.follower ? -3 : 0; col < 3; col++) {
                    const int gl = GLs[6 * (size_t)(nd - 1) + 3 + lin], gc = GLs[6 * (size_t)(nd - 1) + 3 + col];
                    if (gl == 0 || gc == 0) continue;
                    m.push_back(gl > 0 ? (gc > 0 ? GFA_AA : GFA_AB) : (gc > 0 ? GFA_BA : GFA_BB));
                    r.push_back(abs(gl) - 1); c.push_back(abs(gc) - 1);
                }
}

bool GfaHost::SetGlobalSize() {
    std::vector<int> m, r, c;
    CollectLoadPattern(m, r, c);
    if (gfa_set_dofs(h, GLs.data(), n_GL_free, n_GL_fixed, (int64_t)m.size(), m.data(), r.data(), c.data()) != GFA_OK) return fail(gfa_last_error());
    // ShellLoad element sets (the reference ignores elements of other types in the set, Shell_1-only virtual)
    if (!shell_loads.empty()) {
        std::vector<int> ptr(1, 0), elems, area;
        for (const GfaShellLoad& l : shell_loads) {
            if (l.element_set < 1 || l.element_set > (int)element_sets.size()) return fail("ShellLoad refers to an ElementSet that does not exist");
            for (int el1 : element_sets[l.element_set - 1])
                if (el1 >= 1 && el1 <= number_elements() && elem_type[el1 - 1] == GFA_SHELL_1) elems.push_back(el1 - 1);
            ptr.push_back((int)elems.size());
            area.push_back(l.area_update ? 1 : 0);
        }
        if (gfa_set_shell_loads(h, (int32_t)shell_loads.size(), ptr.data(), elems.data(), area.data()) != GFA_OK) return fail(gfa_last_error());
    }
    if (!pipe_loads.empty()) {          // PipeLoad::Check refuses a set with another element type (PipeLoad.cpp:91-106); so does the library
        std::vector<int> ptr(1, 0), elems;
        for (const GfaPipeLoad& l : pipe_loads) {
            if (l.element_set < 1 || l.element_set > (int)element_sets.size()) return fail("PipeLoad refers to an ElementSet that does not exist");
            for (int el1 : element_sets[l.element_set - 1]) elems.push_back(el1 - 1);
            ptr.push_back((int)elems.size());
        }
        if (gfa_set_pipe_loads(h, (int32_t)pipe_loads.size(), ptr.data(), elems.data()) != GFA_OK) return fail(gfa_last_error());
    }
    return true;
}

double GfaHost::LoadFactor() const { return (last_converged_time + current_time_step) / end_time; }

bool GfaHost::MountLocal() {
    gfa_step_t st;
    st.displacements = displacements.data(); st.displacements_on_device = 0;
    st.gravity_factor = g_exist ? LoadFactor() : 0.0;
    if (gfa_assemble(h, &st) != GFA_OK) return fail(gfa_last_error());
    return true;
}

// Dynamic::CalculateNewmarkCoeff (Dynamic.cpp:582-590)
void GfaHost::CalculateNewmarkCoeff(double dt) {
    a1 = 1.0 / (dt * dt * beta_new);
    a2 = 1.0 / (dt * beta_new);
    a3 = 1.0 / (2.0 * beta_new) - 1.0;
    a4 = gamma_new / (dt * beta_new);
    a5 = 1.0 - gamma_new / beta_new;
    a6 = dt * (1.0 - gamma_new / (2.0 * beta_new));
}

bool GfaHost::SetKinematics(const double* vel, const double* accel, const double* copy_vel, const double* copy_accel) {
    if (gfa_set_kinematics(h, vel, accel, copy_vel, copy_accel) != GFA_OK) return fail(gfa_last_error());
    return true;
}

// Dynamic::UpdateDyn (Dynamic.cpp:480-556) on the device copy of the nodal arrays
bool GfaHost::UpdateDyn() {
    const gfa_dynamic_t d = { a1, a2, a3, a4, a5, a6, alpha, beta, 0 };
    if (gfa_update_dyn(h, displacements.data(), &d) != GFA_OK) return fail(gfa_last_error());
    return true;
}

// One Newton iteration of Dynamic::Solve up to MountSparse (Dynamic.cpp:323-340)
bool GfaHost::MountLocalDynamic(bool update_rayleigh) {
    gfa_step_t st;
    st.displacements = displacements.data(); st.displacements_on_device = 0;
    st.gravity_factor = g_exist ? LoadFactor() : 0.0;
    const gfa_dynamic_t d = { a1, a2, a3, a4, a5, a6, alpha, beta, update_rayleigh ? 1 : 0 };
    if (gfa_assemble_dynamic(h, &st, &d) != GFA_OK) return fail(gfa_last_error());
    return true;
}

bool GfaHost::GetKinematics(std::vector<double>& vel, std::vector<double>& accel) {
    vel.assign(6 * (size_t)number_nodes(), 0.0); accel.assign(6 * (size_t)number_nodes(), 0.0);
    if (gfa_kinematics(h, vel.data(), accel.data(), nullptr, nullptr) != GFA_OK) return fail(gfa_last_error());
    return true;
}

// NodalLoad::Mount (NodalLoad.cpp:322-401) for numeric tables and the global CS
// convention Q = rows E1,E2,E3: f, m -> Q^T f; pseudo-moment m -> Xi^T m; stiffness -V(alpha, m).
bool GfaHost::MountLoads() {
    std::vector<int> tr[4], tc[4], ia, ib;
    std::vector<double> tv[4], va, vb;
    const double t = last_converged_time + current_time_step;
    for (const GfaNodalLoad& l : loads) {
        const std::vector<int>& set = node_sets[l.node_set - 1];
        double mult[6];
        for (int k = 0; k < 6; k++) { int cnt = 0; for (int nd : set) cnt += active_GL[6 * (size_t)(nd - 1) + k]; mult[k] = 1.0 / cnt; }
        const double* Q = &cs[9 * (size_t)(l.cs - 1)];
        if (l.follower) {       // NodalFollowerLoad::Mount (NodalFollowerLoad.cpp:243-325)
            if (copy_cache.empty()) {
                copy_cache.resize(6 * (size_t)number_nodes());
                if (gfa_copy_coordinates(h, copy_cache.data()) != GFA_OK) return fail(gfa_last_error());
            }
            for (int nd : set) {
                auto rod = [](const double* a, double& g, double* Qm, double* Xi) {
                    const double A[9] = { 0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0 };
                    g = 4.0 / (4.0 + (a[0] * a[0] + a[1] * a[1] + a[2] * a[2]));
                    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
                        double aa = 0.0; for (int k = 0; k < 3; k++) aa += A[3 * i + k] * A[3 * k + j];
                        Qm[3 * i + j] = (i == j ? 1.0 : 0.0) + g * (A[3 * i + j] + 0.5 * aa);
                        Xi[3 * i + j] = g * ((i == j ? 1.0 : 0.0) + 0.5 * A[3 * i + j]);
                    }
                };
                double g, Qc[9], Xc[9], Qd[9], Xi[9], Qi[9];
                rod(&copy_cache[6 * (size_t)(nd - 1) + 3], g, Qc, Xc);
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double v = 0.0; for (int k = 0; k < 3; k++) v += Qc[3 * i + k] * Q[3 * j + k]; Qi[3 * i + j] = v; }   // Qi = Q(copy) * transp(CS Q)
                const double* a = &displacements[6 * (size_t)(nd - 1) + 3];
                rod(a, g, Qd, Xi);
                double fl[3], ml[3], Qf[3], Qm[3], fip[3], mip[3];
                for (int k = 0; k < 3; k++) { fl[k] = mult[k] * l.GetValueAt(t, 1 + k); ml[k] = mult[3 + k] * l.GetValueAt(t, 4 + k); }
                for (int i = 0; i < 3; i++) { Qf[i] = Qi[3 * i] * fl[0] + Qi[3 * i + 1] * fl[1] + Qi[3 * i + 2] * fl[2]; Qm[i] = Qi[3 * i] * ml[0] + Qi[3 * i + 1] * ml[1] + Qi[3 * i + 2] * ml[2]; }
                for (int i = 0; i < 3; i++) { fip[i] = Qd[3 * i] * Qf[0] + Qd[3 * i + 1] * Qf[1] + Qd[3 * i + 2] * Qf[2]; mip[i] = Xi[3 * i] * Qm[0] + Xi[3 * i + 1] * Qm[1] + Xi[3 * i + 2] * Qm[2]; }
                const double Sf[9] = { 0, -fip[2], fip[1], fip[2], 0, -fip[0], -fip[1], fip[0], 0 };
                const double Sm[9] = { 0, -Qm[2], Qm[1], Qm[2], 0, -Qm[0], -Qm[1], Qm[0], 0 };
                double dq[36];
                for (int q = 0; q < 36; q++) dq[q] = 0.0;
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
                    double k12 = 0.0, xo = 0.0;
                    for (int k = 0; k < 3; k++) { k12 += Sf[3 * i + k] * Xi[3 * k + j]; xo += Xi[3 * i + k] * Qm[k]; }
                    dq[6 * i + 3 + j] = -1.0 * k12;                                        // K12 = -skew(fip) Xi
                    dq[6 * (3 + i) + 3 + j] = -0.5 * g * (Sm[3 * i + j] + xo * a[j]);      // K22 = -g/2 (skew(Qi m) + Xi (Qi m) alpha^T)
                }
                const int* gl = &GLs[6 * (size_t)(nd - 1)];
                for (int lin = 0; lin < 6; lin++) {
                    if (gl[lin] == 0) continue;
                    const double v = -1.0 * (lin < 3 ? fip[lin] : mip[lin - 3]);
                    if (gl[lin] > 0) { ia.push_back(gl[lin] - 1); va.push_back(v); } else { ib.push_back(-gl[lin] - 1); vb.push_back(v); }
                    for (int col = 0; col < 6; col++) {
                        const int g1 = gl[lin], g2 = gl[col];
                        if (g2 == 0) continue;
                        const int w = g1 > 0 ? (g2 > 0 ? GFA_AA : GFA_AB) : (g2 > 0 ? GFA_BA : GFA_BB);
                        tr[w].push_back(abs(g1) - 1); tc[w].push_back(abs(g2) - 1); tv[w].push_back(-1.0 * dq[6 * lin + col]);
                    }
                }
            }
            continue;
        }
        for (int nd : set) {
            double fl[3], ml[3], f[3], m[3];
            for (int k = 0; k < 3; k++) { fl[k] = mult[k] * l.GetValueAt(t, 1 + k); ml[k] = mult[3 + k] * l.GetValueAt(t, 4 + k); }
            for (int k = 0; k < 3; k++) { f[k] = Q[k] * fl[0] + Q[3 + k] * fl[1] + Q[6 + k] * fl[2]; m[k] = Q[k] * ml[0] + Q[3 + k] * ml[1] + Q[6 + k] * ml[2]; }
            const double* a = &displacements[6 * (size_t)(nd - 1) + 3];
            const double al = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            const double g = 4.0 / (4.0 + al * al);
            const double A[9] = { 0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0 };
            double Xi[9], mx[3];
            for (int q = 0; q < 9; q++) Xi[q] = g * ((q % 4 == 0 ? 1.0 : 0.0) + 0.5 * A[q]);
            for (int k = 0; k < 3; k++) mx[k] = Xi[k] * m[0] + Xi[3 + k] * m[1] + Xi[6 + k] * m[2];
            const double h2 = 0.5 * g, h4 = -0.25 * g * g, h8 = -0.5 * g * g;
            const double xt[3] = { a[1] * mx[2] - a[2] * mx[1], a[2] * mx[0] - a[0] * mx[2], a[0] * mx[1] - a[1] * mx[0] };
            const double S[9] = { 0, -mx[2], mx[1], mx[2], 0, -mx[0], -mx[1], mx[0], 0 };
            double V[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[3 * i + j] = (h8 * mx[i] - h4 * xt[i]) * a[j] + h2 * S[3 * i + j];
            const int* gl = &GLs[6 * (size_t)(nd - 1)];
            for (int lin = 0; lin < 6; lin++) {
                if (gl[lin] == 0) continue;
                const double v = -1.0 * (lin < 3 ? f[lin] : mx[lin - 3]);
                if (gl[lin] > 0) { ia.push_back(gl[lin] - 1); va.push_back(v); } else { ib.push_back(-gl[lin] - 1); vb.push_back(v); }
            }
            for (int lin = 0; lin < 3; lin++)
                for (int col = 0; col < 3; col++) {
                    const int g1 = gl[3 + lin], g2 = gl[3 + col];
                    if (g1 == 0 || g2 == 0) continue;
                    const int w = g1 > 0 ? (g2 > 0 ? GFA_AA : GFA_AB) : (g2 > 0 ? GFA_BA : GFA_BB);
                    tr[w].push_back(abs(g1) - 1); tc[w].push_back(abs(g2) - 1); tv[w].push_back(-1.0 * V[3 * lin + col]);
                }
        }
    }
    // ShellLoad::Mount -> Shell_1::MountShellSpecialLoads (ShellLoad.cpp:133-148, Shell_1.cpp:1392-1467): the follower
    // pressure is element arithmetic and runs on the device (registered in SetGlobalSize); the host only evaluates
    // the load's time table, as ShellLoad::GetValueAt does
    if (!shell_loads.empty()) {
        std::vector<double> pressures;
        for (const GfaShellLoad& l : shell_loads) pressures.push_back(l.GetValueAt(t));
        if (gfa_apply_shell_loads(h, pressures.data()) != GFA_OK) return fail(gfa_last_error());
    }
    // PipeLoad::Mount -> Pipe_1::MountPipeSpecialLoads (PipeLoad.cpp:117-133, Pipe_1.cpp:1443-1494): on the device too
    if (!pipe_loads.empty()) {
        std::vector<double> p0i;
        for (const GfaPipeLoad& l : pipe_loads) p0i.push_back(l.GetValueAt(t, 0));
        if (gfa_apply_pipe_loads(h, p0i.data()) != GFA_OK) return fail(gfa_last_error());
    }
    for (int w = 0; w < 4; w++)
        if (!tv[w].empty() && gfa_add_host_triplets(h, w, (int64_t)tv[w].size(), tr[w].data(), tc[w].data(), tv[w].data()) != GFA_OK) return fail(gfa_last_error());
    if (!va.empty() && gfa_add_host_vector(h, GFA_P_A, (int64_t)va.size(), ia.data(), va.data()) != GFA_OK) return fail(gfa_last_error());
    if (!vb.empty() && gfa_add_host_vector(h, GFA_P_B, (int64_t)vb.size(), ib.data(), vb.data()) != GFA_OK) return fail(gfa_last_error());
    return true;
}

void GfaHost::UpdateDisps(const double* x_A) {
    for (size_t k = 0; k < GLs.size(); k++) if (GLs[k] > 0) displacements[k] += x_A[GLs[k] - 1];
}

bool GfaHost::SaveConfiguration() {
    if (gfa_commit_state(h) != GFA_OK) return fail(gfa_last_error());
    copy_cache.clear();                                                  // copy_coordinates moved
    std::fill(displacements.begin(), displacements.end(), 0.0);          // Solution::Zeros of the next increment
    return true;
}

bool GfaHost::GetCSR(int which, std::vector<int>& outer, std::vector<int>& inner, std::vector<double>& values) {
    int32_t rows, cols; int64_t nnz;
    if (gfa_csr_dims(h, which, &rows, &cols, &nnz) != GFA_OK) return fail(gfa_last_error());
    outer.assign((size_t)rows + 1, 0); inner.assign((size_t)nnz, 0); values.assign((size_t)nnz, 0.0);
    if (gfa_csr_pattern(h, which, outer.data(), inner.data()) != GFA_OK) return fail(gfa_last_error());
    if (gfa_csr_values(h, which, values.data()) != GFA_OK) return fail(gfa_last_error());
    return true;
}

// What WriteResults / WriteMonitor read from the elements after Mount (Shell_1.cpp:624-707,
// Beam_1.cpp:444-497, Monitor.cpp:494): one record per element of the type, see include/gfa.h
bool GfaHost::GetGaussPointResults(int element_type, std::vector<double>& out) {
    const int stride = gfa_results_stride(element_type);
    if (stride == 0) return fail("element type keeps no Gauss-point results");
    size_t n = 0;
    for (int t : elem_type) if (t == element_type) n++;
    out.assign(n * stride, 0.0);
    if (n == 0) return true;
    const int64_t got = gfa_gauss_point_results(h, element_type, out.data(), (int64_t)out.size());
    if (got < 0) return fail(gfa_last_error());
    out.resize((size_t)got * stride);
    return true;
}

bool GfaHost::GetVector(int which, std::vector<double>& v) {
    v.assign(which == GFA_P_B ? n_GL_fixed : n_GL_free, 0.0);
    if (gfa_vector(h, which, v.data()) != GFA_OK) return fail(gfa_last_error());
    return true;
}
