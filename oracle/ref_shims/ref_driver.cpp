// TEST INFRASTRUCTURE (oracle) -- C entry points that drive the UNMODIFIED
// reference classes (Node, Beam_1, Shell_1, Solid_1, Solution, ...) compiled
// from /root/reference/src exactly the way Static::Solve does
// (reference Static.cpp:161-163 set-up, :203-212 per Newton iteration):
//   DOFsActive -> SetGlobalDOFs -> SetGlobalSize            (ref_setup_dofs)
//   Clear -> MountLocal -> MountElementLoads -> MountLoads
//         -> MountGlobal -> MountSparse                     (ref_assemble)
//   Node::SaveConfiguration + Element::SaveLagrange          (ref_commit,
//                                      reference Solution.cpp:426-454)
// Phase timers use std::chrono around the same calls the reference brackets
// with its own probes (Solution.cpp:229-247,253-264,324-348,853-862).
// Loaded by tests/ and bench.py through ctypes; never by the product.
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <vector>
#include <omp.h>

#include "Database.h"
#include "Node.h"
#include "Element.h"
#include "Beam_1.h"
#include "Pipe_1.h"
#include "PipeSection.h"
#include "Shell_1.h"
#include "Solid_1.h"
#include "Hooke.h"
#include "SecTube.h"
#include "SecRectangle.h"
#include "ShellSectionHomogeneous.h"
#include "CoordinateSystem.h"
#include "NodeSet.h"
#include "NodalConstraint.h"
#include "NodalLoad.h"
#include "ShellLoad.h"
#include "PipeLoad.h"
#include "NodalFollowerLoad.h"
#include "ElementSet.h"
#include "Environment.h"
#include "Solution.h"
#include "Dynamic.h"
#include "LagrangeSave.h"
#include "ConvergenceCriteria.h"

extern Database db;

// SURVEY.md 8c hazard: the reference calls unqualified abs() on doubles.
static_assert(std::is_same<decltype(abs(1.5)), double>::value,
	"abs(double) must not resolve to the C int abs()");

namespace {

// A concrete Solution so that the reference's shared assembly steps
// (non-virtual members of Solution) can be called directly.
struct OracleSolution : public Solution
{
	bool Read(FILE*) { return true; }
	void Write(FILE*) {}
	bool Solve() { return true; }
};

Solution* g_sol = NULL;          // OracleSolution (static steps) or the reference's own Dynamic
Dynamic* g_dyn = NULL;

template <class T> T** grow(T** arr, int n_old)
{
	T** n = new T*[n_old + 1];
	for (int i = 0; i < n_old; i++) n[i] = arr[i];
	delete[] arr;
	return n;
}

FILE* text_stream(const char* s) { return fmemopen((void*)s, strlen(s), "r"); }

SparseMatrix* pick(int which)
{
	switch (which)
	{
	case 0: return &db.global_stiffness_AA;
	case 1: return &db.global_stiffness_AB;
	case 2: return &db.global_stiffness_BA;
	case 3: return &db.global_stiffness_BB;
	}
	return NULL;
}

double now_s()
{
	using namespace std::chrono;
	return duration_cast<duration<double> >(high_resolution_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" {

// Start a new model.  Previous objects are abandoned (the oracle process is
// short-lived); only counters and arrays are reset.
int ref_reset()
{
	db.number_nodes = 0; db.nodes = NULL;
	db.number_elements = 0; db.elements = NULL;
	db.number_materials = 0; db.materials = NULL;
	db.number_sections = 0; db.sections = NULL;
	db.number_shell_sections = 0; db.shell_sections = NULL;
	db.number_pipe_sections = 0; db.pipe_sections = NULL; db.pipe_sections_exist = false;
	db.number_CS = 0; db.CS = NULL;
	db.number_node_sets = 0; db.node_sets = NULL;
	db.number_constraints = 0; db.constraints = NULL;
	db.number_loads = 0; db.loads = NULL;
	db.number_element_sets = 0; db.element_sets = NULL;
	db.environment = NULL; db.environment_exist = false;
	db.n_GL_free = 0; db.n_GL_fixed = 0;
	g_dyn = NULL;
	if (!g_sol || dynamic_cast<Dynamic*>(g_sol))
	{
		g_sol = new OracleSolution();
		g_sol->solution_number = 1;
		g_sol->start_time = 0.0;
		g_sol->end_time = 1.0;
		db.solution = new Solution*[1];
		db.solution[0] = g_sol;
		db.number_solutions = 1;
	}
	db.current_solution_number = 1;
	db.last_converged_time = 0.0;
	db.current_time_step = 1.0;
	return 0;
}

int ref_set_threads(int n) { omp_set_num_threads(n); return omp_get_max_threads(); }
int ref_max_threads() { return omp_get_max_threads(); }

int ref_set_nodes(int n, const double* xyz)
{
	db.nodes = new Node*[n];
	db.number_nodes = n;
	for (int i = 0; i < n; i++)
	{
		Node* nd = new Node(db.number_GLs_node);
		nd->number = i + 1;
		for (int k = 0; k < 3; k++)
		{
			nd->ref_coordinates[k] = xyz[3 * i + k];
			nd->copy_coordinates[k] = xyz[3 * i + k];
		}
		db.nodes[i] = nd;
	}
	return 0;
}

int ref_add_hooke(double E, double nu, double rho)
{
	Hooke* h = new Hooke();
	h->number = db.number_materials + 1;
	h->E = E; h->nu = nu; h->rho = rho;
	db.materials = grow(db.materials, db.number_materials);
	db.materials[db.number_materials++] = h;
	return h->number;
}

// kind 0: Rectangle(B=a,H=b)   kind 1: Tube(De=a,Di=b); PreCalc is the reference's.
int ref_add_section(int kind, double a, double b)
{
	Section* s = NULL;
	if (kind == 0) { SecRectangle* r = new SecRectangle(); r->b = a; r->h = b; s = r; }
	else if (kind == 1) { SecTube* t = new SecTube(); t->De = a; t->Di = b; s = t; }
	else return -1;
	s->number = db.number_sections + 1;
	s->PreCalc();
	db.sections = grow(db.sections, db.number_sections);
	db.sections[db.number_sections++] = s;
	return s->number;
}
int ref_get_section(int id, double* out6)
{
	Section* s = db.sections[id - 1];
	out6[0] = s->A; out6[1] = s->I11; out6[2] = s->I22; out6[3] = s->I12; out6[4] = s->I33; out6[5] = s->It;
	return 0;
}

// PipeSection constants EA EI GJ GA Rho CDt CDn CAt CAn De Di (reference PipeSection.h:13-23)
int ref_add_pipe_section(const double* v11)
{
	PipeSection* s = new PipeSection();
	s->number = db.number_pipe_sections + 1;
	s->EA = v11[0]; s->EI = v11[1]; s->GJ = v11[2]; s->GA = v11[3]; s->Rho = v11[4];
	s->CDt = v11[5]; s->CDn = v11[6]; s->CAt = v11[7]; s->CAn = v11[8]; s->De = v11[9]; s->Di = v11[10];
	db.pipe_sections = grow(db.pipe_sections, db.number_pipe_sections);
	db.pipe_sections[db.number_pipe_sections++] = s;
	db.pipe_sections_exist = true;
	return s->number;
}

int ref_add_shell_section(double thickness)
{
	ShellSectionHomogeneous* s = new ShellSectionHomogeneous();
	s->number = db.number_shell_sections + 1;
	s->thickness = thickness;
	db.shell_sections = grow(db.shell_sections, db.number_shell_sections);
	db.shell_sections[db.number_shell_sections++] = s;
	return s->number;
}

// Goes through the reference's own reader so that normalisation and the Q
// matrix are the reference's (CoordinateSystem.cpp:23-96).
int ref_add_cs(const double* e1, const double* e3)
{
	char buf[512];
	snprintf(buf, sizeof(buf), "CS %d E1 %.17g %.17g %.17g E3 %.17g %.17g %.17g",
		db.number_CS + 1, e1[0], e1[1], e1[2], e3[0], e3[1], e3[2]);
	FILE* f = text_stream(buf);
	CoordinateSystem* cs = new CoordinateSystem();
	bool ok = cs->Read(f);
	fclose(f);
	if (!ok) return -1;
	db.CS = grow(db.CS, db.number_CS);
	db.CS[db.number_CS++] = cs;
	return cs->number;
}
int ref_get_cs(int id, double* e123)
{
	CoordinateSystem* c = db.CS[id - 1];
	for (int k = 0; k < 3; k++) { e123[k] = (*c->E1)(k, 0); e123[3 + k] = (*c->E2)(k, 0); e123[6 + k] = (*c->E3)(k, 0); }
	return 0;
}

// type: 1 Beam_1 (3 nodes), 2 Pipe_1 (3 nodes), 3 Shell_1 (6 nodes), 7 Solid_1 (8 nodes)
// (ids as listed in reference Element.h:8-15).  conn is 1-based, packed.
int ref_set_elements(int n, const int* type, const int* mat, const int* sec, const int* cs,
	const int* conn, const double* pretension)
{
	db.elements = new Element*[n];
	db.number_elements = n;
	long p = 0;
	for (int e = 0; e < n; e++)
	{
		Element* el = NULL;
		int nn = 0;
		if (type[e] == 1) { Beam_1* b = new Beam_1(); b->T0 = pretension ? pretension[e] : 0.0; el = b; nn = 3; }
		else if (type[e] == 2) { el = new Pipe_1(); nn = 3; }
		else if (type[e] == 3) { el = new Shell_1(); nn = 6; }
		else if (type[e] == 7) { el = new Solid_1(); nn = 8; }
		else return -1;
		el->number = e + 1;
		el->material = mat[e];
		el->section = sec[e];
		el->cs = cs[e];
		for (int k = 0; k < nn; k++) el->nodes[k] = conn[p + k];
		p += nn;
		db.elements[e] = el;
	}
	return 0;
}

int ref_set_gravity(double gx, double gy, double gz)
{
	Environment* env = new Environment();
	env->g_exist = true;
	env->G(0, 0) = gx; env->G(1, 0) = gy; env->G(2, 0) = gz;
	env->bool_g.SetDefault(true);
	db.environment = env;
	db.environment_exist = true;
	return 0;
}

static int add_node_set(int n, const int* nodes)
{
	NodeSet* ns = new NodeSet();
	ns->number = db.number_node_sets + 1;
	ns->n_nodes = n;
	ns->list = true;
	ns->node_list = new int[n];
	for (int i = 0; i < n; i++) ns->node_list[i] = nodes[i];
	db.node_sets = grow(db.node_sets, db.number_node_sets);
	db.node_sets[db.number_node_sets++] = ns;
	return ns->number;
}

// mask bit k set => DOF k (UX,UY,UZ,ROTX,ROTY,ROTZ) constrained in every step.
int ref_add_nodal_constraint(int n, const int* nodes, int mask)
{
	NodalConstraint* c = new NodalConstraint();
	c->number = db.number_constraints + 1;
	c->node_set = add_node_set(n, nodes);
	BoolTable* t[6] = { &c->UX_table, &c->UY_table, &c->UZ_table, &c->ROTX_table, &c->ROTY_table, &c->ROTZ_table };
	for (int k = 0; k < 6; k++) t[k]->SetDefault(((mask >> k) & 1) != 0);
	db.constraints = grow(db.constraints, db.number_constraints);
	db.constraints[db.number_constraints++] = c;
	return c->number;
}

// table rows: time FX FY FZ MX MY MZ  (reference NodalLoad.cpp:41-86 format)
int ref_add_nodal_load(int n, const int* nodes, int cs, int n_times, const double* table7)
{
	std::vector<char> buf(256 + 200 * (size_t)n_times);
	int set_id = add_node_set(n, nodes);
	int w = snprintf(buf.data(), buf.size(), "%d NodeSet %d CS %d NTimes %d\n", db.number_loads + 1, set_id, cs, n_times);
	for (int r = 0; r < n_times; r++)
	{
		for (int k = 0; k < 7; k++)
			w += snprintf(buf.data() + w, buf.size() - w, "%.17g ", table7[7 * r + k]);
		w += snprintf(buf.data() + w, buf.size() - w, "\n");
	}
	FILE* f = text_stream(buf.data());
	NodalLoad* l = new NodalLoad();
	bool ok = l->Read(f);
	fclose(f);
	if (!ok) return -1;
	db.loads = grow(db.loads, db.number_loads);
	db.loads[db.number_loads++] = l;
	return l->number;
}

// NodalFollowerLoad (reference NodalFollowerLoad.cpp:56-112 format, same table as NodalLoad): forces and moments that
// follow the node's rotation.  PreCalc allocates its per-node buffers (:144-151).
int ref_add_nodal_follower_load(int n, const int* nodes, int cs, int n_times, const double* table7)
{
	std::vector<char> buf(256 + 200 * (size_t)n_times);
	int set_id = add_node_set(n, nodes);
	int w = snprintf(buf.data(), buf.size(), "%d NodeSet %d CS %d NTimes %d\n", db.number_loads + 1, set_id, cs, n_times);
	for (int r = 0; r < n_times; r++)
	{
		for (int k = 0; k < 7; k++)
			w += snprintf(buf.data() + w, buf.size() - w, "%.17g ", table7[7 * r + k]);
		w += snprintf(buf.data() + w, buf.size() - w, "\n");
	}
	FILE* f = text_stream(buf.data());
	NodalFollowerLoad* l = new NodalFollowerLoad();
	bool ok = l->Read(f);
	fclose(f);
	if (!ok) return -1;
	l->PreCalc();
	db.loads = grow(db.loads, db.number_loads);
	db.loads[db.number_loads++] = l;
	return l->number;
}

// ShellLoad over an ElementSet (reference ShellLoad.cpp:28-87 format): follower pressure on Shell_1 elements,
// table rows: time pressure.  Goes through the reference's own reader; Shell_1::MountShellSpecialLoads
// (Shell_1.cpp:1392-1467) then folds it into the element block during MountLoads.
int ref_add_shell_load(int n_el, const int* elements, int area_update, int n_times, const double* table2)
{
	ElementSet* es = new ElementSet();
	es->number = db.number_element_sets + 1;
	es->n_el = n_el;
	es->list = true;
	es->el_list = new int[n_el];
	for (int i = 0; i < n_el; i++) es->el_list[i] = elements[i];
	db.element_sets = grow(db.element_sets, db.number_element_sets);
	db.element_sets[db.number_element_sets++] = es;
	std::vector<char> buf(256 + 100 * (size_t)n_times);
	int w = snprintf(buf.data(), buf.size(), "%d ElementSet %d AreaUpdate %d NTimes %d\n", db.number_loads + 1, es->number, area_update ? 1 : 0, n_times);
	for (int r = 0; r < n_times; r++)
		w += snprintf(buf.data() + w, buf.size() - w, "%.17g %.17g\n", table2[2 * r], table2[2 * r + 1]);
	FILE* f = text_stream(buf.data());
	ShellLoad* l = new ShellLoad();
	bool ok = l->Read(f);
	fclose(f);
	if (!ok) return -1;
	db.loads = grow(db.loads, db.number_loads);
	db.loads[db.number_loads++] = l;
	return l->number;
}

// PipeLoad over an ElementSet (reference PipeLoad.cpp:44-88 format): internal / external pressure on Pipe_1 elements,
// table rows: time P0I P0E RhoI RhoE.  Goes through the reference's own reader; Pipe_1::MountPipeSpecialLoads
// (Pipe_1.cpp:1443-1494) folds it into the element block during MountLoads.
int ref_add_pipe_load(int n_el, const int* elements, int n_times, const double* table5)
{
	ElementSet* es = new ElementSet();
	es->number = db.number_element_sets + 1;
	es->n_el = n_el;
	es->list = true;
	es->el_list = new int[n_el];
	for (int i = 0; i < n_el; i++) es->el_list[i] = elements[i];
	db.element_sets = grow(db.element_sets, db.number_element_sets);
	db.element_sets[db.number_element_sets++] = es;
	std::vector<char> buf(256 + 200 * (size_t)n_times);
	int w = snprintf(buf.data(), buf.size(), "%d ElementSet %d NTimes %d\n", db.number_loads + 1, es->number, n_times);
	for (int r = 0; r < n_times; r++)
	{
		for (int k = 0; k < 5; k++)
			w += snprintf(buf.data() + w, buf.size() - w, "%.17g ", table5[5 * r + k]);
		w += snprintf(buf.data() + w, buf.size() - w, "\n");
	}
	FILE* f = text_stream(buf.data());
	PipeLoad* l = new PipeLoad();
	bool ok = l->Read(f);
	fclose(f);
	if (!ok) return -1;
	db.loads = grow(db.loads, db.number_loads);
	db.loads[db.number_loads++] = l;
	return l->number;
}

int ref_check()
{
	for (int i = 0; i < db.number_elements; i++)
		if (!db.elements[i]->Check()) return i + 1;
	return 0;
}

int ref_precalc()
{
	for (int i = 0; i < db.number_elements; i++)
		db.elements[i]->PreCalc();
	return 0;
}

int ref_setup_dofs()
{
	g_sol->DOFsActive();
	g_sol->SetGlobalDOFs();
	g_sol->SetGlobalSize();
	return 0;
}
int ref_n_free() { return db.n_GL_free; }
int ref_n_fixed() { return db.n_GL_fixed; }
int ref_get_gls(int* gls)
{
	for (int i = 0; i < db.number_nodes; i++)
		for (int k = 0; k < 6; k++) gls[6 * i + k] = db.nodes[i]->GLs[k];
	return 0;
}

int ref_set_time(double last_converged, double step)
{
	db.last_converged_time = last_converged;
	db.current_time_step = step;
	return 0;
}

int ref_set_displacements(const double* d)
{
	for (int i = 0; i < db.number_nodes; i++)
		for (int k = 0; k < 6; k++) db.nodes[i]->displacements[k] = d[6 * i + k];
	return 0;
}
int ref_get_copy_coordinates(double* c)
{
	for (int i = 0; i < db.number_nodes; i++)
		for (int k = 0; k < 6; k++) c[6 * i + k] = db.nodes[i]->copy_coordinates[k];
	return 0;
}

// seconds[0..3] = MountLocal, MountElementLoads, MountGlobal, MountSparse;
// seconds[4] = Clear + MountLoads (not part of the like-for-like sum).
int ref_assemble(int with_loads, double* seconds)
{
	double t0 = now_s();
	g_sol->Clear();
	double t1 = now_s();
	g_sol->MountLocal();
	double t2 = now_s();
	g_sol->MountElementLoads();
	double t3 = now_s();
	if (with_loads) g_sol->MountLoads();
	double t4 = now_s();
	g_sol->MountGlobal();
	double t5 = now_s();
	g_sol->MountSparse();
	double t6 = now_s();
	if (seconds)
	{
		seconds[0] = t2 - t1; seconds[1] = t3 - t2; seconds[2] = t5 - t4; seconds[3] = t6 - t5;
		seconds[4] = (t1 - t0) + (t4 - t3);
	}
	return 0;
}

// MountLocal + MountElementLoads only (the OpenMP-parallel part).
int ref_mount_local(double* seconds)
{
	double t1 = now_s();
	g_sol->MountLocal();
	double t2 = now_s();
	g_sol->MountElementLoads();
	double t3 = now_s();
	if (seconds) { seconds[0] = t2 - t1; seconds[1] = t3 - t2; }
	return 0;
}

long ref_triplet_count(int which) { return (long)pick(which)->tripletList.size(); }
int ref_csr_rows(int which) { return (int)pick(which)->m_matrix.rows(); }
int ref_csr_cols(int which) { return (int)pick(which)->m_matrix.cols(); }
long ref_csr_nnz(int which) { return (long)pick(which)->m_matrix.nonZeros(); }
int ref_csr_get(int which, int* outer, int* inner, double* val)
{
	SparseMatrix* m = pick(which);
	long nr = m->m_matrix.rows(), nz = m->m_matrix.nonZeros();
	if (outer) memcpy(outer, m->m_matrix.outerIndexPtr(), sizeof(int) * (size_t)(nr + 1));
	if (inner) memcpy(inner, m->m_matrix.innerIndexPtr(), sizeof(int) * (size_t)nz);
	if (val) memcpy(val, m->m_matrix.valuePtr(), sizeof(double) * (size_t)nz);
	return 0;
}
int ref_get_vectors(double* PA, double* IA, double* PB)
{
	for (int i = 0; i < db.n_GL_free; i++) { if (PA) PA[i] = db.global_P_A(i, 0); if (IA) IA[i] = db.global_I_A(i, 0); }
	for (int i = 0; i < db.n_GL_fixed; i++) if (PB) PB[i] = db.global_P_B(i, 0);
	return 0;
}

// Element block after Mount + MountElementLoads: K row-major n x n in the
// element's own local DOF order, P = P_loading, and the strain energy.
int ref_get_element(int e, double* K, double* P, double* energy)
{
	Element* el = db.elements[e];
	Matrix* k = NULL; Matrix* p = NULL;
	if (Beam_1* b = dynamic_cast<Beam_1*>(el)) { k = b->stiffness; p = b->P_loading; }
	else if (Pipe_1* q = dynamic_cast<Pipe_1*>(el)) { k = q->stiffness; p = q->P_loading; }
	else if (Shell_1* s = dynamic_cast<Shell_1*>(el)) { k = s->stiffness; p = s->P_loading; }
	else return el->nDOFs;
	int n = el->nDOFs;
	for (int i = 0; i < n; i++)
	{
		if (P) P[i] = (*p)(i, 0);
		if (K) for (int j = 0; j < n; j++) K[i * n + j] = (*k)(i, j);
	}
	if (energy) *energy = el->strain_energy;
	return n;
}

// Committed Gauss-point state in the layout the C-ABI uploads/downloads:
//  Shell_1 : per point  Q_i(9, row-major) z_x1_i(3) z_x2_i(3) kappa_r1_i(3) kappa_r2_i(3)   (3 x 21)
//  Beam_1  : per point  Q_i(9, row-major) dz_i(3) kappa_i_ref(3)                               (2 x 15)
int ref_get_state(int e, double* out)
{
	Element* el = db.elements[e];
	int w = 0;
	if (Shell_1* s = dynamic_cast<Shell_1*>(el))
	{
		for (int g = 0; g < 3; g++)
		{
			for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[w++] = (*s->Q_i[g])(i, j);
			for (int i = 0; i < 3; i++) out[w++] = (*s->z_x1_i[g])(i, 0);
			for (int i = 0; i < 3; i++) out[w++] = (*s->z_x2_i[g])(i, 0);
			for (int i = 0; i < 3; i++) out[w++] = (*s->kappa_r1_i[g])(i, 0);
			for (int i = 0; i < 3; i++) out[w++] = (*s->kappa_r2_i[g])(i, 0);
		}
	}
	else if (Beam_1* b = dynamic_cast<Beam_1*>(el))
	{
		for (int g = 0; g < 2; g++)
		{
			for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[w++] = (*b->lag_save->Q_i[g])(i, j);
			for (int i = 0; i < 3; i++) out[w++] = (*b->lag_save->dz_i[g])(i, 0);
			for (int i = 0; i < 3; i++) out[w++] = (*b->lag_save->kappa_i_ref[g])(i, 0);
		}
	}
	else if (Pipe_1* q = dynamic_cast<Pipe_1*>(el))
	{
		for (int g = 0; g < 2; g++)
		{
			for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[w++] = (*q->lag_save->Q_i[g])(i, j);
			for (int i = 0; i < 3; i++) out[w++] = (*q->lag_save->dz_i[g])(i, 0);
			for (int i = 0; i < 3; i++) out[w++] = (*q->lag_save->kappa_i_ref[g])(i, 0);
		}
	}
	return w;
}

// Gauss-point results the elements keep after Mount for WriteResults / WriteMonitor, in the layout
// of gfa_gauss_point_results (include/gfa.h):
//  Shell_1 : strain_energy, 3 x [eta_r1 eta_r2 kappa_r1 kappa_r2 n_r1 n_r2 m_r1 m_r2]   (73)
//  Beam_1  : strain_energy, 2 x [epsilon_r(6) sigma_r(6)]                                 (25)
int ref_get_results(int e, double* out)
{
	Element* el = db.elements[e];
	int w = 0;
	if (Shell_1* s = dynamic_cast<Shell_1*>(el))
	{
		out[w++] = s->strain_energy;
		for (int g = 0; g < 3; g++)
		{
			Matrix* v[8] = { s->eta_r1[g], s->eta_r2[g], s->kappa_r1[g], s->kappa_r2[g], s->n_r1[g], s->n_r2[g], s->m_r1[g], s->m_r2[g] };
			for (int k = 0; k < 8; k++) for (int i = 0; i < 3; i++) out[w++] = (*v[k])(i, 0);
		}
	}
	else if (Beam_1* b = dynamic_cast<Beam_1*>(el))
	{
		out[w++] = b->strain_energy;
		for (int g = 0; g < 2; g++)
		{
			for (int i = 0; i < 6; i++) out[w++] = (*b->epsilon_r[g])(i, 0);
			for (int i = 0; i < 6; i++) out[w++] = (*b->sigma_r[g])(i, 0);
		}
	}
	else if (Pipe_1* q = dynamic_cast<Pipe_1*>(el))
	{
		out[w++] = q->strain_energy;
		for (int g = 0; g < 2; g++)
		{
			for (int i = 0; i < 6; i++) out[w++] = (*q->epsilon_r[g])(i, 0);
			for (int i = 0; i < 6; i++) out[w++] = (*q->sigma_r[g])(i, 0);
		}
	}
	return w;
}

// What Solution::SaveConfiguration does for nodes and elements
// (reference Solution.cpp:426-454), followed by Zeros() of the increments
// as the next time increment would (Static.cpp:191).
int ref_commit()
{
	for (int i = 0; i < db.number_nodes; i++)
		db.nodes[i]->SaveConfiguration();
#pragma omp parallel for
	for (int i = 0; i < db.number_elements; i++)
		db.elements[i]->SaveLagrange();
	for (int i = 0; i < db.number_nodes; i++)
		for (int k = 0; k < 6; k++) db.nodes[i]->displacements[k] = 0.0;
	return 0;
}

// ---- Dynamic (Newmark) path: the steps Dynamic::Solve runs per time step and per Newton iteration
// (reference Dynamic.cpp:303-340), on the reference's own Dynamic object so that the typeid checks in
// Element::MountMass / MountDamping (Beam_1.cpp:1566, Shell_1.cpp:2478) see a Dynamic solution.
int ref_dynamic_begin(double beta_new, double gamma_new, double rayleigh_alpha, double rayleigh_beta, int update)
{
	g_dyn = new Dynamic();
	g_dyn->solution_number = 1;
	g_dyn->start_time = 0.0;
	g_dyn->end_time = 1.0;
	g_dyn->beta_new = beta_new; g_dyn->gamma_new = gamma_new;
	g_dyn->alpha = rayleigh_alpha; g_dyn->beta = rayleigh_beta; g_dyn->update = update;
	db.solution[0] = g_dyn;
	g_sol = g_dyn;
	return 0;
}
// Dynamic::CalculateNewmarkCoeff (Dynamic.cpp:582-590); a6[0..5] = a1..a6
int ref_newmark(double time_step, double* a6)
{
	if (!g_dyn) return -1;
	g_dyn->CalculateNewmarkCoeff(time_step);
	a6[0] = g_dyn->a1; a6[1] = g_dyn->a2; a6[2] = g_dyn->a3; a6[3] = g_dyn->a4; a6[4] = g_dyn->a5; a6[5] = g_dyn->a6;
	return 0;
}
// Node::vel / accel / copy_vel / copy_accel, [n_nodes*6] each; NULL pointers are skipped
int ref_set_kinematics(const double* vel, const double* accel, const double* copy_vel, const double* copy_accel)
{
	for (int i = 0; i < db.number_nodes; i++)
		for (int k = 0; k < 6; k++)
		{
			if (vel) db.nodes[i]->vel[k] = vel[6 * i + k];
			if (accel) db.nodes[i]->accel[k] = accel[6 * i + k];
			if (copy_vel) db.nodes[i]->copy_vel[k] = copy_vel[6 * i + k];
			if (copy_accel) db.nodes[i]->copy_accel[k] = copy_accel[6 * i + k];
		}
	return 0;
}
int ref_get_kinematics(double* vel, double* accel, double* copy_vel, double* copy_accel)
{
	for (int i = 0; i < db.number_nodes; i++)
		for (int k = 0; k < 6; k++)
		{
			if (vel) vel[6 * i + k] = db.nodes[i]->vel[k];
			if (accel) accel[6 * i + k] = db.nodes[i]->accel[k];
			if (copy_vel) copy_vel[6 * i + k] = db.nodes[i]->copy_vel[k];
			if (copy_accel) copy_accel[6 * i + k] = db.nodes[i]->copy_accel[k];
		}
	return 0;
}
// Dynamic::UpdateDyn (Dynamic.cpp:480-580)
int ref_update_dyn()
{
	if (!g_dyn) return -1;
	g_dyn->UpdateDyn();
	return 0;
}
// One Newton iteration of Dynamic::Solve (Dynamic.cpp:323-340) up to MountSparse
int ref_assemble_dynamic(int with_loads, int update_rayleigh)
{
	if (!g_dyn) return -1;
	g_sol->Clear();
	g_sol->MountLocal();
	g_sol->MountElementLoads();
	if (with_loads) g_sol->MountLoads();
	g_sol->MountMass();
	g_sol->MountDamping(update_rayleigh != 0);
	g_sol->MountDyn();
	g_sol->MountGlobal();
	g_sol->MountSparse();
	return 0;
}
// Committed Rodrigues vector alpha_i of every Gauss point (Shell_1: 3 x 3, Beam_1: 2 x 3)
int ref_get_alpha_i(int e, double* out)
{
	Element* el = db.elements[e];
	int w = 0;
	if (Shell_1* s = dynamic_cast<Shell_1*>(el))
		for (int g = 0; g < 3; g++) for (int i = 0; i < 3; i++) out[w++] = (*s->alpha_i[g])(i, 0);
	else if (Beam_1* b = dynamic_cast<Beam_1*>(el))
		for (int g = 0; g < 2; g++) for (int i = 0; i < 3; i++) out[w++] = (*b->lag_save->alpha_i[g])(i, 0);
	else if (Pipe_1* q = dynamic_cast<Pipe_1*>(el))
		for (int g = 0; g < 2; g++) for (int i = 0; i < 3; i++) out[w++] = (*q->lag_save->alpha_i[g])(i, 0);
	return w;
}

// The Newton-loop steps either side of the assembly, through the reference's own code:
// Static.cpp:210-217 (sign flip, imposed displacements), ConvergenceCriteria::EstablishResidualCriteria /
// CheckResidualConvergence (node_force, node_moment), Solution::UpdateDisps and CheckGLConvergence
// (node_disp, node_rot).  out4 = node_force, node_moment (residual) or node_disp, node_rot (update), diverged, 0.
int ref_residual(const double* XB, int* out4)
{
	db.global_P_A = -1.0*db.global_P_A;
	if (XB)
	{
		for (int i = 0; i < db.n_GL_fixed; i++) db.global_X_B(i, 0) = XB[i];
		db.global_P_A = db.global_P_A - 1.0*(db.global_stiffness_AB*db.global_X_B);
	}
	ConvergenceCriteria* c = db.conv_criteria;
	c->diverged = false; c->node_force = 0; c->node_moment = 0;
	c->EstablishResidualCriteria();
	c->CheckResidualConvergence();
	out4[0] = c->node_force; out4[1] = c->node_moment; out4[2] = c->diverged ? 1 : 0; out4[3] = 0;
	return 0;
}
int ref_update_displacements(const double* xA, int* out4, double* disp_out)
{
	for (int i = 0; i < db.n_GL_free; i++) db.global_P_A(i, 0) = xA[i];
	g_sol->UpdateDisps();
	ConvergenceCriteria* c = db.conv_criteria;
	c->diverged = false; c->node_disp = 0; c->node_rot = 0;
	c->CheckGLConvergence();
	out4[0] = c->node_disp; out4[1] = c->node_rot; out4[2] = c->diverged ? 1 : 0; out4[3] = 0;
	for (int i = 0; i < db.number_nodes; i++) for (int k = 0; k < 6; k++) disp_out[6 * i + k] = db.nodes[i]->displacements[k];
	return 0;
}

} // extern "C"
