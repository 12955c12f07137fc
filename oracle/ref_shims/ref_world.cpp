// TEST INFRASTRUCTURE (oracle) -- the global `db`, a reduced Database and
// abort-stubs for the non-virtual symbols the UNMODIFIED reference sources
// reference but that the assembly path never reaches (SURVEY.md section 8c,
// "link closure").  Nothing here is product code.
//
// * Database(): only the defaults the assembly path reads
//   (reference Database.cpp:185-340 sets the same values).
// * myprintf: console tee reduced to an optional stderr echo.
// * TryComment: restated comment skipper (reference IO.cpp:683-752).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Database.h"
#include "ConvergenceCriteria.h"
#include "PSYCoupling.h"
#include "BodyGeometry.h"
#include "SplineElement.h"
#include "Spline.h"
#include "InitialCondition.h"
#include "ConfigurationSave.h"
#include "ConcomitantSolution.h"
#include "GeneralContactSearch.h"
#include "Monitor.h"
#include "PostFiles.h"
#include "SuperNode.h"
#include "MathCode.h"

Database db;

static int g_echo = 0;
extern "C" void ref_set_echo(int on) { g_echo = on; }

Database::Database()
{
	number_GLs_node = 6;
	snprintf(version, sizeof(version), "oracle");

	number_solutions = number_nodes = number_super_nodes = number_points = number_arcs = 0;
	number_elements = number_particles = number_IC = number_materials = number_sections = 0;
	number_pipe_sections = number_shell_sections = number_CS = number_RB_data = 0;
	number_analytical_surfaces = number_surfaces = number_splines = number_line_regions = 0;
	number_surface_regions = number_contacts = number_node_sets = number_super_node_sets = 0;
	number_surface_sets = number_element_sets = number_loads = number_displacements = 0;
	number_constraints = number_special_constraints = number_section_details = 0;
	number_aerodynamicdata = number_cad_data = number_contactinterfaces = number_boundaries = 0;
	number_body_geometries = number_geometries = 0;

	solution = NULL; nodes = NULL; super_nodes = NULL; points = NULL; arcs = NULL;
	elements = NULL; particles = NULL; IC = NULL; materials = NULL; sections = NULL;
	pipe_sections = NULL; shell_sections = NULL; CS = NULL; RB_data = NULL;
	environment = NULL; monitor = NULL; analytical_surfaces = NULL; surfaces = NULL;
	splines = NULL; line_regions = NULL; surface_regions = NULL; contacts = NULL;
	node_sets = NULL; super_node_sets = NULL; surface_sets = NULL; element_sets = NULL;
	loads = NULL; displacements = NULL; constraints = NULL; special_constraints = NULL;
	section_details = NULL; aerodynamic_data = NULL; cad_data = NULL;
	contactinterfaces = NULL; boundaries = NULL; body_geometries = NULL; geometries = NULL;
	bem = NULL; gcs = NULL; config_save = NULL; concomitant_solution = NULL; psy_coupling = NULL;

	conv_criteria = new ConvergenceCriteria();
	post_files = NULL;        // output writers are not linked into the oracle
	solver_options = NULL;
	execution_data = NULL;

	solution_exist = nodes_exist = super_nodes_exist = points_exist = arcs_exist = false;
	elements_exist = particles_exist = IC_exist = materials_exist = sections_exist = false;
	pipe_sections_exist = shell_sections_exist = CS_exist = RB_data_exist = false;
	environment_exist = monitor_exist = false;
	solver_options_exist = true;
	analytical_surfaces_exist = surfaces_exist = line_regions_exist = surface_regions_exist = false;
	contacts_exist = node_sets_exist = super_node_sets_exist = surface_sets_exist = false;
	element_sets_exist = loads_exist = displacements_exist = constraints_exist = false;
	special_constraints_exist = section_details_exist = aerodynamic_data_exist = false;
	cad_data_exist = contactinterfaces_exist = boundaries_exist = false;
	body_geometries_exist = geometries_exist = false;
	bem_exist = gcs_exist = config_save_exist = concomitant_solution_exist = false;
	psy_coupling_exist = false;

	flag_nGL_changed = true;
	n_GL_free = 0;
	n_GL_fixed = 0;
	last_converged_time = 0.0;
	current_time_step = 0.0;
	current_solution_number = 0;
	current_iteration_number = 0;
	plot_times = false;
	size_AA = size_BB = size_AB = 0;
	n_element_results = 15;
	console_output = NULL;
}

// The oracle keeps one model per process image and lets the OS reclaim it.
Database::~Database() {}

int Database::myprintf(const char* format, ...)
{
	if (!g_echo) return 0;
	va_list args;
	va_start(args, format);
	int r = vfprintf(stderr, format, args);
	va_end(args);
	return r;
}

// Skip `// ...` and `/* ... */` runs in front of the next token.
void TryComment(FILE* f)
{
	for (;;)
	{
		long mark = ftell(f);
		char tok[10000];
		if (fscanf(f, "%9999s", tok) != 1) { fseek(f, mark, SEEK_SET); return; }
		if (tok[0] == '/' && tok[1] == '/')
		{
			int c;
			// rewind to just after the token start is not needed: consume the line
			while ((c = fgetc(f)) != EOF && c != '\n') {}
			continue;
		}
		if (tok[0] == '/' && tok[1] == '*')
		{
			if (strstr(tok + 2, "*/")) continue;
			int prev = 0, c;
			while ((c = fgetc(f)) != EOF)
			{
				if (prev == '*' && c == '/') break;
				prev = c;
			}
			continue;
		}
		fseek(f, mark, SEEK_SET);
		return;
	}
}

// ---- never reached from Solution::Mount* in the in-scope configurations ----
static void unreachable(const char* who)
{
	fprintf(stderr, "oracle: %s is outside the assembly path and was reached\n", who);
	abort();
}
#define ORACLE_STUB(sig, name) sig { unreachable(name); }

ORACLE_STUB(void PSYCoupling::SetConstraints(), "PSYCoupling::SetConstraints")
ORACLE_STUB(void PSYCoupling::Couple(), "PSYCoupling::Couple")
ORACLE_STUB(void BodyGeometry::SaveLagrange(), "BodyGeometry::SaveLagrange")
ORACLE_STUB(void SplineElement::SaveConfiguration(), "SplineElement::SaveConfiguration")
ORACLE_STUB(void SplineElement::FillNodes(), "SplineElement::FillNodes")
ORACLE_STUB(void SplineElement::UpdateBox(), "SplineElement::UpdateBox")
ORACLE_STUB(void Spline::SaveConfiguration(), "Spline::SaveConfiguration")
ORACLE_STUB(void InitialCondition::ComputeInitialCondition(), "InitialCondition::ComputeInitialCondition")
ORACLE_STUB(void ConfigurationSave::ExportConfiguration(double), "ConfigurationSave::ExportConfiguration")
ORACLE_STUB(void ConcomitantSolution::UpdateConcomitantSolution(double), "ConcomitantSolution::UpdateConcomitantSolution")
bool GeneralContactSearch::HaveErrors() { unreachable("GeneralContactSearch::HaveErrors"); return false; }
ORACLE_STUB(void GeneralContactSearch::MountContacts(), "GeneralContactSearch::MountContacts")
double GeneralContactSearch::TimeStepControl() { unreachable("GeneralContactSearch::TimeStepControl"); return 0; }
ORACLE_STUB(void GeneralContactSearch::SaveConfiguration(), "GeneralContactSearch::SaveConfiguration")
ORACLE_STUB(void GeneralContactSearch::MountContactsGlobal(), "GeneralContactSearch::MountContactsGlobal")
ORACLE_STUB(void GeneralContactSearch::SolutionStepInitialCheck(), "GeneralContactSearch::SolutionStepInitialCheck")
ORACLE_STUB(void Monitor::UpdateMonitor(double), "Monitor::UpdateMonitor")
ORACLE_STUB(void PostFiles::UpdateSinglePartPostFiles(int, double, int), "PostFiles::UpdateSinglePartPostFiles")
ORACLE_STUB(void PostFiles::WriteConfigurationResults(int, double, int), "PostFiles::WriteConfigurationResults")
ORACLE_STUB(void SuperNode::SaveConfiguration(), "SuperNode::SaveConfiguration")

// exprtk-backed load expressions: numeric tables only in the oracle.
MathCode::MathCode() {}
MathCode::MathCode(int) { unreachable("MathCode"); }
MathCode::~MathCode() {}
double MathCode::GetValueAt(double, int) { unreachable("MathCode::GetValueAt"); return 0; }
bool MathCode::Read(FILE*) { unreachable("MathCode::Read"); return false; }
void MathCode::Write(FILE*) {}
