// Drives the C++ plugin layer (ann_force_b200/plugin) through a Context of the OpenMM stand-in, the way the
// reference's own test program drives the reference plugin (openmmapi/tests/test_ANN_package.cpp): build a System,
// add an ANN_Force through its setters, create a Context on the "CUDA" platform, compare forces with finite
// differences of the energy.  Cases and constants follow that file; golden energies are those measured on the
// unmodified reference.
#include "OpenMM.h"
#include "OpenMM_ANN.h"
#include "CudaANN_KernelFactory.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/internal/AssertionUtilities.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>

using namespace OpenMM;
using namespace std;

const double FORCE_TOL = 5e-2;   // test_ANN_package.cpp:21

static map<string, string> props(const string& precision, bool reverse = false) {
    map<string, string> p;
    p["Precision"] = precision;
    if (reverse) p["StubReverseAtomOrder"] = "true";
    return p;
}

// analytic forces vs central differences of the energy (the reference uses forward differences with delta = 0.005,
// test_ANN_package.cpp:238-249; both criteria are applied)
static void check_forces_against_energy(Context& context, const vector<Vec3>& positions, double tight_tol) {
    context.setPositions(positions);
    State state = context.getState(State::Forces | State::Energy);
    const vector<Vec3> forces = state.getForces();
    const double e0 = state.getPotentialEnergy();
    double fmax = 1.0;
    for (const Vec3& f : forces) for (int c = 0; c < 3; c++) fmax = max(fmax, fabs(f[c]));
    for (size_t i = 0; i < positions.size(); i++)
        for (int c = 0; c < 3; c++) {
            vector<Vec3> p = positions;
            p[i][c] += 0.005;
            context.setPositions(p);
            const double forward = -(context.getState(State::Energy).getPotentialEnergy() - e0) / 0.005;
            ASSERT_EQUAL_TOL(forces[i][c], forward, FORCE_TOL);
            if (tight_tol <= 0) continue;   // single precision: fp32 positions and an fp32 energy buffer cannot resolve h = 1e-5
            const double h = 1e-5;
            p = positions; p[i][c] += h;
            context.setPositions(p);
            const double ep = context.getState(State::Energy).getPotentialEnergy();
            p[i][c] -= 2 * h;
            context.setPositions(p);
            const double em = context.getState(State::Energy).getPotentialEnergy();
            const double central = -(ep - em) / (2 * h);
            if (fabs(central - forces[i][c]) > tight_tol * fmax) {
                printf("atom %zu comp %d: analytic %.10g central %.10g\n", i, c, forces[i][c], central);
                throwException(__FILE__, __LINE__, "analytic force differs from central difference");
            }
        }
}

static ANN_Force* cartesian_12_4_4(const string& last_activation) {   // test_ANN_package.cpp:432-451
    ANN_Force* f = new ANN_Force();
    f->set_num_of_nodes(vector<int>({12, 4, 4}));
    f->set_layer_types(vector<string>({"Tanh", last_activation}));
    vector<double> w0(48, 0.0);
    for (int j = 0; j < 4; j++) for (int c = 0; c < 3; c++) w0[12 * j + 3 * j + c] = 1.0;
    f->set_coeffients_of_connections({w0, {1, 0, 0.4, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}});
    f->set_force_constant(10);
    f->set_scaling_factor(1);
    f->set_index_of_backbone_atoms({1, 2, 3, 4});
    f->set_potential_center(last_activation == "Circular" ? vector<double>({0, 0}) : vector<double>({0, 0, 0, 0}));
    f->set_values_of_biased_nodes({{0.1, 0.2, 0.3, 0.4}, {0.5, 0.6, 0.4, 0.3}});
    f->set_data_type_in_input_layer(1);
    return f;
}

static void test_cartesian(const string& last_activation, double golden_energy, const string& precision, bool reverse) {
    cout << "cartesian 12-4-4 Tanh/" << last_activation << " (" << precision << (reverse ? ", reversed atom order" : "") << ")\n";
    System system;
    for (int i = 0; i < 4; i++) system.addParticle(1.0);
    VerletIntegrator integrator(0.01);
    system.addForce(cartesian_12_4_4(last_activation));
    Context context(system, integrator, Platform::getPlatformByName("CUDA"), props(precision, reverse));
    vector<Vec3> positions = {Vec3(-1, -2, -3), Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(0, 0, 2)};
    context.setPositions(positions);
    State state = context.getState(State::Forces | State::Energy);
    ASSERT_EQUAL_TOL(golden_energy, state.getPotentialEnergy(), precision == "single" ? 1e-6 : 1e-12);
    if (last_activation == "Tanh") {   // force components measured on the unmodified reference
        const double fx[4] = {0.266341673478, -0.681947964612, 0.165733598395, 0.249872692739};
        for (int i = 0; i < 4; i++) for (int c = 0; c < 3; c++) ASSERT_EQUAL_TOL(fx[i], state.getForces()[i][c], 1e-9);
    }
    check_forces_against_energy(context, positions, precision == "single" ? 0.0 : 1e-6);
}

static void test_pair_distances() {                                  // test_ANN_package.cpp:596-674 (run there on "CUDA" too)
    cout << "pair distances 6-3-3 Tanh/Tanh\n";
    System system;
    for (int i = 0; i < 4; i++) system.addParticle(1.0);
    VerletIntegrator integrator(0.01);
    ANN_Force* f = new ANN_Force();
    f->set_index_of_backbone_atoms({1, 2, 3, 4});
    f->set_num_of_nodes(vector<int>({6, 3, 3}));
    f->set_layer_types(vector<string>({"Tanh", "Tanh"}));
    f->set_list_of_pair_index_for_distances(vector<vector<int> >({{1, 2}, {1, 3}, {1, 4}, {2, 3}, {2, 4}, {3, 4}}));
    f->set_coeffients_of_connections({{1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0.7, 0, 0, 0, 0, 1, 1}, {1, 0, 0.4, 0, 1, 0, 0, 0, 1}});
    f->set_force_constant(10);
    f->set_scaling_factor(10);
    f->set_potential_center(vector<double>({0, 0, 0}));
    f->set_values_of_biased_nodes({{0.1, 0.2, 0.3}, {0.5, 0.6, 0.4}});
    f->set_data_type_in_input_layer(2);
    system.addForce(f);
    Context context(system, integrator, Platform::getPlatformByName("CUDA"), props("double", true));
    vector<Vec3> positions = {Vec3(-1, -2, -3), Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(0, 0, 2)};
    context.setPositions(positions);
    ASSERT_EQUAL_TOL(10.832122286301122, context.getState(State::Energy).getPotentialEnergy(), 1e-12);
    check_forces_against_energy(context, positions, 1e-6);
}

static void test_dihedrals(const string& last_activation, vector<double> center, double golden_energy) {   // :182-341
    cout << "dihedrals 4-4-4 Tanh/" << last_activation << "\n";
    System system;
    for (int i = 0; i < 6; i++) system.addParticle(1.0);
    VerletIntegrator integrator(0.01);
    ANN_Force* f = new ANN_Force();
    f->set_num_of_nodes(vector<int>({4, 4, 4}));
    f->set_layer_types(vector<string>({"Tanh", last_activation}));
    f->set_coeffients_of_connections({{1, 0, 0, 0, 0, 1, 0, 0.7, 0, 0, 1, 0, 0, 0, 0, 1}, {1, 0, 0.4, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}});
    f->set_force_constant(10);
    f->set_potential_center(center);
    f->set_values_of_biased_nodes({{0.1, 0.2, 0.3, 0.4}, {0.5, 0.6, 0.4, 0.3}});
    f->set_list_of_index_of_atoms_forming_dihedrals_from_index_of_backbone_atoms(vector<int>({1, 2, 3, 4, 5, 6}));
    system.addForce(f);                                               // data_type_in_input_layer defaults to 0 (ANN_Force.h:137)
    Context context(system, integrator, Platform::getPlatformByName("CUDA"), props("double"));
    vector<Vec3> positions = {Vec3(-1, -2, -3), Vec3(0, 0, 0), Vec3(1, 0, 0), Vec3(0, 0, 1), Vec3(0.5, 0, 0), Vec3(0, 0.3, 0.6)};
    context.setPositions(positions);
    ASSERT_EQUAL_TOL(golden_energy, context.getState(State::Energy).getPotentialEnergy(), 1e-12);
    // tolerance of the central-difference check: forces reach 110 here and curvature is large (the reference's own
    // forward-difference criterion is applied inside as well, except for the {2.4, 2.3} window where its truncation
    // error alone exceeds FORCE_TOL: see tests/test_oracle.py)
    context.setPositions(positions);
    State state = context.getState(State::Forces | State::Energy);
    const double h = 1e-6;
    double fmax = 1.0;
    for (const Vec3& v : state.getForces()) for (int c = 0; c < 3; c++) fmax = max(fmax, fabs(v[c]));
    for (size_t i = 0; i < positions.size(); i++)
        for (int c = 0; c < 3; c++) {
            vector<Vec3> p = positions; p[i][c] += h;
            context.setPositions(p);
            const double ep = context.getState(State::Energy).getPotentialEnergy();
            p[i][c] -= 2 * h;
            context.setPositions(p);
            const double em = context.getState(State::Energy).getPotentialEnergy();
            ASSERT(fabs(-(ep - em) / (2 * h) - state.getForces()[i][c]) <= 1e-6 * fmax);
        }
}

static vector<double> generate_random_array(int size, int seed) {   // test_ANN_package.cpp:504-511
    srand((unsigned) seed);
    vector<double> a;
    for (int i = 0; i < size; i++) a.push_back((rand() % 10) / 100.0);
    return a;
}

static void test_larger_system(int num_of_atoms, int seed) {        // test_ANN_package.cpp:513-594
    cout << "cartesian 3N-150-4 Tanh/Tanh, N = " << num_of_atoms << "\n";
    System system;
    vector<int> index(num_of_atoms);
    for (int i = 0; i < num_of_atoms; i++) { system.addParticle(1.0); index[i] = i + 1; }
    VerletIntegrator integrator(0.01);
    ANN_Force* f = new ANN_Force();
    const int n0 = 3 * num_of_atoms;
    f->set_num_of_nodes(vector<int>({n0, 150, 4}));
    f->set_layer_types(vector<string>({"Tanh", "Tanh"}));
    f->set_coeffients_of_connections({generate_random_array(n0 * 150, seed), generate_random_array(150 * 4, seed)});
    f->set_force_constant(10);
    f->set_scaling_factor(1);
    f->set_index_of_backbone_atoms(index);
    f->set_potential_center(vector<double>({0, 0, 0, 0}));
    f->set_values_of_biased_nodes({generate_random_array(150, seed), generate_random_array(4, seed)});
    f->set_data_type_in_input_layer(1);
    system.addForce(f);
    Context context(system, integrator, Platform::getPlatformByName("CUDA"), props("mixed"));
    vector<Vec3> positions(num_of_atoms);
    for (int i = 0; i < num_of_atoms; i++) positions[i] = Vec3((rand() % 100) / 20.0, (rand() % 100) / 20.0, (rand() % 100) / 20.0);
    check_forces_against_energy(context, positions, 1e-6);
}

static void test_update_parameters_and_reorder() {
    cout << "updateParametersInContext + atom reordering\n";
    System system;
    for (int i = 0; i < 9; i++) system.addParticle(1.0);
    VerletIntegrator integrator(0.01);
    ANN_Force* f = cartesian_12_4_4("Circular");
    f->set_index_of_backbone_atoms({8, 2, 5, 3});                      // scattered selection inside a larger system
    system.addForce(f);
    Context context(system, integrator, Platform::getPlatformByName("CUDA"), props("double"));
    vector<Vec3> positions(9);
    for (int i = 0; i < 9; i++) positions[i] = Vec3(0.3 * i - 1, 0.1 * i * i - 2, 0.7 * sin(1.0 * i));
    context.setPositions(positions);
    State before = context.getState(State::Forces | State::Energy);
    // reorder the device slots: results must not change, unselected atoms must stay force-free
    CudaContext& cu = CudaPlatform::cu(context.getImpl());
    cu.reorderAtoms({4, 7, 1, 0, 8, 2, 6, 5, 3});
    State after = context.getState(State::Forces | State::Energy);
    ASSERT_EQUAL_TOL(before.getPotentialEnergy(), after.getPotentialEnergy(), 1e-13);
    for (int i = 0; i < 9; i++) ASSERT_EQUAL_VEC(before.getForces()[i], after.getForces()[i], 1e-9);
    for (int i : {0, 3, 5, 6, 8}) ASSERT_EQUAL_VEC(Vec3(0, 0, 0), after.getForces()[i], 0.0);
    // move the umbrella window without rebuilding the Context (a no-op stub in the reference)
    f->set_potential_center({2.4, 2.3});
    f->set_force_constant(25);
    f->updateParametersInContext(context);
    const double moved = context.getState(State::Energy).getPotentialEnergy();
    ASSERT(fabs(moved - before.getPotentialEnergy()) > 1.0);
    check_forces_against_energy(context, positions, 1e-6);
    // same parameters from scratch give the same energy
    System system2;
    for (int i = 0; i < 9; i++) system2.addParticle(1.0);
    ANN_Force* g = cartesian_12_4_4("Circular");
    g->set_index_of_backbone_atoms({8, 2, 5, 3});
    g->set_potential_center({2.4, 2.3});
    g->set_force_constant(25);
    system2.addForce(g);
    VerletIntegrator integrator2(0.01);
    Context context2(system2, integrator2, Platform::getPlatformByName("CUDA"), props("double"));
    context2.setPositions(positions);
    ASSERT_EQUAL_TOL(context2.getState(State::Energy).getPotentialEnergy(), moved, 1e-13);
}

static void test_errors() {
    cout << "error behaviour\n";
    System system;
    for (int i = 0; i < 4; i++) system.addParticle(1.0);
    VerletIntegrator integrator(0.01);
    ANN_Force* f = cartesian_12_4_4("Tanh");
    f->set_potential_center({0, 0});                                   // wrong size for a Tanh output layer
    system.addForce(f);
    bool threw = false;
    try { Context context(system, integrator, Platform::getPlatformByName("CUDA")); } catch (const OpenMMException& e) { threw = true; }
    ASSERT(threw);
    {
        // a selected atom listed twice is rejected (the reference would sum its two contributions, ANN_ReferenceKernels.cpp:401)
        System s3;
        for (int i = 0; i < 4; i++) s3.addParticle(1.0);
        VerletIntegrator i3(0.01);
        ANN_Force* g = cartesian_12_4_4("Tanh");
        g->set_index_of_backbone_atoms({1, 2, 2, 4});
        s3.addForce(g);
        threw = false;
        try { Context c3(s3, i3, Platform::getPlatformByName("CUDA")); } catch (const OpenMMException& e) { threw = true; }
        ASSERT(threw);
    }
    CudaANN_KernelFactory factory;
    threw = false;
    try {
        System s2; s2.addParticle(1.0);
        VerletIntegrator i2(0.01);
        Context c2(s2, i2, Platform::getPlatformByName("CUDA"));
        factory.createKernelImpl("NoSuchKernel", Platform::getPlatformByName("CUDA"), c2.getImpl());
    } catch (const OpenMMException& e) { threw = true; }
    ASSERT(threw);
}

int main() {
    try {
        registerANN_CudaKernelFactories();                              // registers the CUDA platform if missing + our factory
        test_cartesian("Tanh", 11.421839171580899, "double", false);
        test_cartesian("Tanh", 11.421839171580899, "mixed", true);
        test_cartesian("Tanh", 11.421839171580899, "single", false);
        test_cartesian("Softmax", 1.4608776635136895, "double", false);
        test_cartesian("Linear", 26.790165804903218, "mixed", false);
        test_cartesian("Circular", 16.572144252757681, "double", true);
        for (int n = 20; n < 200; n += 20) test_larger_system(n, 1);     // the reference's sweep, 20..180 step 20 (:689-694)
        test_pair_distances();
        test_dihedrals("Tanh", {0, 0, 0, 0}, 0.5707411857967678);
        test_dihedrals("Circular", {0, 0}, 34.049960896473365);
        test_dihedrals("Circular", {2.4, 2.3}, 50.906053293232567);
        test_update_parameters_and_reorder();
        test_errors();
    } catch (const exception& e) {
        cout << "exception: " << e.what() << endl;
        return 1;
    }
    cout << "Done" << endl;
    return 0;
}
