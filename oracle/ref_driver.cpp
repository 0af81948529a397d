// TEST INFRASTRUCTURE — not part of the product.
//
// C-callable driver around the reference's own, unmodified CPU kernel
//   /root/reference/platforms/reference/ANN_ReferenceKernels.cpp   (ReferenceCalcANN_ForceKernel)
//   /root/reference/openmmapi/src/ANN_Force.cpp                    (the parameter container)
// which oracle/Makefile compiles from where they lie into oracle/_ref/libannref_L<depth>.so.
// Nothing of the reference is copied here: this file only *calls* its public methods
//   initialize                                   ANN_ReferenceKernels.h:31
//   calculateForceAndEnergy                      ANN_ReferenceKernels.h:56
//   calculate_output_of_each_layer / back_prop   ANN_ReferenceKernels.h:67,77
// The only liberty taken is `#define private public` so that the umbrella centre of an already
// initialised kernel can be changed per replica without re-running initialize() (which leaks the
// weight matrices, ANN_ReferenceKernels.cpp:46-48,76-79).
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the result.
#define private public
#include "ANN_ReferenceKernels.h"
#undef private
#include "openmm/reference/ReferencePlatform.h"

#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace OpenMM;
using std::vector;

extern "C" {

struct annref_net {
    int num_layers;                 // must equal the NUM_OF_LAYERS this library was compiled with
    const int* num_of_nodes;        // [num_layers]
    const char* const* layer_types; // [num_layers-1]  "Linear" | "Tanh" | "Circular" | "Softmax"
    const double* const* coeff;     // [num_layers-1]  row-major [n_out][n_in]
    const double* const* bias;      // [num_layers-1]
    const double* potential_center;
    int num_center;
    double force_constant;
    double scaling_factor;
    int data_type_in_input_layer;   // 0 dihedral cos/sin, 1 Cartesian, 2 pair distances
    const int* index_of_backbone_atoms;   // 1-based, [num_backbone]
    int num_backbone;
    const int* dihedrals;           // 1-based quadruples, [4*num_dihedrals]
    int num_dihedrals;
    const int* pairs;               // 1-based pairs, [2*num_pairs]
    int num_pairs;
};

int annref_num_layers() { return NUM_OF_LAYERS; }

} // extern "C"

namespace {

ReferencePlatform& thePlatform() {
    static ReferencePlatform platform;
    return platform;
}

struct Instance {
    System system;
    ANN_Force force;
    ReferenceCalcANN_ForceKernel kernel;
    Instance() : kernel("CalcANN_Force", thePlatform(), system) {}
};

// Feed the reference's own setters (ANN_Force.cpp:13-169) and initialise its kernel.
bool configure(Instance& inst, const annref_net* net) {
    if (net->num_layers != NUM_OF_LAYERS) return false;
    const int L = net->num_layers;
    inst.force.set_num_of_nodes(vector<int>(net->num_of_nodes, net->num_of_nodes + L));
    vector<std::string> types;
    vector<vector<double> > coeff, bias;
    for (int l = 0; l + 1 < L; l++) {
        const int n_in = net->num_of_nodes[l], n_out = net->num_of_nodes[l + 1];
        types.push_back(net->layer_types[l]);
        coeff.push_back(vector<double>(net->coeff[l], net->coeff[l] + (size_t) n_in * n_out));
        bias.push_back(vector<double>(net->bias[l], net->bias[l] + n_out));
    }
    inst.force.set_layer_types(types);
    inst.force.set_coeffients_of_connections(coeff);
    inst.force.set_values_of_biased_nodes(bias);
    inst.force.set_potential_center(vector<double>(net->potential_center, net->potential_center + net->num_center));
    inst.force.set_force_constant(net->force_constant);
    inst.force.set_scaling_factor(net->scaling_factor);
    inst.force.set_data_type_in_input_layer(net->data_type_in_input_layer);
    inst.force.set_index_of_backbone_atoms(vector<int>(net->index_of_backbone_atoms,
                                                       net->index_of_backbone_atoms + net->num_backbone));
    if (net->num_dihedrals > 0) {
        vector<vector<int> > quads;
        for (int d = 0; d < net->num_dihedrals; d++)
            quads.push_back(vector<int>(net->dihedrals + 4 * d, net->dihedrals + 4 * d + 4));
        inst.force.set_list_of_index_of_atoms_forming_dihedrals(quads);
    }
    if (net->num_pairs > 0) {
        vector<vector<int> > pairs;
        for (int p = 0; p < net->num_pairs; p++)
            pairs.push_back(vector<int>(net->pairs + 2 * p, net->pairs + 2 * p + 2));
        inst.force.set_list_of_pair_index_for_distances(pairs);
    }
    inst.kernel.initialize(inst.system, inst.force);
    return true;
}

} // namespace

extern "C" {

// R independent evaluations ("replicas") of the same network.
//   positions [R][num_atoms][3] fp64 (nm);   centers [R][num_center] or NULL (use the net's centre)
//   forces    [R][num_atoms][3] fp64, OVERWRITTEN with this force's contribution
//   energies  [R]
// `threads` host threads each own a private kernel instance and take a contiguous block of
// replicas; the reference itself is single-threaded.  Returns 0, or -1 on a depth mismatch.
// *seconds receives the wall time of the evaluation loops only (set-up and copies excluded).
int annref_eval_batch(const annref_net* net, int num_atoms, int R, const double* positions, const double* centers,
                      double* forces, double* energies, int threads, double* seconds) {
    if (net->num_layers != NUM_OF_LAYERS) return -1;
    if (threads < 1) threads = 1;
    if (threads > R) threads = R > 0 ? R : 1;
    vector<Instance*> instances(threads);
    for (int t = 0; t < threads; t++) {
        instances[t] = new Instance();
        configure(*instances[t], net);
    }
    // stage inputs/outputs in the containers the reference kernel takes
    vector<vector<RealVec> > pos(R, vector<RealVec>(num_atoms)), frc(R, vector<RealVec>(num_atoms));
    for (int r = 0; r < R; r++)
        for (int a = 0; a < num_atoms; a++) {
            const double* p = positions + ((size_t) r * num_atoms + a) * 3;
            pos[r][a] = RealVec(p[0], p[1], p[2]);
        }
    auto work = [&](int t) {
        const int begin = (int) ((long long) R * t / threads), end = (int) ((long long) R * (t + 1) / threads);
        ReferenceCalcANN_ForceKernel& kernel = instances[t]->kernel;
        for (int r = begin; r < end; r++) {
            if (centers != nullptr)
                kernel.potential_center.assign(centers + (size_t) r * net->num_center,
                                               centers + (size_t) (r + 1) * net->num_center);
            energies[r] = kernel.calculateForceAndEnergy(pos[r], frc[r]);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        work(0);
    } else {
        vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds != nullptr) *seconds = std::chrono::duration<double>(t1 - t0).count();
    for (int r = 0; r < R; r++)
        for (int a = 0; a < num_atoms; a++)
            for (int c = 0; c < 3; c++)
                forces[((size_t) r * num_atoms + a) * 3 + c] = frc[r][a][c];
    for (Instance* inst : instances) delete inst;
    return 0;
}

// The two unit-level calls the reference's own known-answer tests make
// (test_ANN_package.cpp:98-173): forward over an explicit input vector, then back_prop.
//   layer_outputs / layer_derivs : concatenation over layers 0..L-1, sum(num_of_nodes) doubles each
int annref_forward_backward(const annref_net* net, const double* input, double* layer_outputs, double* layer_derivs) {
    Instance inst;
    if (!configure(inst, net)) return -1;
    vector<RealOpenMM> in(input, input + net->num_of_nodes[0]);
    inst.kernel.calculate_output_of_each_layer(in);
    vector<vector<double> > derivs;
    inst.kernel.back_prop(derivs);
    const vector<vector<double> >& out = inst.kernel.get_output_of_each_layer();
    size_t o = 0;
    for (int l = 0; l < net->num_layers; l++)
        for (int j = 0; j < net->num_of_nodes[l]; j++, o++) {
            layer_outputs[o] = out[l][j];
            layer_derivs[o] = derivs[l][j];
        }
    return 0;
}

// cos / sin of one dihedral (test_ANN_package.cpp:80-95; ANN_ReferenceKernels.cpp:599-636), 0-based atoms.
void annref_cos_sin_four_atoms(const double* xyz12, double* cos_value, double* sin_value) {
    System system;
    ReferenceCalcANN_ForceKernel kernel("", thePlatform(), system);
    vector<RealVec> pos(4);
    for (int a = 0; a < 4; a++) pos[a] = RealVec(xyz12[3 * a], xyz12[3 * a + 1], xyz12[3 * a + 2]);
    RealOpenMM c, s;
    kernel.get_cos_and_sin_for_four_atoms(0, 1, 2, 3, pos, c, s);
    *cos_value = c;
    *sin_value = s;
}

} // extern "C"
