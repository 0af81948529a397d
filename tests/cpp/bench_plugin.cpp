// Per-step cost of the bias force THROUGH THE C++ PLUGIN LAYER (ANN_ForceImpl::calcForcesAndEnergy ->
// CudaCalcANN_ForceKernel::execute -> annf_eval_openmm), i.e. what OpenMM pays per MD step.  BASELINE config 2:
// alanine dipeptide, 22 atoms, 7 backbone heavy atoms selected, 21-40-4 Tanh/Circular, mixed precision.
// Prints one JSON object.  usage: bench_plugin [steps] [c2|c1|c1pairs]     (c1 = 21-40-2 Tanh/Tanh, the shape the legacy CUDA
// plugin supports; c1pairs = the same network fed with the 21 pair distances of the 7 atoms, the legacy plugin's other input
// mode: compare with tools/legacy_restatement [steps] [pairs])
#include "OpenMM.h"
#include "OpenMM_ANN.h"
#include "CudaANN_KernelFactory.h"
#include "openmm/cuda/CudaPlatform.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

using namespace OpenMM;
using namespace std;

static double lcg(unsigned long long& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return ((s >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0; }

int main(int argc, char** argv) {
    const int steps = argc > 1 ? atoi(argv[1]) : 20000;
    const bool pairs = argc > 2 && string(argv[2]) == "c1pairs";
    const bool c1 = pairs || (argc > 2 && string(argv[2]) == "c1");
    registerANN_CudaKernelFactories();
    System system;
    const int num_atoms = 22;
    for (int i = 0; i < num_atoms; i++) system.addParticle(12.0);
    ANN_Force* f = new ANN_Force();
    const vector<int> nodes = {21, 40, c1 ? 2 : 4};
    f->set_num_of_nodes(nodes);
    f->set_layer_types({"Tanh", c1 ? "Tanh" : "Circular"});
    unsigned long long seed = 1234;
    vector<vector<double> > coeff(2), bias(2);
    for (int l = 0; l < 2; l++) {
        const double bound = sqrt(6.0 / (nodes[l] + nodes[l + 1]));
        for (int i = 0; i < nodes[l] * nodes[l + 1]; i++) coeff[l].push_back(bound * lcg(seed));
        for (int i = 0; i < nodes[l + 1]; i++) bias[l].push_back(0.1 * lcg(seed));
    }
    f->set_coeffients_of_connections(coeff);
    f->set_values_of_biased_nodes(bias);
    f->set_potential_center(c1 ? vector<double>({0.3, 0.3}) : vector<double>({0.7, -1.1}));
    f->set_force_constant(500);
    f->set_scaling_factor(0.5);
    const vector<int> backbone = {2, 5, 7, 9, 15, 17, 19};         // ACE-ALA-NME backbone heavy atoms
    f->set_index_of_backbone_atoms(backbone);
    f->set_data_type_in_input_layer(pairs ? 2 : 1);
    if (pairs) {
        vector<vector<int> > list;
        for (size_t i = 0; i < backbone.size(); i++)
            for (size_t j = i + 1; j < backbone.size(); j++) list.push_back({backbone[i], backbone[j]});   // 7 choose 2 = 21 inputs
        f->set_list_of_pair_index_for_distances(list);
    }
    system.addForce(f);
    VerletIntegrator integrator(0.002);
    map<string, string> props; props["Precision"] = "mixed";
    Context context(system, integrator, Platform::getPlatformByName("CUDA"), props);
    vector<Vec3> positions(num_atoms);
    for (int i = 0; i < num_atoms; i++) positions[i] = Vec3(0.5 * lcg(seed), 0.5 * lcg(seed), 0.5 * lcg(seed));
    context.setPositions(positions);
    ContextImpl& impl = context.getImpl();
    ForceImpl* force_impl = impl.getForceImpls()[0];
    CudaContext& cu = CudaPlatform::cu(impl);
    const double e_check = context.getState(State::Energy).getPotentialEnergy();
    for (int i = 0; i < 200; i++) force_impl->calcForcesAndEnergy(impl, true, true, 0xFFFFFFFF);
    cudaStreamSynchronize(cu.getCurrentStream());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto t0 = chrono::steady_clock::now();
    cudaEventRecord(e0, cu.getCurrentStream());
    for (int i = 0; i < steps; i++) force_impl->calcForcesAndEnergy(impl, true, true, 0xFFFFFFFF);
    cudaEventRecord(e1, cu.getCurrentStream());
    cudaStreamSynchronize(cu.getCurrentStream());
    auto t1 = chrono::steady_clock::now();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double wall_us = chrono::duration<double, micro>(t1 - t0).count() / steps;
    // the same calls replayed from a CUDA graph of 100 steps (what an integrator that graphs its step would see)
    cudaGraph_t graph; cudaGraphExec_t exec;
    cudaStreamBeginCapture(cu.getCurrentStream(), cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 100; i++) force_impl->calcForcesAndEnergy(impl, true, true, 0xFFFFFFFF);
    cudaStreamEndCapture(cu.getCurrentStream(), &graph);
    float ms_graph = -1.f;
    if (cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphLaunch(exec, cu.getCurrentStream());
        cudaStreamSynchronize(cu.getCurrentStream());
        cudaEventRecord(e0, cu.getCurrentStream());
        for (int i = 0; i < steps / 100; i++) cudaGraphLaunch(exec, cu.getCurrentStream());
        cudaEventRecord(e1, cu.getCurrentStream());
        cudaStreamSynchronize(cu.getCurrentStream());
        cudaEventElapsedTime(&ms_graph, e0, e1);
    }
    printf("{\"network\": \"%s\", \"steps\": %d, \"gpu_us_per_step\": %.4f, \"wall_us_per_step\": %.4f, \"graph_us_per_step\": %.4f, "
           "\"energy\": %.12g, \"path\": \"ANN_ForceImpl::calcForcesAndEnergy -> CudaCalcANN_ForceKernel::execute -> annf_eval_openmm\"}\n",
           pairs ? "21-40-2 Tanh/Tanh, 21 pair distances" : (c1 ? "21-40-2 Tanh/Tanh" : "21-40-4 Tanh/Circular"), steps, 1e3 * ms / steps, wall_us,
           ms_graph < 0 ? -1.0 : 1e3 * ms_graph / (steps / 100 * 100), e_check);
    return 0;
}
