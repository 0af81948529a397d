// ANN_Force serialization proxy: a force with every field set goes through
// SerializationProxy -> SerializationNode -> SerializationProxy and comes back identical, including bit-exact doubles,
// the 0-based storage of dihedral / pair lists and the force group; the proxy is registered at library load.
// CPU only (no GPU, no kernels).  Exit code 0 = pass.
#include "OpenMM.h"
#include "OpenMM_ANN.h"
#include "openmm/serialization/ANN_ForceProxy.h"
#include <cmath>
#include <cstdio>
#include <string>
#include <typeinfo>
#include <vector>

using namespace OpenMM;
using namespace std;

#define CHECK(cond) do { if (!(cond)) { printf("FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } } while (0)

int main() {
    ANN_Force force;
    force.setForceGroup(3);
    force.set_num_of_nodes({12, 5, 4});
    force.set_layer_types({"Tanh", "Circular"});
    vector<vector<double> > coeff(2), bias(2);
    double v = 0.1;
    for (int i = 0; i < 60; i++) { coeff[0].push_back(sin(v) / 3.0); v += 0.7; }
    for (int i = 0; i < 20; i++) { coeff[1].push_back(cos(v) * 1e-7); v += 0.3; }
    for (int i = 0; i < 5; i++) bias[0].push_back(0.01 * i - 1.0 / 3.0);
    for (int i = 0; i < 4; i++) bias[1].push_back(1e300 * (i + 1));
    force.set_coeffients_of_connections(coeff);
    force.set_values_of_biased_nodes(bias);
    force.set_potential_center({2.4, -2.3});
    force.set_force_constant(512.25);
    force.set_scaling_factor(0.1 + 0.2);                       // not exactly representable: must survive bit for bit
    force.set_index_of_backbone_atoms({2, 5, 7, 9});
    force.set_data_type_in_input_layer(1);
    force.set_list_of_index_of_atoms_forming_dihedrals({{1, 2, 3, 4}, {2, 3, 4, 5}});
    force.set_list_of_pair_index_for_distances({{1, 3}, {2, 4}, {1, 5}});

    const SerializationProxy& proxy = SerializationProxy::getProxy(typeid(ANN_Force));      // registered at load
    CHECK(proxy.getTypeName() == "ANN_Force");
    CHECK(&SerializationProxy::getProxy("ANN_Force") == &proxy);
    SerializationNode node;
    node.setName("Force");
    proxy.serialize(&force, node);
    CHECK(node.getIntProperty("version") == 1);
    ANN_Force* copy = reinterpret_cast<ANN_Force*>(proxy.deserialize(node));
    CHECK(copy->getForceGroup() == 3);
    CHECK(copy->get_num_of_nodes() == force.get_num_of_nodes());
    CHECK(copy->get_layer_types() == force.get_layer_types());
    CHECK(copy->get_coeffients_of_connections() == force.get_coeffients_of_connections());        // exact doubles
    CHECK(copy->get_values_of_biased_nodes() == force.get_values_of_biased_nodes());
    CHECK(copy->get_potential_center() == force.get_potential_center());
    CHECK(copy->get_force_constant() == force.get_force_constant());
    CHECK(copy->get_scaling_factor() == force.get_scaling_factor());
    CHECK(copy->get_index_of_backbone_atoms() == force.get_index_of_backbone_atoms());
    CHECK(copy->get_data_type_in_input_layer() == 1);
    CHECK(copy->get_list_of_index_of_atoms_forming_dihedrals() == force.get_list_of_index_of_atoms_forming_dihedrals());
    CHECK(copy->get_list_of_index_of_atoms_forming_dihedrals()[1][0] == 1);                       // stored 0-based
    CHECK(copy->get_list_of_pair_index_for_distances() == force.get_list_of_pair_index_for_distances());
    // a second generation is identical to the first (no drift through the 1-based setters)
    SerializationNode node2;
    proxy.serialize(copy, node2);
    ANN_Force* copy2 = reinterpret_cast<ANN_Force*>(proxy.deserialize(node2));
    CHECK(copy2->get_list_of_pair_index_for_distances() == force.get_list_of_pair_index_for_distances());
    CHECK(copy2->get_coeffients_of_connections() == force.get_coeffients_of_connections());
    // unknown versions are refused
    node.setIntProperty("version", 2);
    bool threw = false;
    try { proxy.deserialize(node); } catch (const OpenMMException&) { threw = true; }
    CHECK(threw);
    delete copy;
    delete copy2;
    printf("serialization ok\n");
    return 0;
}
