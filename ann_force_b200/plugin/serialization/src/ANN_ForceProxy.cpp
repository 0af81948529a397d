// ANN_Force <-> SerializationNode.  Every field the force carries (openmmapi/include/openmm/ANN_Force.h) is written with
// the meaning its getter reports: backbone atoms 1-based as given, dihedral / pair lists 0-based as stored; deserialize()
// therefore feeds the latter back through the 1-based setters.
#include "openmm/serialization/ANN_ForceProxy.h"
#include "openmm/ANN_Force.h"
#include <string>
#include <vector>

using namespace OpenMM;
using std::string;
using std::vector;

namespace {

void writeInts(SerializationNode& parent, const string& name, const vector<int>& values) {
    SerializationNode& node = parent.createChildNode(name);
    for (int v : values) node.createChildNode("i").setIntProperty("v", v);
}
vector<int> readInts(const SerializationNode& parent, const string& name) {
    vector<int> values;
    for (const SerializationNode& c : parent.getChildNode(name).getChildren()) values.push_back(c.getIntProperty("v"));
    return values;
}
void writeDoubles(SerializationNode& node, const vector<double>& values) {
    for (double v : values) node.createChildNode("d").setDoubleProperty("v", v);
}
vector<double> readDoubles(const SerializationNode& node) {
    vector<double> values;
    for (const SerializationNode& c : node.getChildren()) values.push_back(c.getDoubleProperty("v"));
    return values;
}

}  // namespace

ANN_ForceProxy::ANN_ForceProxy() : SerializationProxy("ANN_Force") {}

void ANN_ForceProxy::serialize(const void* object, SerializationNode& node) const {
    const ANN_Force& force = *reinterpret_cast<const ANN_Force*>(object);
    node.setIntProperty("version", 1);
    node.setIntProperty("forceGroup", force.getForceGroup());
    node.setDoubleProperty("forceConstant", force.get_force_constant());
    node.setDoubleProperty("scalingFactor", force.get_scaling_factor());
    node.setIntProperty("dataTypeInInputLayer", force.get_data_type_in_input_layer());
    writeInts(node, "NumOfNodes", force.get_num_of_nodes());
    writeInts(node, "IndexOfBackboneAtoms", force.get_index_of_backbone_atoms());
    SerializationNode& types = node.createChildNode("LayerTypes");
    for (const string& t : force.get_layer_types()) types.createChildNode("Layer").setStringProperty("type", t);
    SerializationNode& coeff = node.createChildNode("CoefficientsOfConnections");
    for (const vector<double>& w : force.get_coeffients_of_connections()) writeDoubles(coeff.createChildNode("Connection"), w);
    SerializationNode& bias = node.createChildNode("ValuesOfBiasedNodes");
    for (const vector<double>& b : force.get_values_of_biased_nodes()) writeDoubles(bias.createChildNode("Connection"), b);
    writeDoubles(node.createChildNode("PotentialCenter"), force.get_potential_center());
    SerializationNode& dihedrals = node.createChildNode("Dihedrals");
    for (const vector<int>& q : force.get_list_of_index_of_atoms_forming_dihedrals()) writeInts(dihedrals, "Dihedral", q);
    SerializationNode& pairs = node.createChildNode("PairsForDistances");
    for (const vector<int>& q : force.get_list_of_pair_index_for_distances()) writeInts(pairs, "Pair", q);
}

void* ANN_ForceProxy::deserialize(const SerializationNode& node) const {
    if (node.getIntProperty("version") != 1) throw OpenMMException("Unsupported version number");
    ANN_Force* force = new ANN_Force();
    try {
        force->setForceGroup(node.getIntProperty("forceGroup", 0));
        force->set_force_constant(node.getDoubleProperty("forceConstant"));
        force->set_scaling_factor(node.getDoubleProperty("scalingFactor"));
        force->set_data_type_in_input_layer(node.getIntProperty("dataTypeInInputLayer"));
        force->set_num_of_nodes(readInts(node, "NumOfNodes"));
        force->set_index_of_backbone_atoms(readInts(node, "IndexOfBackboneAtoms"));
        vector<string> types;
        for (const SerializationNode& c : node.getChildNode("LayerTypes").getChildren()) types.push_back(c.getStringProperty("type"));
        force->set_layer_types(types);
        vector<vector<double> > coeff, bias;
        for (const SerializationNode& c : node.getChildNode("CoefficientsOfConnections").getChildren()) coeff.push_back(readDoubles(c));
        for (const SerializationNode& c : node.getChildNode("ValuesOfBiasedNodes").getChildren()) bias.push_back(readDoubles(c));
        force->set_coeffients_of_connections(coeff);
        force->set_values_of_biased_nodes(bias);
        force->set_potential_center(readDoubles(node.getChildNode("PotentialCenter")));
        // stored 0-based; the setters take 1-based indices and subtract one (ANN_Force.cpp:111,164 of the reference)
        vector<vector<int> > dihedrals, pairs;
        for (const SerializationNode& c : node.getChildNode("Dihedrals").getChildren()) {
            vector<int> q;
            for (const SerializationNode& i : c.getChildren()) q.push_back(i.getIntProperty("v") + 1);
            dihedrals.push_back(q);
        }
        for (const SerializationNode& c : node.getChildNode("PairsForDistances").getChildren()) {
            vector<int> q;
            for (const SerializationNode& i : c.getChildren()) q.push_back(i.getIntProperty("v") + 1);
            pairs.push_back(q);
        }
        if (!dihedrals.empty()) force->set_list_of_index_of_atoms_forming_dihedrals(dihedrals);
        if (!pairs.empty()) force->set_list_of_pair_index_for_distances(pairs);
    } catch (...) {
        delete force;
        throw;
    }
    return force;
}
