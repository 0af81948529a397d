#ifndef OPENMM_STUB_SERIALIZATIONNODE_H_
#define OPENMM_STUB_SERIALIZATIONNODE_H_
// Stand-in for OpenMM's SerializationNode (the in-memory tree OpenMM's XmlSerializer reads and writes): named nodes with
// string-valued properties and ordered children.  Same accessor names and semantics as OpenMM 7.x; doubles are stored
// with 17 significant digits so that a round trip is exact.
#include "openmm/OpenMMException.h"
#include "openmm/internal/windowsExport.h"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace OpenMM {

class SerializationNode {
public:
    const std::string& getName() const { return name; }
    void setName(const std::string& n) { name = n; }
    const std::vector<SerializationNode>& getChildren() const { return children; }
    std::vector<SerializationNode>& getChildren() { return children; }
    const SerializationNode& getChildNode(const std::string& n) const {
        for (const SerializationNode& c : children) if (c.name == n) return c;
        throw OpenMMException("Unknown child node '" + n + "' in node '" + name + "'");
    }
    SerializationNode& createChildNode(const std::string& n) {
        children.push_back(SerializationNode());
        children.back().setName(n);
        return children.back();
    }
    bool hasProperty(const std::string& n) const { return properties.count(n) != 0; }
    const std::string& getStringProperty(const std::string& n) const {
        std::map<std::string, std::string>::const_iterator it = properties.find(n);
        if (it == properties.end()) throw OpenMMException("Unknown property '" + n + "' in node '" + name + "'");
        return it->second;
    }
    SerializationNode& setStringProperty(const std::string& n, const std::string& v) { properties[n] = v; return *this; }
    int getIntProperty(const std::string& n) const { return atoi(getStringProperty(n).c_str()); }
    int getIntProperty(const std::string& n, int fallback) const { return hasProperty(n) ? getIntProperty(n) : fallback; }
    SerializationNode& setIntProperty(const std::string& n, int v) { return setStringProperty(n, std::to_string(v)); }
    double getDoubleProperty(const std::string& n) const { return strtod(getStringProperty(n).c_str(), nullptr); }
    SerializationNode& setDoubleProperty(const std::string& n, double v) {
        char buf[40];
        snprintf(buf, sizeof(buf), "%.17g", v);
        return setStringProperty(n, buf);
    }
private:
    std::string name;
    std::map<std::string, std::string> properties;
    std::vector<SerializationNode> children;
};

} // namespace OpenMM
#endif
