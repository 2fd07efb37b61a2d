// gmsh.h stand-in: the subset of the Gmsh SDK 4.1.4 C++ API that the reference calls,
// implemented on gmshlite (dgfem-acoustic_b200/host/gmshlite.h). TEST INFRASTRUCTURE: it exists only so that
// the reference's own translation units (/root/reference/src/*.cpp) can be compiled, unmodified, into
// oracle/_ref/dgalerkin_ref. Extras controlled by environment variables:
//   GMSHLITE_ORDER=p     gmsh::open elevates the (order-1) mesh to order p, like `gmsh -order p`
//   GMSHLITE_QUIET=1     silences gmsh::logger::write
// gmsh::view::write(tag, file, append) writes `<file>.<ViewName>.bin` (see gmsh_shim.cpp for the layout).
#pragma once
#include <cmath>
#include <string>
#include <utility>
#include <vector>

namespace gmsh {

typedef std::vector<std::pair<int, int> > vectorpair;

void initialize();
void finalize();
void open(const std::string& fileName);

namespace option {
void setNumber(const std::string& name, const double value);
}

namespace logger {
void write(const std::string& message, const std::string& level = "info");
}

namespace model {
int getDimension();
void getPhysicalGroups(vectorpair& dimTags, const int dim = -1);
void getPhysicalName(const int dim, const int tag, std::string& name);
int addDiscreteEntity(const int dim, const int tag = -1, const std::vector<int>& boundary = std::vector<int>());
void list(std::vector<std::string>& names);

namespace mesh {
void getElementTypes(std::vector<int>& elementTypes, const int dim = -1, const int tag = -1);
void getElementProperties(const int elementType, std::string& elementName, int& dim, int& order, int& numNodes,
                          std::vector<double>& parametricCoord);
void getElementsByType(const int elementType, std::vector<int>& elementTags, std::vector<int>& nodeTags, const int tag = -1);
void getJacobians(const int elementType, const std::string& integrationType, std::vector<double>& jacobians,
                  std::vector<double>& determinants, std::vector<double>& points, const int tag = -1);
void getBasisFunctions(const int elementType, const std::string& integrationType, const std::string& functionSpaceType,
                       std::vector<double>& integrationPoints, int& numComponents, std::vector<double>& basisFunctions);
int getElementType(const std::string& familyName, const int order, const bool serendip = false);
void getElementEdgeNodes(const int elementType, std::vector<int>& nodes, const int tag = -1, const bool primary = false);
void getElementFaceNodes(const int elementType, const int faceType, std::vector<int>& nodes, const int tag = -1,
                         const bool primary = false);
void setElementsByType(const int dim, const int tag, const int elementType, const std::vector<int>& elementTags,
                       const std::vector<int>& nodeTags);
void getBarycenters(const int elementType, const int tag, const bool fast, const bool primary, std::vector<double>& barycenters);
void getNode(const int nodeTag, std::vector<double>& coord, std::vector<double>& parametricCoord);
void getNodesForPhysicalGroup(const int dim, const int tag, std::vector<int>& nodeTags, std::vector<double>& coord);
}  // namespace mesh
}  // namespace model

namespace view {
int add(const std::string& name, const int tag = -1);
void addModelData(const int tag, const int step, const std::string& modelName, const std::string& dataType,
                  const std::vector<int>& tags, const std::vector<std::vector<double> >& data, const double time = 0.,
                  const int numComponents = -1, const int partition = 0);
void addListData(const int tag, const std::string& dataType, const int numEle, const std::vector<double>& data);
void write(const int tag, const std::string& fileName, const bool append = false);
}  // namespace view

}  // namespace gmsh
