// Implementation of the gmsh.h stand-in on gmshlite. TEST INFRASTRUCTURE (see gmsh.h).
#include "gmsh.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>

#include "gmshlite.h"

namespace {
gml::Model g_model;
bool g_quiet = false;

struct ViewStep { int step; double time; int ncomp; std::vector<double> data; };
struct View { std::string name; std::vector<int> tags; std::vector<ViewStep> steps; };
std::map<int, View> g_views;

int degreeOf(const std::string& integrationType) {  // "Gauss<degree>"
    if (integrationType.rfind("Gauss", 0) != 0) throw std::runtime_error("gmsh shim: unknown integration type " + integrationType);
    return std::atoi(integrationType.c_str() + 5);
}
gml::Quadrature ruleFor(int elementType, const std::string& integrationType) {
    int dim, order;
    if (!gml::elementTypeInfo(elementType, dim, order)) throw std::runtime_error("gmsh shim: unknown element type");
    return gml::integrationRule(dim, degreeOf(integrationType), order, g_model.dimension() == 3);
}
}  // namespace

namespace gmsh {

void initialize() { g_quiet = std::getenv("GMSHLITE_QUIET") != nullptr; }
void finalize() {}
void open(const std::string& fileName) {
    g_model = gml::readMsh(fileName);
    if (const char* o = std::getenv("GMSHLITE_ORDER")) gml::elevate(g_model, std::atoi(o));
    if (const char* w = std::getenv("GMSHLITE_WARP")) {  // "amp,k": the curved stand-in geometry (gml::warp), same as dgf_warp_model
        double amp = 0, k = 0, c[3] = {0, 0, 0}, R = 0;
        const int got = std::sscanf(w, "%lf,%lf,%lf,%lf,%lf,%lf", &amp, &k, &c[0], &c[1], &c[2], &R);
        if (got == 6) gml::warpLocal(g_model, amp, k, c, R);  // "amp,k,cx,cy,cz,R": a curved patch in a straight-sided mesh
        else if (got >= 2) gml::warp(g_model, amp, k);
    }
}
namespace option { void setNumber(const std::string&, const double) {} }
namespace logger {
void write(const std::string& message, const std::string&) { if (!g_quiet) std::printf("Info    : %s\n", message.c_str()); }
}

namespace model {
int getDimension() { return g_model.dimension(); }
void getPhysicalGroups(vectorpair& dimTags, const int dim) {
    dimTags.clear();
    for (int d = 0; d <= 3; ++d) {
        if (dim >= 0 && d != dim) continue;
        for (int t : g_model.physicalGroups(d)) dimTags.push_back(std::make_pair(d, t));
    }
}
void getPhysicalName(const int dim, const int tag, std::string& name) { name = g_model.physicalName(dim, tag); }
int addDiscreteEntity(const int dim, const int, const std::vector<int>&) { return g_model.addDiscreteEntity(dim); }
void list(std::vector<std::string>& names) { names.assign(1, g_model.name); }

namespace mesh {
void getElementTypes(std::vector<int>& elementTypes, const int dim, const int) { elementTypes = g_model.elementTypes(dim); }

void getElementProperties(const int elementType, std::string& elementName, int& dim, int& order, int& numNodes,
                          std::vector<double>& parametricCoord) {
    if (!gml::elementTypeInfo(elementType, dim, order)) throw std::runtime_error("gmsh shim: unknown element type");
    const gml::RefElement& re = gml::refElement(dim, order);
    elementName = re.name;
    numNodes = re.np;
    parametricCoord.clear();
    for (int n = 0; n < re.np; ++n)
        for (int c = 0; c < dim; ++c) parametricCoord.push_back(re.uvw[3 * n + c]);
}

void getElementsByType(const int elementType, std::vector<int>& elementTags, std::vector<int>& nodeTags, const int tag) {
    g_model.elementsByType(elementType, elementTags, nodeTags, tag);
}

void getJacobians(const int elementType, const std::string& integrationType, std::vector<double>& jacobians,
                  std::vector<double>& determinants, std::vector<double>& points, const int tag) {
    int dim, order;
    gml::elementTypeInfo(elementType, dim, order);
    const gml::RefElement& re = gml::refElement(dim, order);
    const gml::RefElement& lin = gml::refElement(dim, 1);
    const gml::Quadrature q = ruleFor(elementType, integrationType);
    std::vector<int> tags, nodes;
    g_model.elementsByType(elementType, tags, nodes, tag);
    const size_t ne = tags.size();
    jacobians.resize(ne * q.n * 9);
    determinants.resize(ne * q.n);
    points.resize(ne * q.n * 3);
    std::vector<double> phi(g_model.curved ? re.np : lin.np);
    for (size_t e = 0; e < ne; ++e) {
        double jac[9], det;
        if (!g_model.curved) gml::affineJacobian(g_model, dim, &nodes[e * re.np], jac, det);
        for (int g = 0; g < q.n; ++g) {
            if (g_model.curved) gml::isoJacobian(g_model, dim, order, &nodes[e * re.np], &q.pts[4 * g], jac, det);  // per point
            std::copy(jac, jac + 9, &jacobians[(e * q.n + g) * 9]);
            determinants[e * q.n + g] = det;
            const gml::RefElement& map = g_model.curved ? re : lin;
            map.basis(&q.pts[4 * g], phi.data());
            for (int x = 0; x < 3; ++x) {
                double s = 0;
                for (int v = 0; v < map.np; ++v) s += phi[v] * g_model.node(nodes[e * re.np + v])[x];
                points[(e * q.n + g) * 3 + x] = s;
            }
        }
    }
}

void getBasisFunctions(const int elementType, const std::string& integrationType, const std::string& functionSpaceType,
                       std::vector<double>& integrationPoints, int& numComponents, std::vector<double>& basisFunctions) {
    int dim, order;
    gml::elementTypeInfo(elementType, dim, order);
    const gml::RefElement& re = gml::refElement(dim, order);
    const gml::Quadrature q = ruleFor(elementType, integrationType);
    integrationPoints = q.pts;
    if (functionSpaceType == "Lagrange") {
        numComponents = 1;
        basisFunctions.resize((size_t)q.n * re.np);
        for (int g = 0; g < q.n; ++g) re.basis(&q.pts[4 * g], &basisFunctions[(size_t)g * re.np]);
    } else if (functionSpaceType == "GradLagrange") {
        numComponents = 3;
        basisFunctions.resize((size_t)q.n * re.np * 3);
        for (int g = 0; g < q.n; ++g) re.gradBasis(&q.pts[4 * g], &basisFunctions[(size_t)g * re.np * 3]);
    } else {
        throw std::runtime_error("gmsh shim: unsupported function space " + functionSpaceType);
    }
}

int getElementType(const std::string& familyName, const int order, const bool) {
    const int dim = familyName == "point" ? 0 : familyName == "line" ? 1 : familyName == "triangle" ? 2 : familyName == "tetrahedron" ? 3 : -1;
    if (dim < 0) throw std::runtime_error("gmsh shim: unknown family " + familyName);
    return gml::elementType(dim, order);
}

static void faceNodesOf(const int elementType, std::vector<int>& out) {
    int dim, order;
    gml::elementTypeInfo(elementType, dim, order);
    const gml::RefElement& re = gml::refElement(dim, order);
    std::vector<int> tags, nodes;
    g_model.elementsByType(elementType, tags, nodes, -1);
    out.clear();
    for (size_t e = 0; e < tags.size(); ++e)
        for (int k = 0; k < re.nFaces * re.nfp; ++k) out.push_back(nodes[e * re.np + re.faceNodes[k]]);
}
void getElementEdgeNodes(const int elementType, std::vector<int>& nodes, const int, const bool) { faceNodesOf(elementType, nodes); }
void getElementFaceNodes(const int elementType, const int, std::vector<int>& nodes, const int, const bool) { faceNodesOf(elementType, nodes); }

void setElementsByType(const int dim, const int tag, const int elementType, const std::vector<int>&, const std::vector<int>& nodeTags) {
    g_model.addElements(dim, tag, elementType, nodeTags);
}

void getBarycenters(const int elementType, const int tag, const bool, const bool, std::vector<double>& barycenters) {
    int dim, order;
    gml::elementTypeInfo(elementType, dim, order);
    const gml::RefElement& re = gml::refElement(dim, order);
    std::vector<int> tags, nodes;
    g_model.elementsByType(elementType, tags, nodes, tag);
    barycenters.assign(tags.size() * 3, 0.0);
    for (size_t e = 0; e < tags.size(); ++e) {
        for (int v = 0; v <= dim; ++v)
            for (int x = 0; x < 3; ++x) barycenters[3 * e + x] += g_model.node(nodes[e * re.np + v])[x];
        for (int x = 0; x < 3; ++x) barycenters[3 * e + x] /= (dim + 1);
    }
}

void getNode(const int nodeTag, std::vector<double>& coord, std::vector<double>& parametricCoord) {
    coord.assign(g_model.node(nodeTag), g_model.node(nodeTag) + 3);
    parametricCoord.clear();
}

void getNodesForPhysicalGroup(const int dim, const int tag, std::vector<int>& nodeTags, std::vector<double>& coord) {
    g_model.nodesForPhysicalGroup(dim, tag, nodeTags);
    coord.clear();
    for (int t : nodeTags) coord.insert(coord.end(), g_model.node(t), g_model.node(t) + 3);
}
}  // namespace mesh
}  // namespace model

namespace view {
int add(const std::string& name, const int tag) {
    int t = tag;
    if (t < 0) { t = 1; while (g_views.count(t)) ++t; }
    g_views[t].name = name;
    return t;
}
void addModelData(const int tag, const int step, const std::string&, const std::string&, const std::vector<int>& tags,
                  const std::vector<std::vector<double> >& data, const double time, const int numComponents, const int) {
    View& v = g_views[tag];
    v.tags = tags;
    ViewStep s;
    s.step = step;
    s.time = time;
    s.ncomp = numComponents;
    for (const auto& row : data) s.data.insert(s.data.end(), row.begin(), row.end());
    v.steps.push_back(std::move(s));
}
void addListData(const int, const std::string&, const int, const std::vector<double>&) {}
// Layout of <file>.<ViewName>.bin : int32 nSteps, int32 nElements, int32 valuesPerElement; then per step:
// int32 step, float64 time, float64 data[nElements*valuesPerElement] (element-major, the order addModelData received).
void write(const int tag, const std::string& fileName, const bool) {
    auto it = g_views.find(tag);
    if (it == g_views.end() || it->second.steps.empty()) return;
    const View& v = it->second;
    const std::string path = fileName + "." + v.name + ".bin";
    std::FILE* fp = std::fopen(path.c_str(), "wb");
    if (!fp) throw std::runtime_error("gmsh shim: cannot write " + path);
    const int32_t ns = (int32_t)v.steps.size(), ne = (int32_t)v.tags.size(), per = ne ? (int32_t)(v.steps[0].data.size() / ne) : 0;
    std::fwrite(&ns, 4, 1, fp); std::fwrite(&ne, 4, 1, fp); std::fwrite(&per, 4, 1, fp);
    for (const auto& s : v.steps) {
        const int32_t st = s.step;
        std::fwrite(&st, 4, 1, fp);
        std::fwrite(&s.time, 8, 1, fp);
        std::fwrite(s.data.data(), 8, s.data.size(), fp);
    }
    std::fclose(fp);
}
}  // namespace view

}  // namespace gmsh

// The reference's unused lapack:: namespace (src/utils.cpp:5-91) references these Fortran symbols; they are
// never called, so aborting stubs are enough to link.
extern "C" {
void dgetrf_(int*, int*, double*, int*, int*, int*) { std::abort(); }
void dgetri_(int*, double*, int*, int*, double*, int*, int*) { std::abort(); }
void dgesv_(int*, int*, double*, int*, int*, double*, int*, int*) { std::abort(); }
double dlange_(char*, int*, int*, double*, int*, double*) { std::abort(); }
double ddot_(int*, double*, int*, double*, int*) { std::abort(); }
void dgemv_(char&, int&, int&, double&, double*, int&, double*, int&, double&, double*, int&) { std::abort(); }
}
