// gmshlite — a small, self-contained stand-in for the slice of the Gmsh SDK 4.1.4 API that the
// reference front end uses: MSH ASCII reader (4.0, 4.1, 2.2), equispaced Lagrange simplex
// elements of order 1..6 (line / triangle / tetrahedron) in Gmsh's node ordering, exact quadrature of a
// requested degree, Jacobians with Gmsh's lower-dimensional completion, straight-sided order elevation
// and a structured Kuhn-tetrahedra cube generator.
//
// Gmsh itself is not available in this environment, so conventions that cannot be observed here
// (high-order node numbering inside faces/volumes, the position of the first Gauss point) follow the
// published Gmsh conventions as closely as they are known; DESIGN.md lists them. None of them enters
// the hot path other than through the arrays handed to the C ABI (include/dgb.h).
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace gml {

// ---------------------------------------------------------------------------------------------
// Reference elements
// ---------------------------------------------------------------------------------------------
struct RefElement {
    int dim = 0, order = 0, np = 0, type = 0;
    std::string name;
    std::vector<std::array<int, 4>> bary;  // integer barycentric coordinates (sum == order)
    std::vector<double> uvw;               // np*3 parametric coordinates (line: u in [-1,1])
    int nFaces = 0, nfp = 0;               // (dim-1)-faces per element, nodes per face
    std::vector<int> faceNodes;            // [nFaces][nfp] local node ids in the face's own Gmsh order
    // Lagrange basis at one parametric point (long double inside, rounded once)
    void basis(const double* uvw_pt, double* phi) const;
    void gradBasis(const double* uvw_pt, double* dphi /* np*3 */) const;
};
const RefElement& refElement(int dim, int order);
int elementType(int dim, int order);                 // Gmsh element type id
bool elementTypeInfo(int type, int& dim, int& order);

struct Quadrature {
    int n = 0;
    std::vector<double> pts;  // n*4 : u, v, w, weight
};
// Conical-product Gauss–Jacobi rule, exact for polynomials of total degree <= degree on the
// reference simplex of dimension dim (dim 0: one point of weight 1). Points are sorted so that
// the first one is the closest to vertex 0.
const Quadrature& gaussRule(int dim, int degree);
// The rule the stand-in hands out for "Gauss<degree>" on an element of dimension dim and Lagrange order `order`.
// Identical to gaussRule except for triangles used as faces of a 3D mesh (faceOf3D): there the reference derives
// the raw sign of the face normal from d(phi0,phi1)/d(u,v) at the rule's FIRST point (Mesh.cpp:183-188), which
// for Gmsh's own tables cannot be observed here. The reference scheme is stable only when
// fc * orientation(first owner) = +1 with fc = (order == 1 ? +1 : -1) (Mesh.cpp:210-211), so the stand-in lists
// first a point where that Jacobian sign equals fc. Only the ORDER of the points changes, not the rule.
Quadrature integrationRule(int dim, int degree, int order, bool faceOf3D);

// ---------------------------------------------------------------------------------------------
// Model (what gmsh::open leaves in memory)
// ---------------------------------------------------------------------------------------------
struct Entity {
    int dim = 0, tag = 0;
    std::vector<int> phys;
};
struct ElemBlock {
    int entityTag = 0, entityDim = 0, type = 0;
    std::vector<int> tags;      // element tags
    std::vector<int> nodeTags;  // numNodes(type) per element
};
struct Model {
    std::string name;
    std::vector<double> xyz;  // 3*(maxTag+1), indexed by node tag
    int maxNodeTag = 0;
    int maxElemTag = 0;
    std::map<std::pair<int, int>, std::string> physNames;  // (dim, physTag) -> name
    std::vector<Entity> entities;
    std::vector<ElemBlock> blocks;
    bool curved = false;  // high-order nodes moved off the straight-sided positions (warp): Jacobians vary inside an element

    int dimension() const;  // highest dimension that carries elements
    std::vector<int> elementTypes(int dim) const;
    // concatenation over all blocks of that type (entityTag < 0) or of one entity
    void elementsByType(int type, std::vector<int>& tags, std::vector<int>& nodeTags, int entityTag = -1) const;
    std::vector<int> physicalGroups(int dim) const;  // ascending physical tags that own entities
    std::string physicalName(int dim, int tag) const;
    void nodesForPhysicalGroup(int dim, int physTag, std::vector<int>& nodeTags) const;
    int addDiscreteEntity(int dim);
    void addElements(int dim, int entityTag, int type, const std::vector<int>& nodeTags);
    const double* node(int tag) const { return &xyz[3 * (size_t)tag]; }
};

// MSH ASCII: 4.0 (the only format in doc/**/*.msh), 4.1 (what current Gmsh writes) and 2.2 (entities are
// synthesised from the elements' elementary / physical tags). Binary files are rejected. Throws std::runtime_error.
Model readMsh(const std::string& path);
// Straight-sided elevation of an order-1 model to `order` (the stand-in for `gmsh -order p`).
void elevate(Model& m, int order);
// n^3 cells x 6 Kuhn tetrahedra on [lo,hi]^3, already at `order`, cells in Morton order, boundary
// triangles in a physical group "Boundary" (tag 1). Node tags are lattice indices (no hashing).
Model makeCube(int n, double lo, double hi, int order, bool withBoundaryElements = true);
Model makeSquare(int n, double lo, double hi, int order, bool withBoundaryElements = true);

// Jacobians as Gmsh returns them: 9 doubles per element, index u*3+x = d x_x / d u_u, with Gmsh's
// completion for elements of dimension < 3; det as Gmsh defines it (signed for tets, a norm below).
// All elements are straight-sided, so one Jacobian per element is returned (constant over the points).
void affineJacobian(const Model& m, int dim, const int* vertexTags, double jac[9], double& det);
// The same for an isoparametric (possibly curved) element of Lagrange order `order` at the parametric point uvw:
// tangents d x / d u_u = sum_n x_n d phi_n / d u_u, then the same completion and determinant conventions.
void isoJacobian(const Model& m, int dim, int order, const int* nodeTags, const double* uvw, double jac[9], double& det);
// Smooth displacement of EVERY node, x += amp * (sin(k y + 0.3), sin(k z + 0.7), sin(k x + 1.1)) restricted to the model's
// dimension (2D models stay in their plane): turns a straight-sided order-p mesh into a conforming curved isoparametric one
// (the stand-in for meshing a curved geometry with `gmsh -order p`). Sets Model::curved.
void warp(Model& m, double amp, double k);
// Sets Model::curved if some high-order node of some element is not where the affine map of the element's vertices puts it
// (a mesh file written by `gmsh -order p` on a curved geometry); returns the flag.
bool detectCurved(Model& m);
// The same displacement multiplied by the window (1 - (r/R)^2)^2 around `center` (zero for r >= R): only the elements that
// reach into the ball become curved, the rest of the mesh stays exactly straight-sided (a curved layer in a straight mesh).
void warpLocal(Model& m, double amp, double k, const double center[3], double radius);

}  // namespace gml
