// gmshlite implementation — see gmshlite.h. Everything here is set-up code that runs once.
#include "gmshlite.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace gml {

// =============================================================================================
// Element type table (Gmsh ids: lines 1,8,26,27,28,62; triangles 2,9,21,23,25,42; tets 4,11,29,30,31,71)
// =============================================================================================
static const int kTypeTable[4][7] = {
    {15, 15, 15, 15, 15, 15, 15},
    {0, 1, 8, 26, 27, 28, 62},
    {0, 2, 9, 21, 23, 25, 42},
    {0, 4, 11, 29, 30, 31, 71},
};

int elementType(int dim, int order) {
    if (dim == 0) return 15;
    if (dim < 0 || dim > 3 || order < 1 || order > 6) throw std::runtime_error("gmshlite: unsupported element dim/order");
    return kTypeTable[dim][order];
}

bool elementTypeInfo(int type, int& dim, int& order) {
    if (type == 15) { dim = 0; order = 1; return true; }
    for (int d = 1; d <= 3; ++d)
        for (int p = 1; p <= 6; ++p)
            if (kTypeTable[d][p] == type) { dim = d; order = p; return true; }
    return false;
}

// =============================================================================================
// Node lattices in Gmsh ordering (vertices, edges, faces, interior — recursively)
// =============================================================================================
static void triLattice(int o, int s, std::vector<std::array<int, 2>>& out) {
    if (s < 0) return;
    if (s == 0) { out.push_back({o, o}); return; }
    out.push_back({o, o});
    out.push_back({o + s, o});
    out.push_back({o, o + s});
    for (int k = 1; k < s; ++k) out.push_back({o + k, o});
    for (int k = 1; k < s; ++k) out.push_back({o + s - k, o + k});
    for (int k = 1; k < s; ++k) out.push_back({o, o + s - k});
    if (s >= 3) triLattice(o + 1, s - 3, out);
}

static const int kTetEdges[6][2] = {{0, 1}, {1, 2}, {2, 0}, {3, 0}, {3, 2}, {3, 1}};
static const int kTetFaces[4][3] = {{0, 2, 1}, {0, 1, 3}, {0, 3, 2}, {3, 1, 2}};
static const int kTriEdges[3][2] = {{0, 1}, {1, 2}, {2, 0}};

static void tetLattice(int o, int s, std::vector<std::array<int, 3>>& out) {
    if (s < 0) return;
    if (s == 0) { out.push_back({o, o, o}); return; }
    const std::array<int, 3> V[4] = {{o, o, o}, {o + s, o, o}, {o, o + s, o}, {o, o, o + s}};
    for (int v = 0; v < 4; ++v) out.push_back(V[v]);
    for (auto& e : kTetEdges) {
        const auto &A = V[e[0]], &B = V[e[1]];
        for (int k = 1; k < s; ++k)
            out.push_back({A[0] + k * (B[0] - A[0]) / s, A[1] + k * (B[1] - A[1]) / s, A[2] + k * (B[2] - A[2]) / s});
    }
    if (s >= 3) {
        std::vector<std::array<int, 2>> in;
        triLattice(1, s - 3, in);
        for (auto& f : kTetFaces) {
            const auto &A = V[f[0]], &B = V[f[1]], &C = V[f[2]];
            for (auto& ab : in) {
                std::array<int, 3> P;
                for (int c = 0; c < 3; ++c) P[c] = A[c] + ab[0] * ((B[c] - A[c]) / s) + ab[1] * ((C[c] - A[c]) / s);
                out.push_back(P);
            }
        }
    }
    if (s >= 4) tetLattice(o + 1, s - 4, out);
}

static RefElement buildRef(int dim, int order) {
    RefElement r;
    r.dim = dim;
    r.order = order;
    r.type = elementType(dim, order);
    const int p = order;
    if (dim == 0) {
        r.name = "Point";
        r.np = 1;
        r.bary.push_back({1, 0, 0, 0});
        r.uvw = {0, 0, 0};
        r.nFaces = 0;
        r.nfp = 0;
        return r;
    }
    if (dim == 1) {
        r.name = "Line " + std::to_string(p);
        std::vector<int> lat = {0, p};
        for (int i = 1; i < p; ++i) lat.push_back(i);
        for (int i : lat) {
            r.bary.push_back({p - i, i, 0, 0});
            r.uvw.insert(r.uvw.end(), {-1.0 + 2.0 * i / p, 0.0, 0.0});
        }
        r.np = (int)lat.size();
        // Reference quirk: "faces" of a line are single nodes; getElementEdgeNodes returns all
        // p+1 nodes of the element, so there are p+1 one-node faces per element (sane only for p == 1).
        r.nfp = 1;
        r.nFaces = r.np;
        for (int i = 0; i < r.np; ++i) r.faceNodes.push_back(i);
        return r;
    }
    if (dim == 2) {
        r.name = "Triangle " + std::to_string(p);
        std::vector<std::array<int, 2>> lat;
        triLattice(0, p, lat);
        r.np = (int)lat.size();
        std::map<std::array<int, 2>, int> idx;
        for (int n = 0; n < r.np; ++n) {
            idx[lat[n]] = n;
            r.bary.push_back({p - lat[n][0] - lat[n][1], lat[n][0], lat[n][1], 0});
            r.uvw.insert(r.uvw.end(), {(double)lat[n][0] / p, (double)lat[n][1] / p, 0.0});
        }
        r.nFaces = 3;
        r.nfp = p + 1;
        const std::array<int, 2> V[3] = {{0, 0}, {p, 0}, {0, p}};
        std::vector<int> line = {0, p};
        for (int i = 1; i < p; ++i) line.push_back(i);
        for (auto& e : kTriEdges)
            for (int a : line) {
                std::array<int, 2> P = {V[e[0]][0] + a * ((V[e[1]][0] - V[e[0]][0]) / p), V[e[0]][1] + a * ((V[e[1]][1] - V[e[0]][1]) / p)};
                r.faceNodes.push_back(idx.at(P));
            }
        return r;
    }
    r.name = "Tetrahedron " + std::to_string(p);
    std::vector<std::array<int, 3>> lat;
    tetLattice(0, p, lat);
    r.np = (int)lat.size();
    std::map<std::array<int, 3>, int> idx;
    for (int n = 0; n < r.np; ++n) {
        idx[lat[n]] = n;
        r.bary.push_back({p - lat[n][0] - lat[n][1] - lat[n][2], lat[n][0], lat[n][1], lat[n][2]});
        r.uvw.insert(r.uvw.end(), {(double)lat[n][0] / p, (double)lat[n][1] / p, (double)lat[n][2] / p});
    }
    r.nFaces = 4;
    r.nfp = (p + 1) * (p + 2) / 2;
    std::vector<std::array<int, 2>> tri;
    triLattice(0, p, tri);
    const std::array<int, 3> V[4] = {{0, 0, 0}, {p, 0, 0}, {0, p, 0}, {0, 0, p}};
    for (auto& f : kTetFaces)
        for (auto& ab : tri) {
            std::array<int, 3> P;
            for (int c = 0; c < 3; ++c)
                P[c] = V[f[0]][c] + ab[0] * ((V[f[1]][c] - V[f[0]][c]) / p) + ab[1] * ((V[f[2]][c] - V[f[0]][c]) / p);
            r.faceNodes.push_back(idx.at(P));
        }
    return r;
}

const RefElement& refElement(int dim, int order) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, RefElement> cache;
    std::lock_guard<std::mutex> lk(mu);
    if (dim == 0) order = 1;
    auto key = std::make_pair(dim, order);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, buildRef(dim, order)).first;
    return it->second;
}

// ---------------------------------------------------------------------------------------------
// Equispaced Lagrange basis in closed form:  phi_alpha(lambda) = prod_k prod_{m<alpha_k} (p*lambda_k - m)/(m+1)
// ---------------------------------------------------------------------------------------------
static inline long double facVal(int p, int a, long double lam) {
    long double r = 1.0L;
    for (int m = 0; m < a; ++m) r *= (p * lam - m) / (long double)(m + 1);
    return r;
}
static inline long double facDer(int p, int a, long double lam) {
    long double s = 0.0L;
    for (int m = 0; m < a; ++m) {
        long double t = (long double)p / (long double)(m + 1);
        for (int m2 = 0; m2 < a; ++m2)
            if (m2 != m) t *= (p * lam - m2) / (long double)(m2 + 1);
        s += t;
    }
    return s;
}
static inline void barycentric(int dim, const double* x, long double lam[4]) {
    lam[0] = lam[1] = lam[2] = lam[3] = 0.0L;
    if (dim == 0) { lam[0] = 1.0L; return; }
    if (dim == 1) { lam[0] = (1.0L - (long double)x[0]) / 2; lam[1] = (1.0L + (long double)x[0]) / 2; return; }
    long double s = 1.0L;
    for (int c = 0; c < dim; ++c) { lam[c + 1] = x[c]; s -= (long double)x[c]; }
    lam[0] = s;
}

void RefElement::basis(const double* x, double* phi) const {
    long double lam[4];
    barycentric(dim, x, lam);
    for (int n = 0; n < np; ++n) {
        long double v = 1.0L;
        for (int k = 0; k <= dim; ++k) v *= facVal(order, bary[n][k], lam[k]);
        phi[n] = (double)v;
    }
}

void RefElement::gradBasis(const double* x, double* dphi) const {
    long double lam[4];
    barycentric(dim, x, lam);
    for (int n = 0; n < np; ++n) {
        long double f[4], df[4];
        for (int k = 0; k <= dim; ++k) { f[k] = facVal(order, bary[n][k], lam[k]); df[k] = facDer(order, bary[n][k], lam[k]); }
        long double dl[4];  // d phi / d lambda_k
        for (int k = 0; k <= dim; ++k) {
            long double t = df[k];
            for (int k2 = 0; k2 <= dim; ++k2)
                if (k2 != k) t *= f[k2];
            dl[k] = t;
        }
        dphi[3 * n + 0] = dphi[3 * n + 1] = dphi[3 * n + 2] = 0.0;
        if (dim == 1) dphi[3 * n] = (double)((dl[1] - dl[0]) / 2);
        else
            for (int c = 0; c < dim; ++c) dphi[3 * n + c] = (double)(dl[c + 1] - dl[0]);
    }
}

// =============================================================================================
// Quadrature: Gauss–Jacobi (alpha, 0) rules and their conical product on the simplex
// =============================================================================================
static long double jacobiP(int n, int a, long double x) {  // P_n^{(a,0)}(x), standard normalisation
    if (n == 0) return 1.0L;
    long double p0 = 1.0L, p1 = (a + 1) + (a + 2) * (x - 1) / 2;
    for (int k = 1; k < n; ++k) {
        long double c = 2.0L * k + a;  // 2k + alpha + beta
        long double a1 = 2.0L * (k + 1) * (k + a + 1) * c;
        long double a2 = (c + 1) * ((c + 2) * c * x + (long double)a * a);
        long double a3 = 2.0L * (k + a) * k * (c + 2);
        long double p2 = (a2 * p1 - a3 * p0) / a1;
        p0 = p1;
        p1 = p2;
    }
    return p1;
}
static long double jacobiPab(int n, int a, int b, long double x) {  // general (a,b), used for the derivative
    if (n == 0) return 1.0L;
    long double p0 = 1.0L, p1 = (a + 1) + (a + b + 2) * (x - 1) / 2;
    for (int k = 1; k < n; ++k) {
        long double c = 2.0L * k + a + b;
        long double a1 = 2.0L * (k + 1) * (k + a + b + 1) * c;
        long double a2 = (c + 1) * ((c + 2) * c * x + (long double)a * a - (long double)b * b);
        long double a3 = 2.0L * (k + a) * (k + b) * (c + 2);
        long double p2 = (a2 * p1 - a3 * p0) / a1;
        p0 = p1;
        p1 = p2;
    }
    return p1;
}
static long double jacobiDP(int n, int a, long double x) {  // d/dx P_n^{(a,0)} = (n+a+1)/2 P_{n-1}^{(a+1,1)}
    if (n == 0) return 0.0L;
    return (n + a + 1) / 2.0L * jacobiPab(n - 1, a + 1, 1, x);
}

static void gaussJacobi(int n, int a, std::vector<long double>& x, std::vector<long double>& w) {
    x.clear();
    w.clear();
    const int grid = 40000;
    long double prevX = -1.0L, prevV = jacobiP(n, a, prevX);
    for (int i = 1; i <= grid && (int)x.size() < n; ++i) {
        long double cx = -1.0L + 2.0L * i / grid, cv = jacobiP(n, a, cx);
        if ((prevV < 0) != (cv < 0)) {
            long double lo = prevX, hi = cx, flo = prevV;
            for (int it = 0; it < 200; ++it) {
                long double mid = (lo + hi) / 2, fm = jacobiP(n, a, mid);
                if ((fm < 0) == (flo < 0)) { lo = mid; flo = fm; } else hi = mid;
            }
            long double r = (lo + hi) / 2;
            for (int it = 0; it < 4; ++it) r -= jacobiP(n, a, r) / jacobiDP(n, a, r);
            x.push_back(r);
        }
        prevX = cx;
        prevV = cv;
    }
    if ((int)x.size() != n) throw std::runtime_error("gmshlite: Gauss-Jacobi root search failed");
    for (int i = 0; i < n; ++i) {
        long double d = jacobiDP(n, a, x[i]);
        w.push_back(std::pow(2.0L, a + 1) / ((1 - x[i] * x[i]) * d * d));
    }
}

static Quadrature buildRule(int dim, int degree) {
    Quadrature q;
    if (dim == 0) { q.n = 1; q.pts = {0, 0, 0, 1}; return q; }
    int n = degree / 2 + 1;  // 2n-1 >= degree
    std::vector<long double> xr, wr, xs, ws, xt, wt;
    gaussJacobi(n, 0, xr, wr);
    if (dim == 1) {
        for (int i = 0; i < n; ++i) q.pts.insert(q.pts.end(), {(double)xr[i], 0.0, 0.0, (double)wr[i]});
    } else if (dim == 2) {
        gaussJacobi(n, 1, xs, ws);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                long double v = (1 + xs[j]) / 2, u = (1 + xr[i]) / 2 * (1 - v);
                q.pts.insert(q.pts.end(), {(double)u, (double)v, 0.0, (double)(wr[i] * ws[j] / 8)});
            }
    } else {
        gaussJacobi(n, 1, xs, ws);
        gaussJacobi(n, 2, xt, wt);
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    long double w = (1 + xt[k]) / 2, v = (1 + xs[j]) / 2 * (1 - w), u = (1 + xr[i]) / 2 * (1 - v - w);
                    q.pts.insert(q.pts.end(), {(double)u, (double)v, (double)w, (double)(wr[i] * ws[j] * wt[k] / 64)});
                }
    }
    q.n = (int)q.pts.size() / 4;
    return q;
}

const Quadrature& gaussRule(int dim, int degree) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, Quadrature> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(dim, dim == 0 ? 0 : degree);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, buildRule(dim, degree)).first;
    return it->second;
}

Quadrature integrationRule(int dim, int degree, int order, bool faceOf3D) {
    Quadrature q = gaussRule(dim, degree);
    if (!(faceOf3D && dim == 2)) return q;
    const int fc = order != 1 ? -1 : 1;
    const RefElement& rf = refElement(2, order);
    std::vector<double> dphi((size_t)rf.np * 3);
    for (int g = 0; g < q.n; ++g) {
        rf.gradBasis(&q.pts[4 * g], dphi.data());
        const double det = dphi[0] * dphi[4] - dphi[1] * dphi[3];
        if (det * fc > 0) {
            for (int k = 0; k < 4; ++k) std::swap(q.pts[k], q.pts[4 * g + k]);
            break;
        }
    }
    return q;
}

// =============================================================================================
// Model queries
// =============================================================================================
int Model::dimension() const {
    int d = -1;
    for (auto& b : blocks) if (!b.tags.empty()) d = std::max(d, b.entityDim);
    return d;
}

std::vector<int> Model::elementTypes(int dim) const {
    std::vector<int> t;
    for (auto& b : blocks)
        if (b.entityDim == dim && !b.tags.empty() && std::find(t.begin(), t.end(), b.type) == t.end()) t.push_back(b.type);
    return t;
}

void Model::elementsByType(int type, std::vector<int>& tags, std::vector<int>& nodeTags, int entityTag) const {
    tags.clear();
    nodeTags.clear();
    int tdim, tord;
    if (!elementTypeInfo(type, tdim, tord)) return;
    for (auto& b : blocks) {
        if (b.type != type) continue;
        if (entityTag >= 0 && !(b.entityTag == entityTag && b.entityDim == tdim)) continue;
        tags.insert(tags.end(), b.tags.begin(), b.tags.end());
        nodeTags.insert(nodeTags.end(), b.nodeTags.begin(), b.nodeTags.end());
    }
}

std::vector<int> Model::physicalGroups(int dim) const {
    std::vector<int> g;
    for (auto& e : entities)
        if (e.dim == dim)
            for (int p : e.phys)
                if (std::find(g.begin(), g.end(), p) == g.end()) g.push_back(p);
    std::sort(g.begin(), g.end());
    return g;
}

std::string Model::physicalName(int dim, int tag) const {
    auto it = physNames.find({dim, tag});
    return it == physNames.end() ? std::string() : it->second;
}

void Model::nodesForPhysicalGroup(int dim, int physTag, std::vector<int>& nodeTags) const {
    nodeTags.clear();
    std::vector<char> seen(maxNodeTag + 1, 0);
    for (auto& e : entities) {
        if (e.dim != dim || std::find(e.phys.begin(), e.phys.end(), physTag) == e.phys.end()) continue;
        for (auto& b : blocks) {
            if (b.entityDim != dim || b.entityTag != e.tag) continue;
            for (int t : b.nodeTags)
                if (!seen[t]) { seen[t] = 1; nodeTags.push_back(t); }
        }
    }
}

int Model::addDiscreteEntity(int dim) {
    int tag = 0;
    for (auto& e : entities) if (e.dim == dim) tag = std::max(tag, e.tag);
    for (auto& b : blocks) if (b.entityDim == dim) tag = std::max(tag, b.entityTag);
    Entity e;
    e.dim = dim;
    e.tag = tag + 1;
    entities.push_back(e);
    return e.tag;
}

void Model::addElements(int dim, int entityTag, int type, const std::vector<int>& nodeTags) {
    int tdim, tord;
    if (!elementTypeInfo(type, tdim, tord) || tdim != dim) throw std::runtime_error("gmshlite: addElements type/dim mismatch");
    const int nn = refElement(tdim, tord).np;
    ElemBlock b;
    b.entityTag = entityTag;
    b.entityDim = dim;
    b.type = type;
    b.nodeTags = nodeTags;
    const size_t ne = nodeTags.size() / nn;
    b.tags.resize(ne);
    for (size_t i = 0; i < ne; ++i) b.tags[i] = ++maxElemTag;
    blocks.push_back(std::move(b));
}

// =============================================================================================
// MSH ASCII reader (4.0, 4.1, 2.2)
// =============================================================================================
Model readMsh(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("gmshlite: cannot open " + path);
    Model m;
    m.name = path;
    {
        auto slash = path.find_last_of('/');
        std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
        auto dot = base.find_last_of('.');
        m.name = dot == std::string::npos ? base : base.substr(0, dot);
    }
    std::string line;
    auto expect = [&](const char* what) {
        if (!std::getline(in, line)) throw std::runtime_error(std::string("gmshlite: unexpected end of file in ") + what);
    };
    std::vector<std::pair<int, std::vector<double>>> nodeStage;
    int fmt = 40;  // 22: MSH 2.2, 40: MSH 4.0 (the reference's sample meshes), 41: MSH 4.1 (what current Gmsh writes)
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        if (line == "$MeshFormat") {
            expect("MeshFormat");
            double ver;
            int ftype, dsize;
            std::istringstream(line) >> ver >> ftype >> dsize;
            if (ftype != 0) throw std::runtime_error("gmshlite: binary MSH files are not supported (got '" + line + "')");
            if (ver >= 2.0 && ver < 3.0) fmt = 22;
            else if (ver >= 4.0 && ver < 4.05) fmt = 40;
            else if (ver >= 4.05 && ver < 4.2) fmt = 41;
            else throw std::runtime_error("gmshlite: MSH ASCII versions 2.x, 4.0 and 4.1 are supported (got '" + line + "')");
        } else if (line == "$PhysicalNames") {
            expect("PhysicalNames");
            int n = std::stoi(line);
            for (int i = 0; i < n; ++i) {
                expect("PhysicalNames");
                std::istringstream ss(line);
                int d, t;
                ss >> d >> t;
                auto q0 = line.find('"'), q1 = line.rfind('"');
                std::string nm = (q0 != std::string::npos && q1 > q0) ? line.substr(q0 + 1, q1 - q0 - 1) : std::string();
                m.physNames[{d, t}] = nm;
            }
        } else if (line == "$Entities") {
            expect("Entities");
            size_t cnt[4];
            std::istringstream(line) >> cnt[0] >> cnt[1] >> cnt[2] >> cnt[3];
            for (int d = 0; d < 4; ++d)
                for (size_t i = 0; i < cnt[d]; ++i) {
                    expect("Entities");
                    std::istringstream ss(line);
                    Entity e;
                    e.dim = d;
                    double bb;
                    ss >> e.tag;
                    for (int k = 0; k < (fmt == 41 && d == 0 ? 3 : 6); ++k) ss >> bb;  // 4.1: points carry X Y Z, the rest a bounding box
                    size_t np = 0;
                    ss >> np;
                    e.phys.resize(np);
                    for (size_t k = 0; k < np; ++k) ss >> e.phys[k];
                    m.entities.push_back(e);
                }
        } else if (line == "$Nodes" && fmt == 22) {  // MSH 2.2: numNodes, then "tag x y z"
            expect("Nodes");
            const size_t nn = std::stoul(line);
            for (size_t i = 0; i < nn; ++i) {
                expect("Nodes");
                std::istringstream ss(line);
                int tag;
                double x, y, z;
                ss >> tag >> x >> y >> z;
                if (tag > m.maxNodeTag) m.maxNodeTag = tag;
                nodeStage.push_back({tag, {x, y, z}});
            }
            m.xyz.assign(3 * (size_t)(m.maxNodeTag + 1), 0.0);
            for (auto& n : nodeStage) std::copy(n.second.begin(), n.second.end(), m.xyz.begin() + 3 * (size_t)n.first);
        } else if (line == "$Nodes" && fmt == 41) {  // MSH 4.1: per block the node tags first, then the coordinates
            expect("Nodes");
            size_t nb, nn;
            std::istringstream(line) >> nb >> nn;
            for (size_t b = 0; b < nb; ++b) {
                expect("Nodes");
                int ed, et, par;
                size_t cnt;
                std::istringstream(line) >> ed >> et >> par >> cnt;
                std::vector<int> tags(cnt);
                for (size_t i = 0; i < cnt; ++i) { expect("Nodes"); tags[i] = std::stoi(line); }
                for (size_t i = 0; i < cnt; ++i) {
                    expect("Nodes");
                    std::istringstream ss(line);
                    double x, y, z;
                    ss >> x >> y >> z;
                    if (tags[i] > m.maxNodeTag) m.maxNodeTag = tags[i];
                    nodeStage.push_back({tags[i], {x, y, z}});
                }
            }
            m.xyz.assign(3 * (size_t)(m.maxNodeTag + 1), 0.0);
            for (auto& n : nodeStage) std::copy(n.second.begin(), n.second.end(), m.xyz.begin() + 3 * (size_t)n.first);
        } else if (line == "$Nodes") {
            expect("Nodes");
            size_t nb, nn;
            std::istringstream(line) >> nb >> nn;
            for (size_t b = 0; b < nb; ++b) {
                expect("Nodes");
                int et, ed, par;
                size_t cnt;
                std::istringstream(line) >> et >> ed >> par >> cnt;
                for (size_t i = 0; i < cnt; ++i) {
                    expect("Nodes");
                    std::istringstream ss(line);
                    int tag;
                    double x, y, z;
                    ss >> tag >> x >> y >> z;
                    if (tag > m.maxNodeTag) m.maxNodeTag = tag;
                    nodeStage.push_back({tag, {x, y, z}});
                }
            }
            m.xyz.assign(3 * (size_t)(m.maxNodeTag + 1), 0.0);
            for (auto& n : nodeStage) std::copy(n.second.begin(), n.second.end(), m.xyz.begin() + 3 * (size_t)n.first);
        } else if (line == "$Elements" && fmt == 22) {
            // MSH 2.2: "number type ntags <physical elementary ...> nodes"; no $Entities section: one entity per (dimension,
            // elementary tag) carrying the physical tag, one block per run of elements of the same entity and type
            expect("Elements");
            const size_t ne = std::stoul(line);
            for (size_t i = 0; i < ne; ++i) {
                expect("Elements");
                std::istringstream ss(line);
                int tag, type, ntags, phys = 0, elem = 0, t;
                ss >> tag >> type >> ntags;
                for (int k = 0; k < ntags; ++k) { ss >> t; if (k == 0) phys = t; if (k == 1) elem = t; }
                int tdim, tord;
                if (!elementTypeInfo(type, tdim, tord)) {
                    if (type == 15) continue;  // point elements carry nothing the solver uses
                    throw std::runtime_error("gmshlite: unsupported element type " + std::to_string(type));
                }
                const int nn = refElement(tdim, tord).np;
                if (m.blocks.empty() || m.blocks.back().entityTag != elem || m.blocks.back().entityDim != tdim || m.blocks.back().type != type) {
                    ElemBlock blk;
                    blk.entityTag = elem; blk.entityDim = tdim; blk.type = type;
                    m.blocks.push_back(std::move(blk));
                    bool known = false;
                    for (Entity& e : m.entities)
                        if (e.dim == tdim && e.tag == elem) {
                            known = true;
                            if (phys != 0 && std::find(e.phys.begin(), e.phys.end(), phys) == e.phys.end()) e.phys.push_back(phys);
                        }
                    if (!known) {
                        Entity e;
                        e.dim = tdim; e.tag = elem;
                        if (phys != 0) e.phys.push_back(phys);
                        m.entities.push_back(e);
                    }
                }
                ElemBlock& blk = m.blocks.back();
                blk.tags.push_back(tag);
                if (tag > m.maxElemTag) m.maxElemTag = tag;
                for (int k = 0; k < nn; ++k) { ss >> t; blk.nodeTags.push_back(t); }
            }
        } else if (line == "$Elements") {
            expect("Elements");
            size_t nb, ne;
            std::istringstream(line) >> nb >> ne;
            for (size_t b = 0; b < nb; ++b) {
                expect("Elements");
                ElemBlock blk;
                size_t cnt;
                if (fmt == 41) std::istringstream(line) >> blk.entityDim >> blk.entityTag >> blk.type >> cnt;  // 4.1 swapped the first two
                else std::istringstream(line) >> blk.entityTag >> blk.entityDim >> blk.type >> cnt;
                int tdim, tord;
                if (!elementTypeInfo(blk.type, tdim, tord))
                    throw std::runtime_error("gmshlite: unsupported element type " + std::to_string(blk.type));
                const int nn = refElement(tdim, tord).np;
                blk.tags.reserve(cnt);
                blk.nodeTags.reserve(cnt * nn);
                for (size_t i = 0; i < cnt; ++i) {
                    expect("Elements");
                    std::istringstream ss(line);
                    int tag, nt;
                    ss >> tag;
                    blk.tags.push_back(tag);
                    if (tag > m.maxElemTag) m.maxElemTag = tag;
                    for (int k = 0; k < nn; ++k) { ss >> nt; blk.nodeTags.push_back(nt); }
                }
                m.blocks.push_back(std::move(blk));
            }
        }
    }
    if (m.blocks.empty()) throw std::runtime_error("gmshlite: no $Elements section in " + path);
    return m;
}

// =============================================================================================
// Straight-sided order elevation
// =============================================================================================
namespace {
struct KeyHash {
    size_t operator()(const std::array<int, 8>& k) const {
        uint64_t h = 1469598103934665603ull;
        for (int v : k) { h ^= (uint32_t)v; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
}  // namespace

void elevate(Model& m, int order) {
    if (order == 1) return;
    if (order < 1 || order > 6) throw std::runtime_error("gmshlite: order must be in 1..6");
    std::unordered_map<std::array<int, 8>, int, KeyHash> created;
    for (auto& b : m.blocks) {
        int dim, ord;
        elementTypeInfo(b.type, dim, ord);
        if (dim == 0) continue;
        if (ord != 1) throw std::runtime_error("gmshlite: elevate expects an order-1 mesh");
        const RefElement& hi = refElement(dim, order);
        const int nv = dim + 1;
        const size_t ne = b.tags.size();
        std::vector<int> nt(ne * hi.np);
        for (size_t e = 0; e < ne; ++e) {
            const int* v = &b.nodeTags[e * nv];
            for (int n = 0; n < hi.np; ++n) {
                // key: (vertex tag, weight) pairs with non-zero weight, sorted by tag
                std::array<std::pair<int, int>, 4> pr;
                int cnt = 0;
                for (int k = 0; k < nv; ++k)
                    if (hi.bary[n][k] > 0) pr[cnt++] = {v[k], hi.bary[n][k]};
                if (cnt == 1) { nt[e * hi.np + n] = pr[0].first; continue; }
                std::sort(pr.begin(), pr.begin() + cnt);
                std::array<int, 8> key{};
                for (int k = 0; k < cnt; ++k) { key[2 * k] = pr[k].first; key[2 * k + 1] = pr[k].second; }
                auto it = created.find(key);
                if (it == created.end()) {
                    int tag = ++m.maxNodeTag;
                    double x[3] = {0, 0, 0};
                    for (int k = 0; k < cnt; ++k)
                        for (int c = 0; c < 3; ++c) x[c] += (double)pr[k].second / order * m.xyz[3 * (size_t)pr[k].first + c];
                    m.xyz.insert(m.xyz.end(), x, x + 3);
                    it = created.emplace(key, tag).first;
                }
                nt[e * hi.np + n] = it->second;
            }
        }
        b.nodeTags.swap(nt);
        b.type = hi.type;
    }
}

// =============================================================================================
// Structured cube: n^3 cells x 6 Kuhn tetrahedra, cells in Morton order
// =============================================================================================
static inline uint64_t spread3(uint64_t v) {
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

Model makeCube(int n, double lo, double hi, int order, bool withBoundaryElements) {
    if (n < 1 || order < 1 || order > 6) throw std::runtime_error("gmshlite: makeCube bad arguments");
    Model m;
    m.name = "cube" + std::to_string(n);
    const int p = order;
    const int64_t L = (int64_t)p * n + 1;  // lattice points per direction
    if (L * L * L > 2000000000LL) throw std::runtime_error("gmshlite: makeCube lattice exceeds 32-bit node tags");
    m.maxNodeTag = (int)(L * L * L);
    m.xyz.assign(3 * (size_t)(m.maxNodeTag + 1), 0.0);
    const double h = (hi - lo) / ((double)p * n);
    for (int64_t k = 0; k < L; ++k)
        for (int64_t j = 0; j < L; ++j)
            for (int64_t i = 0; i < L; ++i) {
                size_t t = 1 + (size_t)(i + L * (j + L * k));
                m.xyz[3 * t + 0] = (i == L - 1) ? hi : lo + h * i;
                m.xyz[3 * t + 1] = (j == L - 1) ? hi : lo + h * j;
                m.xyz[3 * t + 2] = (k == L - 1) ? hi : lo + h * k;
            }
    auto tagOf = [&](int64_t i, int64_t j, int64_t k) { return (int)(1 + i + L * (j + L * k)); };

    m.physNames[{2, 1}] = "Boundary";
    m.physNames[{3, 2}] = "Domain";
    Entity vol;  vol.dim = 3; vol.tag = 1; vol.phys = {2};
    Entity surf; surf.dim = 2; surf.tag = 1; surf.phys = {1};
    m.entities = {surf, vol};

    // Kuhn tetrahedra of the unit cell: one per permutation of the axes, made positively oriented.
    static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
    int kuhn[6][4][3];
    for (int t = 0; t < 6; ++t) {
        int v[4][3] = {{0, 0, 0}};
        for (int s = 0; s < 3; ++s) {
            for (int c = 0; c < 3; ++c) v[s + 1][c] = v[s][c];
            v[s + 1][perms[t][s]] += 1;
        }
        int a[3], b[3], c[3];
        for (int q = 0; q < 3; ++q) { a[q] = v[1][q] - v[0][q]; b[q] = v[2][q] - v[0][q]; c[q] = v[3][q] - v[0][q]; }
        int det = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
        if (det < 0) for (int q = 0; q < 3; ++q) std::swap(v[2][q], v[3][q]);
        std::memcpy(kuhn[t], v, sizeof(v));
    }

    std::vector<std::pair<uint64_t, std::array<int, 3>>> cells;
    cells.reserve((size_t)n * n * n);
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) cells.push_back({spread3(i) | spread3(j) << 1 | spread3(k) << 2, {i, j, k}});
    std::sort(cells.begin(), cells.end());

    const RefElement& re = refElement(3, p);
    ElemBlock tets;
    tets.entityTag = 1;
    tets.entityDim = 3;
    tets.type = re.type;
    const size_t K = cells.size() * 6;
    tets.tags.resize(K);
    tets.nodeTags.resize(K * re.np);
    const RefElement& rf = refElement(2, p);
    ElemBlock tris;
    tris.entityTag = 1;
    tris.entityDim = 2;
    tris.type = rf.type;

    size_t e = 0;
    for (auto& cell : cells) {
        const int ci = cell.second[0], cj = cell.second[1], ck = cell.second[2];
        for (int t = 0; t < 6; ++t, ++e) {
            int64_t V[4][3];
            for (int v = 0; v < 4; ++v) {
                V[v][0] = (int64_t)p * (ci + kuhn[t][v][0]);
                V[v][1] = (int64_t)p * (cj + kuhn[t][v][1]);
                V[v][2] = (int64_t)p * (ck + kuhn[t][v][2]);
            }
            for (int nn = 0; nn < re.np; ++nn) {
                int64_t P[3] = {0, 0, 0};
                for (int v = 0; v < 4; ++v)
                    for (int c = 0; c < 3; ++c) P[c] += re.bary[nn][v] * V[v][c];
                tets.nodeTags[e * re.np + nn] = tagOf(P[0] / p, P[1] / p, P[2] / p);
            }
            if (withBoundaryElements) {
                for (int lf = 0; lf < 4; ++lf) {
                    const int* fn = &re.faceNodes[lf * re.nfp];
                    bool onB = false;
                    for (int c = 0; c < 3 && !onB; ++c) {
                        int64_t a0 = V[kTetFaces[lf][0]][c], a1 = V[kTetFaces[lf][1]][c], a2 = V[kTetFaces[lf][2]][c];
                        if (a0 == a1 && a1 == a2 && (a0 == 0 || a0 == (int64_t)p * n)) onB = true;
                    }
                    if (!onB) continue;
                    for (int q = 0; q < re.nfp; ++q) tris.nodeTags.push_back(tets.nodeTags[e * re.np + fn[q]]);
                }
            }
        }
    }
    int tag = 0;
    if (withBoundaryElements) {
        tris.tags.resize(tris.nodeTags.size() / rf.np);
        for (auto& t : tris.tags) t = ++tag;
        m.blocks.push_back(std::move(tris));
    }
    for (auto& t : tets.tags) t = ++tag;
    m.maxElemTag = tag;
    m.blocks.push_back(std::move(tets));
    return m;
}

// Structured square [lo,hi]^2 in the z = 0 plane: n^2 cells x 2 positively oriented triangles (cells in Morton order), order p,
// the boundary edges as line elements of the physical group "Boundary" (the 2D twin of makeCube, for the order sweep of
// north_star config 2 on a refined square).
Model makeSquare(int n, double lo, double hi, int order, bool withBoundaryElements) {
    if (n < 1 || order < 1 || order > 6) throw std::runtime_error("gmshlite: makeSquare bad arguments");
    Model m;
    m.name = "square" + std::to_string(n);
    const int p = order;
    const int64_t L = (int64_t)p * n + 1;
    if (L * L > 2000000000LL) throw std::runtime_error("gmshlite: makeSquare lattice exceeds 32-bit node tags");
    m.maxNodeTag = (int)(L * L);
    m.xyz.assign(3 * (size_t)(m.maxNodeTag + 1), 0.0);
    const double h = (hi - lo) / ((double)p * n);
    for (int64_t j = 0; j < L; ++j)
        for (int64_t i = 0; i < L; ++i) {
            const size_t t = 1 + (size_t)(i + L * j);
            m.xyz[3 * t + 0] = (i == L - 1) ? hi : lo + h * i;
            m.xyz[3 * t + 1] = (j == L - 1) ? hi : lo + h * j;
        }
    auto tagOf = [&](int64_t i, int64_t j) { return (int)(1 + i + L * j); };
    m.physNames[{1, 1}] = "Boundary";
    m.physNames[{2, 2}] = "Domain";
    Entity surf; surf.dim = 2; surf.tag = 1; surf.phys = {2};
    Entity curve; curve.dim = 1; curve.tag = 1; curve.phys = {1};
    m.entities = {curve, surf};
    static const int tri[2][3][2] = {{{0, 0}, {1, 0}, {1, 1}}, {{0, 0}, {1, 1}, {0, 1}}};
    std::vector<std::pair<uint64_t, std::array<int, 2>>> cells;
    cells.reserve((size_t)n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) cells.push_back({spread3(i) | spread3(j) << 1, {i, j}});
    std::sort(cells.begin(), cells.end());
    const RefElement& re = refElement(2, p);
    const RefElement& rl = refElement(1, p);
    ElemBlock tris, lines;
    tris.entityTag = 1; tris.entityDim = 2; tris.type = re.type;
    lines.entityTag = 1; lines.entityDim = 1; lines.type = rl.type;
    const size_t K = cells.size() * 2;
    tris.tags.resize(K);
    tris.nodeTags.resize(K * re.np);
    size_t e = 0;
    for (auto& cell : cells) {
        const int ci = cell.second[0], cj = cell.second[1];
        for (int t = 0; t < 2; ++t, ++e) {
            int64_t V[3][2];
            for (int v = 0; v < 3; ++v) { V[v][0] = (int64_t)p * (ci + tri[t][v][0]); V[v][1] = (int64_t)p * (cj + tri[t][v][1]); }
            for (int nn = 0; nn < re.np; ++nn) {
                int64_t P[2] = {0, 0};
                for (int v = 0; v < 3; ++v)
                    for (int c = 0; c < 2; ++c) P[c] += re.bary[nn][v] * V[v][c];
                tris.nodeTags[e * re.np + nn] = tagOf(P[0] / p, P[1] / p);
            }
            if (withBoundaryElements)
                for (int lf = 0; lf < 3; ++lf) {
                    bool onB = false;
                    for (int c = 0; c < 2 && !onB; ++c) {
                        const int64_t a0 = V[kTriEdges[lf][0]][c], a1 = V[kTriEdges[lf][1]][c];
                        if (a0 == a1 && (a0 == 0 || a0 == (int64_t)p * n)) onB = true;
                    }
                    if (!onB) continue;
                    const int* fn = &re.faceNodes[lf * re.nfp];
                    for (int q = 0; q < re.nfp; ++q) lines.nodeTags.push_back(tris.nodeTags[e * re.np + fn[q]]);
                }
        }
    }
    int tag = 0;
    if (withBoundaryElements) {
        lines.tags.resize(lines.nodeTags.size() / rl.np);
        for (auto& t : lines.tags) t = ++tag;
        m.blocks.push_back(std::move(lines));
    }
    for (auto& t : tris.tags) t = ++tag;
    m.maxElemTag = tag;
    m.blocks.push_back(std::move(tris));
    return m;
}

// =============================================================================================
// Jacobians (Gmsh convention, straight-sided elements)
// =============================================================================================
void affineJacobian(const Model& m, int dim, const int* v, double jac[9], double& det) {
    for (int i = 0; i < 9; ++i) jac[i] = 0.0;
    if (dim == 0) { jac[0] = jac[4] = jac[8] = 1.0; det = 1.0; return; }
    const double* x0 = m.node(v[0]);
    if (dim == 3) {
        for (int u = 0; u < 3; ++u) {
            const double* xu = m.node(v[u + 1]);
            for (int x = 0; x < 3; ++x) jac[u * 3 + x] = xu[x] - x0[x];
        }
        det = jac[0] * (jac[4] * jac[8] - jac[5] * jac[7]) - jac[1] * (jac[3] * jac[8] - jac[5] * jac[6]) +
              jac[2] * (jac[3] * jac[7] - jac[4] * jac[6]);
        return;
    }
    if (dim == 2) {
        const double *x1 = m.node(v[1]), *x2 = m.node(v[2]);
        for (int x = 0; x < 3; ++x) { jac[x] = x1[x] - x0[x]; jac[3 + x] = x2[x] - x0[x]; }
        double c[3] = {jac[1] * jac[5] - jac[2] * jac[4], jac[2] * jac[3] - jac[0] * jac[5], jac[0] * jac[4] - jac[1] * jac[3]};
        det = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int x = 0; x < 3; ++x) jac[6 + x] = c[x] / det;
        return;
    }
    // dim == 1 : tangent, then Gmsh's completion (a unit vector "to the right" of the tangent, then their cross product)
    const double* x1 = m.node(v[1]);
    double a[3];
    for (int x = 0; x < 3; ++x) { a[x] = 0.5 * (x1[x] - x0[x]); jac[x] = a[x]; }
    det = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    double b[3];
    if ((std::fabs(a[0]) >= std::fabs(a[1]) && std::fabs(a[0]) >= std::fabs(a[2])) ||
        (std::fabs(a[1]) >= std::fabs(a[0]) && std::fabs(a[1]) >= std::fabs(a[2]))) {
        b[0] = a[1]; b[1] = -a[0]; b[2] = 0.0;
    } else {
        b[0] = 0.0; b[1] = a[2]; b[2] = -a[1];
    }
    double nb = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    for (int x = 0; x < 3; ++x) b[x] /= nb;
    double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double nc = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (int x = 0; x < 3; ++x) { jac[3 + x] = b[x]; jac[6 + x] = c[x] / nc; }
}

// completion of the tangent vectors (rows 0..dim-1 of jac already set) and determinant, Gmsh's conventions as in affineJacobian
static void completeJacobian(int dim, double jac[9], double& det) {
    if (dim == 3) {
        det = jac[0] * (jac[4] * jac[8] - jac[5] * jac[7]) - jac[1] * (jac[3] * jac[8] - jac[5] * jac[6]) + jac[2] * (jac[3] * jac[7] - jac[4] * jac[6]);
        return;
    }
    if (dim == 2) {
        double c[3] = {jac[1] * jac[5] - jac[2] * jac[4], jac[2] * jac[3] - jac[0] * jac[5], jac[0] * jac[4] - jac[1] * jac[3]};
        det = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int x = 0; x < 3; ++x) jac[6 + x] = c[x] / det;
        return;
    }
    double a[3] = {jac[0], jac[1], jac[2]};
    det = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    double b[3];
    if ((std::fabs(a[0]) >= std::fabs(a[1]) && std::fabs(a[0]) >= std::fabs(a[2])) ||
        (std::fabs(a[1]) >= std::fabs(a[0]) && std::fabs(a[1]) >= std::fabs(a[2]))) {
        b[0] = a[1]; b[1] = -a[0]; b[2] = 0.0;
    } else {
        b[0] = 0.0; b[1] = a[2]; b[2] = -a[1];
    }
    double nb = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    for (int x = 0; x < 3; ++x) b[x] /= nb;
    double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    double nc = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (int x = 0; x < 3; ++x) { jac[3 + x] = b[x]; jac[6 + x] = c[x] / nc; }
}

void isoJacobian(const Model& m, int dim, int order, const int* nodeTags, const double* uvw, double jac[9], double& det) {
    for (int i = 0; i < 9; ++i) jac[i] = 0.0;
    if (dim == 0) { jac[0] = jac[4] = jac[8] = 1.0; det = 1.0; return; }
    const RefElement& re = refElement(dim, order);
    std::vector<double> dphi((size_t)re.np * 3);
    re.gradBasis(uvw, dphi.data());
    for (int n = 0; n < re.np; ++n) {
        const double* xn = m.node(nodeTags[n]);
        for (int u = 0; u < dim; ++u)
            for (int x = 0; x < 3; ++x) jac[u * 3 + x] += xn[x] * dphi[3 * n + u];
    }
    completeJacobian(dim, jac, det);
}

bool detectCurved(Model& m) {
    for (const ElemBlock& blk : m.blocks) {
        int dim, order;
        if (!elementTypeInfo(blk.type, dim, order) || order < 2 || dim < 1) continue;
        const RefElement& re = refElement(dim, order);
        const RefElement& lin = refElement(dim, 1);
        std::vector<double> phi(lin.np);
        for (size_t e = 0; e < blk.tags.size(); ++e) {
            const int* nt = &blk.nodeTags[e * re.np];
            double scale = 0;
            for (int v = 1; v < lin.np; ++v)
                for (int x = 0; x < 3; ++x) scale = std::max(scale, std::fabs(m.node(nt[v])[x] - m.node(nt[0])[x]));
            for (int n = lin.np; n < re.np; ++n) {
                lin.basis(&re.uvw[3 * n], phi.data());
                for (int x = 0; x < 3; ++x) {
                    double sx = 0;
                    for (int v = 0; v < lin.np; ++v) sx += phi[v] * m.node(nt[v])[x];
                    if (std::fabs(sx - m.node(nt[n])[x]) > 1e-10 * scale) { m.curved = true; return true; }
                }
            }
        }
    }
    return m.curved;
}

static void warpImpl(Model& m, double amp, double k, const double* center, double radius) {
    const int dim = m.dimension();
    for (int tag = 0; tag <= m.maxNodeTag; ++tag) {
        double* x = &m.xyz[3 * (size_t)tag];
        const double x0 = x[0], x1 = x[1], x2 = x[2];
        double w = 1.0;
        if (center) {
            const double r2 = ((x0 - center[0]) * (x0 - center[0]) + (x1 - center[1]) * (x1 - center[1]) + (x2 - center[2]) * (x2 - center[2])) / (radius * radius);
            w = r2 < 1.0 ? (1.0 - r2) * (1.0 - r2) : 0.0;
        }
        if (w == 0.0) continue;
        x[0] = x0 + w * (amp * std::sin(k * x1 + 0.3) * (dim >= 2 ? 1.0 : 0.0) + (dim == 3 ? 0.5 * amp * std::sin(k * x2 + 0.2) : 0.0));
        if (dim >= 2) x[1] = x1 + w * amp * std::sin(k * (dim == 3 ? x2 : x0) + 0.7);
        if (dim == 3) x[2] = x2 + w * amp * std::sin(k * x0 + 1.1);
    }
    m.curved = true;
}
void warp(Model& m, double amp, double k) { warpImpl(m, amp, k, nullptr, 0.0); }
void warpLocal(Model& m, double amp, double k, const double center[3], double radius) { warpImpl(m, amp, k, center, radius); }

}  // namespace gml
