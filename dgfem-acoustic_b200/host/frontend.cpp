// Host front end: config parser + mesh set-up with the semantics of the reference's
// src/configParser.cpp and Mesh::Mesh (src/Mesh.cpp:20-435), on top of gmshlite, with hash-based face
// de-duplication instead of the reference's O(K*F) / O(F^2) scans (same numbering).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "dgfront.h"
#include "gmshlite.h"

struct dgf_model {
    gml::Model m;
};

struct dgf_mesh {
    dgb_desc d;
    std::vector<int32_t> elTags, elNodeTags, faceNodes;
    std::vector<double> nodeCoords;
    std::vector<double> elBasisFct, elUGradBasisFct, elWeight, fBasisFct, fWeight;
    std::vector<double> elJacobian, elJacobianDet, fNormal, fJacobianDet;
    std::vector<int32_t> elFId, elFOrientation, fNbrElId, fNToElNId, fBC, fNodeTags;
    std::vector<uint8_t> fIsBoundary;
};

static thread_local std::string g_err;
extern "C" const char* dgf_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------------------------------------
// models
// ---------------------------------------------------------------------------------------------
extern "C" dgf_model* dgf_open_msh(const char* path, int order) {
    try {
        auto* mm = new dgf_model;
        mm->m = gml::readMsh(path);
        if (order > 1) gml::elevate(mm->m, order);
        else gml::detectCurved(mm->m);  // a high-order file of a curved geometry: per-integration-point geometry downstream
        return mm;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

extern "C" dgf_model* dgf_make_cube(int n, double lo, double hi, int order) {
    try {
        auto* mm = new dgf_model;
        mm->m = gml::makeCube(n, lo, hi, order);
        return mm;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

extern "C" dgf_model* dgf_make_square(int n, double lo, double hi, int order) {
    try {
        auto* mm = new dgf_model;
        mm->m = gml::makeSquare(n, lo, hi, order);
        return mm;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

extern "C" int dgf_write_msh(const dgf_model* mm, const char* path) {
    try {
        const gml::Model& m = mm->m;
        std::FILE* fp = std::fopen(path, "w");
        if (!fp) throw std::runtime_error(std::string("cannot open ") + path);
        std::fprintf(fp, "$MeshFormat\n4 0 8\n$EndMeshFormat\n$PhysicalNames\n%zu\n", m.physNames.size());
        for (auto& pn : m.physNames) std::fprintf(fp, "%d %d \"%s\"\n", pn.first.first, pn.first.second, pn.second.c_str());
        std::fprintf(fp, "$EndPhysicalNames\n$Entities\n");
        size_t cnt[4] = {0, 0, 0, 0};
        for (auto& e : m.entities) cnt[e.dim]++;
        std::fprintf(fp, "%zu %zu %zu %zu\n", cnt[0], cnt[1], cnt[2], cnt[3]);
        for (int d = 0; d < 4; ++d)
            for (auto& e : m.entities) {
                if (e.dim != d) continue;
                std::fprintf(fp, "%d 0 0 0 0 0 0 %zu", e.tag, e.phys.size());
                for (int p : e.phys) std::fprintf(fp, " %d", p);
                std::fprintf(fp, d == 0 ? "\n" : " 0\n");
            }
        std::fprintf(fp, "$EndEntities\n$Nodes\n1 %d\n", m.maxNodeTag);
        int topDim = m.dimension(), topTag = 1;
        for (auto& b : m.blocks) if (b.entityDim == topDim) topTag = b.entityTag;
        std::fprintf(fp, "%d %d 0 %d\n", topTag, topDim, m.maxNodeTag);
        for (int t = 1; t <= m.maxNodeTag; ++t) std::fprintf(fp, "%d %.17g %.17g %.17g\n", t, m.node(t)[0], m.node(t)[1], m.node(t)[2]);
        size_t ne = 0;
        for (auto& b : m.blocks) ne += b.tags.size();
        std::fprintf(fp, "$EndNodes\n$Elements\n%zu %zu\n", m.blocks.size(), ne);
        for (auto& b : m.blocks) {
            const size_t nn = b.tags.empty() ? 0 : b.nodeTags.size() / b.tags.size();
            std::fprintf(fp, "%d %d %d %zu\n", b.entityTag, b.entityDim, b.type, b.tags.size());
            for (size_t e = 0; e < b.tags.size(); ++e) {
                std::fprintf(fp, "%d", b.tags[e]);
                for (size_t k = 0; k < nn; ++k) std::fprintf(fp, " %d", b.nodeTags[e * nn + k]);
                std::fprintf(fp, " \n");
            }
        }
        std::fprintf(fp, "$EndElements\n");
        std::fclose(fp);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

extern "C" void dgf_model_free(dgf_model* m) { delete m; }
extern "C" int dgf_warp_model(dgf_model* m, double amp, double k) {
    if (!m) return -1;
    gml::warp(m->m, amp, k);
    return 0;
}
extern "C" int dgf_warp_model_local(dgf_model* m, double amp, double k, double cx, double cy, double cz, double radius) {
    if (!m || !(radius > 0)) return -1;
    const double c[3] = {cx, cy, cz};
    gml::warpLocal(m->m, amp, k, c, radius);
    return 0;
}
extern "C" int dgf_model_dimension(const dgf_model* m) { return m ? m->m.dimension() : -1; }

// ---------------------------------------------------------------------------------------------
// config (src/configParser.cpp:30-150)
// ---------------------------------------------------------------------------------------------
extern "C" void dgf_default_config(dgf_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->timeStart = 0; c->timeEnd = 1; c->timeStep = 0.1; c->timeRate = 0.1;  // configParser.h:9-12
    std::strcpy(c->elementType, "Lagrange");
    std::strcpy(c->timeIntMethod, "Euler1");
    std::strcpy(c->saveFile, "results.msh");
    c->numThreads = 1;
    c->rho0 = 1; c->c0 = 1;
    std::strcpy(c->receiverFile, "receivers.txt");
}

static std::vector<std::string> splitCsv(const std::string& s) {
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) out.push_back(tok);
    return out;
}

extern "C" int dgf_parse_config(const char* path, const dgf_model* model, dgf_config* c) {
    try {
        std::ifstream in(path);
        if (!in.is_open()) throw std::runtime_error(std::string("cannot open config file ") + path);
        dgf_default_config(c);
        std::map<std::string, std::string> kv;
        std::string line;
        while (std::getline(in, line)) {
            line.erase(std::remove_if(line.begin(), line.end(), [](unsigned char ch) { return std::isspace(ch); }), line.end());
            if (line.empty() || line[0] == '#') continue;
            auto eq = line.find('=');
            kv[line.substr(0, eq)] = line.substr(eq + 1);  // no '=': key == value == whole line, like the reference
        }
        auto num = [&](const char* k) {
            auto it = kv.find(k);
            if (it == kv.end()) throw std::runtime_error(std::string("config key missing: ") + k);
            return std::stod(it->second);
        };
        auto str = [&](const char* k) { auto it = kv.find(k); return it == kv.end() ? std::string() : it->second; };
        c->timeStart = num("timeStart");
        c->timeEnd = num("timeEnd");
        c->timeStep = num("timeStep");
        c->timeRate = num("timeRate");
        std::snprintf(c->elementType, sizeof c->elementType, "%s", str("elementType").c_str());
        std::snprintf(c->timeIntMethod, sizeof c->timeIntMethod, "%s", str("timeIntMethod").c_str());
        std::snprintf(c->saveFile, sizeof c->saveFile, "%s", str("saveFile").c_str());
        c->numThreads = (int)num("numThreads");
        if (c->numThreads == 1) c->numThreads = 0;  // configParser.cpp:58 (0 = "let OpenMP decide")
        c->v0[0] = num("v0_x"); c->v0[1] = num("v0_y"); c->v0[2] = num("v0_z");
        c->rho0 = num("rho0");
        c->c0 = num("c0");

        const double pi = M_PI;
        auto push = [&](double pole, double x, double y, double z, double size, double amp, double f, double ph, double dur) {
            if (c->nSources >= DGF_MAX_SOURCES) throw std::runtime_error("too many sources");
            double* s = c->sources[c->nSources++];
            s[0] = pole; s[1] = x; s[2] = y; s[3] = z; s[4] = size; s[5] = amp; s[6] = f; s[7] = ph; s[8] = dur;
        };
        for (auto& it : kv) {  // std::map order == the reference's iteration order
            if (it.first.rfind("source", 0) == 0) {
                auto sep = splitCsv(it.second);
                if (sep.size() < 9) throw std::runtime_error("source needs 9 comma-separated fields: " + it.first);
                double x = std::stod(sep[1]), y = std::stod(sep[2]), z = std::stod(sep[3]), size = std::stod(sep[4]);
                double amp = std::stod(sep[5]), f = std::stod(sep[6]), ph = std::stod(sep[7]), dur = std::stod(sep[8]);
                if (sep[0] == "dipole") {
                    push(1, x - size, y, z, size / 2., amp, f, ph, dur);
                    push(1, x + size, y, z, size / 2., amp, f, ph + pi, dur);
                } else if (sep[0] == "quadrupole") {
                    push(2, x - size, y, z, size / 2., amp, f, ph, dur);
                    push(2, x + size, y, z, size / 2., amp, f, ph, dur);
                    push(2, x, y - size, z, size / 2., amp, f, ph + pi, dur);
                    push(2, x, y + size, z, size / 2., amp, f, ph + pi, dur);
                } else {
                    push(0, x, y, z, size, amp, f, ph, dur);
                }
            } else if (it.first.rfind("initialCondtition", 0) == 0) {  // (sic) configParser.cpp:103
                auto sep = splitCsv(it.second);
                if (sep.size() < 6) throw std::runtime_error("initial condition needs 6 comma-separated fields: " + it.first);
                if (c->nInit >= DGF_MAX_INIT) throw std::runtime_error("too many initial conditions");
                double* q = c->initConditions[c->nInit++];
                q[0] = 0;
                for (int k = 1; k <= 5; ++k) q[k] = std::stod(sep[k]);
            } else if (it.first == "receiverFile") {
                std::snprintf(c->receiverFile, sizeof c->receiverFile, "%s", it.second.c_str());
            } else if (it.first == "receiverWav") {
                std::snprintf(c->receiverWav, sizeof c->receiverWav, "%s", it.second.c_str());
            } else if (it.first.rfind("receiver", 0) == 0) {  // not a key of the reference: it skips it
                auto sep = splitCsv(it.second);
                if (sep.size() < 3) throw std::runtime_error("receiver needs 3 comma-separated coordinates: " + it.first);
                if (c->nReceivers >= DGF_MAX_RECEIVERS) throw std::runtime_error("too many receivers");
                double* r = c->receivers[c->nReceivers++];
                for (int k = 0; k < 3; ++k) r[k] = std::stod(sep[k]);
            }
        }
        // boundary conditions by physical-group name (configParser.cpp:117-132)
        if (model) {
            const int bcDim = model->m.dimension() - 1;
            for (int tag : model->m.physicalGroups(bcDim)) {
                std::string val = str(model->m.physicalName(bcDim, tag).c_str());
                int type = val.rfind("Absorbing", 0) == 0 ? 0 : val.rfind("Reflecting", 0) == 0 ? 1 : -1;
                if (type < 0) continue;  // "Not specified or supported boundary conditions."
                if (c->nPhysBC >= DGF_MAX_PHYS) throw std::runtime_error("too many physical groups");
                c->physBCTag[c->nPhysBC] = tag;
                c->physBCType[c->nPhysBC++] = type;
            }
        }
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// ---------------------------------------------------------------------------------------------
// Mesh::Mesh
// ---------------------------------------------------------------------------------------------
namespace {
struct FaceKey {
    int a, b, c;
    bool operator==(const FaceKey& o) const { return a == o.a && b == o.b && c == o.c; }
};
struct FaceKeyHash {
    size_t operator()(const FaceKey& k) const {
        uint64_t h = (uint64_t)(uint32_t)k.a * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)(uint32_t)k.b + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        h ^= ((uint64_t)(uint32_t)k.c + 0x165667B1ull) * 0x9E3779B185EBCA87ull;
        return (size_t)(h ^ (h >> 29));
    }
};
inline void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// Solve A g = rhs for the dim x dim system with A(r,c) = jac[r*3+c] (see Mesh.cpp:40-72: J^T grad = ugrad)
void solveJT(const double* jac, int dim, const double* rhs, double* g) {
    g[0] = g[1] = g[2] = 0.0;
    if (dim == 1) { g[0] = rhs[0] / jac[0]; return; }
    if (dim == 2) {
        double det = jac[0] * jac[4] - jac[1] * jac[3];
        g[0] = (rhs[0] * jac[4] - jac[1] * rhs[1]) / det;
        g[1] = (jac[0] * rhs[1] - rhs[0] * jac[3]) / det;
        return;
    }
    double A[3][4];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) A[r][c] = jac[r * 3 + c]; A[r][3] = rhs[r]; }
    for (int c = 0; c < 3; ++c) {  // partial pivoting, like Eigen's PartialPivLU
        int piv = c;
        for (int r = c + 1; r < 3; ++r) if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        if (piv != c) for (int k = 0; k < 4; ++k) std::swap(A[c][k], A[piv][k]);
        for (int r = c + 1; r < 3; ++r) {
            double f = A[r][c] / A[c][c];
            for (int k = c; k < 4; ++k) A[r][k] -= f * A[c][k];
        }
    }
    for (int r = 2; r >= 0; --r) {
        double s = A[r][3];
        for (int c = r + 1; c < 3; ++c) s -= A[r][c] * g[c];
        g[r] = s / A[r][r];
    }
}
}  // namespace

extern "C" dgf_mesh* dgf_mesh_build(dgf_model* model, const dgf_config* cfg) {
    try {
        if (!model || !cfg) throw std::runtime_error("null argument");
        gml::Model& gm = model->m;
        auto* M = new dgf_mesh;
        std::unique_ptr<dgf_mesh> guard(M);
        dgb_desc& d = M->d;
        std::memset(&d, 0, sizeof d);

        // ---- elements (Mesh.cpp:25-38) ----
        const int dim = gm.dimension();
        if (dim < 1 || dim > 3) throw std::runtime_error("model has no 1D/2D/3D elements");
        auto types = gm.elementTypes(dim);
        int tdim, order;
        if (types.empty() || !gml::elementTypeInfo(types[0], tdim, order)) throw std::runtime_error("no usable element type");
        const gml::RefElement& re = gml::refElement(dim, order);
        const int Np = re.np;
        {
            std::vector<int> tags, nodes;
            gm.elementsByType(types[0], tags, nodes);
            M->elTags.assign(tags.begin(), tags.end());
            M->elNodeTags.assign(nodes.begin(), nodes.end());
        }
        const int K = (int)M->elTags.size();
        if (gm.curved) {
            // a curved model: straight-sided elements first, curved ones last (stable), so that the engine can run its collapsed
            // kernels on a prefix and the curved-element kernel on the suffix. An element is straight-sided if every node sits
            // where the affine map of its vertices puts it. (A fully warped model keeps its order: all elements are curved.)
            const gml::RefElement& lin = gml::refElement(dim, 1);
            std::vector<double> phi(lin.np);
            std::vector<int> straight, bent;
            for (int el = 0; el < K; ++el) {
                const int* nt = &M->elNodeTags[(size_t)el * Np];
                double scale = 0, dev = 0;
                for (int v = 1; v < lin.np; ++v)
                    for (int x = 0; x < 3; ++x) scale = std::max(scale, std::fabs(gm.node(nt[v])[x] - gm.node(nt[0])[x]));
                for (int n = lin.np; n < Np; ++n) {
                    lin.basis(&re.uvw[3 * n], phi.data());
                    for (int x = 0; x < 3; ++x) {
                        double sx = 0;
                        for (int v = 0; v < lin.np; ++v) sx += phi[v] * gm.node(nt[v])[x];
                        dev = std::max(dev, std::fabs(sx - gm.node(nt[n])[x]));
                    }
                }
                (dev > 1e-12 * scale ? bent : straight).push_back(el);
            }
            if (!bent.empty() && !straight.empty()) {
                std::vector<int32_t> tags2, nodes2;
                straight.insert(straight.end(), bent.begin(), bent.end());
                for (int el : straight) {
                    tags2.push_back(M->elTags[el]);
                    nodes2.insert(nodes2.end(), M->elNodeTags.begin() + (size_t)el * Np, M->elNodeTags.begin() + (size_t)(el + 1) * Np);
                }
                M->elTags.swap(tags2);
                M->elNodeTags.swap(nodes2);
            }
        }
        const gml::Quadrature& q = gml::gaussRule(dim, 2 * order);  // "Gauss" + 2*order (Mesh.cpp:31)
        const int nG = q.n;
        M->elBasisFct.resize((size_t)nG * Np);
        M->elUGradBasisFct.resize((size_t)nG * Np * 3);
        M->elWeight.resize(nG);
        for (int g = 0; g < nG; ++g) {
            re.basis(&q.pts[4 * g], &M->elBasisFct[(size_t)g * Np]);
            re.gradBasis(&q.pts[4 * g], &M->elUGradBasisFct[(size_t)g * Np * 3]);
            M->elWeight[g] = q.pts[4 * g + 3];
        }
        // straight-sided meshes: one Jacobian per element / face (compressed, nGeom = 1); curved ones: one per
        // integration point, the reference's own layout (Mesh.h:37-42, 52-54, 70, 88)
        const bool curved = gm.curved;
        const int gE = curved ? nG : 1;
        M->elJacobian.resize((size_t)K * gE * 9);
        M->elJacobianDet.resize((size_t)K * gE);
        for (int el = 0; el < K; ++el) {
            if (!curved) { gml::affineJacobian(gm, dim, &M->elNodeTags[(size_t)el * Np], &M->elJacobian[(size_t)el * 9], M->elJacobianDet[el]); continue; }
            for (int g = 0; g < nG; ++g)
                gml::isoJacobian(gm, dim, order, &M->elNodeTags[(size_t)el * Np], &q.pts[4 * g], &M->elJacobian[((size_t)el * nG + g) * 9],
                                 M->elJacobianDet[(size_t)el * nG + g]);
        }
        M->nodeCoords.resize((size_t)K * Np * 3);
        for (size_t n = 0; n < (size_t)K * Np; ++n) std::copy(gm.node(M->elNodeTags[n]), gm.node(M->elNodeTags[n]) + 3, &M->nodeCoords[3 * n]);

        // ---- faces (Mesh.cpp:92-128, 679-713) ----
        const int fDim = dim - 1;
        const int Nfp = re.nfp, Nf = re.nFaces;
        if (dim == 1 && order != 1)
            throw std::runtime_error("1D meshes are supported at order 1 only (reference quirk, SURVEY Q8)");
        M->faceNodes.assign(re.faceNodes.begin(), re.faceNodes.end());
        const int nKey = fDim + 1;  // a face is identified by its vertices
        std::unordered_map<FaceKey, int, FaceKeyHash> faceOf;
        faceOf.reserve((size_t)K * Nf / 2 + 16);
        M->elFId.resize((size_t)K * Nf);
        std::vector<int32_t> fOwnerLf;  // [F][2] local face index inside each owner
        for (int el = 0; el < K; ++el)
            for (int lf = 0; lf < Nf; ++lf) {
                int v[3] = {0, 0, 0};
                for (int k = 0; k < nKey; ++k) v[k] = M->elNodeTags[(size_t)el * Np + re.faceNodes[lf * Nfp + k]];
                std::sort(v, v + nKey);
                FaceKey key{v[0], v[1], v[2]};
                auto it = faceOf.find(key);
                int f;
                if (it == faceOf.end()) {
                    f = (int)(M->fNbrElId.size() / 2);
                    faceOf.emplace(key, f);
                    M->fNbrElId.push_back(el);
                    M->fNbrElId.push_back(-1);
                    fOwnerLf.push_back(lf);
                    fOwnerLf.push_back(-1);
                    for (int k = 0; k < Nfp; ++k) M->fNodeTags.push_back(M->elNodeTags[(size_t)el * Np + re.faceNodes[lf * Nfp + k]]);
                } else {
                    f = it->second;
                    if (M->fNbrElId[2 * (size_t)f + 1] >= 0) throw std::runtime_error("non-manifold mesh: a face has more than two owners");
                    M->fNbrElId[2 * (size_t)f + 1] = el;
                    fOwnerLf[2 * (size_t)f + 1] = lf;
                }
                M->elFId[(size_t)el * Nf + lf] = f;
            }
        const int F = (int)(M->fNbrElId.size() / 2);
        { std::unordered_map<FaceKey, int, FaceKeyHash>().swap(faceOf); }

        // ---- face basis, Jacobians, normals (Mesh.cpp:135-194) ----
        const gml::RefElement& rf = gml::refElement(fDim, order);
        const gml::Quadrature qf = gml::integrationRule(fDim, 2 * order, order, dim == 3);  // see gmshlite.h
        const int nGf = qf.n;
        const int fc = (dim == 3 && order != 1) ? -1 : 1;  // Mesh.cpp:210-211
        M->fBasisFct.resize((size_t)nGf * Nfp);
        M->fWeight.resize(nGf);
        std::vector<double> fUGrad0((size_t)Nfp * 3);  // parametric gradients at the first integration point
        for (int g = 0; g < nGf; ++g) {
            rf.basis(&qf.pts[4 * g], &M->fBasisFct[(size_t)g * Nfp]);
            M->fWeight[g] = qf.pts[4 * g + 3];
        }
        if (fDim > 0) rf.gradBasis(&qf.pts[0], fUGrad0.data());
        const int gF = curved ? nGf : 1;
        M->fNormal.resize((size_t)F * gF * 3);
        M->fJacobianDet.resize((size_t)F * gF);
        std::vector<double> fUGrad((size_t)Nfp * 3);
        for (int f = 0; f < F; ++f) {
            double first0[3] = {0, 0, 0};  // physical gradient of face basis function 0 at the first point (Mesh.cpp:176)
            for (int g = 0; g < gF; ++g) {
                double jac[9], det, n[3] = {1, 0, 0};
                if (!curved) gml::affineJacobian(gm, fDim, &M->fNodeTags[(size_t)f * Nfp], jac, det);
                else {
                    gml::isoJacobian(gm, fDim, order, &M->fNodeTags[(size_t)f * Nfp], &qf.pts[4 * g], jac, det);
                    if (fDim > 0) rf.gradBasis(&qf.pts[4 * g], fUGrad.data());
                }
                const double* ug = curved ? fUGrad.data() : fUGrad0.data();
                M->fJacobianDet[(size_t)f * gF + g] = det;
                if (fDim == 1) {
                    double g0[3], zdir[3] = {0, 0, -1};
                    solveJT(jac, dim, &ug[0], g0);
                    cross3(g0, zdir, n);  // Mesh.cpp:174-175
                    if (g == 0) std::copy(g0, g0 + 3, first0);
                    else if (dot3(g0, first0) < 0) for (int x = 0; x < 3; ++x) n[x] = -n[x];  // Mesh.cpp:176-179
                } else if (fDim == 2) {
                    double g0[3], g1[3];
                    solveJT(jac, dim, &ug[0], g0);
                    solveJT(jac, dim, &ug[3], g1);
                    cross3(g0, g1, n);  // Mesh.cpp:183
                    if (g != 0 && dot3(&M->fNormal[(size_t)f * gF * 3], n) < 0) for (int x = 0; x < 3; ++x) n[x] = -n[x];  // Mesh.cpp:184-187
                }
                double nn = std::sqrt(dot3(n, n));
                for (int x = 0; x < 3; ++x) M->fNormal[((size_t)f * gF + g) * 3 + x] = n[x] / nn;
            }
        }

        // ---- face node -> element node maps (Mesh.cpp:253-263) ----
        M->fNToElNId.assign((size_t)F * Nfp * 2, -1);
        for (int f = 0; f < F; ++f)
            for (int side = 0; side < 2; ++side) {
                const int el = M->fNbrElId[2 * (size_t)f + side];
                if (el < 0) continue;
                const int lf = fOwnerLf[2 * (size_t)f + side];
                for (int nf = 0; nf < Nfp; ++nf) {
                    const int tag = M->fNodeTags[(size_t)f * Nfp + nf];
                    int found = -1;
                    for (int k = 0; k < Nfp && found < 0; ++k) {
                        const int ln = re.faceNodes[lf * Nfp + k];
                        if (M->elNodeTags[(size_t)el * Np + ln] == tag) found = ln;
                    }
                    if (found < 0) throw std::runtime_error("face/element node mismatch (non-conforming high-order nodes?)");
                    M->fNToElNId[((size_t)f * Nfp + nf) * 2 + side] = found;
                }
            }

        // ---- orientation (Mesh.cpp:278-294) ----
        M->elFOrientation.resize((size_t)K * Nf);
        const int nv = dim + 1;
        for (int el = 0; el < K; ++el) {
            double bary[3] = {0, 0, 0};
            for (int v = 0; v < nv; ++v)
                for (int x = 0; x < 3; ++x) bary[x] += gm.node(M->elNodeTags[(size_t)el * Np + v])[x];
            for (int x = 0; x < 3; ++x) bary[x] /= nv;
            for (int lf = 0; lf < Nf; ++lf) {
                const double* xn = gm.node(M->elNodeTags[(size_t)el * Np + re.faceNodes[lf * Nfp]]);
                const double* nf = &M->fNormal[(size_t)M->elFId[(size_t)el * Nf + lf] * gF * 3];  // fNormal(f, 0), Mesh.cpp:287
                double dp = 0.0;
                for (int x = 0; x < 3; ++x) dp += (xn[x] - bary[x]) * nf[x];
                M->elFOrientation[(size_t)el * Nf + lf] = dp >= 0 ? 1 : -1;
            }
        }
        // (Mesh.cpp:301-314 is a no-op on every mesh with more than two faces, SURVEY Q2: up stays the lower element index.)

        // ---- boundary faces: outward normal, orientation := 1 (Mesh.cpp:336-353) ----
        M->fIsBoundary.assign(F, 0);
        for (int f = 0; f < F; ++f) {
            if (M->fNbrElId[2 * (size_t)f + 1] >= 0) continue;
            M->fIsBoundary[f] = 1;
            const int el = M->fNbrElId[2 * (size_t)f];
            for (int lf = 0; lf < Nf; ++lf)
                if (M->elFId[(size_t)el * Nf + lf] == f) {
                    const int o = M->elFOrientation[(size_t)el * Nf + lf];
                    for (int x = 0; x < 3 * gF; ++x) M->fNormal[(size_t)f * gF * 3 + x] *= o;  // every integration point, Mesh.cpp:341-345
                    M->elFOrientation[(size_t)el * Nf + lf] = 1;
                }
        }

        // ---- boundary-condition tags by the face's first node (Mesh.cpp:364-384) ----
        M->fBC.assign(F, 0);
        for (int b = 0; b < cfg->nPhysBC; ++b) {
            std::vector<int> nodes;
            gm.nodesForPhysicalGroup(fDim, cfg->physBCTag[b], nodes);
            std::vector<char> in(gm.maxNodeTag + 1, 0);
            for (int t : nodes) in[t] = 1;
            const int val = cfg->physBCType[b] == 1 ? 1 : 0;
            for (int f = 0; f < F; ++f)
                if (M->fIsBoundary[f] && in[M->fNodeTags[(size_t)f * Nfp]]) M->fBC[f] = val;
        }

        d.dim = dim; d.order = order; d.Np = Np; d.Nfp = Nfp; d.Nf = Nf; d.K = K; d.F = F;
        d.nG = nG; d.nGf = nGf; d.nGeomEl = gE; d.nGeomF = gF; d.fc = fc;
        d.elBasisFct = M->elBasisFct.data(); d.elUGradBasisFct = M->elUGradBasisFct.data(); d.elWeight = M->elWeight.data();
        d.fBasisFct = M->fBasisFct.data(); d.fWeight = M->fWeight.data();
        d.elJacobian = M->elJacobian.data(); d.elJacobianDet = M->elJacobianDet.data();
        d.fNormal = M->fNormal.data(); d.fJacobianDet = M->fJacobianDet.data();
        d.elFId = M->elFId.data(); d.elFOrientation = M->elFOrientation.data();
        d.fNbrElId = M->fNbrElId.data(); d.fNToElNId = M->fNToElNId.data();
        d.fIsBoundary = M->fIsBoundary.data(); d.fBC = M->fBC.data();
        d.c0 = cfg->c0; d.rho0 = cfg->rho0; d.v0[0] = cfg->v0[0]; d.v0[1] = cfg->v0[1]; d.v0[2] = cfg->v0[2];
        d.dt = cfg->timeStep;
        return guard.release();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

extern "C" void dgf_mesh_free(dgf_mesh* m) { delete m; }
extern "C" const dgb_desc* dgf_mesh_desc(const dgf_mesh* m) { return &m->d; }
extern "C" const double* dgf_mesh_node_coords(const dgf_mesh* m) { return m->nodeCoords.data(); }
extern "C" const int32_t* dgf_mesh_el_tags(const dgf_mesh* m) { return m->elTags.data(); }
extern "C" const int32_t* dgf_mesh_el_node_tags(const dgf_mesh* m) { return m->elNodeTags.data(); }
extern "C" const int32_t* dgf_mesh_face_nodes(const dgf_mesh* m) { return m->faceNodes.data(); }

// ---------------------------------------------------------------------------------------------
// initial condition / sources / loop header
// ---------------------------------------------------------------------------------------------
extern "C" void dgf_initial_condition(const dgf_mesh* mesh, const dgf_config* cfg, double* u) {
    const size_t N = (size_t)mesh->d.K * mesh->d.Np;
    std::fill(u, u + 4 * N, 0.0);
    for (int i = 0; i < cfg->nInit; ++i) {
        const double x = cfg->initConditions[i][1], y = cfg->initConditions[i][2], z = cfg->initConditions[i][3];
        const double size = cfg->initConditions[i][4], amp = cfg->initConditions[i][5];
        for (size_t n = 0; n < N; ++n) {
            const double* c = &mesh->nodeCoords[3 * n];
            u[n] += amp * std::exp(-((c[0] - x) * (c[0] - x) + (c[1] - y) * (c[1] - y) + (c[2] - z) * (c[2] - z)) / size);
        }
    }
}

extern "C" int dgf_source_nodes(const dgf_mesh* mesh, const dgf_config* cfg, int32_t* offsets, int32_t* nodeIdx) {
    const size_t N = (size_t)mesh->d.K * mesh->d.Np;
    int total = 0;
    for (int s = 0; s < cfg->nSources; ++s) {
        if (offsets) offsets[s] = total;
        const double* S = cfg->sources[s];
        for (size_t n = 0; n < N; ++n) {
            const double* c = &mesh->nodeCoords[3 * n];
            if (std::pow(c[0] - S[1], 2) + std::pow(c[1] - S[2], 2) + std::pow(c[2] - S[3], 2) < std::pow(S[4], 2)) {
                if (nodeIdx) nodeIdx[total] = (int32_t)n;
                ++total;
            }
        }
    }
    if (offsets) offsets[cfg->nSources] = total;
    return total;
}

extern "C" int dgf_time_loop(const dgf_config* cfg, int32_t* snapshotSteps, int capacity, int* nSnapshots) {
    int nsnap = 0, steps = 0;
    double step = 0, tDisplay = 0;
    for (double t = cfg->timeStart; t <= cfg->timeEnd; t += cfg->timeStep, tDisplay += cfg->timeStep, ++step) {
        if (tDisplay >= cfg->timeRate || step == 0) {
            tDisplay = 0;
            if (snapshotSteps && nsnap < capacity) snapshotSteps[nsnap] = (int32_t)step;
            ++nsnap;
        }
        ++steps;
    }
    if (nSnapshots) *nSnapshots = nsnap;
    return steps;
}

extern "C" int dgf_nearest_node(const dgf_mesh* mesh, double x, double y, double z) {
    const size_t N = (size_t)mesh->d.K * mesh->d.Np;
    double best = 1e300;
    int arg = -1;
    for (size_t n = 0; n < N; ++n) {
        const double* c = &mesh->nodeCoords[3 * n];
        double r = (c[0] - x) * (c[0] - x) + (c[1] - y) * (c[1] - y) + (c[2] - z) * (c[2] - z);
        if (r < best) { best = r; arg = (int)n; }
    }
    return arg;
}

// ---------------------------------------------------------------------------------------------
// receivers: point location + Lagrange interpolation weights, time-series writers
// ---------------------------------------------------------------------------------------------
extern "C" int dgf_locate_point(const dgf_mesh* mesh, double x, double y, double z, double* weights, double* uvwOut, int* outside) {
    try {
        if (!mesh || !weights) throw std::runtime_error("dgf_locate_point: null argument");
        const dgb_desc& d = mesh->d;
        const int dim = d.dim, Np = d.Np;
        const gml::RefElement& ref = gml::refElement(dim, d.order);
        const double X[3] = {x, y, z};
        // du = argmin | r - sum_a du_a J_a | for the dim tangent vectors J_a (rows of 3): normal equations, Cramer
        auto lsq = [dim](const double* J, const double* r, double* du) {
            double G[3][3], b[3];
            for (int a = 0; a < dim; ++a) {
                b[a] = dot3(&J[a * 3], r);
                for (int c = 0; c < dim; ++c) G[a][c] = dot3(&J[a * 3], &J[c * 3]);
            }
            if (dim == 1) du[0] = b[0] / G[0][0];
            else if (dim == 2) {
                const double det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
                du[0] = (b[0] * G[1][1] - G[0][1] * b[1]) / det;
                du[1] = (G[0][0] * b[1] - b[0] * G[1][0]) / det;
            } else {
                const double det = G[0][0] * (G[1][1] * G[2][2] - G[1][2] * G[2][1]) - G[0][1] * (G[1][0] * G[2][2] - G[1][2] * G[2][0]) +
                                   G[0][2] * (G[1][0] * G[2][1] - G[1][1] * G[2][0]);
                for (int k = 0; k < 3; ++k) {
                    double A[3][3];
                    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) A[a][c] = c == k ? b[a] : G[a][c];
                    du[k] = (A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                             A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0])) / det;
                }
            }
        };
        // how far outside the reference element: line on [-1,1], unit triangle / tetrahedron (0 inside)
        auto violation = [dim](const double* u) {
            double lmin;
            if (dim == 1) lmin = std::min(0.5 * (1 - u[0]), 0.5 * (1 + u[0]));
            else {
                double l0 = 1;
                lmin = 1e300;
                for (int a = 0; a < dim; ++a) { l0 -= u[a]; lmin = std::min(lmin, u[a]); }
                lmin = std::min(lmin, l0);
            }
            return lmin >= -1e-12 ? 0.0 : -lmin;
        };
        const bool curved = d.nGeomEl > 1;  // one Jacobian per integration point: the elements may be curved
        std::vector<double> phi(Np), dphi((size_t)Np * 3);
        int best = -1;
        double bestViol = 1e300, bestU[3] = {0, 0, 0};
        for (int el = 0; el < d.K; ++el) {
            const double* xn = &mesh->nodeCoords[(size_t)el * Np * 3];
            double u[3] = {0, 0, 0}, J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            // straight-sided map through the vertices (the first dim+1 nodes): x = x_0 + sum_a (u_a - u0_a) J_a
            for (int a = 0; a < dim; ++a)
                for (int c = 0; c < 3; ++c) J[a * 3 + c] = (xn[3 * (a + 1) + c] - xn[c]) * (dim == 1 ? 0.5 : 1.0);
            const double r0[3] = {X[0] - xn[0], X[1] - xn[1], X[2] - xn[2]};
            double du[3] = {0, 0, 0};
            lsq(J, r0, du);
            for (int a = 0; a < dim; ++a) u[a] = ref.uvw[a] + du[a];
            double viol = violation(u);
            if (curved && viol < 0.5) {
                // Newton on the isoparametric map x(u) = sum_n phi_n(u) x_n from the straight-sided guess
                for (int it = 0; it < 20; ++it) {
                    ref.basis(u, phi.data());
                    ref.gradBasis(u, dphi.data());
                    double r[3] = {X[0], X[1], X[2]};
                    for (int k = 0; k < 9; ++k) J[k] = 0.0;
                    for (int n = 0; n < Np; ++n)
                        for (int c = 0; c < 3; ++c) {
                            r[c] -= phi[n] * xn[3 * n + c];
                            for (int a = 0; a < dim; ++a) J[a * 3 + c] += dphi[3 * n + a] * xn[3 * n + c];
                        }
                    lsq(J, r, du);
                    double step = 0;
                    for (int a = 0; a < dim; ++a) { u[a] += du[a]; step = std::max(step, std::fabs(du[a])); }
                    if (step < 1e-14) break;
                }
                viol = violation(u);
            }
            if (viol < bestViol) {
                bestViol = viol; best = el;
                std::copy(u, u + 3, bestU);
                if (viol == 0.0) break;  // ascending scan: the lowest element id that contains the point
            }
        }
        if (best < 0) throw std::runtime_error("dgf_locate_point: empty mesh");
        ref.basis(bestU, weights);
        if (uvwOut) std::copy(bestU, bestU + 3, uvwOut);
        if (outside) *outside = bestViol > 0.0;
        return best;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

extern "C" int dgf_write_receivers(const char* path, int nrecv, const double* xyz, int nsteps, double tStart, double dt, const double* rec) {
    try {
        if (!path || nrecv < 0 || nsteps < 0 || (nrecv > 0 && (!xyz || (nsteps > 0 && !rec)))) throw std::runtime_error("dgf_write_receivers: bad arguments");
        std::FILE* fp = std::fopen(path, "w");
        if (!fp) throw std::runtime_error(std::string("cannot open ") + path);
        std::fprintf(fp, "# receivers: %d, steps: %d, columns: t then (p vx vy vz) per receiver\n", nrecv, nsteps);
        for (int j = 0; j < nrecv; ++j) std::fprintf(fp, "# receiver %d at %.17g %.17g %.17g\n", j, xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]);
        double t = tStart;
        for (int k = 0; k < nsteps; ++k, t += dt) {  // the same accumulation as the loop header, solver.cpp:216
            std::fprintf(fp, "%.17g", t);
            for (int j = 0; j < nrecv; ++j)
                for (int q = 0; q < 4; ++q) std::fprintf(fp, " %.17g", rec[((size_t)k * nrecv + j) * 4 + q]);
            std::fprintf(fp, "\n");
        }
        std::fclose(fp);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

extern "C" int dgf_write_wav(const char* path, int nrecv, int receiver, int field, int nsteps, double dt, int rate, const double* rec) {
    try {
        if (!path || !rec || receiver < 0 || receiver >= nrecv || field < 0 || field > 3 || nsteps < 1 || !(dt > 0 || rate > 0))
            throw std::runtime_error("dgf_write_wav: bad arguments");
        const uint32_t sr = rate > 0 ? (uint32_t)rate : (uint32_t)std::llround(1.0 / dt);
        double peak = 0;
        for (int k = 0; k < nsteps; ++k) peak = std::max(peak, std::fabs(rec[((size_t)k * nrecv + receiver) * 4 + field]));
        std::vector<int16_t> pcm(nsteps);
        for (int k = 0; k < nsteps; ++k)
            pcm[k] = peak > 0 ? (int16_t)std::lround(32767.0 * rec[((size_t)k * nrecv + receiver) * 4 + field] / peak) : (int16_t)0;
        std::FILE* fp = std::fopen(path, "wb");
        if (!fp) throw std::runtime_error(std::string("cannot open ") + path);
        const uint32_t dataBytes = (uint32_t)nsteps * 2, riff = 36 + dataBytes, fmtLen = 16, byteRate = sr * 2;
        const uint16_t pcmTag = 1, channels = 1, blockAlign = 2, bits = 16;
        std::fwrite("RIFF", 1, 4, fp); std::fwrite(&riff, 4, 1, fp); std::fwrite("WAVEfmt ", 1, 8, fp);
        std::fwrite(&fmtLen, 4, 1, fp); std::fwrite(&pcmTag, 2, 1, fp); std::fwrite(&channels, 2, 1, fp);
        std::fwrite(&sr, 4, 1, fp); std::fwrite(&byteRate, 4, 1, fp); std::fwrite(&blockAlign, 2, 1, fp); std::fwrite(&bits, 2, 1, fp);
        std::fwrite("data", 1, 4, fp); std::fwrite(&dataBytes, 4, 1, fp);
        std::fwrite(pcm.data(), 2, pcm.size(), fp);
        std::fclose(fp);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// ---------------------------------------------------------------------------------------------
// Gmsh-compatible post-processing output ($ElementNodeData, MSH 4.0 ASCII)
// ---------------------------------------------------------------------------------------------
extern "C" int dgf_write_views(const char* path, const dgf_model* model, const dgf_mesh* mesh, const dgf_config* cfg,
                               int nSnap, const int32_t* snapStep, const double* snapTime, const double* snapU) {
    try {
        const dgb_desc& d = mesh->d;
        const size_t N = (size_t)d.K * d.Np;
        std::FILE* fp = std::fopen(path, "a");
        if (!fp) throw std::runtime_error(std::string("cannot open ") + path);
        std::fprintf(fp, "$MeshFormat\n4 0 8\n$EndMeshFormat\n");
        const char* names[3] = {"Pressure", "Density", "Velocity"};
        const int comps[3] = {1, 1, 3};
        for (int view = 0; view < 3; ++view)
            for (int s = 0; s < nSnap; ++s) {
                const double* U = snapU + (size_t)s * 4 * N;
                std::fprintf(fp, "$ElementNodeData\n1\n\"%s\"\n1\n%.16g\n3\n%d\n%d\n%d\n", names[view], snapTime[s], snapStep[s],
                             comps[view], d.K);
                for (int el = 0; el < d.K; ++el) {
                    std::fprintf(fp, "%d %d", mesh->elTags[el], d.Np);
                    for (int n = 0; n < d.Np; ++n) {
                        const size_t i = (size_t)el * d.Np + n;
                        if (view == 0) std::fprintf(fp, " %.16g", U[i]);
                        else if (view == 1) std::fprintf(fp, " %.16g", U[i] / (cfg->c0 * cfg->c0));  // solver.cpp:230
                        else std::fprintf(fp, " %.16g %.16g %.16g", U[N + i], U[2 * N + i], U[3 * N + i]);
                    }
                    std::fprintf(fp, "\n");
                }
                std::fprintf(fp, "$EndElementNodeData\n");
            }
        std::fclose(fp);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
