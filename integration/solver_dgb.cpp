// solver_dgb.cpp — drop-in replacement for the reference's src/solver.cpp: same namespace, same two signatures
// (include/solver.h:15,24), the time loops of src/solver.cpp:61-161 and 171-292 re-expressed as calls into the C ABI of
// include/dgb.h. Everything above this file (src/dgalerkin.cpp, src/configParser.cpp, the Gmsh-based Mesh constructor of
// src/Mesh.cpp) is the reference's own code, unmodified.
//
// Build (oracle/Makefile, target _ref/dgalerkin_dgb): the reference's dgalerkin.cpp, Mesh.cpp, configParser.cpp, utils.cpp +
// this file + -ldgb. A maintainer of the reference would replace solver.cpp by this file in src/CMakeLists.txt and add `dgb`
// to TARGET_LINK_LIBRARIES.
//
// Mesh keeps the arrays the hot path needs private (include/Mesh.h:139-223); a maintainer would add ten trivial getters.
// This file reads them directly instead (the `private` keyword is lifted for the one include), so that the reference's
// headers stay untouched.
#include <gmsh.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define private public
#include "Mesh.h"
#undef private
#include "configParser.h"
#include "dgb.h"
#include "solver.h"

namespace solver {

namespace {

void fail(const char* what) {
    gmsh::logger::write(std::string(what) + ": " + dgb_last_error(), "error");
    std::fprintf(stderr, "Error   : %s: %s\n", what, dgb_last_error());
    std::exit(EXIT_FAILURE);
}

// Flat copies of what the reference's Mesh computed: the ONLY data preparation of the binding.
dgb_handle* upload(Mesh& mesh, const Config& config) {
    dgb_desc d{};
    d.dim = mesh.m_elDim;      d.order = mesh.m_elOrder;   d.Np = mesh.m_elNumNodes;
    d.Nfp = mesh.m_fNumNodes;  d.Nf = mesh.m_fNumPerEl;    d.K = mesh.m_elNum;       d.F = mesh.m_fNum;
    d.nG = mesh.m_elNumIntPts; d.nGf = mesh.m_fNumIntPts;
    d.nGeomEl = d.nG;          d.nGeomF = d.nGf;  // the reference stores one Jacobian / normal per integration point
    d.fc = mesh.fc;
    std::vector<double> wEl(d.nG), wF(d.nGf);     // m_*IntParamCoords are (u, v, w, weight) per point
    for (int g = 0; g < d.nG; ++g) wEl[g] = mesh.elWeight(g);
    for (int g = 0; g < d.nGf; ++g) wF[g] = mesh.fWeight(g);
    d.elBasisFct = mesh.m_elBasisFcts.data();  d.elUGradBasisFct = mesh.m_elUGradBasisFcts.data();  d.elWeight = wEl.data();
    d.fBasisFct = mesh.m_fBasisFcts.data();    d.fWeight = wF.data();
    d.elJacobian = mesh.m_elJacobians.data();  d.elJacobianDet = mesh.m_elJacobianDets.data();
    d.fNormal = mesh.m_fNormals.data();        d.fJacobianDet = mesh.m_fJacobianDets.data();
    d.elFId = mesh.m_elFIds.data();            d.elFOrientation = mesh.m_elFOrientation.data();
    // vector<vector<int>> -> flat [F][2] / [F][Nfp][2] with -1 where a boundary face has no second owner
    std::vector<int32_t> nbr(2 * (size_t)d.F, -1), map(2 * (size_t)d.F * d.Nfp, -1);
    std::vector<uint8_t> isB(d.F);
    for (int f = 0; f < d.F; ++f) {
        const int n = (int)mesh.m_fNbrElIds[f].size();
        for (int s = 0; s < n; ++s) nbr[2 * (size_t)f + s] = mesh.m_fNbrElIds[f][s];
        for (int i = 0; i < d.Nfp; ++i)
            for (int s = 0; s < n; ++s) map[((size_t)f * d.Nfp + i) * 2 + s] = mesh.m_fNToElNIds[f][i * n + s];
        isB[f] = mesh.m_fIsBoundary[f] ? 1 : 0;
    }
    d.fNbrElId = nbr.data();  d.fNToElNId = map.data();  d.fIsBoundary = isB.data();  d.fBC = mesh.m_fBC.data();
    d.c0 = config.c0;  d.rho0 = config.rho0;
    for (int x = 0; x < 3; ++x) d.v0[x] = config.v0[x];
    d.dt = config.timeStep;
    dgb_handle* h = nullptr;
    if (dgb_create(&d, &h) != DGB_OK) fail("dgb_create");
    return h;  // everything was copied: the temporaries may go out of scope
}

void march(std::vector<std::vector<double>>& u, Mesh& mesh, Config config, int integrator) {
    const int elNumNodes = mesh.getElNumNodes();
    const size_t N = (size_t)mesh.getNumNodes();
    std::vector<int> elTags(&mesh.elTag(0), &mesh.elTag(0) + mesh.getElNum());

    /** Gmsh save init (src/solver.cpp:185-191) */
    std::vector<std::string> g_names;
    gmsh::model::list(g_names);
    int gp_viewTag = gmsh::view::add("Pressure");
    int gv_viewTag = gmsh::view::add("Velocity");
    int grho_viewTag = gmsh::view::add("Density");
    std::vector<std::vector<double>> g_p(mesh.getElNum(), std::vector<double>(elNumNodes));
    std::vector<std::vector<double>> g_rho(mesh.getElNum(), std::vector<double>(elNumNodes));
    std::vector<std::vector<double>> g_v(mesh.getElNum(), std::vector<double>(3 * elNumNodes));

    dgb_handle* h = upload(mesh, config);  // replaces mesh.precomputeMassMatrix() and every per-stage Mesh method

    /** Source (src/solver.cpp:197-210): the same node sets, handed over once */
    if (!config.sources.empty()) {
        std::vector<int32_t> offsets(1, 0), idx;
        std::vector<double> amp, freq, phase, duration;
        for (size_t i = 0; i < config.sources.size(); ++i) {
            for (int n = 0; n < mesh.getNumNodes(); n++) {
                std::vector<double> coord, paramCoord;
                gmsh::model::mesh::getNode(mesh.getElNodeTags()[n], coord, paramCoord);
                if (pow(coord[0] - config.sources[i][1], 2) + pow(coord[1] - config.sources[i][2], 2) + pow(coord[2] - config.sources[i][3], 2) <
                    pow(config.sources[i][4], 2))
                    idx.push_back(n);
            }
            offsets.push_back((int32_t)idx.size());
            amp.push_back(config.sources[i][5]);
            freq.push_back(config.sources[i][6]);
            phase.push_back(config.sources[i][7]);
            duration.push_back(config.sources[i][8]);
        }
        if (idx.empty()) idx.push_back(0);
        if (dgb_set_sources(h, (int)config.sources.size(), offsets.data(), idx.data(), amp.data(), freq.data(), phase.data(), duration.data()) != DGB_OK)
            fail("dgb_set_sources");
    }

    std::vector<double> flat(4 * N);
    for (int q = 0; q < 4; ++q) std::copy(u[q].begin(), u[q].end(), flat.begin() + q * N);
    if (dgb_set_state(h, flat.data()) != DGB_OK) fail("dgb_set_state");

    /** Main loop: the reference's own loop header (src/solver.cpp:216-217 / 105-106). It now only decides WHEN a snapshot is
     *  taken; the steps between two snapshots run on the device in one dgb_run call, which accumulates t += dt in double
     *  exactly like this header does (so the source phases are the reference's). */
    auto start = std::chrono::system_clock::now();
    int pending = 0;
    double tPending = config.timeStart;
    for (double t = config.timeStart, step = 0, tDisplay = 0; t <= config.timeEnd; t += config.timeStep, tDisplay += config.timeStep, ++step) {
        if (tDisplay >= config.timeRate || step == 0) {
            tDisplay = 0;
            if (dgb_run(h, integrator, tPending, pending, nullptr) != DGB_OK) fail("dgb_run");
            pending = 0;
            tPending = t;
            if (dgb_get_state(h, flat.data()) != DGB_OK) fail("dgb_get_state");  // device -> host only at the snapshot cadence
            /** [1] Copy solution to match GMSH format (src/solver.cpp:226-238) */
            for (int el = 0; el < mesh.getElNum(); ++el) {
                for (int n = 0; n < mesh.getElNumNodes(); ++n) {
                    const size_t elN = (size_t)el * elNumNodes + n;
                    g_p[el][n] = flat[elN];
                    g_rho[el][n] = flat[elN] / (config.c0 * config.c0);
                    g_v[el][3 * n + 0] = flat[N + elN];
                    g_v[el][3 * n + 1] = flat[2 * N + elN];
                    g_v[el][3 * n + 2] = flat[3 * N + elN];
                }
            }
            gmsh::view::addModelData(gp_viewTag, step, g_names[0], "ElementNodeData", elTags, g_p, t, 1);
            gmsh::view::addModelData(grho_viewTag, step, g_names[0], "ElementNodeData", elTags, g_rho, t, 1);
            gmsh::view::addModelData(gv_viewTag, step, g_names[0], "ElementNodeData", elTags, g_v, t, 3);
            /** [2] Print and compute iteration time */
            auto end = std::chrono::system_clock::now();
            auto elapsed = std::chrono::duration_cast<std::chrono::seconds>(end - start);
            gmsh::logger::write("[" + std::to_string(t) + "/" + std::to_string(config.timeEnd) + "s] Step number : " + std::to_string((int)step) +
                                ", Elapsed time: " + std::to_string(elapsed.count()) + "s");
        }
        ++pending;
    }
    if (dgb_run(h, integrator, tPending, pending, nullptr) != DGB_OK) fail("dgb_run");
    if (dgb_get_state(h, flat.data()) != DGB_OK) fail("dgb_get_state");
    for (int q = 0; q < 4; ++q) std::copy(flat.begin() + q * N, flat.begin() + (q + 1) * N, u[q].begin());  // the caller's u, as the reference leaves it
    gmsh::logger::write(std::string("engine: ") + dgb_version() + ", kernel " + dgb_kernel_name(h) + ", " + std::to_string((long long)dgb_launch_count(h)) + " launches");
    dgb_destroy(h);

    /** Save to file (src/solver.cpp:289-291) */
    gmsh::view::write(gp_viewTag, config.saveFile, true);
    gmsh::view::write(grho_viewTag, config.saveFile, true);
    gmsh::view::write(gv_viewTag, config.saveFile, true);
}

}  // namespace

void forwardEuler(std::vector<std::vector<double>>& u, Mesh& mesh, Config config) { march(u, mesh, config, DGB_EULER1); }
void rungeKutta(std::vector<std::vector<double>>& u, Mesh& mesh, Config config) { march(u, mesh, config, DGB_RUNGE_KUTTA); }

}  // namespace solver
