// dgalerkin mesh.msh config.conf — the reference's command line (src/dgalerkin.cpp:11-63) on top of the host front
// end (include/dgfront.h) and the CUDA engine (include/dgb.h). The state lives on the GPU for the whole run; it comes
// back to the host only at the reference's snapshot cadence (solver.cpp:222-238) and the views are written once at
// the end (solver.cpp:289-291).
//
// Extras the reference ignores: environment DGB_ORDER=p elevates an order-1 mesh (the stand-in for `gmsh -order p`).
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dgb.h"
#include "dgfront.h"

int main(int argc, char** argv) {
    if (argc != 3) return E2BIG;  // dgalerkin.cpp:20
    const int order = std::getenv("DGB_ORDER") ? std::atoi(std::getenv("DGB_ORDER")) : 1;
    dgf_model* model = dgf_open_msh(argv[1], order);
    if (!model) { std::fprintf(stderr, "Error   : %s\n", dgf_last_error()); return EXIT_FAILURE; }
    dgf_config cfg;
    if (dgf_parse_config(argv[2], model, &cfg) != 0) { std::fprintf(stderr, "Error   : %s\n", dgf_last_error()); return EXIT_FAILURE; }
    std::printf("Info    : Config loaded : %s\n", argv[2]);
    std::remove(cfg.saveFile);  // configParser.cpp:115
    dgf_mesh* mesh = dgf_mesh_build(model, &cfg);
    if (!mesh) { std::fprintf(stderr, "Error   : %s\n", dgf_last_error()); return EXIT_FAILURE; }
    const dgb_desc* d = dgf_mesh_desc(mesh);
    const size_t N = (size_t)d->K * d->Np;
    std::printf("Info    : Number of Elements : %d, order %d, nodes per element %d, faces %d\n", d->K, d->order, d->Np, d->F);

    int integrator;
    if (std::strcmp(cfg.timeIntMethod, "Euler1") == 0) integrator = DGB_EULER1;
    else if (std::strcmp(cfg.timeIntMethod, "Runge-Kutta") == 0) integrator = DGB_RUNGE_KUTTA;
    else return EXIT_SUCCESS;  // dgalerkin.cpp:55-58: any other method silently does nothing

    std::vector<double> u(4 * N);
    dgf_initial_condition(mesh, &cfg, u.data());
    dgb_handle* h = nullptr;
    if (dgb_create(d, &h) != DGB_OK) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }
    if (cfg.nSources > 0) {
        std::vector<int32_t> off(cfg.nSources + 1);
        const int total = dgf_source_nodes(mesh, &cfg, off.data(), nullptr);
        std::vector<int32_t> idx(total > 0 ? total : 1);
        dgf_source_nodes(mesh, &cfg, off.data(), idx.data());
        std::vector<double> amp, freq, phase, dur;
        for (int s = 0; s < cfg.nSources; ++s) {
            amp.push_back(cfg.sources[s][5]); freq.push_back(cfg.sources[s][6]);
            phase.push_back(cfg.sources[s][7]); dur.push_back(cfg.sources[s][8]);
        }
        if (dgb_set_sources(h, cfg.nSources, off.data(), idx.data(), amp.data(), freq.data(), phase.data(), dur.data()) != DGB_OK) {
            std::fprintf(stderr, "Error   : %s\n", dgb_last_error());
            return EXIT_FAILURE;
        }
    }
    // receivers (keys `receiver<name> = x, y, z`, ignored by the reference's parser): located once, sampled on the device
    std::vector<double> rcvRec;
    if (cfg.nReceivers > 0) {
        std::vector<int32_t> rEl(cfg.nReceivers);
        std::vector<double> rW((size_t)cfg.nReceivers * d->Np);
        for (int j = 0; j < cfg.nReceivers; ++j) {
            int outside = 0;
            rEl[j] = dgf_locate_point(mesh, cfg.receivers[j][0], cfg.receivers[j][1], cfg.receivers[j][2], &rW[(size_t)j * d->Np], nullptr, &outside);
            if (rEl[j] < 0) { std::fprintf(stderr, "Error   : %s\n", dgf_last_error()); return EXIT_FAILURE; }
            if (outside) std::printf("Warning : receiver %d lies outside the mesh, extrapolating from element %d\n", j, rEl[j]);
        }
        if (dgb_set_receivers(h, cfg.nReceivers, rEl.data(), rW.data()) != DGB_OK) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }
    }
    auto drainReceivers = [&](int nsteps) {
        if (cfg.nReceivers == 0 || nsteps == 0) return true;
        const size_t at = rcvRec.size();
        rcvRec.resize(at + (size_t)nsteps * cfg.nReceivers * 4);
        int got = 0;
        if (dgb_get_receivers(h, rcvRec.data() + at, nsteps, &got) != DGB_OK) { rcvRec.resize(at); return false; }
        rcvRec.resize(at + (size_t)got * cfg.nReceivers * 4);
        return true;
    };
    if (dgb_set_state(h, u.data()) != DGB_OK) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }

    std::vector<int32_t> snapStep;
    std::vector<double> snapTime, snapU;
    // snapshots leave the GPU asynchronously (dgb_snapshot_begin / _end, two pinned buffers): the copy of snapshot k overlaps the
    // steps up to snapshot k+1. If that is not available (partitioned handle, no pinned memory) the plain dgb_get_state is used.
    double* pinned[2] = {static_cast<double*>(dgb_host_alloc(4 * N * sizeof(double))), static_cast<double*>(dgb_host_alloc(4 * N * sizeof(double)))};
    bool async = pinned[0] && pinned[1];
    int inflight = -1;
    auto collect = [&]() {  // the snapshot in flight has landed: append it
        if (inflight < 0) return true;
        if (dgb_snapshot_end(h) != DGB_OK) return false;
        snapU.insert(snapU.end(), pinned[inflight], pinned[inflight] + 4 * N);
        inflight = -1;
        return true;
    };
    const auto start = std::chrono::system_clock::now();
    int pending = 0;
    double tPending = cfg.timeStart, step = 0, tDisplay = 0;
    for (double t = cfg.timeStart; t <= cfg.timeEnd; t += cfg.timeStep, tDisplay += cfg.timeStep, ++step) {  // solver.cpp:216-217
        if (tDisplay >= cfg.timeRate || step == 0) {
            tDisplay = 0;
            if (dgb_run(h, integrator, tPending, pending, nullptr) != DGB_OK) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }
            const int buf = (int)(snapStep.size() & 1);
            const bool ok = collect();  // the previous snapshot had the whole chunk to reach the host
            if (ok && async && dgb_snapshot_begin(h, pinned[buf]) == DGB_OK) inflight = buf;
            else {
                async = false;
                if (!ok || dgb_get_state(h, u.data()) != DGB_OK) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }
                snapU.insert(snapU.end(), u.begin(), u.end());
            }
            if (!drainReceivers(pending)) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }
            pending = 0;
            tPending = t;
            snapStep.push_back((int32_t)step);
            snapTime.push_back(t);
            const auto el = std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now() - start);
            std::printf("Info    : [%f/%fs] Step number : %d, Elapsed time: %llds\n", t, cfg.timeEnd, (int)step, (long long)el.count());
        }
        ++pending;
    }
    if (dgb_run(h, integrator, tPending, pending, nullptr) != DGB_OK) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }
    if (!collect() || !drainReceivers(pending)) { std::fprintf(stderr, "Error   : %s\n", dgb_last_error()); return EXIT_FAILURE; }
    if (cfg.nReceivers > 0) {
        const int nrec = (int)(rcvRec.size() / ((size_t)cfg.nReceivers * 4));
        if (dgf_write_receivers(cfg.receiverFile, cfg.nReceivers, &cfg.receivers[0][0], nrec, cfg.timeStart, cfg.timeStep, rcvRec.data()) != 0)
            std::fprintf(stderr, "Error   : %s\n", dgf_last_error());
        else std::printf("Info    : %d receiver(s), %d samples -> %s\n", cfg.nReceivers, nrec, cfg.receiverFile);
        for (int j = 0; cfg.receiverWav[0] && j < cfg.nReceivers && nrec > 0; ++j) {  // receiver audio, like the reference's assets/
            const std::string wav = std::string(cfg.receiverWav) + std::to_string(j) + ".wav";
            if (dgf_write_wav(wav.c_str(), cfg.nReceivers, j, 0, nrec, cfg.timeStep, 0, rcvRec.data()) != 0) std::fprintf(stderr, "Error   : %s\n", dgf_last_error());
        }
    }
    std::printf("Info    : %lld kernel launches, last chunk %.3f ms on the device\n", (long long)dgb_launch_count(h), dgb_last_run_ms(h));
    const int viewsRc = dgf_write_views(cfg.saveFile, model, mesh, &cfg, (int)snapStep.size(), snapStep.data(), snapTime.data(), snapU.data());
    if (viewsRc != 0) std::fprintf(stderr, "Error   : %s\n", dgf_last_error());
    dgb_destroy(h);
    dgb_host_free(pinned[0]);
    dgb_host_free(pinned[1]);
    dgf_mesh_free(mesh);
    dgf_model_free(model);
    return viewsRc == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
