/* dgfront.h — C API of the host front end that stands in for the reference's Gmsh-based set-up
 * (src/dgalerkin.cpp, src/configParser.cpp, Mesh::Mesh in src/Mesh.cpp:20-435) in an environment without
 * the Gmsh SDK. It produces exactly the arrays the reference's Mesh object holds and exposes them as a
 * dgb_desc (include/dgb.h). Host-only code; runs once per simulation.
 */
#ifndef DGFRONT_H
#define DGFRONT_H

#include <stdint.h>
#include "dgb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dgf_model dgf_model; /* what gmsh::open leaves in memory */
typedef struct dgf_mesh dgf_mesh;   /* the reference's Mesh object, as plain arrays */

#define DGF_MAX_SOURCES 64
#define DGF_MAX_INIT 32
#define DGF_MAX_PHYS 64
#define DGF_MAX_RECEIVERS 64

/* struct Config, include/configParser.h:7-41 */
typedef struct dgf_config {
    double timeStart, timeEnd, timeStep, timeRate;
    char elementType[64];
    char timeIntMethod[64];
    char saveFile[512];
    int32_t numThreads;
    double v0[3], rho0, c0;
    int32_t nSources;
    double sources[DGF_MAX_SOURCES][9]; /* pole, x, y, z, size, amp, freq, phase, duration (configParser.cpp:81-100) */
    int32_t nInit;
    double initConditions[DGF_MAX_INIT][6]; /* 0, x, y, z, size, amp (configParser.cpp:110) */
    int32_t nPhysBC;
    int32_t physBCTag[DGF_MAX_PHYS];  /* ascending physical tag (std::map order, configParser.h:23) */
    int32_t physBCType[DGF_MAX_PHYS]; /* 0 "Absorbing", 1 "Reflecting" */
    /* Receivers — keys the reference's parser ignores (it only looks at `source*`,
     * `initialCondtition*` and the physical-group names, configParser.cpp:65-132), so one file serves both programs:
     *   receiver<name> = x, y, z      (std::map order of the keys)      receiverFile = path   (default receivers.txt)
     *   receiverWav = prefix          (optional: the pressure at receiver j as 16-bit PCM, <prefix><j>.wav, rate 1/timeStep) */
    int32_t nReceivers;
    double receivers[DGF_MAX_RECEIVERS][3];
    char receiverFile[512];
    char receiverWav[512];
} dgf_config;

const char* dgf_last_error(void);

/* gmsh::open + (optionally) `gmsh -order p`: order <= 1 keeps the file's order. MSH ASCII 4.0 (the reference's sample meshes),
 * 4.1 (current Gmsh) and 2.2 are read; binary files are rejected with an error. */
dgf_model* dgf_open_msh(const char* path, int order);
/* synthetic cube of n^3 x 6 Kuhn tetrahedra on [lo,hi]^3 at `order` (BASELINE config 5) */
dgf_model* dgf_make_cube(int n, double lo, double hi, int order);
/* Structured square [lo,hi]^2: n^2 cells x 2 triangles of the given order, boundary edges in the physical group "Boundary"
 * (the refined square of BASELINE config 2). */
dgf_model* dgf_make_square(int n, double lo, double hi, int order);
/* writes the model as MSH 4.0 ASCII (the format of the reference's doc meshes), e.g. to feed the reference itself */
int dgf_write_msh(const dgf_model* m, const char* path);
/* Curved stand-in geometry: moves EVERY node by a smooth field of amplitude amp and wave number k, which
 * turns the straight-sided order-p model into a conforming curved isoparametric one (what `gmsh -order p` produces along
 * curved boundaries). dgf_mesh_build then stores one Jacobian / normal per integration point (nGeomEl = nG, nGeomF = nGf). */
int dgf_warp_model(dgf_model* m, double amp, double k);
/* The same inside a ball only (window (1-(r/R)^2)^2): a curved patch in an otherwise straight-sided mesh. dgf_mesh_build then
 * numbers the straight-sided elements first and the curved ones last (the engine runs its collapsed kernels on the first
 * group and the curved-element kernel on the second). */
int dgf_warp_model_local(dgf_model* m, double amp, double k, double cx, double cy, double cz, double radius);
void dgf_model_free(dgf_model* m);
int dgf_model_dimension(const dgf_model* m);

/* config::parseConfig (src/configParser.cpp:30-150); the model provides the physical-group names */
int dgf_parse_config(const char* path, const dgf_model* model, dgf_config* out);
/* a default-initialised config (include/configParser.h defaults) for programmatic use */
void dgf_default_config(dgf_config* out);

/* Mesh::Mesh (src/Mesh.cpp:20-435) */
dgf_mesh* dgf_mesh_build(dgf_model* model, const dgf_config* cfg);
void dgf_mesh_free(dgf_mesh* mesh);
const dgb_desc* dgf_mesh_desc(const dgf_mesh* mesh); /* pointers stay valid until dgf_mesh_free */
const double* dgf_mesh_node_coords(const dgf_mesh* mesh); /* [K*Np][3], what gmsh::model::mesh::getNode returns per DG node */
const int32_t* dgf_mesh_el_tags(const dgf_mesh* mesh);     /* [K] */
const int32_t* dgf_mesh_el_node_tags(const dgf_mesh* mesh); /* [K*Np] */
const int32_t* dgf_mesh_face_nodes(const dgf_mesh* mesh);   /* [Nf][Nfp] local node ids of each local face */

/* initial condition, src/dgalerkin.cpp:36-50 :  u[0][n] += amp*exp(-|x_n - x0|^2/size) ; u is [4][K*Np], zeroed first */
void dgf_initial_condition(const dgf_mesh* mesh, const dgf_config* cfg, double* u);
/* source node sets, src/solver.cpp:197-210 : nodes with |x_n - x_s|^2 < size^2.
 * Returns the total count; offsets has nSources+1 entries; nodeIdx may be NULL to query the size. */
int dgf_source_nodes(const dgf_mesh* mesh, const dgf_config* cfg, int32_t* offsets, int32_t* nodeIdx);
/* Replays the floating-point loop header of src/solver.cpp:216-223. Returns the number of executed steps;
 * if snapshotSteps != NULL it receives up to capacity step indices at which the reference takes a snapshot
 * and *nSnapshots their count. */
int dgf_time_loop(const dgf_config* cfg, int32_t* snapshotSteps, int capacity, int* nSnapshots);
/* nearest DG node to a point (probe placement helper; probes are a new capability) */
int dgf_nearest_node(const dgf_mesh* mesh, double x, double y, double z);

/* Receivers: finds the element that contains (x,y,z) — the lowest element id if the point lies on
 * a shared face / edge / vertex; if the point is outside the mesh, the element it violates least, with *outside = 1 —
 * and evaluates that element's Np Lagrange basis functions at the point (weights[Np]; uvw[3] optional = the
 * parametric coordinates). Straight-sided elements: the affine inverse map; curved ones: Newton on the isoparametric map.
 * Returns the element id, -1 on error. */
int dgf_locate_point(const dgf_mesh* mesh, double x, double y, double z, double* weights, double* uvw, int* outside);
/* Receiver time series as text: one line per step, `t  p vx vy vz` per receiver; a header names the points.
 * rec is [nsteps][nrecv][4] as dgb_get_receivers returns it; t_k accumulates t += dt like the loop header. */
int dgf_write_receivers(const char* path, int nrecv, const double* xyz /* [nrecv][3] */, int nsteps, double tStart,
                        double dt, const double* rec);
/* One field of one receiver as 16-bit mono PCM WAV, peak-normalised (the receiver audio of the reference's assets/);
 * the sample rate is round(1/dt) unless rate > 0. field: 0 p, 1..3 velocity. */
int dgf_write_wav(const char* path, int nrecv, int receiver, int field, int nsteps, double dt, int rate, const double* rec);

/* Gmsh-compatible output: appends $ElementNodeData views to `path` (MSH 4.0 ASCII, after a copy of the mesh
 * on first use), the way gmsh::view::write(tag, saveFile, append=true) does at src/solver.cpp:289-291. */
int dgf_write_views(const char* path, const dgf_model* model, const dgf_mesh* mesh, const dgf_config* cfg,
                    int nSnap, const int32_t* snapStep, const double* snapTime, const double* snapU /* [nSnap][4][K*Np] */);

/* ---- domain decomposition, host-side planning shared with the engine ---------------- */
/* recursive coordinate bisection of the element centroids into nparts parts (balanced to +-1 element) */
int dgf_partition_rcb(const dgf_mesh* mesh, int nparts, int32_t* elPart /* [K] */);
/* METIS k-way partition of the element dual graph (Gmsh's own partitioner is METIS as well); edgeCut (optional) = number of cut
   faces. Returns 0, -1 on failure, -2 if the library was built without METIS (libmetis_static.a of the CUDA toolkit). */
int dgf_partition_metis(const dgf_mesh* mesh, int nparts, int32_t* elPart /* [K] */, int64_t* edgeCut);
typedef struct dgf_plan dgf_plan;
dgf_plan* dgf_plan_create(const dgf_mesh* mesh, const int32_t* elPart, int rank, int nranks);
void dgf_plan_free(dgf_plan* plan);
/* sizes[6] = Kown, Kinterior, Khalo, npeers, nsend, nranks */
void dgf_plan_sizes(const dgf_plan* plan, int32_t* sizes);
/* localToGlobal[Kown+Khalo], peers[npeers], recvOffset[npeers+1], sendOffset[npeers+1], sendElems[nsend] (local ids) */
void dgf_plan_arrays(const dgf_plan* plan, int32_t* localToGlobal, int32_t* peers, int32_t* recvOffset, int32_t* sendOffset,
                     int32_t* sendElems);

#ifdef __cplusplus
}
#endif
#endif /* DGFRONT_H */
