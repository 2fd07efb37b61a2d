/* dgb.h — C ABI of the B200-native DG time-marching engine ("dgb" = DG on Blackwell).
 *
 * This is the drop-in boundary for the hot path of povanberg/DGFEM-Acoustic. It replaces the body of
 *     void solver::rungeKutta (std::vector<std::vector<double>>& u, Mesh& mesh, Config config)   include/solver.h:24
 *     void solver::forwardEuler(std::vector<std::vector<double>>& u, Mesh& mesh, Config config)   include/solver.h:15
 * (src/solver.cpp:61-161, 171-292) and everything they call per stage:
 *     solver::numStep            src/solver.cpp:35-52
 *     Mesh::updateFlux           src/Mesh.cpp:569-674
 *     Mesh::precomputeFlux       src/Mesh.cpp:500-539
 *     Mesh::getElFlux            src/Mesh.cpp:548-557
 *     Mesh::getElStiffVector     src/Mesh.cpp:476-489
 *     Mesh::precomputeMassMatrix src/Mesh.cpp:440-466
 *     eigen::linEq/minus/plusTimes  src/utils.cpp:118-148
 * The front end (CLI, config parser, Mesh constructor) stays where it is: it fills a dgb_desc with the arrays
 * the reference's Mesh object already holds, uploads them once with dgb_create(), and then drives the time
 * loop with dgb_run(); the solution stays resident in HBM between calls. INTEGRATION.md shows the binding.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer (copied during the call); the
 * handle owns all device memory; every function returns 0 on success or a negative dgb_status and never
 * throws; a handle is not thread-safe. There is NO CPU fallback: without a CUDA device dgb_create() fails
 * with DGB_ERR_CUDA.
 */
#ifndef DGB_H
#define DGB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dgb_handle dgb_handle;

typedef enum dgb_status {
    DGB_OK = 0,
    DGB_ERR_ARG = -1,         /* null pointer / inconsistent sizes */
    DGB_ERR_UNSUPPORTED = -2, /* unknown element, order > 6, curved geometry on a partitioned handle ... */
    DGB_ERR_CUDA = -3,        /* CUDA runtime / no device */
    DGB_ERR_NCCL = -4,
    DGB_ERR_STATE = -5        /* call order (e.g. run before set_state) */
} dgb_status;

typedef enum dgb_integrator {
    DGB_EULER1 = 0,      /* config.timeIntMethod == "Euler1"       src/dgalerkin.cpp:55 */
    DGB_RUNGE_KUTTA = 1  /* config.timeIntMethod == "Runge-Kutta"  src/dgalerkin.cpp:57 */
} dgb_integrator;

/* Everything the hot path reads, in the layouts of the reference's Mesh (include/Mesh.h line numbers given).
 * Index names: el element, f unique face, lf local face, g quadrature point, i/n node, x physical, u parametric. */
typedef struct dgb_desc {
    int32_t dim;     /* m_elDim        Mesh.h:140 */
    int32_t order;   /* m_elOrder      Mesh.h:143 */
    int32_t Np;      /* m_elNumNodes   Mesh.h:144 */
    int32_t Nfp;     /* m_fNumNodes    Mesh.h:177 */
    int32_t Nf;      /* m_fNumPerEl    Mesh.h:178 */
    int32_t K;       /* m_elNum        Mesh.h:146 */
    int32_t F;       /* m_fNum         Mesh.h:180 */
    int32_t nG;      /* m_elNumIntPts  Mesh.h:145 */
    int32_t nGf;     /* m_fNumIntPts   Mesh.h:181 */
    int32_t nGeomEl; /* points stored per element in elJacobian/elJacobianDet: nG (reference layout) or 1 (affine, compressed) */
    int32_t nGeomF;  /* points stored per face in fNormal/fJacobianDet: nGf (reference layout) or 1 */
    int32_t fc;      /* numerical-flux sign, Mesh.h:139, Mesh.cpp:210-211 */

    const double* elBasisFct;      /* [nG][Np]        Mesh.h:46  */
    const double* elUGradBasisFct; /* [nG][Np][3]     Mesh.h:49  */
    const double* elWeight;        /* [nG]            Mesh.h:43  */
    const double* fBasisFct;       /* [nGf][Nfp]      Mesh.h:85  */
    const double* fWeight;         /* [nGf]           Mesh.h:82  */

    const double* elJacobian;    /* [K][nGeomEl][9], index u*3+x = dx_x/du_u   Mesh.h:37 */
    const double* elJacobianDet; /* [K][nGeomEl]                               Mesh.h:40 */
    const double* fNormal;       /* [F][nGeomF][3]  (after the boundary flip, Mesh.cpp:336-349)  Mesh.h:88 */
    const double* fJacobianDet;  /* [F][nGeomF]                                Mesh.h:70 */

    const int32_t* elFId;          /* [K][Nf]                                   Mesh.h:91  */
    const int32_t* elFOrientation; /* [K][Nf]  +1/-1                            Mesh.h:100 */
    const int32_t* fNbrElId;       /* [F][2]   second entry -1 on boundary faces Mesh.h:94 */
    const int32_t* fNToElNId;      /* [F][Nfp][2]  (-1 where there is no second owner)  Mesh.h:97 */
    const uint8_t* fIsBoundary;    /* [F]                                       Mesh.h:218 */
    const int32_t* fBC;            /* [F]  1 reflecting, anything else absorbing  Mesh.h:219, Mesh.cpp:360-384 */

    double c0, rho0, v0[3]; /* Config, include/configParser.h:26-28 */
    double dt;              /* config.timeStep */
} dgb_desc;

/* ---- life cycle -------------------------------------------------------------------------------------- */
/* Builds the reference-element operators (Dw^u, LIFT), converts to the device layout and uploads everything once.
 * Straight-sided meshes run the collapsed operator kernels; a mesh with curved elements (Jacobians / normals that vary
 * between the integration points; the desc must then use the reference layout nGeomEl == nG, nGeomF == nGf) runs the
 * reference's own quadrature loops on the device (csrc/stage_curved.cu). Single GPU (current CUDA device). */
int dgb_create(const dgb_desc* desc, dgb_handle** out);

/* Multi-GPU: every rank passes the SAME global desc plus elPart[K] (owner rank of every element); the handle
 * keeps the rank's elements and the one-layer halo and exchanges halo traces with NCCL once per stage.
 * nccl_unique_id is the 128-byte ncclUniqueId obtained by rank 0 from dgb_nccl_unique_id() and distributed
 * by the caller (torch.distributed / MPI / a file). nranks == 1 degenerates to dgb_create(). */
int dgb_create_partitioned(const dgb_desc* desc, const int32_t* elPart, int rank, int nranks,
                           const void* nccl_unique_id, dgb_handle** out);
int dgb_nccl_unique_id(void* out128);
void dgb_destroy(dgb_handle* h);

/* ---- state: the reference's u[eq][el*Np+n], eq = p,vx,vy,vz (src/dgalerkin.cpp:36) ------------------- */
/* Partitioned handles read/write only the entries of the elements they own (global indexing is kept). */
int dgb_set_state(dgb_handle* h, const double* u /* [4][K*Np] */);
int dgb_get_state(dgb_handle* h, double* u /* [4][K*Np] */);

/* Asynchronous snapshot (the reference copies the state into its Gmsh views at every timeRate, src/solver.cpp:222-238; SURVEY.md
 * §8 f2: "so that snapshots do not stall the GPU"). dgb_snapshot_begin copies the state into a device-side snapshot buffer on
 * the compute stream — the next dgb_run may start at once — and starts the device->host copy into u_host on a second stream;
 * dgb_snapshot_end waits for that copy, after which u_host holds what dgb_get_state would have returned at the time of
 * dgb_snapshot_begin. One snapshot in flight per handle (a second begin waits for the first copy); u_host should be pinned
 * (dgb_host_alloc) for the copy to overlap the computation. Single-GPU handles; partitioned ones return DGB_ERR_UNSUPPORTED. */
int dgb_snapshot_begin(dgb_handle* h, double* u_host /* [4][K*Np] */);
int dgb_snapshot_end(dgb_handle* h);
void* dgb_host_alloc(uint64_t bytes); /* page-locked host memory, NULL on failure */
void dgb_host_free(void* p);

/* ---- sources (src/solver.cpp:197-210, 248-256) and probes (new capability) --------------------------- */
/* Source s overwrites u[0][nodeIdx[offsets[s] .. offsets[s+1])] with amp*sin(2*pi*freq*t+phase) at the start
 * of every step while t < duration. The sine is evaluated on the host in the reference's expression. */
int dgb_set_sources(dgb_handle* h, int nsrc, const int32_t* offsets, const int32_t* nodeIdx,
                    const double* amp, const double* freq, const double* phase, const double* duration);
/* Probe j records (p,vx,vy,vz) at DG node nodeIdx[j] at the START of every step (same instant as the
 * reference's snapshots, src/solver.cpp:222-238). */
int dgb_set_probes(dgb_handle* h, int nprobe, const int32_t* nodeIdx);
/* out[step][probe][4]; returns the number of recorded steps in *nsteps and clears the record. Probes owned
 * by another rank are returned as 0. If capacity_steps is smaller than the number of recorded steps nothing is
 * copied or cleared: the call returns DGB_ERR_ARG with the required capacity in *nsteps. */
int dgb_get_probes(dgb_handle* h, double* out, int capacity_steps, int* nsteps);

/* Receivers (new capability): receiver j records the four fields interpolated at a point inside
 * element el[j], value_q = sum_n weights[j][n] * u[q][el[j]*Np + n] with weights = the element's Lagrange basis at
 * the point (dgf_locate_point in include/dgfront.h computes both), at the START of every step like the probes.
 * el uses global element ids; receivers whose element another rank owns are returned as 0. */
int dgb_set_receivers(dgb_handle* h, int nrecv, const int32_t* el, const double* weights /* [nrecv][Np] */);
/* out[step][receiver][4]; returns the number of recorded steps in *nsteps and clears the record; a too small
 * capacity_steps is refused like in dgb_get_probes. */
int dgb_get_receivers(dgb_handle* h, double* out, int capacity_steps, int* nsteps);

/* ---- time marching ----------------------------------------------------------------------------------- */
/* Advances nsteps steps starting at time t_start, accumulating t += dt in double exactly like the loop
 * header src/solver.cpp:216-217; *t_end (optional) receives the accumulated time. The state never leaves
 * the device. */
int dgb_run(dgb_handle* h, int integrator, double t_start, int nsteps, double* t_end);
/* One operator evaluation rhs = L(u) = M^-1 (S(u) - F(u)) (what numStep applies with dt = 1, beta = 0),
 * host in / host out; used by the parity tests. */
int dgb_eval_rhs(dgb_handle* h, const double* u, double* rhs);

/* ---- instrumentation --------------------------------------------------------------------------------- */
int dgb_set_stream(dgb_handle* h, void* cuda_stream);   /* run on the caller's stream (e.g. torch's current one) */
int dgb_synchronize(dgb_handle* h);
double dgb_last_run_ms(dgb_handle* h);        /* CUDA-event time of the last dgb_run */
double dgb_last_stage_kernel_ms(dgb_handle* h); /* mean CUDA-event duration of the stage kernel in the last run */
int64_t dgb_launch_count(dgb_handle* h);      /* kernels launched by this handle so far */
/* FP64 (DFMA) issue peak of the handle's device in TFLOP/s, measured on the spot (~10 ms): the denominator of the FP64
 * roofline fraction bench.py reports, from the same box in the same run. */
double dgb_measure_fp64_tflops(dgb_handle* h);
/* Options (all integers):
 *   "kernel"   0 automatic (default: a Bernstein-Bezier kernel wherever one exists — 7 on triangles of order 1 / 2 and tetrahedra
 *              of order 1, 6 on triangles of order 3..6 and tetrahedra of order 2..6 — else the CUDA-core kernel),
 *              1 generic CUDA cores, 2 tiled DMMA, 3 warp-specialised DMMA (zero mean flow),
 *              4 / 5 first-generation Bernstein-Bezier kernels (tetrahedra of order 2..5; 5 = face-sequential schedule;
 *              "bb_tile": 32 / 16 / 8 elements per CTA), 6 second-generation Bernstein-Bezier kernel (csrc/stage_bb2.cu:
 *              tetrahedra and triangles of order 1..6), 7 the same arithmetic with one thread per element
 *              (csrc/stage_bbe.cu: triangles of order 1..3, tetrahedra of order 1 / 2). A kernel that does not exist for the
 *              mesh is refused with DGB_ERR_UNSUPPORTED.
 *              With a Bernstein kernel the state is kept as Bernstein coefficients on the device; dgb_set_state /
 *              dgb_get_state / probes / receivers / sources convert.
 *   "overlap"  -1 automatic (default), 0 none, 1 same-stage, 2 next-stage (NCCL exchange only)
 *   "sm_reserve" SMs left to NCCL during overlapped launches
 *   "exchange" (COLLECTIVE, partitioned handles) 0 ncclSend/ncclRecv; 1 direct stores into the peers' halo slots over NVLink
 *              through CUDA IPC mappings + epoch flags (csrc/halo_p2p.cu: three small launches per stage); 2 the same stores
 *              issued by the stage kernel itself (TMA bulk stores from shared memory, last CTA signals, border tiles wait:
 *              no extra launch; second-generation Bernstein kernel, other kernels behave as with 1). Default: 2 if every rank
 *              can map its peers' arrays, else 0 (environment DGB_EXCHANGE overrides). With 1 / 2 destroy is collective.
 *   "p2p_timeout_ms" how long a rank waits for a peer's halo before dgb_run fails
 *   "graph"    -1 automatic / 0 / 1 CUDA graph of an RK4 step (one GPU, no sources / probes / receivers)
 *   "time_stages" 0 / 1 */
int dgb_set_option(dgb_handle* h, const char* key, int value);
/* Reads back an option ("kernel" reports the active kernel's id, "exchange" the mode in effect, "representation" 0 nodal /
 * 1 / 2 Bernstein layouts, "overlap", "graph", "bb_tile", "sm_reserve", "p2p_timeout_ms"). */
int dgb_get_option(dgb_handle* h, const char* key, int* value);
const char* dgb_kernel_name(dgb_handle* h);   /* which stage kernel the handle selected */
const char* dgb_last_error(void);
const char* dgb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DGB_H */
