// C ABI of the engine (include/dgb.h): operator construction, upload, device-resident time loop, halo exchange.
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time (see NcclApi)

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "bb_setup.h"
#include "curved_setup.h"
#include "dgb_internal.h"
#include "dgb_launch.h"
#include "partition.h"
#include "tile_cfg.h"

using namespace dgb;

namespace {

thread_local std::string g_err;

struct DgbException : std::runtime_error {
    int code;
    DgbException(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                                         \
    do {                                                                                                         \
        cudaError_t _e = (expr);                                                                                 \
        if (_e != cudaSuccess)                                                                                   \
            throw DgbException(DGB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)
// NCCL is bound with dlopen the first time a partitioned handle (or a unique id) is requested: single-GPU users need
// no NCCL at all, and a process that already carries an NCCL (e.g. the one bundled with PyTorch, same SONAME
// libnccl.so.2) keeps using exactly that one instead of getting a second copy mapped over it.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.ok) return api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw DgbException(DGB_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) {
        void* p = dlsym(h, n);
        if (!p) throw DgbException(DGB_ERR_NCCL, std::string("libnccl lacks ") + n);
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = true;
    return api;
}

#define NCCL_CHECK(expr)                                                                                         \
    do {                                                                                                         \
        ncclResult_t _r = (expr);                                                                                \
        if (_r != ncclSuccess) throw DgbException(DGB_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
    } while (0)

template <typename T>
T* devAlloc(size_t n) {
    T* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}
template <typename T>
T* devUpload(const std::vector<T>& v) {
    T* p = devAlloc<T>(v.size());
    if (!v.empty()) CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

// ---- small dense helpers in extended precision (set-up only) ----
typedef long double real;
void invertDense(std::vector<real>& A, int n) {
    std::vector<real> B((size_t)n * n, 0);
    for (int i = 0; i < n; ++i) B[(size_t)i * n + i] = 1;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsl(A[(size_t)r * n + c]) > fabsl(A[(size_t)piv * n + c])) piv = r;
        if (A[(size_t)piv * n + c] == 0) throw DgbException(DGB_ERR_ARG, "singular reference mass matrix");
        if (piv != c)
            for (int k = 0; k < n; ++k) { std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]); std::swap(B[(size_t)c * n + k], B[(size_t)piv * n + k]); }
        const real d = 1 / A[(size_t)c * n + c];
        for (int k = 0; k < n; ++k) { A[(size_t)c * n + k] *= d; B[(size_t)c * n + k] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const real f = A[(size_t)r * n + c];
            if (f == 0) continue;
            for (int k = 0; k < n; ++k) { A[(size_t)r * n + k] -= f * A[(size_t)c * n + k]; B[(size_t)r * n + k] -= f * B[(size_t)c * n + k]; }
        }
    }
    A.swap(B);
}

}  // namespace

struct dgb_handle {
    DeviceMesh M{};
    dgb_desc hd{};  // scalar members of the global desc
    int Kglobal = 0, Np = 0;
    int device = 0;  // the CUDA device the handle was created on; every entry point makes it current
    bool partitioned = false;
    int nranks = 1;
    PartitionPlan plan;
    double *U = nullptr, *ACC = nullptr, *YA = nullptr, *YB = nullptr;
    cudaStream_t stream = nullptr, commStream = nullptr;
    bool ownStream = true;
    cudaEvent_t evStart = nullptr, evStop = nullptr, evBorder = nullptr, evRecv = nullptr;
    std::vector<cudaEvent_t> stageEv;  // pairs, for per-launch timing of the stage kernel
    int stageEvUsed = 0;
    StageKernel generic, tiled, ws, bbKernel, bbSeqKernel, bb2Kernel, bbeKernel, active;
    // Bernstein-Bezier mode (dgb_set_option("kernel", 4)): the state arrays hold Bernstein coefficients; V / V^-1 convert
    int bbMode = 0;                  // representation of the resident state: 0 nodal, 1 coefficients [field][el][mesh node order] (stage_bb.cu),
                                     // 2 coefficients [el][canonical index][field] (stage_bb2.cu)
    int bbTile = 32;                 // elements per CTA of the Bernstein kernels (32, 16, 8)
    std::string bbWhyNot;            // why the Bernstein path is unavailable for this mesh (empty: available)
    // curved (non-affine) meshes: the reference's own tables on the device + inverse element mass matrices; every
    // element then goes through stage_curved.cu
    bool curved = false;
    int firstCurved = 0;             // elements [0, firstCurved) are straight-sided and keep the collapsed kernels
    std::string curvedName;          // kernel name reported for a curved handle
    CurvedMesh CM{};
    std::vector<void*> curvedAllocs;
    double *dV = nullptr, *dVinv = nullptr;
    double *dVC = nullptr, *dVinvC = nullptr;  // the same with the coefficient index in canonical order: VC[n][i], VinvC[i][n]
    uint8_t* dBBTab = nullptr;                 // DeviceMesh::bbTab, followed by the mesh node -> canonical index permutation
    size_t bbPermOffset = 0;
    std::vector<uint8_t> permG2C;
    double *dProbeWBB2 = nullptr, *dRecvWBB2 = nullptr;
    std::vector<double> hostV;       // [Np][Np], u_n = sum_m V[n][m] c_m (mesh node order)
    // Bernstein twins of the probes / receivers (a nodal value is a weighted sum of the element's coefficients) and sources
    int32_t* dProbeElBB = nullptr;
    double *dProbeWBB = nullptr, *dRecvWBB = nullptr;
    std::vector<int32_t> srcElOff;   // per source: its range in dSrcElList
    int32_t *dSrcElList = nullptr, *dSrcNodeOff = nullptr, *dSrcNodeLocal = nullptr;
    // Measured on B200 (profiles/r02_order_sweep.json): the second-generation Bernstein kernel beats the CUDA-core kernel on
    // tetrahedra from order 2 on and the dense DMMA kernels at orders 3 / 4, with and without mean flow
    // measured (profiles/r02_order_sweep.json): Bernstein kernels win wherever they exist — one thread per element (stage_bbe) on
    // triangles of orders 1 / 2 and tetrahedra of order 1, the (element, field) pipeline (stage_bb2) on triangles of orders 3..6
    // and tetrahedra of orders 2..5
    StageKernel autoKernel() const {
        if (preferBB2 && bbeKernel.launch && ((M.dim == 2 && M.order <= 2) || (M.dim == 3 && M.order == 1))) return bbeKernel;
        if (preferBB2 && bb2Kernel.launch) return bb2Kernel;
        return ws.launch ? ws : tiled.launch ? tiled : generic;
    }
    bool preferBB2 = true;
    int overlap = -1;   // 0: stage, then exchange; 1: border, [exchange || interior]; 2: interior(s+1) || exchange(s), then border; -1: automatic
    int smReserve = 4;  // SMs left to the NCCL kernels while an overlapped interior launch of a persistent kernel runs
    bool recvPending = false;  // overlap 2: the halo exchange of the previous stage has not been waited for yet
    // CUDA graph of one RK4 step (4 stage launches) for launch-bound runs: small meshes on one GPU, no sources, no probes
    int useGraph = -1;         // -1 automatic, 0 never, 1 whenever possible
    cudaGraphExec_t stepGraph = nullptr;
    double stepGraphDt = 0;
    StageLaunchFn stepGraphKernel = nullptr;
    const double* stepGraphPtr[3] = {nullptr, nullptr, nullptr};  // U / YA / YB the graph was captured with (Euler runs swap the names)
    int timeStages = 1;
    // sources
    std::vector<int32_t> srcOff;
    int32_t* dSrcIdx = nullptr;
    std::vector<double> srcAmp, srcFreq, srcPhase, srcDur;
    // probes
    int nprobe = 0;
    int32_t* dProbeIdx = nullptr;
    double* dProbeRec = nullptr;
    int probeCap = 0, probeCount = 0;
    // receivers (interpolated inside an element)
    int nrecv = 0;
    int32_t* dRecvEl = nullptr;
    double *dRecvW = nullptr, *dRecvRec = nullptr;
    int recvCap = 0, recvCount = 0;
    bool stateSet = false;
    double lastRunMs = 0, lastStageMs = 0;
    int64_t launches = 0;
    // halo exchange
    ncclComm_t comm = nullptr;
    double *sendBuf = nullptr, *recvBuf = nullptr;
    int32_t* dSendElems = nullptr;
    double* hostStage = nullptr;  // pinned, [4][stride], partitioned handles only (pageable caller buffers)
    int32_t* dL2G = nullptr;      // localToGlobal on the device (pinned caller buffers: the GPU gathers / scatters over PCIe)
    // asynchronous snapshot: device-side copy of the state + a copy stream
    double* dSnap = nullptr;
    cudaStream_t snapStream = nullptr;
    cudaEvent_t evSnapReady = nullptr, evSnapDone = nullptr;
    bool snapPending = false;
    // direct peer-to-peer halo exchange (dgb_set_option("exchange", 1), halo_p2p.cu): U / YA / YB and the epoch flags live in
    // ONE allocation ("arena") that every peer maps through CUDA IPC
    int exchangeMode = 0;            // 0: ncclSend/ncclRecv, 1: stores into the peers' halo slots + epoch flags, 2: the same fused into the stage kernel
    char* arena = nullptr;           // owns U, YA, YB once P2P is set up
    double* phys[3] = {nullptr, nullptr, nullptr};  // allocation identity of the three exchanged arrays (the names U/YA swap in Euler runs)
    struct PeerMap { void* opened = nullptr; char* base = nullptr; int64_t stride = 0; size_t arrayBytes = 0, flagOffset = 0; };
    std::vector<PeerMap> peerMap;    // per plan.peers[i]
    size_t arenaArrayBytes = 0, arenaFlagOffset = 0;
    unsigned long long* dFlags = nullptr;  // [nranks] inside the arena, slot r is written by rank r
    unsigned long long epoch = 0;
    int32_t *dSendPeer = nullptr, *dSendSlot = nullptr;
    int* p2pErr = nullptr;           // pinned + mapped: the wait kernel reports a timeout here
    // exchange fused into the stage kernel (exchange = 2, stage_bb2.cu)
    int32_t *dPushOff = nullptr, *dPushPeer = nullptr, *dPushSlot = nullptr;
    unsigned int* dDone = nullptr;
    FusedHalo hostFused{};
    FusedHalo* dFused = nullptr;
    bool haloInFlight = false;       // a fused stage has been launched whose incoming halo has not been waited for by a wait kernel
    int p2pTimeoutMs = 20000;
};

namespace {

// Makes the handle's device current for the duration of an API call (a process may hold handles on several GPUs).
struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(const dgb_handle* h) {
        if (!h) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != h->device && cudaSetDevice(h->device) == cudaSuccess) prev = cur;
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

void freeHandle(dgb_handle* h) {
    if (!h) return;
    DeviceScope onDevice(h);
    auto F = [](void* p) { if (p) cudaFree(p); };
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->commStream) cudaStreamSynchronize(h->commStream);
    if (h->stepGraph) cudaGraphExecDestroy(h->stepGraph);
    if (h->hostStage) cudaFreeHost(h->hostStage);
    if (h->snapStream) { cudaStreamSynchronize(h->snapStream); cudaStreamDestroy(h->snapStream); }
    for (auto e : {h->evSnapReady, h->evSnapDone}) if (e) cudaEventDestroy(e);
    F(h->dSnap);
    for (auto& pm : h->peerMap) if (pm.opened) cudaIpcCloseMemHandle(pm.opened);
    if (h->arena && h->comm && h->stream) {
        // peers still map this rank's arena: every rank closes its mappings (above) before anybody frees (collective destroy).
        // The communicator must still be alive here: it is destroyed only after this barrier.
        char* scratch = h->arena + h->arenaFlagOffset + 128;  // the second half of the flag block is unused
        const ncclResult_t r = nccl().AllGather(scratch + h->plan.rank, scratch, 1, ncclChar, h->comm, h->stream);
        if (r == ncclSuccess) cudaStreamSynchronize(h->stream);
        else fprintf(stderr, "dgb_destroy: arena barrier failed (%s); a peer may still map this rank's arena\n", nccl().GetErrorString(r));
    }
    if (h->comm) { nccl().CommDestroy(h->comm); h->comm = nullptr; }
    if (h->p2pErr) cudaFreeHost(h->p2pErr);
    if (h->arena) { F(h->arena); h->U = h->YA = h->YB = nullptr; }  // the arena owns the three arrays
    for (void* p : h->curvedAllocs) F(p);
    F(h->dL2G); F(h->dSendPeer); F(h->dSendSlot); F(h->dPushOff); F(h->dPushPeer); F(h->dPushSlot); F(h->dDone); F(h->dFused); F(h->dV); F(h->dVinv); F(h->dVC); F(h->dVinvC); F(h->dBBTab); F(h->dProbeWBB2); F(h->dRecvWBB2); F(h->dProbeElBB); F(h->dProbeWBB); F(h->dRecvWBB); F(h->dSrcElList); F(h->dSrcNodeOff); F(h->dSrcNodeLocal);
    F(h->U); F(h->ACC); F(h->YA); F(h->YB);
    F(h->M.DwT); F(h->M.nLiftT); F(h->M.tiledOps); F(h->M.faceNodes); F(h->M.nbrMaps);
    F(h->M.Ginv); F(h->M.fgeo); F(h->M.fnbr); F(h->M.fflags);
    F(h->dSrcIdx); F(h->dProbeIdx); F(h->dProbeRec); F(h->dRecvEl); F(h->dRecvW); F(h->dRecvRec); F(h->sendBuf); F(h->recvBuf); F(h->dSendElems);
    for (auto e : h->stageEv) cudaEventDestroy(e);
    for (auto e : {h->evStart, h->evStop, h->evBorder, h->evRecv}) if (e) cudaEventDestroy(e);
    if (h->ownStream && h->stream) cudaStreamDestroy(h->stream);
    if (h->commStream) cudaStreamDestroy(h->commStream);
    delete h;
}

void validate(const dgb_desc* d) {
    if (!d) throw DgbException(DGB_ERR_ARG, "desc is null");
    const void* ptrs[] = {d->elBasisFct, d->elUGradBasisFct, d->elWeight, d->fBasisFct, d->fWeight, d->elJacobian, d->elJacobianDet,
                          d->fNormal, d->fJacobianDet, d->elFId, d->elFOrientation, d->fNbrElId, d->fNToElNId, d->fIsBoundary, d->fBC};
    for (const void* p : ptrs) if (!p) throw DgbException(DGB_ERR_ARG, "desc has a null array pointer");
    if (d->dim < 1 || d->dim > 3 || d->order < 1 || d->order > 6) throw DgbException(DGB_ERR_UNSUPPORTED, "dim must be 1..3 and order 1..6");
    if (d->K < 1 || d->F < 1 || d->Np < 1 || d->Nfp < 1 || d->Nf < 1 || d->nG < 1 || d->nGf < 1) throw DgbException(DGB_ERR_ARG, "non-positive size in desc");
    const int p = d->order;
    const int np = d->dim == 1 ? p + 1 : d->dim == 2 ? (p + 1) * (p + 2) / 2 : (p + 1) * (p + 2) * (p + 3) / 6;
    const int nfp = d->dim == 1 ? 1 : d->dim == 2 ? p + 1 : (p + 1) * (p + 2) / 2;
    const int nf = d->dim == 1 ? p + 1 : d->dim + 1;
    if (d->Np != np || d->Nfp != nfp || d->Nf != nf) throw DgbException(DGB_ERR_UNSUPPORTED, "Np/Nfp/Nf do not describe a simplex Lagrange element of this order");
    if (!(d->nGeomEl == 1 || d->nGeomEl == d->nG) || !(d->nGeomF == 1 || d->nGeomF == d->nGf)) throw DgbException(DGB_ERR_ARG, "nGeomEl/nGeomF must be 1 or nG/nGf");
    if (d->Np > 255) throw DgbException(DGB_ERR_UNSUPPORTED, "more than 255 nodes per element");
    if (!(d->rho0 > 0) || !(d->c0 > 0)) throw DgbException(DGB_ERR_ARG, "rho0 and c0 must be positive");
}

struct HostOperators {
    std::vector<double> DwT, nLiftT, tiledOps;
    std::vector<int32_t> faceNodes;
    std::vector<real> Mf;
};

// Reference-element operators from the tables the reference's Mesh holds:
//   Mref_ij = sum_g w_g phi_i phi_j,  K^u_ij = sum_g w_g dphi_i/du phi_j,  Dw^u = Mref^-1 K^u,
//   Mf_nm = sum_g wf_g phif_n phif_m,  LIFT_lf = Mref^-1[:, faceNodes(lf)] Mf
HostOperators buildOperators(const dgb_desc* d) {
    const int Np = d->Np, Nfp = d->Nfp, Nf = d->Nf, dim = d->dim, nG = d->nG, nGf = d->nGf;
    HostOperators H;
    std::vector<real> Minv((size_t)Np * Np, 0);
    for (int g = 0; g < nG; ++g)
        for (int i = 0; i < Np; ++i) {
            const real wi = (real)d->elWeight[g] * d->elBasisFct[(size_t)g * Np + i];
            for (int j = 0; j < Np; ++j) Minv[(size_t)i * Np + j] += wi * d->elBasisFct[(size_t)g * Np + j];
        }
    invertDense(Minv, Np);
    H.DwT.assign((size_t)dim * Np * Np, 0.0);
    std::vector<real> Dw((size_t)dim * Np * Np, 0);
    for (int u = 0; u < dim; ++u) {
        std::vector<real> Ku((size_t)Np * Np, 0);
        for (int g = 0; g < nG; ++g)
            for (int i = 0; i < Np; ++i) {
                const real wi = (real)d->elWeight[g] * d->elUGradBasisFct[((size_t)g * Np + i) * 3 + u];
                for (int j = 0; j < Np; ++j) Ku[(size_t)i * Np + j] += wi * d->elBasisFct[(size_t)g * Np + j];
            }
        for (int i = 0; i < Np; ++i)
            for (int j = 0; j < Np; ++j) {
                real s = 0;
                for (int k = 0; k < Np; ++k) s += Minv[(size_t)i * Np + k] * Ku[(size_t)k * Np + j];
                Dw[((size_t)u * Np + i) * Np + j] = s;
                H.DwT[((size_t)u * Np + j) * Np + i] = (double)s;
            }
    }
    H.Mf.assign((size_t)Nfp * Nfp, 0);
    for (int g = 0; g < nGf; ++g)
        for (int n = 0; n < Nfp; ++n) {
            const real wn = (real)d->fWeight[g] * d->fBasisFct[(size_t)g * Nfp + n];
            for (int m = 0; m < Nfp; ++m) H.Mf[(size_t)n * Nfp + m] += wn * d->fBasisFct[(size_t)g * Nfp + m];
        }
    // local face-node table: element 0 is the first owner of each of its faces
    H.faceNodes.resize((size_t)Nf * Nfp);
    for (int lf = 0; lf < Nf; ++lf) {
        const int f = d->elFId[lf];
        if (d->fNbrElId[2 * (size_t)f] != 0) throw DgbException(DGB_ERR_ARG, "fNbrElId: element 0 must be the first owner of its faces");
        for (int m = 0; m < Nfp; ++m) H.faceNodes[lf * Nfp + m] = d->fNToElNId[((size_t)f * Nfp + m) * 2];
    }
    const int NFL = Nf * Nfp;
    H.nLiftT.assign((size_t)NFL * Np, 0.0);
    for (int i = 0; i < Np; ++i)
        for (int lf = 0; lf < Nf; ++lf)
            for (int m = 0; m < Nfp; ++m) {
                real s = 0;
                for (int n = 0; n < Nfp; ++n) s += Minv[(size_t)i * Np + H.faceNodes[lf * Nfp + n]] * H.Mf[(size_t)n * Nfp + m];
                H.nLiftT[((size_t)lf * Nfp + m) * Np + i] = (double)(-s);
            }
    // shared-memory image of the tiled DMMA kernel (stage_tiled.cu / tile_cfg.h)
    int npp = 0, ldq = 0, ldf = 0;
    if (tiledLayout(dim, d->order, &npp, &ldq, &ldf)) {
        H.tiledOps.assign((size_t)3 * npp * ldq + (size_t)npp * ldf, 0.0);
        for (int u = 0; u < 3; ++u)
            for (int i = 0; i < Np; ++i)
                for (int j = 0; j < Np; ++j) H.tiledOps[((size_t)u * npp + i) * ldq + j] = H.DwT[((size_t)u * Np + j) * Np + i];
        double* L = H.tiledOps.data() + (size_t)3 * npp * ldq;
        for (int i = 0; i < Np; ++i)
            for (int l = 0; l < NFL; ++l) L[(size_t)i * ldf + l] = H.nLiftT[(size_t)l * Np + i];
    }
    return H;
}

int representationOf(const dgb_handle* h, const StageKernel& k);
bool setupP2P(dgb_handle* h, bool required);

void createImpl(const dgb_desc* d, const int32_t* elPart, int rank, int nranks, const void* ncclId, dgb_handle** out) {
    if (!out) throw DgbException(DGB_ERR_ARG, "out is null");
    *out = nullptr;
    validate(d);
    if (nranks < 1 || rank < 0 || rank >= nranks) throw DgbException(DGB_ERR_ARG, "bad rank/nranks");
    if (nranks > 1 && (!elPart || !ncclId)) throw DgbException(DGB_ERR_ARG, "partitioned create needs elPart and an NCCL id");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        throw DgbException(DGB_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    const std::vector<uint8_t> elCurved = curvedElements(d);
    const bool curved = std::find(elCurved.begin(), elCurved.end(), (uint8_t)1) != elCurved.end();
    if (curved && nranks > 1) throw DgbException(DGB_ERR_UNSUPPORTED, "curved (non-affine) elements are not supported on partitioned handles yet");
    if (curved && (d->nGeomEl != d->nG || d->nGeomF != d->nGf))
        throw DgbException(DGB_ERR_ARG, "curved elements need one Jacobian / normal per integration point (nGeomEl == nG, nGeomF == nGf)");
    HostOperators H = buildOperators(d);

    dgb_handle* h = new dgb_handle;
    try {
        const int Np = d->Np, Nfp = d->Nfp, Nf = d->Nf, dim = d->dim, K = d->K;
        h->hd = *d;
        CUDA_CHECK(cudaGetDevice(&h->device));
        h->Kglobal = K;
        h->Np = Np;
        h->partitioned = nranks > 1;
        h->nranks = nranks;
        if (h->partitioned) {
            h->plan = makePartitionPlan(K, Nf, d->elFId, d->fNbrElId, elPart, rank, nranks);
        } else {
            h->plan.Kown = h->plan.Kinterior = K;
            h->plan.Khalo = 0;
        }
        const PartitionPlan& P = h->plan;
        auto toGlobal = [&](int l) { return h->partitioned ? P.localToGlobal[l] : l; };
        auto toLocal = [&](int g) { return h->partitioned ? P.globalToLocal[g] : g; };
        DeviceMesh& M = h->M;
        M.dim = dim; M.order = d->order; M.Np = Np; M.Nfp = Nfp; M.Nf = Nf; M.L = dim * Np + Nf * Nfp;
        M.Kown = P.Kown; M.Ktot = P.Kown + P.Khalo;
        M.stride = (int64_t)M.Ktot * Np;
        M.c0 = d->c0; M.rho0 = d->rho0; M.v0[0] = d->v0[0]; M.v0[1] = d->v0[1]; M.v0[2] = d->v0[2];

        // ---- per-element / per-face geometry in the device layout ----
        const int gE = d->nGeomEl, gF = d->nGeomF;
        std::vector<double> Ginv((size_t)M.Kown * dim * dim), fgeo((size_t)M.Kown * Nf * 4);
        std::vector<int32_t> fnbr((size_t)M.Kown * Nf), fflags((size_t)M.Kown * Nf);
        std::map<std::vector<uint8_t>, int> mapIds;
        std::vector<uint8_t> maps;
        std::vector<int> pos(Np, -1);
        std::map<std::vector<int>, bool> checkedPerm;
        for (int l = 0; l < M.Kown; ++l) {
            const int el = toGlobal(l);
            // inverse of A(r,c) = dx_c/du_r  ->  Ginv[x][u] = du_u/dx_x
            const double* J = &d->elJacobian[(size_t)el * gE * 9];
            real A[3][3], B[3][3];
            for (int r = 0; r < dim; ++r) for (int c = 0; c < dim; ++c) A[r][c] = J[r * 3 + c];
            if (dim == 1) B[0][0] = 1 / A[0][0];
            else if (dim == 2) {
                const real det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
                B[0][0] = A[1][1] / det; B[0][1] = -A[0][1] / det; B[1][0] = -A[1][0] / det; B[1][1] = A[0][0] / det;
            } else {
                const real det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                                 A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        const int r1 = (c + 1) % 3, r2 = (c + 2) % 3, c1 = (r + 1) % 3, c2 = (r + 2) % 3;
                        B[r][c] = (A[r1][c1] * A[r2][c2] - A[r1][c2] * A[r2][c1]) / det;
                    }
            }
            for (int x = 0; x < dim; ++x) for (int u = 0; u < dim; ++u) Ginv[(size_t)l * dim * dim + x * dim + u] = (double)B[x][u];
            const double detE = d->elJacobianDet[(size_t)el * gE];
            for (int lf = 0; lf < Nf; ++lf) {
                const int f = d->elFId[(size_t)el * Nf + lf];
                const int side = d->fNbrElId[2 * (size_t)f] == el ? 0 : 1;
                if (d->fNbrElId[2 * (size_t)f + side] != el) throw DgbException(DGB_ERR_ARG, "elFId/fNbrElId are inconsistent");
                const int o = d->elFOrientation[(size_t)el * Nf + lf];
                double* fg = &fgeo[((size_t)l * Nf + lf) * 4];
                for (int x = 0; x < 3; ++x) fg[x] = o * d->fNormal[(size_t)f * gF * 3 + x];
                fg[3] = d->fJacobianDet[(size_t)f * gF] / detE;
                int flags;
                if (d->fIsBoundary[f]) {
                    flags = d->fBC[f] == 1 ? FACE_REFLECTING : FACE_ABSORBING;
                    fnbr[(size_t)l * Nf + lf] = -1;
                } else {
                    const int nbG = d->fNbrElId[2 * (size_t)f + (1 - side)];
                    if (nbG < 0) throw DgbException(DGB_ERR_ARG, "interior face without a second owner");
                    const int nbL = toLocal(nbG);
                    if (nbL < 0) throw DgbException(DGB_ERR_STATE, "partition plan misses a halo element");
                    fnbr[(size_t)l * Nf + lf] = nbL;
                    const int tau = d->fc * o * (side == 0 ? 1 : -1);
                    flags = FACE_INTERIOR | (tau < 0 ? FLAG_TAU_NEG : 0);
                }
                // face-node pairing in the element's own face-node order
                std::fill(pos.begin(), pos.end(), -1);
                for (int m = 0; m < Nfp; ++m) pos[H.faceNodes[lf * Nfp + m]] = m;
                std::vector<uint8_t> mp(Nfp, 0);
                std::vector<int> perm(Nfp);
                for (int n = 0; n < Nfp; ++n) {
                    const int own = d->fNToElNId[((size_t)f * Nfp + n) * 2 + side];
                    if (own < 0 || own >= Np || pos[own] < 0) throw DgbException(DGB_ERR_ARG, "fNToElNId does not match the local face-node table");
                    perm[n] = pos[own];
                    const int nb = d->fIsBoundary[f] ? own : d->fNToElNId[((size_t)f * Nfp + n) * 2 + (1 - side)];
                    mp[pos[own]] = (uint8_t)nb;
                }
                // the element-local LIFT assumes the face mass matrix is invariant under this renumbering
                if (!checkedPerm.count(perm)) {
                    for (int a = 0; a < Nfp; ++a)
                        for (int b = 0; b < Nfp; ++b)
                            if (fabsl(H.Mf[(size_t)a * Nfp + b] - H.Mf[(size_t)perm[a] * Nfp + perm[b]]) > 1e-14L)
                                throw DgbException(DGB_ERR_UNSUPPORTED, "face node ordering is not a symmetry of the face element");
                    checkedPerm[perm] = true;
                }
                auto it = mapIds.find(mp);
                if (it == mapIds.end()) {
                    it = mapIds.emplace(mp, (int)mapIds.size()).first;
                    maps.insert(maps.end(), mp.begin(), mp.end());
                }
                if (it->second >= (1 << 20)) throw DgbException(DGB_ERR_UNSUPPORTED, "too many distinct face pairings");
                fflags[(size_t)l * Nf + lf] = flags | (it->second << FLAG_MAP_SHIFT);
            }
        }
        M.nMaps = (int)mapIds.size();
        M.DwT = devUpload(H.DwT);
        M.nLiftT = devUpload(H.nLiftT);
        M.tiledOps = devUpload(H.tiledOps);
        M.faceNodes = devUpload(H.faceNodes);
        M.nbrMaps = devUpload(maps);
        M.Ginv = devUpload(Ginv);
        M.fgeo = devUpload(fgeo);
        M.fnbr = devUpload(fnbr);
        M.fflags = devUpload(fflags);

        const size_t stateN = (size_t)4 * M.stride;
        h->U = devAlloc<double>(stateN);
        h->ACC = devAlloc<double>(stateN);
        h->YA = devAlloc<double>(stateN);
        h->YB = devAlloc<double>(stateN);
        CUDA_CHECK(cudaMemset(h->U, 0, stateN * sizeof(double)));
        CUDA_CHECK(cudaMemset(h->ACC, 0, stateN * sizeof(double)));
        CUDA_CHECK(cudaMemset(h->YA, 0, stateN * sizeof(double)));
        CUDA_CHECK(cudaMemset(h->YB, 0, stateN * sizeof(double)));
        CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreate(&h->evStart));
        CUDA_CHECK(cudaEventCreate(&h->evStop));
        h->stageEv.resize(128);
        for (auto& e : h->stageEv) CUDA_CHECK(cudaEventCreate(&e));

        h->generic = selectGenericKernel(dim, d->order);
        h->tiled = selectTiledKernel(dim, d->order);
        if (!h->generic.launch) throw DgbException(DGB_ERR_UNSUPPORTED, "no stage kernel for this dim/order");
        if (M.v0[0] == 0.0 && M.v0[1] == 0.0 && M.v0[2] == 0.0) h->ws = selectWsKernel(dim, d->order);
        h->active = h->autoKernel();
        if (curved) {
            // nothing collapses on curved elements: upload the reference's own tables and the inverse element mass matrices
            h->curved = true;
            CurvedMesh& C = h->CM;
            auto up = [&](const auto* src, size_t n) {
                typedef typename std::remove_const<typename std::remove_pointer<decltype(src)>::type>::type T;
                T* p = devAlloc<T>(n);
                h->curvedAllocs.push_back(p);
                CUDA_CHECK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
                return (const T*)p;
            };
            const size_t nG = d->nG, nGf = d->nGf, F = d->F;
            C.dim = dim; C.Np = Np; C.Nfp = Nfp; C.Nf = Nf; C.K = K; C.F = d->F; C.nG = d->nG; C.nGf = d->nGf; C.fc = d->fc;
            C.elBasis = up(d->elBasisFct, nG * Np); C.elUGrad = up(d->elUGradBasisFct, nG * Np * 3); C.elWeight = up(d->elWeight, nG);
            C.fBasis = up(d->fBasisFct, nGf * Nfp); C.fWeight = up(d->fWeight, nGf);
            C.elJac = up(d->elJacobian, (size_t)K * nG * 9); C.elDet = up(d->elJacobianDet, (size_t)K * nG);
            C.fNormal = up(d->fNormal, F * nGf * 3); C.fDet = up(d->fJacobianDet, F * nGf);
            C.elFId = up(d->elFId, (size_t)K * Nf); C.elFOrientation = up(d->elFOrientation, (size_t)K * Nf);
            C.fNbrElId = up(d->fNbrElId, F * 2); C.fNToElNId = up(d->fNToElNId, F * Nfp * 2);
            C.fIsBoundary = up(d->fIsBoundary, F); C.fBC = up(d->fBC, F);
            // straight-sided elements keep the collapsed kernels when the curved ones form a suffix of the numbering
            h->firstCurved = curvedSuffixStart(elCurved);
            const std::vector<double> Minv = curvedInverseMass(d, h->firstCurved);
            C.Minv = up(Minv.data(), Minv.size());
            C.firstCurved = h->firstCurved;
            C.c0 = d->c0; C.rho0 = d->rho0; C.v0[0] = d->v0[0]; C.v0[1] = d->v0[1]; C.v0[2] = d->v0[2];
            C.stride = M.stride;
            h->curvedName = h->firstCurved > 0 ? std::string(h->active.name) + " + stage_curved" : std::string("stage_curved");
        }
        // Bernstein-Bezier path (opt-in): conversion matrices, permutation tables, self-check of the closed-form lift
        h->bbKernel = curved ? StageKernel{} : selectBBKernel(dim, d->order);
        const StageKernel bb2Candidate = curved ? StageKernel{} : selectBB2Kernel(dim, d->order);
        if (curved) h->bbWhyNot = "curved elements";
        else if (!h->bbKernel.launch && !bb2Candidate.launch) h->bbWhyNot = "no Bernstein-Bezier kernel for this dimension / order (tetrahedra and triangles of orders 1..6)";
        else {
            try {
                const bb::Setup S = bb::buildSetup(d);
                double dev = 1.0;
                switch (d->order) {
                    case 1: dev = bb::liftDeviation<1>(S); break;
                    case 2: dev = bb::liftDeviation<2>(S); break;
                    case 3: dev = bb::liftDeviation<3>(S); break;
                    case 4: dev = bb::liftDeviation<4>(S); break;
                    case 5: dev = bb::liftDeviation<5>(S); break;
                    case 6: dev = bb::liftDeviation<6>(S); break;
                    default: break;
                }
                if (!(dev < 1e-11)) throw std::runtime_error("closed-form lift deviates from the dense one by " + std::to_string(dev));
                for (int lf = 0; lf < Nf; ++lf)
                    for (int m = 0; m < Nfp; ++m)
                        if (S.faceNodes[(size_t)lf * Nfp + m] != H.faceNodes[(size_t)lf * Nfp + m]) throw std::runtime_error("face-node tables disagree");
                h->hostV = S.V;
                h->dV = devUpload(S.V);
                h->dVinv = devUpload(S.Vinv);
                if (h->bbKernel.launch) {  // first generation (tetrahedra): permutation tables in constant memory
                    setBBTables(d->order, S.T);
                    CUDA_CHECK(cudaGetLastError());
                    h->bbSeqKernel = selectBBKernel(dim, d->order, 1);
                }
                // second generation: canonical coefficient order, interleaved fields
                h->bb2Kernel = bb2Candidate;
                h->bbeKernel = selectBBEKernel(dim, d->order);  // same representation, one thread per element (lowest orders)
                {
                    h->permG2C.assign(Np, 0);
                    for (int i = 0; i < Np; ++i) h->permG2C[S.T.permC2G[i]] = (uint8_t)i;
                    std::vector<uint8_t> tab((size_t)4 * Nfp + (size_t)M.nMaps * 4 * Nfp + Np, 0);
                    for (int J = 0; J < Nf; ++J)
                        for (int b = 0; b < Nfp; ++b) {
                            tab[(size_t)J * Nfp + b] = (uint8_t)S.ownIdx[(size_t)J * Nfp + b];
                            M.bbOwn[J][b] = tab[(size_t)J * Nfp + b];
                            for (int mp = 0; mp < M.nMaps; ++mp)
                                tab[(size_t)4 * Nfp + ((size_t)mp * 4 + J) * Nfp + b] = h->permG2C[maps[(size_t)mp * Nfp + S.T.facePos[J][b]]];
                        }
                    h->bbPermOffset = (size_t)4 * Nfp + (size_t)M.nMaps * 4 * Nfp;
                    std::copy(h->permG2C.begin(), h->permG2C.end(), tab.begin() + h->bbPermOffset);
                    // the neighbour table once more with padded, 16-byte aligned rows (the trace gathers of stage_bb2 load a row with 128-bit loads)
                    const size_t RS = (size_t)(Nfp + 15) / 16 * 16;
                    const size_t nbr16Offset = (tab.size() + 15) / 16 * 16;
                    tab.resize(nbr16Offset + (size_t)M.nMaps * 4 * RS, 0);
                    for (int mp = 0; mp < M.nMaps; ++mp)
                        for (int J = 0; J < 4; ++J)
                            for (int b = 0; b < Nfp; ++b)
                                tab[nbr16Offset + ((size_t)mp * 4 + J) * RS + b] = tab[(size_t)4 * Nfp + ((size_t)mp * 4 + J) * Nfp + b];
                    h->dBBTab = devUpload(tab);
                    M.bbTab = h->dBBTab;
                    M.bbNbr16 = h->dBBTab + nbr16Offset;  // cudaMalloc returns 256-byte aligned memory
                    for (int J = 0; J < Nf; ++J) M.bbFaceLf[J] = S.T.faceLf[J];
                    std::vector<double> VC((size_t)Np * Np), VinvC((size_t)Np * Np);
                    for (int n = 0; n < Np; ++n)
                        for (int i = 0; i < Np; ++i) {
                            VC[(size_t)n * Np + i] = S.V[(size_t)n * Np + S.T.permC2G[i]];
                            VinvC[(size_t)i * Np + n] = S.Vinv[(size_t)S.T.permC2G[i] * Np + n];
                        }
                    h->dVC = devUpload(VC);
                    h->dVinvC = devUpload(VinvC);
                }
            } catch (const std::exception& e) {
                h->bbWhyNot = e.what();
                h->bbKernel = StageKernel{};
                h->bb2Kernel = StageKernel{};
                h->bbeKernel = StageKernel{};
            }
        }

        if (!curved) {  // the automatic choice may be a Bernstein kernel: the (still empty) state then lives in its representation
            h->active = h->autoKernel();
            h->bbMode = representationOf(h, h->active);
        }
        if (h->partitioned) {
            CUDA_CHECK(cudaStreamCreateWithFlags(&h->commStream, cudaStreamNonBlocking));
            CUDA_CHECK(cudaEventCreateWithFlags(&h->evBorder, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&h->evRecv, cudaEventDisableTiming));
            ncclUniqueId id;
            std::memcpy(&id, ncclId, sizeof(id));
            NCCL_CHECK(nccl().CommInitRank(&h->comm, nranks, id, rank));
            h->sendBuf = devAlloc<double>((size_t)4 * P.sendElems.size() * Np);
            h->dSendElems = devUpload(P.sendElems);
            // Default exchange: direct peer-to-peer stores, fused into the stage kernel where the kernel supports it, if every
            // rank can map its peers' state arrays (CUDA IPC); otherwise NCCL send/recv. DGB_EXCHANGE=0/1/2 overrides.
            int want = 2;
            if (const char* e = getenv("DGB_EXCHANGE")) want = std::max(0, std::min(2, atoi(e)));
            if (want >= 1 && setupP2P(h, false)) h->exchangeMode = want;
        }
        *out = h;
    } catch (...) {
        freeHandle(h);
        throw;
    }
}

// ---------------------------------------------------------------------------------------------
// time loop
// ---------------------------------------------------------------------------------------------
int overlapMode(const dgb_handle* h);

void launchStage(dgb_handle* h, StageArgs A, int eBegin, int eEnd, bool timed) {
    if (eEnd <= eBegin) return;
    A.eBegin = eBegin;
    A.eEnd = eEnd;
    A.smReserve = (h->partitioned && h->exchangeMode == 0 && overlapMode(h) && eEnd <= h->plan.Kinterior) ? h->smReserve : 0;  // interior launches beside NCCL kernels only
    if (A.fx == nullptr) { A.fxWhich = 0; A.fxEpochWait = A.fxEpochSignal = 0; }
    const bool t = timed && h->timeStages && h->stageEvUsed + 2 <= (int)h->stageEv.size();
    if (t) cudaEventRecord(h->stageEv[h->stageEvUsed], h->stream);
    int nLaunched = 1;
    if (h->curved) {  // straight-sided prefix through the collapsed kernel, curved suffix through the reference's quadrature loops
        nLaunched = 0;
        StageArgs B = A;
        B.eEnd = std::min(eEnd, h->firstCurved);
        if (B.eEnd > B.eBegin) { h->active.launch(h->M, B, h->stream); ++nLaunched; }
        B = A;
        B.eBegin = std::max(eBegin, h->firstCurved);
        if (B.eEnd > B.eBegin) { launchCurved(h->CM, B, h->stream); ++nLaunched; }
    } else h->active.launch(h->M, A, h->stream);
    if (t) { cudaEventRecord(h->stageEv[h->stageEvUsed + 1], h->stream); h->stageEvUsed += 2; }
    h->launches += nLaunched;
}

void packHalo(dgb_handle* h, const double* produced, int nSendEl) {
    if (h->bbMode == 2) launchPackElementsBB2(produced, h->Np, h->dSendElems, nSendEl, h->sendBuf, h->stream);
    else launchPackElements(produced, h->M.stride, h->Np, h->dSendElems, nSendEl, h->sendBuf, h->stream);
}

// Halo exchange of array y (owned border elements -> the peers' halo slots), SURVEY §8 e1.
void exchangeHalo(dgb_handle* h, double* y, cudaStream_t s) {
    const PartitionPlan& P = h->plan;
    const int Np = h->Np;
    const int64_t S = h->M.stride;
    NCCL_CHECK(nccl().GroupStart());
    for (size_t i = 0; i < P.peers.size(); ++i) {
        const int peer = P.peers[i];
        const int nSend = P.sendOffset[i + 1] - P.sendOffset[i], nRecv = P.recvOffset[i + 1] - P.recvOffset[i];
        if (h->bbMode == 2) {  // interleaved coefficients: an element is one run of 4*Np doubles, one message per peer and direction
            const int64_t run = 4ll * Np;
            if (nSend > 0) NCCL_CHECK(nccl().Send(h->sendBuf + (int64_t)P.sendOffset[i] * run, (size_t)nSend * run, ncclDouble, peer, h->comm, s));
            if (nRecv > 0) NCCL_CHECK(nccl().Recv(y + (int64_t)(P.Kown + P.recvOffset[i]) * run, (size_t)nRecv * run, ncclDouble, peer, h->comm, s));
            continue;
        }
        const int64_t totalSend = (int64_t)P.sendElems.size() * Np;
        for (int q = 0; q < 4; ++q) {
            if (nSend > 0)
                NCCL_CHECK(nccl().Send(h->sendBuf + q * totalSend + (int64_t)P.sendOffset[i] * Np, (size_t)nSend * Np, ncclDouble, peer, h->comm, s));
            if (nRecv > 0)
                NCCL_CHECK(nccl().Recv(y + q * S + (int64_t)(P.Kown + P.recvOffset[i]) * Np, (size_t)nRecv * Np, ncclDouble, peer, h->comm, s));
        }
    }
    NCCL_CHECK(nccl().GroupEnd());
}

// ---- direct peer-to-peer exchange (opt-in) ---------------------------------------------------------------------------
// Collective over the handle's communicator: every rank moves U / YA / YB into one arena, publishes its CUDA IPC handle,
// the offset of the arena inside the underlying allocation, its array stride and, per peer, the first halo slot that
// peer's elements occupy; then maps the arenas of its peers.
struct P2PRecord {
    cudaIpcMemHandle_t mem;
    int64_t offset;      // arena - base of the allocation the IPC handle names
    int64_t stride;      // Ktot * Np of the publishing rank
    int64_t arrayBytes, flagOffset;
    int32_t slot0[MAX_PEERS];  // [r] first local element slot of rank r's halo elements (-1: not a peer)
    int32_t count[MAX_PEERS];  // [r] number of halo elements expected from rank r
    int32_t device;
    int32_t ok;          // this rank got as far as the record describes
};

// Returns false (after undoing everything, on every rank alike) if some rank cannot take part — no CUDA IPC in this
// environment, no memory for the arena — so that the caller can fall back to NCCL; throws if `required`. Every rank makes
// the same sequence of collective calls whatever happens locally.
bool setupP2P(dgb_handle* h, bool required) {
    if (!h->partitioned) throw DgbException(DGB_ERR_UNSUPPORTED, "exchange = 1 / 2 needs a partitioned handle");
    if (!h->peerMap.empty() || h->arena) return true;
    if (h->nranks > MAX_PEERS) {
        if (required) throw DgbException(DGB_ERR_UNSUPPORTED, "direct exchange supports at most 16 ranks");
        return false;
    }
    const PartitionPlan& P = h->plan;
    const int rank = P.rank, nranks = h->nranks;
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (h->commStream) CUDA_CHECK(cudaStreamSynchronize(h->commStream));
    // 1. arena: three state arrays (256-byte aligned) + one flag per rank — local work, failures are recorded, not thrown
    const size_t stateBytes = (size_t)4 * h->M.stride * sizeof(double);
    const size_t arrBytes = (stateBytes + 255) / 256 * 256;
    const size_t flagOff = 3 * arrBytes, total = flagOff + 256;
    char* arena = nullptr;
    double* old[3] = {h->U, h->YA, h->YB};
    std::vector<P2PRecord> rec(nranks);
    P2PRecord& me = rec[rank];
    std::memset(&me, 0, sizeof(me));
    std::string why;
    try {
        CUDA_CHECK(cudaMalloc(&arena, total));
        CUDA_CHECK(cudaMemset(arena, 0, total));
        for (int k = 0; k < 3; ++k) CUDA_CHECK(cudaMemcpy(arena + k * arrBytes, old[k], stateBytes, cudaMemcpyDeviceToDevice));
        CUDA_CHECK(cudaIpcGetMemHandle(&me.mem, arena));
        me.ok = 1;
    } catch (const DgbException& e) {
        why = e.what();
        cudaGetLastError();
    }
    if (me.ok) {
        // offset of the arena inside the allocation the handle names (cudaMalloc may sub-allocate small requests)
        typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        unsigned long long base = 0;
        size_t sz = 0;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn &&
            reinterpret_cast<GetRangeFn>(fn)(&base, &sz, (unsigned long long)(uintptr_t)arena) == 0 && base)
            me.offset = (int64_t)((unsigned long long)(uintptr_t)arena - base);
        else
            cudaGetLastError();
    }
    me.stride = h->M.stride;
    me.arrayBytes = (int64_t)arrBytes;
    me.flagOffset = (int64_t)flagOff;
    for (int r = 0; r < MAX_PEERS; ++r) { me.slot0[r] = -1; me.count[r] = 0; }
    for (size_t i = 0; i < P.peers.size(); ++i) {
        me.slot0[P.peers[i]] = P.Kown + P.recvOffset[i];
        me.count[P.peers[i]] = P.recvOffset[i + 1] - P.recvOffset[i];
    }
    CUDA_CHECK(cudaGetDevice(&me.device));
    // 2. publish (every rank, whatever happened above)
    P2PRecord* dRec = devAlloc<P2PRecord>(nranks);
    auto gather = [&]() {
        CUDA_CHECK(cudaMemcpy(dRec + rank, &me, sizeof(me), cudaMemcpyHostToDevice));
        NCCL_CHECK(nccl().AllGather(dRec + rank, dRec, sizeof(P2PRecord), ncclChar, h->comm, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        CUDA_CHECK(cudaMemcpy(rec.data(), dRec, sizeof(P2PRecord) * nranks, cudaMemcpyDeviceToHost));
    };
    auto allOk = [&]() { for (int r = 0; r < nranks; ++r) if (!rec[r].ok) return false; return true; };
    std::vector<dgb_handle::PeerMap> maps(P.peers.size());
    auto undo = [&]() {
        for (auto& pm : maps) if (pm.opened) { cudaIpcCloseMemHandle(pm.opened); pm.opened = nullptr; }
        if (arena) cudaFree(arena);
        arena = nullptr;
        cudaFree(dRec);
        cudaGetLastError();
    };
    try {
        gather();
    } catch (...) {
        undo();
        throw;
    }
    if (!allOk()) {
        undo();
        if (required) throw DgbException(DGB_ERR_UNSUPPORTED, "direct exchange unavailable: " + (why.empty() ? std::string("a peer could not export its state arrays") : why));
        return false;
    }
    // 3. map the peers; a failure is again agreed on collectively
    std::vector<int32_t> sendPeer(P.sendElems.size()), sendSlot(P.sendElems.size());
    P2PRecord mine = me;
    try {
        for (size_t i = 0; i < P.peers.size(); ++i) {
            const P2PRecord& pr = rec[P.peers[i]];
            const int nSend = P.sendOffset[i + 1] - P.sendOffset[i];
            if (pr.count[rank] != nSend || (nSend > 0 && pr.slot0[rank] < 0)) throw DgbException(DGB_ERR_STATE, "direct exchange: send / receive plans of two ranks disagree");
            CUDA_CHECK(cudaIpcOpenMemHandle(&maps[i].opened, pr.mem, cudaIpcMemLazyEnablePeerAccess));
            maps[i].base = static_cast<char*>(maps[i].opened) + pr.offset;
            maps[i].stride = pr.stride;
            maps[i].arrayBytes = (size_t)pr.arrayBytes;
            maps[i].flagOffset = (size_t)pr.flagOffset;
            const int64_t expect = ((int64_t)4 * pr.stride * (int64_t)sizeof(double) + 255) / 256 * 256;
            if (pr.arrayBytes != expect || pr.flagOffset != 3 * pr.arrayBytes) throw DgbException(DGB_ERR_STATE, "direct exchange: inconsistent arena layout");
            for (int k = P.sendOffset[i]; k < P.sendOffset[i + 1]; ++k) { sendPeer[k] = (int32_t)i; sendSlot[k] = pr.slot0[rank] + (k - P.sendOffset[i]); }
        }
    } catch (const DgbException& e) {
        why = e.what();
        mine.ok = 0;
        cudaGetLastError();
    }
    me = mine;
    try {
        gather();
    } catch (...) {
        undo();
        throw;
    }
    if (!allOk()) {
        undo();
        if (required) throw DgbException(DGB_ERR_UNSUPPORTED, "direct exchange unavailable: " + (why.empty() ? std::string("a peer could not map this rank's state arrays") : why));
        return false;
    }
    cudaFree(dRec);
    h->dSendPeer = devUpload(sendPeer);
    h->dSendSlot = devUpload(sendSlot);
    CUDA_CHECK(cudaHostAlloc(&h->p2pErr, sizeof(int), cudaHostAllocMapped));
    *h->p2pErr = 0;
    // 4. commit: the arena replaces the three separate arrays
    for (int k = 0; k < 3; ++k) cudaFree(old[k]);
    h->arena = arena;
    h->arenaArrayBytes = arrBytes;
    h->arenaFlagOffset = flagOff;
    h->U = h->phys[0] = reinterpret_cast<double*>(arena);
    h->YA = h->phys[1] = reinterpret_cast<double*>(arena + arrBytes);
    h->YB = h->phys[2] = reinterpret_cast<double*>(arena + 2 * arrBytes);
    h->dFlags = reinterpret_cast<unsigned long long*>(arena + flagOff);
    h->peerMap.swap(maps);
    h->epoch = 0;
    // 5. tables of the exchange fused into the stage kernel (exchange = 2)
    {
        const int nBorder = P.Kown - P.Kinterior;
        std::vector<std::vector<std::pair<int32_t, int32_t>>> targets(nBorder);
        for (size_t k = 0; k < P.sendElems.size(); ++k) {
            const int b = P.sendElems[k] - P.Kinterior;
            if (b < 0 || b >= nBorder) throw DgbException(DGB_ERR_STATE, "direct exchange: a send element is not a border element");
            targets[b].push_back({sendPeer[k], sendSlot[k]});
        }
        std::vector<int32_t> off(nBorder + 1, 0), pp, ps;
        for (int b = 0; b < nBorder; ++b) {
            for (auto& t : targets[b]) { pp.push_back(t.first); ps.push_back(t.second); }
            off[b + 1] = (int32_t)pp.size();
        }
        h->dPushOff = devUpload(off);
        h->dPushPeer = devUpload(pp);
        h->dPushSlot = devUpload(ps);
        h->dDone = devAlloc<unsigned int>(1);
        CUDA_CHECK(cudaMemset(h->dDone, 0, sizeof(unsigned int)));
        FusedHalo F{};
        F.Kinterior = P.Kinterior;
        F.nPeers = (int)P.peers.size();
        F.pushOff = h->dPushOff; F.pushPeer = h->dPushPeer; F.pushSlot = h->dPushSlot;
        for (size_t i = 0; i < P.peers.size(); ++i) {
            const dgb_handle::PeerMap& pm = h->peerMap[i];
            for (int k = 0; k < 3; ++k) F.arr[k][i] = reinterpret_cast<double*>(pm.base + (size_t)k * pm.arrayBytes);
            F.peerFlag[i] = reinterpret_cast<unsigned long long*>(pm.base + pm.flagOffset) + P.rank;
            F.waitRank[i] = P.peers[i];
        }
        F.myFlags = h->dFlags;
        F.doneCounter = h->dDone;
        F.timeoutNs = (unsigned long long)h->p2pTimeoutMs * 1000000ull;
        int* errDev = nullptr;
        CUDA_CHECK(cudaHostGetDevicePointer(&errDev, h->p2pErr, 0));
        F.err = errDev;
        h->hostFused = F;
        h->dFused = devAlloc<FusedHalo>(1);
        CUDA_CHECK(cudaMemcpy(h->dFused, &F, sizeof(F), cudaMemcpyHostToDevice));
    }
    return true;
}

// push the owned cut-adjacent elements of `produced` into the peers' halo slots and raise this rank's flag there
void pushHalo(dgb_handle* h, double* produced) {
    const PartitionPlan& P = h->plan;
    int which = -1;
    for (int k = 0; k < 3; ++k) if (produced == h->phys[k]) which = k;
    if (which < 0) throw DgbException(DGB_ERR_STATE, "direct exchange: produced array is not one of the exchanged arrays");
    PeerTargets T{};
    PeerFlags F{};
    F.n = (int)P.peers.size();
    for (size_t i = 0; i < P.peers.size(); ++i) {
        const dgb_handle::PeerMap& pm = h->peerMap[i];
        T.arr[i] = reinterpret_cast<double*>(pm.base + (size_t)which * pm.arrayBytes);
        T.stride[i] = pm.stride;
        F.flag[i] = reinterpret_cast<unsigned long long*>(pm.base + pm.flagOffset) + P.rank;
    }
    ++h->epoch;
    launchPushHalo(produced, h->M.stride, h->Np, h->dSendElems, h->dSendPeer, h->dSendSlot, (int)P.sendElems.size(), T, h->stream, h->bbMode == 2);
    launchSignalPeers(F, h->epoch, h->stream);
    h->launches += 2;
}

void waitHalo(dgb_handle* h) {
    const PartitionPlan& P = h->plan;
    PeerWait W{};
    W.n = (int)P.peers.size();
    for (size_t i = 0; i < P.peers.size(); ++i) W.rank[i] = P.peers[i];
    launchWaitPeers(h->dFlags, W, h->epoch, (unsigned long long)h->p2pTimeoutMs * 1000000ull, h->p2pErr, h->stream);
    ++h->launches;
}

// Effective overlap mode. Measured on B200 (config 5, profiles/r01w_scale_*.json): the persistent DMMA kernels pay more for a
// second, small launch per stage (prologue, pipeline fill / drain, static tile lists) than the ~0.09 ms exchange costs, at 2, 4
// and 8 GPUs alike, so they run one launch per stage and exchange afterwards; the light-weight generic kernel overlaps.
int overlapMode(const dgb_handle* h) {
    if (h->exchangeMode >= 1) return h->overlap >= 1 ? 1 : h->overlap == 0 ? 0 : (h->active.launch == h->generic.launch ? 1 : 0);  // no deferred order
    if (h->overlap >= 0) return h->overlap;
    return h->active.launch == h->generic.launch ? 1 : 0;
}

// One stage of a partitioned run. Only the cut-adjacent ("border") elements read halo values, so the exchange of the values
// produced by stage s can run beside the interior elements of stage s+1 (overlap 2):
//   main stream : interior(s+1) | wait recv(s) | border(s+1) | pack(s+1)
//   comm stream :   exchange(s) .............. |             |          | exchange(s+1) ...
// The DMMA kernels are persistent (one CTA per SM, static tile lists), so the interior launch leaves `smReserve` SMs to the
// NCCL kernels. Overlap 1 is the older order (border first, exchange beside the interior of the SAME stage): with the
// persistent kernels its small border launch and the late CTAs behind the NCCL kernels cost more than the exchange it hides.
void runStage(dgb_handle* h, const StageArgs& A, double* produced) {
    const PartitionPlan& P = h->plan;
    if (!h->partitioned) {
        launchStage(h, A, 0, h->M.Kown, true);
        return;
    }
    const int nSendEl = (int)P.sendElems.size();
    const int overlap = overlapMode(h);
    if (h->exchangeMode == 2 && h->active.launch == h->bb2Kernel.launch) {
        // exchange fused into the stage kernel: one launch over all owned elements; the border tiles (last in the interior-first
        // numbering) wait for the peers' flags of the previous stage, push their results into the peers' halo slots, and the
        // last CTA signals. Nothing else is launched.
        int which = -1;
        for (int k = 0; k < 3; ++k) if (produced == h->phys[k]) which = k;
        if (which < 0) throw DgbException(DGB_ERR_STATE, "direct exchange: produced array is not one of the exchanged arrays");
        StageArgs B = A;
        B.fx = h->dFused;
        B.fxWhich = which;
        B.fxEpochWait = h->epoch;
        B.fxEpochSignal = ++h->epoch;
        launchStage(h, B, 0, h->M.Kown, true);
        h->haloInFlight = true;
        return;
    }
    if (h->exchangeMode >= 1) {
        // direct stores into the peers' halo slots: the transfer needs no second stream, it drains over NVLink while the
        // interior elements run (overlap 1) or is simply waited for (overlap 0)
        if (overlap == 1) {
            launchStage(h, A, P.Kinterior, P.Kown, false);
            pushHalo(h, produced);
            launchStage(h, A, 0, P.Kinterior, true);
        } else {
            launchStage(h, A, 0, h->M.Kown, true);
            pushHalo(h, produced);
        }
        waitHalo(h);
        return;
    }
    if (overlap == 2) {
        launchStage(h, A, 0, P.Kinterior, true);
        if (h->recvPending) {
            CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->evRecv, 0));
            h->recvPending = false;
        }
        launchStage(h, A, P.Kinterior, P.Kown, false);
        packHalo(h, produced, nSendEl);
        ++h->launches;
        CUDA_CHECK(cudaEventRecord(h->evBorder, h->stream));
        CUDA_CHECK(cudaStreamWaitEvent(h->commStream, h->evBorder, 0));
        exchangeHalo(h, produced, h->commStream);
        CUDA_CHECK(cudaEventRecord(h->evRecv, h->commStream));
        h->recvPending = true;
    } else if (overlap == 1) {
        launchStage(h, A, P.Kinterior, P.Kown, false);
        packHalo(h, produced, nSendEl);
        ++h->launches;
        CUDA_CHECK(cudaEventRecord(h->evBorder, h->stream));
        CUDA_CHECK(cudaStreamWaitEvent(h->commStream, h->evBorder, 0));
        exchangeHalo(h, produced, h->commStream);
        CUDA_CHECK(cudaEventRecord(h->evRecv, h->commStream));
        launchStage(h, A, 0, P.Kinterior, true);
        CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->evRecv, 0));
    } else {
        launchStage(h, A, 0, h->M.Kown, true);
        packHalo(h, produced, nSendEl);
        ++h->launches;
        exchangeHalo(h, produced, h->stream);
    }
}

// Closes the deferred exchange of overlap 2 (before anything but a stage launch touches the halo slots)
void finishExchange(dgb_handle* h) {
    if (h->recvPending) {
        CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->evRecv, 0));
        h->recvPending = false;
    }
    if (h->haloInFlight) {  // fused exchange: the peers' pushes of the last stage are awaited by the next stage kernel; here by a wait kernel
        waitHalo(h);
        h->haloInFlight = false;
    }
}

// Representation of the resident state a stage kernel works on (dgb_handle::bbMode)
int representationOf(const dgb_handle* h, const StageKernel& k) {
    if (!k.launch) return 0;
    if (k.launch == h->bb2Kernel.launch || k.launch == h->bbeKernel.launch) return 2;
    if (k.launch == h->bbKernel.launch || k.launch == h->bbSeqKernel.launch) return 1;
    return 0;
}

// Converts what is resident (U, owned + halo elements) when the active kernel changes the representation. ACC is the
// scratch array of the layout-changing conversions (it is dead between steps: the first RK stage overwrites it).
void setRepresentation(dgb_handle* h, int want) {
    if (want == h->bbMode) return;
    finishExchange(h);
    if (h->stateSet) {
        const int64_t S = h->M.stride;
        const int Np = h->Np, K = h->M.Ktot;
        const size_t bytes = (size_t)4 * S * sizeof(double);
        if (h->bbMode == 1) { launchElementMatrix(h->U, h->U, S, Np, K, h->dV, h->stream); ++h->launches; }
        else if (h->bbMode == 2) {
            launchConvertBB2(h->U, h->ACC, S, Np, K, h->dVC, false, h->stream);
            CUDA_CHECK(cudaMemcpyAsync(h->U, h->ACC, bytes, cudaMemcpyDeviceToDevice, h->stream));
            ++h->launches;
        }
        if (want == 1) { launchElementMatrix(h->U, h->U, S, Np, K, h->dVinv, h->stream); ++h->launches; }
        else if (want == 2) {
            launchConvertBB2(h->U, h->ACC, S, Np, K, h->dVinvC, true, h->stream);
            CUDA_CHECK(cudaMemcpyAsync(h->U, h->ACC, bytes, cudaMemcpyDeviceToDevice, h->stream));
            ++h->launches;
        }
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    h->bbMode = want;
}

void runImpl(dgb_handle* h, int integrator, double t, int nsteps, double* tEnd) {
    if (!h) throw DgbException(DGB_ERR_ARG, "handle is null");
    if (!h->stateSet) throw DgbException(DGB_ERR_STATE, "dgb_run before dgb_set_state");
    if (nsteps < 0) throw DgbException(DGB_ERR_ARG, "nsteps < 0");
    if (integrator != DGB_EULER1 && integrator != DGB_RUNGE_KUTTA) throw DgbException(DGB_ERR_ARG, "unknown integrator");
    const double dt = h->hd.dt;
    if (h->nprobe > 0 && h->probeCount + nsteps > h->probeCap) {
        const int cap = std::max(h->probeCount + nsteps, 2 * h->probeCap);
        double* nb = devAlloc<double>((size_t)cap * h->nprobe * 4);
        if (h->probeCount) CUDA_CHECK(cudaMemcpyAsync(nb, h->dProbeRec, (size_t)h->probeCount * h->nprobe * 4 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (h->dProbeRec) cudaFree(h->dProbeRec);
        h->dProbeRec = nb;
        h->probeCap = cap;
    }
    if (h->nrecv > 0 && h->recvCount + nsteps > h->recvCap) {
        const int cap = std::max(h->recvCount + nsteps, 2 * h->recvCap);
        double* nb = devAlloc<double>((size_t)cap * h->nrecv * 4);
        if (h->recvCount) CUDA_CHECK(cudaMemcpyAsync(nb, h->dRecvRec, (size_t)h->recvCount * h->nrecv * 4 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (h->dRecvRec) cudaFree(h->dRecvRec);
        h->dRecvRec = nb;
        h->recvCap = cap;
    }
    h->stageEvUsed = 0;
    CUDA_CHECK(cudaEventRecord(h->evStart, h->stream));
    // Launch-bound case (the reference's own 2D configs: a stage kernel of a few microseconds, 4 000 launches per run): the
    // four stage launches of an RK4 step are captured once into a CUDA graph and replayed. The first step always runs
    // eagerly (kernel attributes are set on first use), t accumulates exactly as in the eager loop (solver.cpp:216).
    const bool graphable = !h->partitioned && h->ownStream && integrator == DGB_RUNGE_KUTTA && h->srcAmp.empty() && h->nprobe == 0 && h->nrecv == 0 &&
                           nsteps >= 4 && (h->useGraph == 1 || (h->useGraph < 0 && (int64_t)h->M.Kown * h->Np <= (1 << 18)));
    int step0 = 0;
    if (graphable) {
        auto stages = [&](bool timed) {
            StageArgs A{};
            A.u = h->U; A.acc = h->ACC; A.dt = dt;
            A.yin = h->U;  A.yout = h->YA; A.mode = MODE_RK1; launchStage(h, A, 0, h->M.Kown, timed);
            A.yin = h->YA; A.yout = h->YB; A.mode = MODE_RK2; launchStage(h, A, 0, h->M.Kown, timed);
            A.yin = h->YB; A.yout = h->YA; A.mode = MODE_RK3; launchStage(h, A, 0, h->M.Kown, timed);
            A.yin = h->YA; A.yout = nullptr; A.mode = MODE_RK4; launchStage(h, A, 0, h->M.Kown, timed);
        };
        stages(true);  // step 0, eager (and timed: dgb_last_stage_kernel_ms)
        t += dt;
        step0 = 1;
        if (!h->stepGraph || h->stepGraphDt != dt || h->stepGraphKernel != h->active.launch || h->stepGraphPtr[0] != h->U ||
            h->stepGraphPtr[1] != h->YA || h->stepGraphPtr[2] != h->YB) {
            if (h->stepGraph) { cudaGraphExecDestroy(h->stepGraph); h->stepGraph = nullptr; }
            cudaGraph_t graph = nullptr;
            const int64_t launchesBefore = h->launches;
            CUDA_CHECK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            stages(false);
            CUDA_CHECK(cudaStreamEndCapture(h->stream, &graph));
            h->launches = launchesBefore;  // captured, not launched
            CUDA_CHECK(cudaGraphInstantiate(&h->stepGraph, graph, 0));
            cudaGraphDestroy(graph);
            h->stepGraphDt = dt;
            h->stepGraphKernel = h->active.launch;
            h->stepGraphPtr[0] = h->U; h->stepGraphPtr[1] = h->YA; h->stepGraphPtr[2] = h->YB;
        }
        for (int step = step0; step < nsteps; ++step, t += dt) {
            CUDA_CHECK(cudaGraphLaunch(h->stepGraph, h->stream));
            h->launches += 4;
        }
        step0 = nsteps;
    }
    for (int step = step0; step < nsteps; ++step, t += dt) {
        if (h->nprobe > 0) {
            if (h->bbMode) launchGatherReceivers(h->U, h->M.stride, h->Np, h->dProbeElBB, h->bbMode == 2 ? h->dProbeWBB2 : h->dProbeWBB, h->nprobe,
                                                 h->dProbeRec + (size_t)h->probeCount * h->nprobe * 4, h->stream, h->bbMode == 2);
            else launchGatherProbes(h->U, h->M.stride, h->dProbeIdx, h->nprobe, h->dProbeRec + (size_t)h->probeCount * h->nprobe * 4, h->stream);
            ++h->probeCount;
            ++h->launches;
        }
        if (h->nrecv > 0) {  // owned elements only: no halo value is read
            launchGatherReceivers(h->U, h->M.stride, h->Np, h->dRecvEl, h->bbMode == 2 ? h->dRecvWBB2 : h->bbMode ? h->dRecvWBB : h->dRecvW, h->nrecv,
                                  h->dRecvRec + (size_t)h->recvCount * h->nrecv * 4, h->stream, h->bbMode == 2);
            ++h->recvCount;
            ++h->launches;
        }
        if (!h->srcAmp.empty()) finishExchange(h);  // sources overwrite halo copies of U too: the exchange of the last stage must have landed
        for (size_t s = 0; s < h->srcAmp.size(); ++s)
            if (t < h->srcDur[s]) {  // solver.cpp:253-255, evaluated on the host in the reference's own expression
                const double val = h->srcAmp[s] * sin(2 * M_PI * h->srcFreq[s] * t + h->srcPhase[s]);
                const int n = h->srcOff[s + 1] - h->srcOff[s];
                if (h->bbMode) launchSetNodesBB(h->U, h->Np, h->dSrcElList + h->srcElOff[s], h->dSrcNodeOff + h->srcElOff[s], h->dSrcNodeLocal,
                                                h->srcElOff[s + 1] - h->srcElOff[s] - 1 /* minus the closing entry */, val, h->dV, h->dVinv, h->stream,
                                                h->bbMode == 2 ? 4 : 1, h->bbMode == 2 ? h->dBBTab + h->bbPermOffset : nullptr);
                else launchSetNodes(h->U, h->dSrcIdx + h->srcOff[s], n, val, h->stream);
                if (n > 0) ++h->launches;
            }
        StageArgs A{};
        A.u = h->U; A.acc = h->ACC; A.dt = dt;
        if (integrator == DGB_EULER1) {
            A.yin = h->U; A.yout = h->YA; A.mode = MODE_EULER;
            runStage(h, A, h->YA);
            std::swap(h->U, h->YA);
            continue;
        }
        A.yin = h->U;  A.yout = h->YA; A.mode = MODE_RK1; runStage(h, A, h->YA);
        A.yin = h->YA; A.yout = h->YB; A.mode = MODE_RK2; runStage(h, A, h->YB);
        A.yin = h->YB; A.yout = h->YA; A.mode = MODE_RK3; runStage(h, A, h->YA);
        A.yin = h->YA; A.yout = nullptr; A.mode = MODE_RK4; runStage(h, A, h->U);
    }
    finishExchange(h);
    CUDA_CHECK(cudaEventRecord(h->evStop, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    CUDA_CHECK(cudaGetLastError());
    if (h->p2pErr && *h->p2pErr) {
        const int who = *h->p2pErr - 1;
        *h->p2pErr = 0;
        throw DgbException(DGB_ERR_STATE, "direct exchange: timed out waiting for the halo of rank " + std::to_string(who));
    }
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, h->evStart, h->evStop));
    h->lastRunMs = ms;
    double sum = 0;
    for (int i = 0; i + 1 < h->stageEvUsed; i += 2) {
        float m2 = 0;
        cudaEventElapsedTime(&m2, h->stageEv[i], h->stageEv[i + 1]);
        sum += m2;
    }
    h->lastStageMs = h->stageEvUsed ? sum / (h->stageEvUsed / 2) : 0.0;
    if (tEnd) *tEnd = t;
}

template <typename Fn>
int guarded(Fn fn) {
    try {
        fn();
        return DGB_OK;
    } catch (const DgbException& e) {
        g_err = e.what();
        return e.code;
    } catch (const UnsupportedError& e) {
        g_err = e.what();
        return DGB_ERR_UNSUPPORTED;
    } catch (const std::exception& e) {
        g_err = e.what();
        return DGB_ERR_ARG;
    } catch (...) {
        g_err = "unknown error";
        return DGB_ERR_ARG;
    }
}

// host <-> device state transfer; identity layout on one GPU, gather/scatter by element when partitioned
// Host threads of the gather / scatter loops: a fair share of the machine per rank, whatever OMP_NUM_THREADS says (torchrun
// sets it to 1, which would make the host side of dgb_set_state / dgb_get_state the slowest part of a partitioned run).
int hostThreads(const dgb_handle* h) {
    const unsigned hw = std::thread::hardware_concurrency();
    const int share = (int)(hw ? hw : 8) / std::max(1, h->nranks);
    return std::max(1, std::min(share, 16));
}

// Device-side address of a caller buffer if it is page-locked host memory (cudaMallocHost / cudaHostRegister /
// dgb_host_alloc), nullptr for pageable memory.
const double* deviceViewOfHost(const double* p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return nullptr;
    return static_cast<const double*>(attr.devicePointer);
}

double* stagingBuffer(dgb_handle* h) {
    if (!h->hostStage) CUDA_CHECK(cudaMallocHost(&h->hostStage, (size_t)4 * h->M.stride * sizeof(double)));
    return h->hostStage;
}

void stateToDevice(dgb_handle* h, const double* u, double* dst) {
    const int Np = h->Np;
    const int64_t Ng = (int64_t)h->Kglobal * Np, S = h->M.stride;
    if (!h->partitioned) {
        if (h->bbMode == 2) {  // nodal values land in ACC (dead between steps), the conversion writes the interleaved coefficients
            CUDA_CHECK(cudaMemcpyAsync(h->ACC, u, (size_t)4 * Ng * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            launchConvertBB2(h->ACC, dst, S, Np, h->M.Ktot, h->dVinvC, true, h->stream);
            ++h->launches;
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
            return;
        }
        CUDA_CHECK(cudaMemcpyAsync(dst, u, (size_t)4 * Ng * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (h->bbMode) { launchElementMatrix(dst, dst, S, Np, h->M.Ktot, h->dVinv, h->stream); ++h->launches; }  // nodal -> Bernstein
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        return;
    }
    const int Ktot = h->M.Ktot;
    if (const double* mapped = deviceViewOfHost(u)) {  // pinned caller buffer: the GPU reads it directly, element by element
        if (!h->dL2G) h->dL2G = devUpload(h->plan.localToGlobal);
        launchGatherState(mapped, Ng, h->dL2G, Ktot, Np, h->bbMode == 2 ? h->ACC : dst, S, h->stream);
        ++h->launches;
        if (h->bbMode == 2) { launchConvertBB2(h->ACC, dst, S, Np, Ktot, h->dVinvC, true, h->stream); ++h->launches; }
        else if (h->bbMode) { launchElementMatrix(dst, dst, S, Np, Ktot, h->dVinv, h->stream); ++h->launches; }
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        return;
    }
    double* stage = stagingBuffer(h);
    const int32_t* l2g = h->plan.localToGlobal.data();
    const int nth = hostThreads(h);
    // field by field: the copy of field q runs while field q+1 is gathered
    for (int q = 0; q < 4; ++q) {
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int l = 0; l < Ktot; ++l)
            std::memcpy(stage + (size_t)q * S + (size_t)l * Np, u + q * Ng + (int64_t)l2g[l] * Np, Np * sizeof(double));
        CUDA_CHECK(cudaMemcpyAsync((h->bbMode == 2 ? h->ACC : dst) + (size_t)q * S, stage + (size_t)q * S, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    if (h->bbMode == 2) { launchConvertBB2(h->ACC, dst, S, Np, Ktot, h->dVinvC, true, h->stream); ++h->launches; }
    else if (h->bbMode) { launchElementMatrix(dst, dst, S, Np, Ktot, h->dVinv, h->stream); ++h->launches; }  // owned + halo elements
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

void stateToHost(dgb_handle* h, const double* src, double* u) {
    const int Np = h->Np;
    const int64_t Ng = (int64_t)h->Kglobal * Np, S = h->M.stride;
    if (h->bbMode) {  // Bernstein -> nodal into ACC (dead between stages: MODE_RK1 overwrites it), then copy from there
        if (h->bbMode == 2) launchConvertBB2(src, h->ACC, S, Np, h->M.Kown, h->dVC, false, h->stream);
        else launchElementMatrix(src, h->ACC, S, Np, h->M.Kown, h->dV, h->stream);
        ++h->launches;
        src = h->ACC;
    }
    if (!h->partitioned) {
        CUDA_CHECK(cudaMemcpyAsync(u, src, (size_t)4 * Ng * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        return;
    }
    const int Kown = h->M.Kown;
    if (double* mapped = const_cast<double*>(deviceViewOfHost(u))) {  // pinned caller buffer: the GPU writes the owned elements into it
        if (!h->dL2G) h->dL2G = devUpload(h->plan.localToGlobal);
        launchScatterState(mapped, Ng, h->dL2G, Kown, Np, src, S, h->stream);
        ++h->launches;
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        return;
    }
    double* stage = stagingBuffer(h);
    const int32_t* l2g = h->plan.localToGlobal.data();
    const int nth = hostThreads(h);
    cudaEvent_t ev[4];
    for (int q = 0; q < 4; ++q) {
        CUDA_CHECK(cudaEventCreateWithFlags(&ev[q], cudaEventDisableTiming));
        CUDA_CHECK(cudaMemcpyAsync(stage + (size_t)q * S, src + (size_t)q * S, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_CHECK(cudaEventRecord(ev[q], h->stream));
    }
    // field by field: field q is scattered while field q+1 is still in flight
    for (int q = 0; q < 4; ++q) {
        CUDA_CHECK(cudaEventSynchronize(ev[q]));
        cudaEventDestroy(ev[q]);
#pragma omp parallel for schedule(static) num_threads(nth)
        for (int l = 0; l < Kown; ++l)  // owned elements only
            std::memcpy(u + q * Ng + (int64_t)l2g[l] * Np, stage + (size_t)q * S + (size_t)l * Np, Np * sizeof(double));
    }
}

// global DG node index -> local (or -1); halo copies included when withHalo
int localNode(const dgb_handle* h, int gnode, bool withHalo) {
    if (gnode < 0 || gnode >= h->Kglobal * h->Np) throw DgbException(DGB_ERR_ARG, "node index out of range");
    if (!h->partitioned) return gnode;
    const int el = gnode / h->Np, n = gnode - el * h->Np;
    const int l = h->plan.globalToLocal[el];
    if (l < 0 || (!withHalo && l >= h->M.Kown)) return -1;
    return l * h->Np + n;
}

}  // namespace

// =============================================================================================
// extern "C"
// =============================================================================================
extern "C" {

const char* dgb_last_error(void) { return g_err.c_str(); }
const char* dgb_version(void) { return "dgb 0.1 (sm_100a)"; }

int dgb_create(const dgb_desc* desc, dgb_handle** out) {
    return guarded([&] { createImpl(desc, nullptr, 0, 1, nullptr, out); });
}

int dgb_create_partitioned(const dgb_desc* desc, const int32_t* elPart, int rank, int nranks, const void* id, dgb_handle** out) {
    return guarded([&] { createImpl(desc, elPart, rank, nranks, id, out); });
}

int dgb_nccl_unique_id(void* out128) {
    return guarded([&] {
        if (!out128) throw DgbException(DGB_ERR_ARG, "out128 is null");
        ncclUniqueId id;
        NCCL_CHECK(nccl().GetUniqueId(&id));
        static_assert(sizeof(id) == 128, "ncclUniqueId is expected to be 128 bytes");
        std::memcpy(out128, &id, sizeof(id));
    });
}

void dgb_destroy(dgb_handle* h) { freeHandle(h); }

int dgb_set_state(dgb_handle* h, const double* u) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || !u) throw DgbException(DGB_ERR_ARG, "null argument");
        stateToDevice(h, u, h->U);
        h->stateSet = true;
    });
}

int dgb_get_state(dgb_handle* h, double* u) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || !u) throw DgbException(DGB_ERR_ARG, "null argument");
        if (!h->stateSet) throw DgbException(DGB_ERR_STATE, "dgb_get_state before dgb_set_state");
        stateToHost(h, h->U, u);
    });
}

int dgb_snapshot_begin(dgb_handle* h, double* u_host) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || !u_host) throw DgbException(DGB_ERR_ARG, "null argument");
        if (!h->stateSet) throw DgbException(DGB_ERR_STATE, "dgb_snapshot_begin before dgb_set_state");
        if (h->partitioned) throw DgbException(DGB_ERR_UNSUPPORTED, "asynchronous snapshots are not available on partitioned handles");
        const size_t n = (size_t)4 * h->M.stride;
        if (!h->dSnap) {
            h->dSnap = devAlloc<double>(n);
            CUDA_CHECK(cudaStreamCreateWithFlags(&h->snapStream, cudaStreamNonBlocking));
            CUDA_CHECK(cudaEventCreateWithFlags(&h->evSnapReady, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&h->evSnapDone, cudaEventDisableTiming));
        }
        if (h->snapPending) CUDA_CHECK(cudaStreamWaitEvent(h->stream, h->evSnapDone, 0));  // the previous copy still reads the buffer
        if (h->bbMode == 2) { launchConvertBB2(h->U, h->dSnap, h->M.stride, h->Np, h->M.Kown, h->dVC, false, h->stream); ++h->launches; }
        else if (h->bbMode) { launchElementMatrix(h->U, h->dSnap, h->M.stride, h->Np, h->M.Kown, h->dV, h->stream); ++h->launches; }  // Bernstein -> nodal
        else CUDA_CHECK(cudaMemcpyAsync(h->dSnap, h->U, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_CHECK(cudaEventRecord(h->evSnapReady, h->stream));
        CUDA_CHECK(cudaStreamWaitEvent(h->snapStream, h->evSnapReady, 0));
        CUDA_CHECK(cudaMemcpyAsync(u_host, h->dSnap, n * sizeof(double), cudaMemcpyDeviceToHost, h->snapStream));
        CUDA_CHECK(cudaEventRecord(h->evSnapDone, h->snapStream));
        h->snapPending = true;
    });
}

int dgb_snapshot_end(dgb_handle* h) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h) throw DgbException(DGB_ERR_ARG, "handle is null");
        if (!h->snapPending) throw DgbException(DGB_ERR_STATE, "dgb_snapshot_end without dgb_snapshot_begin");
        CUDA_CHECK(cudaEventSynchronize(h->evSnapDone));
        h->snapPending = false;
    });
}

void* dgb_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void dgb_host_free(void* p) { if (p) cudaFreeHost(p); }

int dgb_set_sources(dgb_handle* h, int nsrc, const int32_t* offsets, const int32_t* nodeIdx, const double* amp, const double* freq,
                    const double* phase, const double* duration) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || nsrc < 0 || (nsrc > 0 && (!offsets || !amp || !freq || !phase || !duration))) throw DgbException(DGB_ERR_ARG, "bad source arguments");
        std::vector<int32_t> off(1, 0), idx;
        for (int s = 0; s < nsrc; ++s) {
            for (int k = offsets[s]; k < offsets[s + 1]; ++k) {
                const int l = localNode(h, nodeIdx[k], true);  // halo copies are overwritten too, so no extra exchange is needed
                if (l >= 0) idx.push_back(l);
            }
            off.push_back((int32_t)idx.size());
        }
        if (h->dSrcIdx) { cudaFree(h->dSrcIdx); h->dSrcIdx = nullptr; }
        h->dSrcIdx = devUpload(idx);
        // the same node sets grouped by element, for the Bernstein mode (launchSetNodesBB): per source a run of elements, per
        // element a run of local nodes. nodeOff holds one shared, ever-growing offset array (entry b and b+1 bracket element b;
        // a closing entry after the last element of every source keeps the runs of two sources apart)
        {
            std::vector<int32_t> elList, nodeOff, nodeLocal, elOff(1, 0);
            for (int s = 0; s < nsrc; ++s) {
                std::vector<int32_t> nodes(idx.begin() + off[s], idx.begin() + off[s + 1]);
                std::sort(nodes.begin(), nodes.end());
                nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
                for (size_t k = 0; k < nodes.size(); ++k) {
                    const int el = nodes[k] / h->Np;
                    if (k == 0 || el != nodes[k - 1] / h->Np) { elList.push_back(el); nodeOff.push_back((int32_t)nodeLocal.size()); }
                    nodeLocal.push_back(nodes[k] - el * h->Np);
                }
                elList.push_back(-1);  // closing entry: nodeOff[b + 1] of the source's last element
                nodeOff.push_back((int32_t)nodeLocal.size());
                elOff.push_back((int32_t)elList.size());
            }
            for (int32_t** p : {&h->dSrcElList, &h->dSrcNodeOff, &h->dSrcNodeLocal}) if (*p) { cudaFree(*p); *p = nullptr; }
            h->dSrcElList = devUpload(elList);
            h->dSrcNodeOff = devUpload(nodeOff);
            h->dSrcNodeLocal = devUpload(nodeLocal);
            h->srcElOff = elOff;
        }
        h->srcOff = off;
        h->srcAmp.assign(amp, amp + nsrc);
        h->srcFreq.assign(freq, freq + nsrc);
        h->srcPhase.assign(phase, phase + nsrc);
        h->srcDur.assign(duration, duration + nsrc);
    });
}

int dgb_set_probes(dgb_handle* h, int nprobe, const int32_t* nodeIdx) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || nprobe < 0 || (nprobe > 0 && !nodeIdx)) throw DgbException(DGB_ERR_ARG, "bad probe arguments");
        std::vector<int32_t> idx(nprobe);
        for (int j = 0; j < nprobe; ++j) idx[j] = localNode(h, nodeIdx[j], false);
        if (h->dProbeIdx) { cudaFree(h->dProbeIdx); h->dProbeIdx = nullptr; }
        if (h->dProbeRec) { cudaFree(h->dProbeRec); h->dProbeRec = nullptr; }
        h->dProbeIdx = devUpload(idx);
        if (h->dProbeElBB) { cudaFree(h->dProbeElBB); h->dProbeElBB = nullptr; }
        if (h->dProbeWBB) { cudaFree(h->dProbeWBB); h->dProbeWBB = nullptr; }
        if (!h->hostV.empty()) {  // Bernstein mode: the nodal value is row n of V applied to the element's coefficients
            std::vector<int32_t> elBB(nprobe);
            std::vector<double> wBB((size_t)nprobe * h->Np, 0.0);
            for (int j = 0; j < nprobe; ++j) {
                elBB[j] = idx[j] >= 0 ? idx[j] / h->Np : -1;
                if (idx[j] >= 0) std::copy_n(&h->hostV[(size_t)(idx[j] % h->Np) * h->Np], h->Np, &wBB[(size_t)j * h->Np]);
            }
            h->dProbeElBB = devUpload(elBB);
            h->dProbeWBB = devUpload(wBB);
            if (h->dProbeWBB2) { cudaFree(h->dProbeWBB2); h->dProbeWBB2 = nullptr; }
            if (!h->permG2C.empty()) {
                std::vector<double> w2(wBB.size());
                for (int j = 0; j < nprobe; ++j)
                    for (int m = 0; m < h->Np; ++m) w2[(size_t)j * h->Np + h->permG2C[m]] = wBB[(size_t)j * h->Np + m];
                h->dProbeWBB2 = devUpload(w2);
            }
        }
        h->nprobe = nprobe;
        h->probeCap = h->probeCount = 0;
    });
}

int dgb_get_probes(dgb_handle* h, double* out, int capacity_steps, int* nsteps) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || !nsteps) throw DgbException(DGB_ERR_ARG, "null argument");
        if (capacity_steps < h->probeCount) {  // nothing is dropped: the record stays until a large enough buffer reads it
            *nsteps = h->probeCount;
            throw DgbException(DGB_ERR_ARG, "dgb_get_probes: capacity_steps " + std::to_string(capacity_steps) + " < recorded steps " + std::to_string(h->probeCount));
        }
        const int n = h->probeCount;
        if (n > 0 && h->nprobe > 0) {
            if (!out) throw DgbException(DGB_ERR_ARG, "out is null");
            CUDA_CHECK(cudaMemcpyAsync(out, h->dProbeRec, (size_t)n * h->nprobe * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
        *nsteps = n;
        h->probeCount = 0;
    });
}

int dgb_set_receivers(dgb_handle* h, int nrecv, const int32_t* el, const double* weights) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || nrecv < 0 || (nrecv > 0 && (!el || !weights))) throw DgbException(DGB_ERR_ARG, "bad receiver arguments");
        std::vector<int32_t> loc(nrecv);
        for (int j = 0; j < nrecv; ++j) {
            if (el[j] < 0 || el[j] >= h->Kglobal) throw DgbException(DGB_ERR_ARG, "receiver element out of range");
            const int l = h->partitioned ? h->plan.globalToLocal[el[j]] : el[j];
            loc[j] = (l >= 0 && l < h->M.Kown) ? l : -1;  // receivers of other ranks' elements record 0
        }
        std::vector<double> w(weights, weights + (size_t)nrecv * h->Np);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (h->dRecvEl) { cudaFree(h->dRecvEl); h->dRecvEl = nullptr; }
        if (h->dRecvW) { cudaFree(h->dRecvW); h->dRecvW = nullptr; }
        if (h->dRecvRec) { cudaFree(h->dRecvRec); h->dRecvRec = nullptr; }
        h->dRecvEl = devUpload(loc);
        h->dRecvW = devUpload(w);
        if (h->dRecvWBB) { cudaFree(h->dRecvWBB); h->dRecvWBB = nullptr; }
        if (!h->hostV.empty()) {  // Bernstein mode: sum_n w_n u_n = sum_m (V^T w)_m c_m
            std::vector<double> wBB((size_t)nrecv * h->Np, 0.0);
            for (int j = 0; j < nrecv; ++j)
                for (int n = 0; n < h->Np; ++n)
                    for (int m = 0; m < h->Np; ++m) wBB[(size_t)j * h->Np + m] += w[(size_t)j * h->Np + n] * h->hostV[(size_t)n * h->Np + m];
            h->dRecvWBB = devUpload(wBB);
            if (h->dRecvWBB2) { cudaFree(h->dRecvWBB2); h->dRecvWBB2 = nullptr; }
            if (!h->permG2C.empty()) {
                std::vector<double> w2(wBB.size());
                for (int j = 0; j < nrecv; ++j)
                    for (int m = 0; m < h->Np; ++m) w2[(size_t)j * h->Np + h->permG2C[m]] = wBB[(size_t)j * h->Np + m];
                h->dRecvWBB2 = devUpload(w2);
            }
        }
        h->nrecv = nrecv;
        h->recvCap = h->recvCount = 0;
    });
}

int dgb_get_receivers(dgb_handle* h, double* out, int capacity_steps, int* nsteps) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || !nsteps) throw DgbException(DGB_ERR_ARG, "null argument");
        if (capacity_steps < h->recvCount) {
            *nsteps = h->recvCount;
            throw DgbException(DGB_ERR_ARG, "dgb_get_receivers: capacity_steps " + std::to_string(capacity_steps) + " < recorded steps " + std::to_string(h->recvCount));
        }
        const int n = h->recvCount;
        if (n > 0 && h->nrecv > 0) {
            if (!out) throw DgbException(DGB_ERR_ARG, "out is null");
            CUDA_CHECK(cudaMemcpyAsync(out, h->dRecvRec, (size_t)n * h->nrecv * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
        *nsteps = n;
        h->recvCount = 0;
    });
}

int dgb_run(dgb_handle* h, int integrator, double t_start, int nsteps, double* t_end) {
    return guarded([&] { DeviceScope onDevice(h); runImpl(h, integrator, t_start, nsteps, t_end); });
}

int dgb_eval_rhs(dgb_handle* h, const double* u, double* rhs) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || !u || !rhs) throw DgbException(DGB_ERR_ARG, "null argument");
        stateToDevice(h, u, h->YA);
        StageArgs A{};
        A.yin = h->YA; A.u = h->U; A.acc = h->ACC; A.yout = h->YB; A.mode = MODE_RHS; A.dt = 1.0;
        launchStage(h, A, 0, h->M.Kown, false);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        CUDA_CHECK(cudaGetLastError());
        stateToHost(h, h->YB, rhs);
    });
}

int dgb_set_stream(dgb_handle* h, void* cuda_stream) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h) throw DgbException(DGB_ERR_ARG, "handle is null");
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (h->ownStream && h->stream) cudaStreamDestroy(h->stream);
        h->stream = (cudaStream_t)cuda_stream;
        h->ownStream = false;
    });
}

int dgb_synchronize(dgb_handle* h) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h) throw DgbException(DGB_ERR_ARG, "handle is null");
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        if (h->commStream) CUDA_CHECK(cudaStreamSynchronize(h->commStream));
    });
}

double dgb_last_run_ms(dgb_handle* h) { return h ? h->lastRunMs : 0.0; }
double dgb_last_stage_kernel_ms(dgb_handle* h) { return h ? h->lastStageMs : 0.0; }
int64_t dgb_launch_count(dgb_handle* h) { return h ? h->launches : 0; }
const char* dgb_kernel_name(dgb_handle* h) { return !h ? "none" : h->curved ? h->curvedName.c_str() : h->active.name; }

int dgb_set_option(dgb_handle* h, const char* key, int value) {
    return guarded([&] {
        DeviceScope onDevice(h);
        if (!h || !key) throw DgbException(DGB_ERR_ARG, "null argument");
        const std::string k(key);
        if (k == "kernel") {
            if (h->curved && (h->firstCurved == 0 ? value != 0 : value > 3))
                throw DgbException(DGB_ERR_UNSUPPORTED, "curved meshes: the curved elements run the curved-element kernel, the straight-sided ones kernels 0..3");
            if (value == 1) h->active = h->generic;
            else if (value == 2) {
                if (!h->tiled.launch) throw DgbException(DGB_ERR_UNSUPPORTED, "no tiled kernel for this dim/order");
                h->active = h->tiled;
            } else if (value == 3) {
                if (!h->ws.launch) throw DgbException(DGB_ERR_UNSUPPORTED, "no warp-specialised kernel for this dim/order/mean flow");
                h->active = h->ws;
            } else if (value == 4 || value == 5) {  // 5: the face-sequential schedule of the same arithmetic
                if (!h->bbKernel.launch)
                    throw DgbException(DGB_ERR_UNSUPPORTED, "first-generation Bernstein-Bezier kernel unavailable: " +
                                                                (h->bbWhyNot.empty() ? std::string("tetrahedra of orders 2..5 only (kernel 6 covers this mesh)") : h->bbWhyNot));
                h->active = value == 4 ? h->bbKernel : h->bbSeqKernel;
            } else if (value == 6) {
                if (!h->bb2Kernel.launch) throw DgbException(DGB_ERR_UNSUPPORTED, "Bernstein-Bezier kernel unavailable: " + h->bbWhyNot);
                h->active = h->bb2Kernel;
            } else if (value == 7) {
                if (!h->bbeKernel.launch)
                    throw DgbException(DGB_ERR_UNSUPPORTED, "element-per-thread Bernstein-Bezier kernel unavailable: " +
                                                                (h->bbWhyNot.empty() ? std::string("triangles of orders 1..3 and tetrahedra of orders 1 / 2 only") : h->bbWhyNot));
                h->active = h->bbeKernel;
            } else h->active = h->autoKernel();
            // the Bernstein kernel keeps the state as Bernstein coefficients: convert what is resident when the representation changes
            if (h->curved) h->curvedName = h->firstCurved > 0 ? std::string(h->active.name) + " + stage_curved" : std::string("stage_curved");
            setRepresentation(h, representationOf(h, h->active));
        } else if (k == "overlap") {
            if (value < -1 || value > 2) throw DgbException(DGB_ERR_ARG, "overlap must be -1 (automatic), 0, 1 or 2");
            h->overlap = value;
        }
        else if (k == "exchange") {  // collective: every rank of the communicator must make the same call
            if (value < 0 || value > 2)
                throw DgbException(DGB_ERR_ARG, "exchange must be 0 (NCCL send/recv), 1 (direct peer-to-peer stores) or 2 (stores fused into the stage kernel)");
            finishExchange(h);
            if (value >= 1) setupP2P(h, true);
            h->exchangeMode = value;
        } else if (k == "p2p_timeout_ms") {
            h->p2pTimeoutMs = std::max(1, value);
            if (h->dFused) {
                h->hostFused.timeoutNs = (unsigned long long)h->p2pTimeoutMs * 1000000ull;
                CUDA_CHECK(cudaStreamSynchronize(h->stream));
                CUDA_CHECK(cudaMemcpy(h->dFused, &h->hostFused, sizeof(FusedHalo), cudaMemcpyHostToDevice));
            }
        }
        else if (k == "bb_tile") {  // elements per CTA of the Bernstein kernels; takes effect at once if one of them is active
            if (value != 8 && value != 16 && value != 32) throw DgbException(DGB_ERR_ARG, "bb_tile must be 8, 16 or 32");
            if (h->bbKernel.launch) {
                const bool a = h->active.launch == h->bbKernel.launch, b = h->active.launch == h->bbSeqKernel.launch;
                h->bbKernel = selectBBKernel(h->M.dim, h->M.order, 0, value);
                h->bbSeqKernel = selectBBKernel(h->M.dim, h->M.order, 1, value);
                if (a) h->active = h->bbKernel;
                if (b) h->active = h->bbSeqKernel;
            }
            h->bbTile = value;
        }
        else if (k == "sm_reserve") h->smReserve = std::max(0, value);
        else if (k == "graph") h->useGraph = value < 0 ? -1 : (value ? 1 : 0);
        else if (k == "time_stages") h->timeStages = value ? 1 : 0;
        else throw DgbException(DGB_ERR_ARG, "unknown option " + k);
    });
}

double dgb_measure_fp64_tflops(dgb_handle* h) {
    if (!h) return 0.0;
    DeviceScope onDevice(h);
    cudaStreamSynchronize(h->stream);
    return measureFp64Tflops(h->stream);
}

int dgb_get_option(dgb_handle* h, const char* key, int* value) {
    return guarded([&] {
        if (!h || !key || !value) throw DgbException(DGB_ERR_ARG, "null argument");
        const std::string k(key);
        auto is = [&](const StageKernel& s) { return s.launch && h->active.launch == s.launch; };
        if (k == "kernel") *value = is(h->bbeKernel) ? 7 : is(h->bb2Kernel) ? 6 : is(h->bbSeqKernel) ? 5 : is(h->bbKernel) ? 4 : is(h->ws) ? 3 : is(h->tiled) ? 2 : 1;
        else if (k == "exchange") *value = h->partitioned ? h->exchangeMode : 0;
        else if (k == "representation") *value = h->bbMode;
        else if (k == "overlap") *value = h->partitioned ? overlapMode(h) : 0;
        else if (k == "graph") *value = h->useGraph;
        else if (k == "bb_tile") *value = h->bbTile;
        else if (k == "sm_reserve") *value = h->smReserve;
        else if (k == "p2p_timeout_ms") *value = h->p2pTimeoutMs;
        else throw DgbException(DGB_ERR_ARG, "unknown option " + k);
    });
}

}  // extern "C"
