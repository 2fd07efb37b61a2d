// Internal declarations shared by the C-ABI layer (dgb_api.cu) and the kernels (stage_*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/dgb.h"
#include "bb_ops.h"

namespace dgb {

// What a stage launch does with k = dt * L(yin) (the 3-register RK4 that reproduces the
// reference's (k1+2*k2+2*k3+k4)/6.0 expression bit for bit given identical k's, solver.cpp:261-285).
enum StageMode : int {
    MODE_RK1 = 0,    // acc  = k            ; yout = u + 0.5*k      (yin == u)
    MODE_RK2 = 1,    // acc += 2*k          ; yout = u + 0.5*k
    MODE_RK3 = 2,    // acc += 2*k          ; yout = u + 1.0*k
    MODE_RK4 = 3,    // u   += (acc + k)/6.0
    MODE_EULER = 4,  // yout = yin + k                              (solver.cpp:152-153, beta = 1)
    MODE_RHS = 5     // yout = L(yin)  (dt ignored)                 (dgb_eval_rhs)
};

// Face flags (one int32 per element-local face)
constexpr int FACE_INTERIOR = 0, FACE_ABSORBING = 1, FACE_REFLECTING = 2;
constexpr int FLAG_BC_MASK = 0x3, FLAG_TAU_NEG = 0x4, FLAG_MAP_SHIFT = 8;

struct DeviceMesh {
    int dim, order, Np, Nfp, Nf, L;   // L = dim*Np + Nf*Nfp : contraction length of the fused operator
    int Kown, Ktot;                   // owned elements, owned + halo
    int64_t stride;                   // Ktot*Np : distance between two fields of a state array
    // operators (device)
    double* DwT;      // [dim][Np(j)][Np(i)]  transposed differentiation/stiffness operators, Dw^u = Mref^-1 K^u
    double* nLiftT;   // [Nf*Nfp(l)][Np(i)]   -Mref^-1 E_lf Mf, transposed
    double* tiledOps; // [3][NPP][LDQ] Dw^u then [NPP][LDF] -LIFT, zero padded: the shared-memory image of the tiled kernel
    int32_t* faceNodes;  // [Nf][Nfp] element-local node of face node m
    uint8_t* nbrMaps;    // [nMaps][Nfp] neighbour's local node for face node m, de-duplicated patterns
    int nMaps;
    // geometry (device)
    double* Ginv;     // [Kown][dim*dim]  Ginv[x*dim+u] = d u_u / d x_x
    double* fgeo;     // [Kown][Nf][4]    outward unit normal (3), Fscale = detJ_f/detJ_e
    int32_t* fnbr;    // [Kown][Nf]       neighbour element (local numbering, may point into the halo), -1 on the boundary
    int32_t* fflags;  // [Kown][Nf]       bc type | tau sign | map id
    // physics
    double c0, rho0, v0[3];
    // second-generation Bernstein kernel (stage_bb2.cu): byte tables in canonical order — [4][Nfp] coefficient of face J's
    // 2D index b, then [nMaps][4][Nfp] the neighbour's coefficient for (face-pairing map, face J, b) — and the mesh's
    // local face of canonical face J
    const uint8_t* bbTab;
    const uint8_t* bbNbr16;  // [nMaps][4][RS] the second table again with rows padded to RS = 16*ceil(Nfp/16) bytes, 16-byte aligned (one or two 128-bit loads per row)
    uint8_t bbFaceLf[4];
    uint8_t bbOwn[4][28];  // the first table again, in the kernel's parameter space (uniform run-time index)
};

constexpr int MAX_PEERS = 16;
// Halo exchange fused into the stage kernel (stage_bb2.cu, dgb_set_option("exchange", 2)): device-resident tables of a
// partitioned handle. The kernel stores the results of its cut-adjacent elements straight into the peers' halo slots
// (CUDA IPC mappings of the peers' state arrays), the last CTA raises this rank's epoch flag at every peer, and the tiles
// that read halo values wait for the peers' flags of the previous stage.
struct FusedHalo {
    int Kinterior, nPeers;
    const int32_t* pushOff;    // [Kown - Kinterior + 1] range of (peer, slot) targets of border element Kinterior + k
    const int32_t* pushPeer;   // index into the peer tables below
    const int32_t* pushSlot;   // element slot in that peer's arrays
    double* arr[3][MAX_PEERS];                 // the peers' copies of the three exchanged arrays (U, YA, YB by allocation)
    unsigned long long* peerFlag[MAX_PEERS];   // this rank's slot in each peer's flag array
    const unsigned long long* myFlags;         // this rank's flag array, indexed by rank
    int waitRank[MAX_PEERS];
    unsigned int* doneCounter;                 // CTAs that have finished (and whose pushes are complete) in the running launch
    unsigned long long timeoutNs;
    int* err;                                  // pinned + mapped: 1 + rank of a peer that never signalled
};

struct StageArgs {
    const double* yin;   // stage input  [4][stride]
    double* u;           // solution     [4][stride]
    double* acc;         // k-sum        [4][stride]
    double* yout;        // next input / Euler or RHS output
    int eBegin, eEnd;    // element range of this launch
    int mode;
    double dt;
    int smReserve;       // persistent kernels leave this many SMs free (halo exchange kernels running beside an interior launch)
    const FusedHalo* fx;                         // fused halo exchange (nullptr: none)
    int fxWhich;                                 // which of the three exchanged arrays this launch produces
    unsigned long long fxEpochWait, fxEpochSignal;
};

// Returns a printable kernel name; launches on `stream`. kernelChoice: 0 auto, 1 generic, 2 tiled.
typedef void (*StageLaunchFn)(const DeviceMesh&, const StageArgs&, cudaStream_t);
struct StageKernel {
    StageLaunchFn launch = nullptr;
    const char* name = "none";
};
StageKernel selectGenericKernel(int dim, int order);
StageKernel selectTiledKernel(int dim, int order);  // launch == nullptr if no tiled instance exists
StageKernel selectWsKernel(int dim, int order);     // warp-specialised DMMA kernel, zero mean flow only (stage_ws.cu)
StageKernel selectBBKernel(int dim, int order, int variant = 0, int tile = 32);  // Bernstein-Bezier sparse-operator kernel, tetrahedra (stage_bb.cu); the state holds Bernstein coefficients; variant 1: the faces one after the other; tile: elements per CTA (32, 16, 8)
// permutation tables of the Bernstein kernel for one order (constant memory of the current device; they depend only on the
// element's node numbering convention, so one copy per order serves every handle of the process)
void setBBTables(int order, const bb::Tables& T);
// hard source in Bernstein mode: elements elList[0..nEl) get the nodal value `value` at their local nodes nodeLocal[nodeOff[b]..nodeOff[b+1])
// coefStride / perm describe where coefficient m (mesh node order, the order of V / Vinv) of an element lives:
// field[(el*Np + (perm ? perm[m] : m)) * coefStride]
void launchSetNodesBB(double* field, int Np, const int32_t* elList, const int32_t* nodeOff, const int32_t* nodeLocal, int nEl, double value,
                      const double* V, const double* Vinv, cudaStream_t s, int coefStride = 1, const uint8_t* perm = nullptr);
// second-generation Bernstein kernel (stage_bb2.cu): interleaved canonical coefficient layout c[(el*Np + i)*4 + q]
StageKernel selectBB2Kernel(int dim, int order);
// the same representation, one thread per element (stage_bbe.cu): triangles of orders 1..3, tetrahedra of order 1
StageKernel selectBBEKernel(int dim, int order);
void launchConvertBB2(const double* in, double* out, int64_t stride, int Np, int K, const double* mat, bool toBB, cudaStream_t s);
void launchPackElementsBB2(const double* y, int Np, const int32_t* elems, int n, double* buf, cudaStream_t s);
// y = Mat x per element and field over a whole state array (nodal <-> Bernstein conversion); in and out may alias
void launchElementMatrix(const double* in, double* out, int64_t stride, int Np, int K, const double* mat, cudaStream_t s);

struct CurvedMesh;  // curved_setup.h
void launchCurved(const CurvedMesh& C, const StageArgs& A, cudaStream_t s);  // stage_curved.cu: every element through the reference's own loops
void launchSetNodes(double* field, const int32_t* idx, int n, double value, cudaStream_t s);
void launchGatherProbes(const double* u, int64_t stride, const int32_t* idx, int n, double* out, cudaStream_t s);
// interleaved != 0: u is an interleaved coefficient array c[(el*Np + m)*4 + q] (stage_bb2.cu) and w is in its coefficient order
void launchGatherReceivers(const double* u, int64_t stride, int Np, const int32_t* el, const double* w, int n, double* out, cudaStream_t s,
                           int interleaved = 0);
void launchPackElements(const double* y, int64_t stride, int Np, const int32_t* elems, int n, double* buf, cudaStream_t s);
void launchUnpackElements(double* y, int64_t stride, int Np, int firstElem, int n, const double* buf, cudaStream_t s);

// partitioned handles, pinned caller buffers: gather / scatter between the caller's global array and the rank's local order (state_io.cu)
void launchGatherState(const double* globalHost, int64_t Ng, const int32_t* l2g, int K, int Np, double* local, int64_t stride, cudaStream_t s);
void launchScatterState(double* globalHost, int64_t Ng, const int32_t* l2g, int K, int Np, const double* local, int64_t stride, cudaStream_t s);

double measureFp64Tflops(cudaStream_t s);  // instrumentation: DFMA issue peak of the current device (state_io.cu)

// direct peer-to-peer halo exchange (halo_p2p.cu); tables are passed by value as kernel arguments
struct PeerTargets { double* arr[MAX_PEERS]; long long stride[MAX_PEERS]; };   // the peers' copy of the produced array, [4][stride]
struct PeerFlags { unsigned long long* flag[MAX_PEERS]; int n; };              // THIS rank's slot in each peer's flag array
struct PeerWait { int rank[MAX_PEERS]; int n; };                               // ranks whose flags this rank waits for
void launchPushHalo(const double* y, int64_t stride, int Np, const int32_t* sendElems, const int32_t* sendPeer, const int32_t* sendSlot,
                    int nSend, const PeerTargets& T, cudaStream_t s, int interleaved = 0);
void launchSignalPeers(const PeerFlags& F, unsigned long long epoch, cudaStream_t s);
void launchWaitPeers(const unsigned long long* flags, const PeerWait& W, unsigned long long epoch, unsigned long long timeoutNs, int* err,
                     cudaStream_t s);

}  // namespace dgb
