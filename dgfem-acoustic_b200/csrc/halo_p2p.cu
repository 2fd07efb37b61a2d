// Direct peer-to-peer halo exchange over NVLink (the lower-latency alternative to ncclSend/ncclRecv):
// the sender stores the nodal values of its cut-adjacent elements straight into the halo slots of the peers' state
// arrays (CUDA IPC mappings made once by dgb_set_option("exchange", 1)), then raises a per-(sender, receiver) epoch
// flag in the peer's memory; the receiver's stream waits for the flags of all its peers before the next stage reads
// the halo. No staging buffer, no NCCL kernels, three tiny launches per stage.
//
// Ordering. push -> (kernel boundary: all stores of the push kernel are performed at system scope) -> signal kernel:
// fence.sys + st.release.sys of the epoch. wait kernel: ld.acquire.sys until flag >= epoch, then the next stage kernel
// of the same stream reads the halo through freshly invalidated L1. A slot is never overwritten before its reader is
// done: the array a rank pushes into at stage s+1 was last read by the peer at a stage <= s, and the sender's stage
// s+1 cannot start before the peer has pushed (hence finished) its stage s (the neighbour relation is symmetric).
#include "dgb_internal.h"
#include "dgb_launch.h"  // the CPU tests run this file under oracle/cuda_emu.h (DGB_EMULATE)

namespace dgb {
namespace {

__global__ void pushHaloKernel(const double* __restrict__ y, int64_t stride, int Np, const int32_t* __restrict__ sendElems,
                               const int32_t* __restrict__ sendPeer, const int32_t* __restrict__ sendSlot, int nSend, PeerTargets T, int interleaved) {
    const int64_t per = (int64_t)nSend * Np;
    if (interleaved) {  // c[(el*Np + i)*4 + q] (stage_bb2.cu): an element is one run of 4*Np doubles
        const int run = 4 * Np;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * per; i += (int64_t)gridDim.x * blockDim.x) {
            const int k = (int)(i / run), r = (int)(i - (int64_t)k * run);
            T.arr[sendPeer[k]][(int64_t)sendSlot[k] * run + r] = y[(int64_t)sendElems[k] * run + r];
        }
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * per; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i / per);
        const int64_t r = i - q * per;
        const int k = (int)(r / Np), nd = (int)(r - (int64_t)k * Np);
        const int p = sendPeer[k];
        T.arr[p][q * T.stride[p] + (int64_t)sendSlot[k] * Np + nd] = y[q * stride + (int64_t)sendElems[k] * Np + nd];
    }
}

__global__ void signalPeersKernel(PeerFlags F, unsigned long long epoch) {
    const int i = threadIdx.x;
    if (i < F.n) {
#ifdef DGB_EMULATE
        __atomic_store_n(F.flag[i], epoch, __ATOMIC_RELEASE);
#else
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(F.flag[i]), "l"(epoch) : "memory");
#endif
    }
}

// One thread per peer spins on this rank's flag array; gives up after timeoutNs (a peer that died or ran a different
// number of stages) and reports through *err instead of hanging the GPU.
__global__ void waitPeersKernel(const unsigned long long* flags, PeerWait W, unsigned long long epoch, unsigned long long timeoutNs, int* err) {
    const int i = threadIdx.x;
    if (i < W.n) {
        const unsigned long long* f = flags + W.rank[i];
#ifdef DGB_EMULATE
        if (__atomic_load_n(f, __ATOMIC_ACQUIRE) < epoch) *err = 1 + W.rank[i];  // the emulation runs the ranks one after the other: no spinning
#else
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeoutNs) {
                *err = 1 + W.rank[i];
                __threadfence_system();
                break;
            }
            __nanosleep(200);
        }
#endif
    }
}

}  // namespace

void launchPushHalo(const double* y, int64_t stride, int Np, const int32_t* sendElems, const int32_t* sendPeer, const int32_t* sendSlot,
                    int nSend, const PeerTargets& T, cudaStream_t s, int interleaved) {
    const int64_t tot = 4ll * nSend * Np;
    if (tot <= 0) return;
    const unsigned blocks = (unsigned)std::min<int64_t>((tot + 255) / 256, 148 * 8);
    DGB_LAUNCH(pushHaloKernel, blocks, 256, 0, s, y, stride, Np, sendElems, sendPeer, sendSlot, nSend, T, interleaved);
}

void launchSignalPeers(const PeerFlags& F, unsigned long long epoch, cudaStream_t s) {
    if (F.n > 0) DGB_LAUNCH(signalPeersKernel, 1, 32, 0, s, F, epoch);
}

void launchWaitPeers(const unsigned long long* flags, const PeerWait& W, unsigned long long epoch, unsigned long long timeoutNs, int* err,
                     cudaStream_t s) {
    if (W.n > 0) DGB_LAUNCH(waitPeersKernel, 1, 32, 0, s, flags, W, epoch, timeoutNs, err);
}

}  // namespace dgb
