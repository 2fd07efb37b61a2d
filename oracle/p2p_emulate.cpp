// Runs the direct halo exchange kernels (dgfem-acoustic_b200/csrc/halo_p2p.cu, the file itself) on the CPU through cuda_emu.h
// for ANY number of ranks in one process — TEST INFRASTRUCTURE ONLY. The hardware run of this exchange had two ranks, i.e. one
// peer per rank; this harness checks the multi-peer bookkeeping (peer index per send element, destination slot in the peer's
// array, flag slots) that dgb_set_option("exchange", 1) sets up in csrc/dgb_api.cu (setupP2P, pushHalo, waitHalo), on the
// partition plans of csrc/partition.cpp: after one push every halo slot of every rank must hold its owner's values.
#define DGB_EMULATE 1
#include "cuda_emu.h"

#include <stdexcept>
#include <string>
#include <vector>

#include "../dgfem-acoustic_b200/csrc/halo_p2p.cu"
#include "../dgfem-acoustic_b200/csrc/partition.cpp"

using namespace dgb;

namespace {
thread_local std::string g_err;
inline double encode(int q, int globalEl, int node) { return q * 1e9 + globalEl * 100.0 + node; }
}  // namespace

extern "C" {
const char* p2e_last_error(void) { return g_err.c_str(); }

// returns the number of wrong halo entries over all ranks (0 = the exchange delivers everything), -1 on error;
// *maxPeers receives the largest number of peers a rank has
int p2e_check(const dgb_desc* d, const int32_t* elPart, int nranks, int* maxPeers) {
    try {
        if (nranks > MAX_PEERS) throw std::runtime_error("too many ranks");
        const int Np = d->Np;
        std::vector<PartitionPlan> plan(nranks);
        std::vector<std::vector<double>> y(nranks);
        std::vector<std::vector<unsigned long long>> flags(nranks, std::vector<unsigned long long>(nranks, 0));
        std::vector<int64_t> stride(nranks);
        std::vector<std::vector<int32_t>> slot0(nranks, std::vector<int32_t>(nranks, -1));  // [r][p]: first slot of p's elements at r
        for (int r = 0; r < nranks; ++r) {
            plan[r] = makePartitionPlan(d->K, d->Nf, d->elFId, d->fNbrElId, elPart, r, nranks);
            const PartitionPlan& P = plan[r];
            stride[r] = (int64_t)(P.Kown + P.Khalo) * Np;
            y[r].assign((size_t)4 * stride[r], -1.0);
            for (int l = 0; l < P.Kown; ++l)
                for (int q = 0; q < 4; ++q)
                    for (int n = 0; n < Np; ++n) y[r][q * stride[r] + (int64_t)l * Np + n] = encode(q, P.localToGlobal[l], n);
            for (size_t i = 0; i < P.peers.size(); ++i) slot0[r][P.peers[i]] = P.Kown + P.recvOffset[i];
        }
        int mp = 0;
        // every rank pushes and signals (setupP2P's tables, pushHalo's arguments)
        for (int r = 0; r < nranks; ++r) {
            const PartitionPlan& P = plan[r];
            mp = std::max(mp, (int)P.peers.size());
            std::vector<int32_t> sendPeer(P.sendElems.size()), sendSlot(P.sendElems.size());
            PeerTargets T{};
            PeerFlags F{};
            F.n = (int)P.peers.size();
            for (size_t i = 0; i < P.peers.size(); ++i) {
                const int peer = P.peers[i];
                if (slot0[peer][r] < 0 && P.sendOffset[i + 1] > P.sendOffset[i]) throw std::runtime_error("plans of two ranks disagree");
                for (int k = P.sendOffset[i]; k < P.sendOffset[i + 1]; ++k) { sendPeer[k] = (int32_t)i; sendSlot[k] = slot0[peer][r] + (k - P.sendOffset[i]); }
                T.arr[i] = y[peer].data();
                T.stride[i] = stride[peer];
                F.flag[i] = flags[peer].data() + r;
            }
            launchPushHalo(y[r].data(), stride[r], Np, P.sendElems.data(), sendPeer.data(), sendSlot.data(), (int)P.sendElems.size(), T, nullptr);
            launchSignalPeers(F, 1ull, nullptr);
        }
        // every rank waits, then its halo is checked
        int wrong = 0;
        for (int r = 0; r < nranks; ++r) {
            const PartitionPlan& P = plan[r];
            PeerWait W{};
            W.n = (int)P.peers.size();
            for (size_t i = 0; i < P.peers.size(); ++i) W.rank[i] = P.peers[i];
            int err = 0;
            launchWaitPeers(flags[r].data(), W, 1ull, 1000000ull, &err, nullptr);
            if (err) ++wrong;
            for (int l = P.Kown; l < P.Kown + P.Khalo; ++l)
                for (int q = 0; q < 4; ++q)
                    for (int n = 0; n < Np; ++n)
                        if (y[r][q * stride[r] + (int64_t)l * Np + n] != encode(q, P.localToGlobal[l], n)) ++wrong;
        }
        if (maxPeers) *maxPeers = mp;
        return wrong;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
}
