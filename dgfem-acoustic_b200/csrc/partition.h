// Host-only domain-decomposition planning: which elements a rank owns, in which local
// order, which one-layer halo it needs and what it has to send to whom. Pure index work on the reference's
// connectivity arrays (elFId / fNbrElId); compiled into both libdgb.so (used by dgb_create_partitioned) and
// libdgfront.so (so that the CPU tests can exercise it without a GPU).
#pragma once
#include <stdint.h>

#include <vector>

namespace dgb {

struct PartitionPlan {
    int rank = 0, nranks = 1;
    int Kown = 0, Kinterior = 0, Khalo = 0;
    std::vector<int32_t> localToGlobal;  // [Kown + Khalo] owned (interior first, then halo-adjacent), then halo grouped by owner
    std::vector<int32_t> globalToLocal;  // [K] -1 if neither owned nor halo
    std::vector<int32_t> peers;          // neighbour ranks, ascending
    std::vector<int32_t> recvOffset;     // [npeers+1] halo elements of peer i are local ids Kown+recvOffset[i] .. Kown+recvOffset[i+1]
    std::vector<int32_t> sendOffset;     // [npeers+1]
    std::vector<int32_t> sendElems;      // local ids of owned elements to pack for peer i, ascending global id
};

// elFId [K][Nf], fNbrElId [F][2], elPart [K]
PartitionPlan makePartitionPlan(int K, int Nf, const int32_t* elFId, const int32_t* fNbrElId, const int32_t* elPart, int rank,
                                int nranks);

// Recursive coordinate bisection of element centroids into nparts (any nparts >= 1), balanced to +-1 element.
void partitionRcb(int K, const double* centroids /* [K][3] */, int nparts, int32_t* elPart);

}  // namespace dgb
