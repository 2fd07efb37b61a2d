// Compile-time layout of the tiled DMMA stage kernel (stage_tiled.cu), shared with the host code that lays
// the operators out for it (dgb_api.cu).
#pragma once

namespace dgb {

// Leading dimension (in doubles) for a K-contiguous operand of mma.m8n8k4: a multiple of 4 with ld % 16 in
// {4, 12}, so that the 16 lanes of a half-warp (8 rows x 4 consecutive k) hit 16 distinct 8-byte bank pairs.
__host__ __device__ constexpr int padLd(int k) {
    int ld = (k + 3) / 4 * 4;
    while (ld % 16 != 4 && ld % 16 != 12) ld += 4;
    return ld;
}

template <int P>
struct TetTile {
    static constexpr int NP = (P + 1) * (P + 2) * (P + 3) / 6, NFP = (P + 1) * (P + 2) / 2, NF = 4, NFL = NF * NFP;
    static constexpr int MT = (NP + 7) / 8, NPP = MT * 8;     // m-tiles of 8 rows per parametric direction
    static constexpr int KTQ = (NP + 3) / 4, LDQ = padLd(NP);  // volume contraction: k-tiles of 4, leading dimension
    static constexpr int KTF = (NFL + 3) / 4, LDF = padLd(NFL);  // lift contraction
    static constexpr int OPD = 3 * NPP * LDQ;                  // doubles: [u][NPP][LDQ]  Dw^u, zero padded
    static constexpr int OPL = NPP * LDF;                      // doubles: [NPP][LDF]     -LIFT, zero padded
    static constexpr int OPS = OPD + OPL;
    static constexpr int WARP_DOUBLES = 16 * LDQ + 16 * LDF;   // per-warp staging: 4 elements x 4 fields columns
};

// sizes for the host: returns false if there is no tiled instance for this order
bool tiledLayout(int dim, int order, int* npp, int* ldq, int* ldf);

}  // namespace dgb
