#include "partition.h"

#include <algorithm>
#include <numeric>
#include <stdexcept>

namespace dgb {

PartitionPlan makePartitionPlan(int K, int Nf, const int32_t* elFId, const int32_t* fNbrElId, const int32_t* elPart, int rank,
                                int nranks) {
    PartitionPlan p;
    p.rank = rank;
    p.nranks = nranks;
    p.globalToLocal.assign(K, -1);
    auto neighbour = [&](int el, int lf) {
        const int f = elFId[(size_t)el * Nf + lf];
        const int a = fNbrElId[2 * (size_t)f], b = fNbrElId[2 * (size_t)f + 1];
        return a == el ? b : a;
    };
    std::vector<int32_t> interior, border;
    std::vector<std::vector<int32_t>> halo(nranks), send(nranks);
    for (int el = 0; el < K; ++el) {
        if (elPart[el] < 0 || elPart[el] >= nranks) throw std::runtime_error("elPart entry out of range");
        if (elPart[el] != rank) continue;
        bool onCut = false;
        for (int lf = 0; lf < Nf; ++lf) {
            const int nb = neighbour(el, lf);
            if (nb < 0 || elPart[nb] == rank) continue;
            onCut = true;
            halo[elPart[nb]].push_back(nb);
            send[elPart[nb]].push_back(el);
        }
        (onCut ? border : interior).push_back(el);
    }
    p.Kinterior = (int)interior.size();
    p.Kown = p.Kinterior + (int)border.size();
    p.localToGlobal = interior;
    p.localToGlobal.insert(p.localToGlobal.end(), border.begin(), border.end());
    for (int l = 0; l < p.Kown; ++l) p.globalToLocal[p.localToGlobal[l]] = l;
    p.recvOffset.push_back(0);
    p.sendOffset.push_back(0);
    for (int r = 0; r < nranks; ++r) {
        auto uniq = [](std::vector<int32_t>& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
        uniq(halo[r]);
        uniq(send[r]);
        if (halo[r].empty() && send[r].empty()) continue;
        p.peers.push_back(r);
        for (int g : halo[r]) {
            p.globalToLocal[g] = (int32_t)p.localToGlobal.size();
            p.localToGlobal.push_back(g);
        }
        for (int g : send[r]) p.sendElems.push_back(p.globalToLocal[g]);
        p.recvOffset.push_back((int32_t)(p.localToGlobal.size() - p.Kown));
        p.sendOffset.push_back((int32_t)p.sendElems.size());
    }
    p.Khalo = (int)p.localToGlobal.size() - p.Kown;
    return p;
}

static void rcbRecurse(std::vector<int32_t>& ids, int lo, int hi, int part0, int nparts, const double* c, int32_t* elPart) {
    if (nparts == 1) {
        for (int i = lo; i < hi; ++i) elPart[ids[i]] = part0;
        return;
    }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = lo; i < hi; ++i)
        for (int x = 0; x < 3; ++x) { mn[x] = std::min(mn[x], c[3 * (size_t)ids[i] + x]); mx[x] = std::max(mx[x], c[3 * (size_t)ids[i] + x]); }
    int ax = 0;
    for (int x = 1; x < 3; ++x) if (mx[x] - mn[x] > mx[ax] - mn[ax]) ax = x;
    const int nl = nparts / 2;
    const int mid = lo + (int)((int64_t)(hi - lo) * nl / nparts);
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int32_t a, int32_t b) {
        const double ca = c[3 * (size_t)a + ax], cb = c[3 * (size_t)b + ax];
        return ca < cb || (ca == cb && a < b);
    });
    rcbRecurse(ids, lo, mid, part0, nl, c, elPart);
    rcbRecurse(ids, mid, hi, part0 + nl, nparts - nl, c, elPart);
}

void partitionRcb(int K, const double* centroids, int nparts, int32_t* elPart) {
    if (nparts < 1) throw std::runtime_error("nparts must be >= 1");
    std::vector<int32_t> ids(K);
    std::iota(ids.begin(), ids.end(), 0);
    rcbRecurse(ids, 0, K, 0, nparts, centroids, elPart);
}

}  // namespace dgb
