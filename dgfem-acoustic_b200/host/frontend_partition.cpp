// C API over the partition planner (csrc/partition.cpp) for the front end and the CPU tests.
#include <cstring>
#include <stdexcept>
#include <vector>

#include "dgfront.h"
#include "partition.h"

struct dgf_plan {
    dgb::PartitionPlan p;
};

extern "C" int dgf_partition_rcb(const dgf_mesh* mesh, int nparts, int32_t* elPart) {
    try {
        const dgb_desc* d = dgf_mesh_desc(mesh);
        const double* xyz = dgf_mesh_node_coords(mesh);
        const int nv = d->dim + 1;
        std::vector<double> c((size_t)d->K * 3, 0.0);
        for (int el = 0; el < d->K; ++el)
            for (int v = 0; v < nv; ++v)  // the first dim+1 nodes of an element are its vertices
                for (int x = 0; x < 3; ++x) c[3 * (size_t)el + x] += xyz[3 * ((size_t)el * d->Np + v) + x] / nv;
        dgb::partitionRcb(d->K, c.data(), nparts, elPart);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// METIS k-way partition of the element dual graph (two elements are adjacent when they share a face), SURVEY.md §8 e1.
// The library is the METIS 5 that ships with the CUDA toolkit (libmetis_static.a, 64-bit idx_t, 32-bit real_t — probed);
// its header is not shipped, so the three entry points are declared here. Built without it, the call reports -2.
#ifdef DGF_HAVE_METIS
extern "C" int METIS_SetDefaultOptions(int64_t* options);
extern "C" int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize, int64_t* adjwgt,
                                   int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* objval, int64_t* part);
#endif

extern "C" int dgf_partition_metis(const dgf_mesh* mesh, int nparts, int32_t* elPart, int64_t* edgeCut) {
#ifdef DGF_HAVE_METIS
    try {
        const dgb_desc* d = dgf_mesh_desc(mesh);
        if (nparts < 1 || !elPart) return -1;
        if (nparts == 1) {
            std::memset(elPart, 0, sizeof(int32_t) * (size_t)d->K);
            if (edgeCut) *edgeCut = 0;
            return 0;
        }
        std::vector<int64_t> xadj((size_t)d->K + 1, 0), adj;
        adj.reserve((size_t)d->K * d->Nf);
        for (int el = 0; el < d->K; ++el) {
            for (int lf = 0; lf < d->Nf; ++lf) {
                const int f = d->elFId[(size_t)el * d->Nf + lf];
                const int a = d->fNbrElId[2 * (size_t)f], b = d->fNbrElId[2 * (size_t)f + 1];
                const int nb = a == el ? b : a;
                if (nb >= 0 && nb != el) adj.push_back(nb);
            }
            xadj[(size_t)el + 1] = (int64_t)adj.size();
        }
        int64_t nv = d->K, ncon = 1, np = nparts, objval = 0;
        std::vector<int64_t> part((size_t)d->K, 0), options(40);
        METIS_SetDefaultOptions(options.data());
        const int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &np, nullptr, nullptr, options.data(),
                                           &objval, part.data());
        if (rc != 1) return -1;  // METIS_OK == 1
        for (int el = 0; el < d->K; ++el) elPart[el] = (int32_t)part[(size_t)el];
        if (edgeCut) *edgeCut = objval;
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
#else
    (void)mesh; (void)nparts; (void)elPart; (void)edgeCut;
    return -2;
#endif
}

extern "C" dgf_plan* dgf_plan_create(const dgf_mesh* mesh, const int32_t* elPart, int rank, int nranks) {
    try {
        const dgb_desc* d = dgf_mesh_desc(mesh);
        auto* pl = new dgf_plan;
        pl->p = dgb::makePartitionPlan(d->K, d->Nf, d->elFId, d->fNbrElId, elPart, rank, nranks);
        return pl;
    } catch (const std::exception&) {
        return nullptr;
    }
}

extern "C" void dgf_plan_free(dgf_plan* plan) { delete plan; }

extern "C" void dgf_plan_sizes(const dgf_plan* plan, int32_t* s) {
    const auto& p = plan->p;
    s[0] = p.Kown; s[1] = p.Kinterior; s[2] = p.Khalo; s[3] = (int32_t)p.peers.size(); s[4] = (int32_t)p.sendElems.size(); s[5] = p.nranks;
}

extern "C" void dgf_plan_arrays(const dgf_plan* plan, int32_t* l2g, int32_t* peers, int32_t* recvOffset, int32_t* sendOffset,
                                int32_t* sendElems) {
    const auto& p = plan->p;
    auto cp = [](int32_t* dst, const std::vector<int32_t>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(int32_t)); };
    cp(l2g, p.localToGlobal); cp(peers, p.peers); cp(recvOffset, p.recvOffset); cp(sendOffset, p.sendOffset); cp(sendElems, p.sendElems);
}
