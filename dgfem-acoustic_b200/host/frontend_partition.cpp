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
